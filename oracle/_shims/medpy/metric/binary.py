"""medpy.metric.binary dc/jc/precision/recall, restated from medpy 0.4 semantics
(bool-cast inputs, count_nonzero, ZeroDivisionError -> 0.0 except jc).  Test infrastructure."""
import numpy


def _b(a):
    return numpy.atleast_1d(numpy.asarray(a).astype(bool))


def dc(result, reference):
    r, g = _b(result), _b(reference)
    inter = numpy.count_nonzero(r & g)
    try:
        return 2.0 * inter / float(numpy.count_nonzero(r) + numpy.count_nonzero(g))
    except ZeroDivisionError:
        return 0.0


def jc(result, reference):
    r, g = _b(result), _b(reference)
    return float(numpy.count_nonzero(r & g)) / float(numpy.count_nonzero(r | g))


def precision(result, reference):
    r, g = _b(result), _b(reference)
    tp = numpy.count_nonzero(r & g)
    fp = numpy.count_nonzero(r & ~g)
    try:
        return tp / float(tp + fp)
    except ZeroDivisionError:
        return 0.0


def recall(result, reference):
    r, g = _b(result), _b(reference)
    tp = numpy.count_nonzero(r & g)
    fn = numpy.count_nonzero(~r & g)
    try:
        return tp / float(tp + fn)
    except ZeroDivisionError:
        return 0.0
