# Drop-in overlay package (see ../README.md).  This REGULAR package wins over the reference's same-named directory (a namespace
# package: the reference ships no __init__.py) wherever it sits on sys.path; the modules it does not replace (the reference's
# csv_handler, map_generator, batch_data_loader_V2, TverskyLoss, ...) are still found because the reference's directory of the
# same name is appended to this package's search path.
import os as _os
import sys as _sys

_here = _os.path.abspath(_os.path.dirname(__file__))
for _p in list(_sys.path):
    _d = _os.path.abspath(_os.path.join(_p or ".", __name__))
    if _d != _here and _os.path.isdir(_d) and _d not in __path__:
        __path__.append(_d)
