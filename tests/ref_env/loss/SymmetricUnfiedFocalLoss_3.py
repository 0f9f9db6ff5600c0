"""`trainer.py:26` imports SYM_UIFIED_FOCAL_LOSS from this module and never uses it; the reference repository does not ship the
file.  Placeholder so that the import line succeeds."""


class SYM_UIFIED_FOCAL_LOSS:
    def __init__(self, *a, **kw):
        raise NotImplementedError("placeholder for a module the reference does not ship")
