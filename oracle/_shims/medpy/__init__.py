# Stand-in for `medpy` (absent from this image).  Test infrastructure only: lets
# oracle/make_golden.py import the reference's scripts/validation_functions.py.
