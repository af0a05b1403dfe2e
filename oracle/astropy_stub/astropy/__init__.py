"""Miniature astropy for running the unmodified reference in this image — test infrastructure only.
See oracle/astropy_stub/README.md."""
__version__ = "7.0.0+agnpy_b200.stub"
IS_AGNPY_B200_STUB = True
