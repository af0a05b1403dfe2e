"""Import-level stand-in for gammapy — test infrastructure only.  It lets the reference's
agnpy/fit/gammapy_wrapper.py be imported so that `*SpectralModel.evaluate(energy, **kwargs)` (kwargs ->
ordered parameters, sed / E^2 output) can be run unmodified.  No fitting machinery."""
IS_AGNPY_B200_STUB = True
from . import modeling  # noqa: E402,F401
