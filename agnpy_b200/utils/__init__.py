"""Import paths of the reference's helper modules: `agnpy.utils.math`, `agnpy.utils.conversion`."""
from . import conversion, math  # noqa: F401
