"""Import path of the reference: `from agnpy.emission_regions import Blob` (emission_regions/blob.py)."""
from .targets import Blob

__all__ = ["Blob"]
