"""astropy.table is only used by agnpy.fit.data (needs sherpa/gammapy as well) — not available."""


class Table:
    def __init__(self, *a, **k):
        raise ImportError("astropy.table is not part of the agnpy_b200 test stub")
