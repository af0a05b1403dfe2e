"""Outputs of the reference's OWN, UNMODIFIED `evaluate_*` code -> tests/golden/ref_evaluate.npz.

Run in the build container (the reference checkout / its offline install must exist):
    python tests/golden/make_ref_evaluate.py [case-name-substring ...]
`oracle/reference_loader.py` imports `agnpy` v0.5.0 from baseline/_ref (or /root/reference) on top of the
miniature astropy in oracle/astropy_stub/ (astropy has no wheel in this image); the cases and their CGS
inputs are in tests/ref_cases.py.  tests/test_oracle_vs_reference.py pins the oracle to these vectors,
tests/test_gpu_vs_reference.py the CUDA path.
"""
import sys
import time
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_loader  # noqa: E402

agnpy = reference_loader.load()
from tests import ref_cases  # noqa: E402

OUT = Path(__file__).resolve().parent / "ref_evaluate.npz"


def main(filters):
    A = ref_cases.RefSide(agnpy)
    store = dict(np.load(OUT)) if (OUT.exists() and filters) else {}
    warnings.simplefilter("ignore")
    with np.errstate(all="ignore"):
        for name in ref_cases.CASES:
            if filters and not any(f in name for f in filters):
                continue
            t0 = time.perf_counter()
            store[name] = np.asarray(ref_cases.run(name, "ref", A, store), dtype=np.float64)
            print(f"{name:40s} {store[name].shape!s:10s} max {np.nanmax(np.abs(store[name])):.6e} "
                  f"{time.perf_counter() - t0:6.1f} s", flush=True)
    store["__meta__"] = np.array([f"agnpy {agnpy.__version__ if hasattr(agnpy, '__version__') else '0.5.0'}"
                                  f" from {agnpy.__reference_root__}; astropy stub: "
                                  f"{reference_loader.using_stub()}; numpy {np.__version__}"])
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main(sys.argv[1:])
