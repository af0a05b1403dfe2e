"""Per-case deviation of the CUDA path from the outputs of the unmodified reference
(tests/golden/ref_evaluate.npz), through the drop-in API: prints one line per case and the maximum.
python scripts/parity_report.py > profiles/r2_parity_vs_reference.txt   (on a GPU box)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from agnpy_b200 import _lib
from tests import cases, ref_cases

store = dict(np.load(cases.GOLDEN / "ref_evaluate.npz"))
store.pop("__meta__", None)
A = ref_cases.ProductSide()
worst = 0.0
print("# case, n values, max |cuda/reference - 1| (exact zeros must agree), mode")
for name in ref_cases.CASES:
    strict = "_ssa" in name
    with _lib.strict_reference_mode(strict):
        got = np.asarray(ref_cases.run(name, "ref", A, dict(store)), dtype=np.float64)
    dev, zeros_ok = cases.max_rel_dev(got, store[name])
    worst = max(worst, dev if not strict else 0.0)
    print(f"{name:36s} {got.size:6d} {dev:10.2e} {'zeros ok' if zeros_ok else 'ZEROS DIFFER'}"
          f"{'  (strict-reference switch: literal SSA attenuation, conditioning 2e-9)' if strict else ''}")
print(f"# maximum over the cases in default mode: {worst:.2e}")
