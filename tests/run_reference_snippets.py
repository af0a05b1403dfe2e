"""Run the reference's documentation snippets (docs/snippets/*.py of the installed reference) with the CUDA
classes substituted (tests/ref_substitution.py; the target classes too): the user code the reference
documents must run unchanged.  `python -m tests.run_reference_snippets [name ...]`; test infrastructure."""
import os
import runpy
import sys
import traceback
import warnings
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("AGNPY_B200_SUBSTITUTE_TARGETS", "1")

SNIPPETS = ("blob_snippet", "synchro_snippet", "ssc_snippet", "ec_snippet", "absorption_snippet",
            "targets_snippet", "p_synchro_snippet")


def main(names):
    from oracle import reference_loader
    from tests import ref_substitution
    warnings.simplefilter("ignore")
    ref_substitution.substitute()
    from agnpy_b200 import _lib
    root = next(c for c in reference_loader.CANDIDATES if (c / "agnpy" / "__init__.py").exists())
    ok = True
    for name in names:
        n0 = _lib.launch_count()
        try:
            runpy.run_path(str(root / "docs" / "snippets" / f"{name}.py"), run_name="__main__")
            print(f"SNIPPET {name:20s} ran, {_lib.launch_count() - n0} kernel launches", flush=True)
        except Exception:
            ok = False
            print(f"SNIPPET {name:20s} FAILED", flush=True)
            traceback.print_exc()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:] or SNIPPETS))
