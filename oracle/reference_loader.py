"""Import the UNMODIFIED reference package (`agnpy` v0.5.0) — TEST INFRASTRUCTURE ONLY.

The reference is pure Python on numpy + astropy; astropy has no wheel in this image, so
`oracle/astropy_stub/` supplies a miniature `astropy` (units with astropy's semantics, CODATA-2018
constants, Planck18 `Distance`, `BlackBody`, a FITS table reader) and a no-op `matplotlib`.  With it the
reference's own test-suite passes (3032 of 3034 non-fit tests; the two failures need a real matplotlib rc
loader), see oracle/astropy_stub/README.md.

`load()` returns the imported `agnpy` module from, in order,
  1. `baseline/_ref/` (the offline `pip install --no-deps --target baseline/_ref` of the reference;
     git-ignored but shipped to the GPU box), or
  2. `/root/reference` (exists only in the build container),
and raises ImportError when neither is there.  If a real astropy is importable it is used instead of the
stub.  Only tests/, tests/golden/make_ref_evaluate.py, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of bench.py may call this; nothing under `agnpy_b200/` does.
"""
import importlib
import importlib.util
import logging
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
STUB = Path(__file__).resolve().parent / "astropy_stub"
CANDIDATES = (ROOT / "baseline" / "_ref", Path("/root/reference"))

_cache = {}


def using_stub():
    astropy = sys.modules.get("astropy")
    return bool(getattr(astropy, "IS_AGNPY_B200_STUB", False))


def _ensure_astropy():
    if "astropy" in sys.modules:
        return
    if importlib.util.find_spec("astropy") is None:
        sys.path.insert(0, str(STUB))
    elif importlib.util.find_spec("matplotlib") is None:
        sys.path.append(str(STUB))          # real astropy, stub matplotlib only
    importlib.import_module("astropy")


def available():
    return any((c / "agnpy" / "__init__.py").exists() for c in CANDIDATES)


def load():
    """the reference's `agnpy` module, imported from its own unmodified sources"""
    if "agnpy" in _cache:
        return _cache["agnpy"]
    for cand in CANDIDATES:
        if (cand / "agnpy" / "__init__.py").exists():
            break
    else:
        raise ImportError("the reference is neither installed under baseline/_ref nor present at "
                          "/root/reference")
    _ensure_astropy()
    if importlib.util.find_spec("matplotlib") is None:
        sys.path.append(str(STUB))
    already = sys.modules.get("agnpy")
    if already is not None and not str(getattr(already, "__file__", "")).startswith(str(cand)):
        raise ImportError(f"another `agnpy` is already imported from {already.__file__}")
    sys.path.insert(0, str(cand))
    try:
        logging.getLogger("agnpy").setLevel(logging.ERROR)   # "sherpa and gammapy are not installed"
        mod = importlib.import_module("agnpy")
    finally:
        sys.path.remove(str(cand))
    mod.__reference_root__ = str(cand)
    _cache["agnpy"] = mod
    return mod
