"""pytest plugin: run the REFERENCE'S OWN test files with the CUDA classes substituted.

    python -m pytest -p tests.ref_substitution <reference>/agnpy/synchrotron/tests/test_synchrotron.py ...

The unmodified reference is imported (oracle/reference_loader.py: baseline/_ref or /root/reference, on the
miniature astropy of the test infrastructure) and the compute classes of the hot path are replaced in its
package namespaces before the reference's test modules are collected:

    agnpy.synchrotron.Synchrotron, .ProtonSynchrotron  -> agnpy_b200.Synchrotron, .ProtonSynchrotron
    agnpy.compton.SynchrotronSelfCompton, .ExternalCompton -> agnpy_b200....
    agnpy.absorption.Absorption                          -> agnpy_b200.Absorption

Everything else the tests touch stays the reference's own: Blob, the spectra classes and their
normalisation class methods, the targets, the literature data files, the comparison utilities -- the scenario
of INTEGRATION.md, where a maintainer keeps the package and swaps the numpy bodies for the C ABI.
`python -m tests.ref_substitution` runs the cases of tests/ref_cases.py the same way (run_reference_cases) and
prints one line per case.  Test infrastructure: used by tests/test_reference_suite_on_cuda.py (gpu)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import reference_loader  # noqa: E402

SUBSTITUTED = {
    "agnpy.synchrotron": ("Synchrotron", "ProtonSynchrotron"),
    "agnpy.compton": ("SynchrotronSelfCompton", "ExternalCompton"),
    "agnpy.absorption": ("Absorption",),
}
if os.environ.get("AGNPY_B200_SUBSTITUTE_TARGETS") == "1":   # the "next" row: u(r), black-body SEDs
    SUBSTITUTED["agnpy.targets"] = ("CMB", "PointSourceBehindJet", "SSDisk", "SphericalShellBLR", "RingDustTorus")
# the module files that define the classes inside those packages
_INNER = {
    "Synchrotron": "agnpy.synchrotron.synchrotron", "ProtonSynchrotron": "agnpy.synchrotron.proton_synchrotron",
    "SynchrotronSelfCompton": "agnpy.compton.synchrotron_self_compton",
    "ExternalCompton": "agnpy.compton.external_compton", "Absorption": "agnpy.absorption.targets_absorption",
    "CMB": "agnpy.targets.targets", "PointSourceBehindJet": "agnpy.targets.targets", "SSDisk": "agnpy.targets.targets",
    "SphericalShellBLR": "agnpy.targets.targets", "RingDustTorus": "agnpy.targets.targets",
}


def substitute():
    import importlib

    import agnpy_b200 as ab
    import agnpy_b200.units as bu
    agnpy = reference_loader.load()
    # results leave the package as Quantities of the units module the caller uses: the astropy branch of
    # the boundary (agnpy_b200/units.py: wrap), here on the astropy the reference was imported with
    import astropy.units as au
    bu._astropy_units, bu.HAVE_ASTROPY = au, True
    done = []
    # the log-log integrator as a callable (utils/math.py:55-119): evaluated on the device
    ref_math = importlib.import_module("agnpy.utils.math")
    ref_math.trapz_loglog = ab.trapz_loglog
    done.append("agnpy.utils.math.trapz_loglog")
    for pkg, names in SUBSTITUTED.items():
        mod = importlib.import_module(pkg)
        for name in names:
            cuda_cls = getattr(ab, name)
            for target in (mod, agnpy, importlib.import_module(_INNER[name])):
                if hasattr(target, name):
                    setattr(target, name, cuda_cls)
            done.append(f"{pkg}.{name}")
    # the fit wrappers imported the compute classes by name: the reference's own sherpa / gammapy wrapper
    # code then drives the CUDA classes (targets stay whatever the packages above hold)
    for wrapper in ("agnpy.fit.sherpa_wrapper", "agnpy.fit.gammapy_wrapper"):
        try:
            w = importlib.import_module(wrapper)
        except Exception:      # needs the sherpa / gammapy shells of the test infrastructure
            continue
        for names in SUBSTITUTED.values():
            for name in names:
                if hasattr(w, name):
                    setattr(w, name, getattr(ab, name))
        done.append(wrapper)
    return done


def run_reference_cases(names=None):
    """every case of tests/ref_cases.py through the REFERENCE side of the case code -- the reference's Blob,
    spectra, targets, fit wrappers and (stand-in) astropy Quantities -- with the CUDA classes substituted,
    against the fixtures the unmodified reference produced: {case: (max rel dev, zeros agree)}"""
    import numpy as np

    from agnpy_b200 import _lib
    from tests import cases, ref_cases
    agnpy = reference_loader.load()
    substitute()
    store = dict(np.load(cases.GOLDEN / "ref_evaluate.npz"))
    store.pop("__meta__", None)
    A = ref_cases.RefSide(agnpy)
    out = {}
    for name in (names or ref_cases.CASES):
        with _lib.strict_reference_mode("_ssa" in name), np.errstate(all="ignore"):
            got = np.asarray(ref_cases.run(name, "ref", A, dict(store)), dtype=np.float64)
        out[name] = cases.max_rel_dev(got, store[name])
    return out


def pytest_configure(config):
    done = substitute()
    config._agnpy_b200_substituted = done


def pytest_report_header(config):
    return ["agnpy_b200 substituted for: " + ", ".join(getattr(config, "_agnpy_b200_substituted", []))]


def pytest_terminal_summary(terminalreporter):
    from agnpy_b200 import _lib
    terminalreporter.write_line(f"agnpy_b200 kernel launches during this session: {_lib.launch_count()}")


if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    res = run_reference_cases(sys.argv[1:] or None)
    from agnpy_b200 import _lib
    worst = 0.0
    for name, (dev, zeros_ok) in res.items():
        worst = max(worst, dev)
        print(f"{name:36s} {dev:10.2e} {'zeros ok' if zeros_ok else 'ZEROS DIFFER'}")
    print(f"# {len(res)} cases through the reference's own calling code on the CUDA classes: max deviation "
          f"{worst:.2e}, {_lib.launch_count()} kernel launches")
