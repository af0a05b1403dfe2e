"""The reference's OWN test files for the hot path, run with the CUDA classes substituted (gpu).

tests/ref_substitution.py replaces Synchrotron, ProtonSynchrotron, SynchrotronSelfCompton,
ExternalCompton and Absorption inside the imported, otherwise unmodified reference (agnpy v0.5.0 from
baseline/_ref) and pytest then collects the reference's test modules as they are: its Blob, spectra,
targets, literature data files and comparison utilities drive this package's kernels.  Runs in a
subprocess (the substitution patches module attributes of the imported reference)."""
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import reference_loader

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

FILES = ["synchrotron/tests/test_synchrotron.py", "synchrotron/tests/test_proton_synchrotron.py",
         "compton/tests/test_compton.py", "absorption/tests/test_absorption.py", "utils/tests/test_utils.py"]
TARGET_FILES = ["targets/tests/test_targets.py"]
# not runnable here: the EBL comparison needs gammapy's own EBL models; of utils/tests/test_utils.py only the
# trapz_loglog test involves this package (the rest tests the reference's own geometry helpers and plots)
NOT_THESE = "not TestEBL and (trapz_log_log or not test_utils)"


def _run(files, targets_too):
    import os
    import re
    ref_root = next(c for c in reference_loader.CANDIDATES if (c / "agnpy" / "__init__.py").exists())
    cmd = [sys.executable, "-m", "pytest", "-p", "tests.ref_substitution", "-q", "-p", "no:cacheprovider",
           "-W", "ignore", "--rootdir", str(ROOT), "-k", NOT_THESE]
    cmd += [str(ref_root / "agnpy" / f) for f in files]
    env = dict(os.environ, AGNPY_B200_SUBSTITUTE_TARGETS="1" if targets_too else "0")
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500, env=env)
    tail = "\n".join(res.stdout.splitlines()[-40:])
    assert res.returncode == 0, f"reference tests on the CUDA classes:\n{tail}\n{res.stderr[-2000:]}"
    launches = re.search(r"agnpy_b200 kernel launches during this session: (\d+)", res.stdout)
    assert launches and int(launches.group(1)) > 100, "the reference's tests did not reach the CUDA kernels"
    return int(re.search(r"(\d+) passed", tail).group(1))


def test_reference_test_files_pass_on_the_cuda_classes():
    """Synchrotron, ProtonSynchrotron, SynchrotronSelfCompton, ExternalCompton, Absorption substituted; the
    reference's Blob, spectra and targets drive them"""
    if not reference_loader.available():
        pytest.skip("the reference install (baseline/_ref) is not here")
    assert _run(FILES, targets_too=False) >= 54 + 81


def test_reference_test_files_pass_with_the_target_classes_substituted_too():
    """in addition CMB, PointSourceBehindJet, SSDisk, SphericalShellBLR, RingDustTorus (u(r), black-body
    SEDs, the container helpers) and the reference's own target tests"""
    if not reference_loader.available():
        pytest.skip("the reference install (baseline/_ref) is not here")
    assert _run(TARGET_FILES + FILES, targets_too=True) >= 96 + 81


def test_reference_calling_code_on_the_cuda_classes_reproduces_the_fixtures():
    """every case of tests/ref_cases.py through the REFERENCE side of the case code -- the reference's Blob,
    spectra classes (passed as classes, as its fit wrappers do), targets, sherpa / gammapy wrapper functions
    and Quantities -- with the CUDA classes substituted: the outputs must be those the unmodified reference
    produced (tests/golden/ref_evaluate.npz) at 1e-9"""
    if not reference_loader.available():
        pytest.skip("the reference install (baseline/_ref) is not here")
    res = subprocess.run([sys.executable, "-m", "tests.ref_substitution"], cwd=ROOT, capture_output=True,
                         text=True, timeout=1500)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    lines = [ln.split() for ln in res.stdout.splitlines() if ln and not ln.startswith("#")]
    assert len(lines) >= 60
    for name, dev, *flag in lines:
        assert flag[:2] == ["zeros", "ok"], f"{name}: exact zeros differ"
        assert float(dev) <= 1e-9, f"{name}: {dev}"
    assert any(name.startswith("fit_sherpa") for name, *_ in lines)


def test_reference_documentation_snippets_run_on_the_cuda_classes():
    """docs/snippets/{blob,synchro,ssc,ec,absorption,targets,p_synchro}_snippet.py of the reference -- the
    usage its documentation shows -- executed unchanged with the compute and target classes substituted"""
    if not reference_loader.available():
        pytest.skip("the reference install (baseline/_ref) is not here")
    ref_root = next(c for c in reference_loader.CANDIDATES if (c / "agnpy" / "__init__.py").exists())
    if not (ref_root / "docs" / "snippets" / "ec_snippet.py").exists():
        pytest.skip("the documentation snippets are not part of this reference install")
    res = subprocess.run([sys.executable, "-m", "tests.run_reference_snippets"], cwd=ROOT, capture_output=True,
                         text=True, timeout=1500)
    ran = [ln for ln in res.stdout.splitlines() if ln.startswith("SNIPPET")]
    assert res.returncode == 0 and len(ran) == 7 and all(" ran, " in ln for ln in ran), \
        "\n".join(ran) + res.stderr[-3000:]
    launches = {ln.split()[1]: int(ln.split()[3]) for ln in ran}
    assert all(launches[k] > 0 for k in ("synchro_snippet", "ssc_snippet", "ec_snippet", "absorption_snippet"))
