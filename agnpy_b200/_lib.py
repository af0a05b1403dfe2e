"""ctypes binding of libagnpy_b200.so (the C ABI of include/agnpy_b200.h).

The product path has no CPU fallback: if the shared library is missing, or there is no
CUDA device, calls raise.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libagnpy_b200.so"
if os.environ.get("AGNPY_B200_CHECKED") == "1":   # debugging build with index checks in the kernels
    LIB_PATH = _PKG / "libagnpy_b200_checked.so"

# constants of include/agnpy_b200.h
NE_POWERLAW, NE_BROKEN_POWERLAW, NE_LOG_PARABOLA = 0, 1, 2
NE_EXP_CUTOFF_POWERLAW, NE_EXP_CUTOFF_BROKEN_POWERLAW, NE_TABLE = 3, 4, 5
INTEG_TRAPZ, INTEG_LOGLOG, INTEG_TRAPZ_PREFIX = 0, 1, 2
INTEG_FOLD_PHI = 0x100  # flag: the phi grid is mirror-symmetric (include/agnpy_b200.h)
EC_ISO_MONO, EC_PS_BEHIND_JET, EC_SS_DISK, EC_BLR, EC_DT = 0, 1, 2, 3, 4
TAU_SS_DISK, TAU_BLR, TAU_DT = 0, 1, 2
TAU_PS_BEHIND_BLOB, TAU_PS_BEHIND_BLOB_MU_S, TAU_BLR_MU_S, TAU_DT_MU_S = 3, 4, 5, 6
BB_DISK, BB_DISK_NORM, BB_DT, BB_DT_NORM = 0, 1, 2, 3

EXPORTS = [
    "agnpy_b200_version", "agnpy_b200_last_error", "agnpy_b200_launch_count",
    "agnpy_b200_fp64_peak", "agnpy_b200_kernel_timing", "agnpy_b200_kernel_timing_read",
    "agnpy_b200_strict_reference", "agnpy_b200_trapz_loglog", "agnpy_b200_trapz_loglog_host",
    "agnpy_b200_synch_sed", "agnpy_b200_synch_sed_host",
    "agnpy_b200_ssc_sed", "agnpy_b200_ssc_sed_host",
    "agnpy_b200_ec_sed", "agnpy_b200_ec_sed_host",
    "agnpy_b200_tau", "agnpy_b200_tau_host",
    "agnpy_b200_tau_synch", "agnpy_b200_tau_synch_host",
    "agnpy_b200_proton_synch_sed", "agnpy_b200_proton_synch_sed_host",
    "agnpy_b200_ssc_loss_rate", "agnpy_b200_ssc_loss_rate_host",
    "agnpy_b200_bb_sed", "agnpy_b200_bb_sed_host",
    "agnpy_b200_target_u", "agnpy_b200_target_u_host",
    "agnpy_b200_ebl_absorption", "agnpy_b200_ebl_absorption_host",
]


class Model(C.Structure):
    """agnpy_model_t"""
    _fields_ = [("z", C.c_double), ("d_L", C.c_double), ("delta_D", C.c_double),
                ("mu_s", C.c_double), ("B", C.c_double), ("R_b", C.c_double),
                ("ne", C.c_double * 8), ("tgt", C.c_double * 8)]


MODEL_DOUBLES = C.sizeof(Model) // 8  # 22
MODEL_DTYPE = np.dtype([("z", "f8"), ("d_L", "f8"), ("delta_D", "f8"), ("mu_s", "f8"),
                        ("B", "f8"), ("R_b", "f8"), ("ne", "f8", (8,)), ("tgt", "f8", (8,))])
assert MODEL_DTYPE.itemsize == C.sizeof(Model)


class AgnpyB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library (built in-tree by `python agnpy_b200/build.py` /
    `__graft_entry__.build()`); raises if it is missing -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise AgnpyB200Error(
            f"{LIB_PATH} not found: build it with `python agnpy_b200/build.py` "
            "(agnpy_b200 has no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    dp, ip, vp = C.c_void_p, C.c_int, C.c_void_p
    L.agnpy_b200_version.restype = C.c_char_p
    L.agnpy_b200_last_error.restype = C.c_char_p
    L.agnpy_b200_launch_count.restype = C.c_longlong
    L.agnpy_b200_fp64_peak.argtypes = [C.c_double, C.POINTER(C.c_double), vp]
    synch = [ip, dp, ip, dp, ip, dp, ip, dp, dp, ip, ip, dp, dp]
    L.agnpy_b200_synch_sed.argtypes = synch + [vp]
    L.agnpy_b200_synch_sed_host.argtypes = synch
    ssc = [ip, dp, ip, dp, ip, dp, ip, dp, ip, dp, dp, ip, ip, dp]
    L.agnpy_b200_ssc_sed.argtypes = ssc + [vp]
    L.agnpy_b200_ssc_sed_host.argtypes = ssc
    ec = [ip, ip, dp, ip, dp, ip, dp, ip, dp, ip, dp, ip, dp, ip, dp]
    L.agnpy_b200_ec_sed.argtypes = ec + [vp]
    L.agnpy_b200_ec_sed_host.argtypes = ec
    tau = [ip, ip, dp, dp, ip, dp, ip, dp, ip, dp, ip, dp]
    L.agnpy_b200_tau.argtypes = tau + [vp]
    L.agnpy_b200_tau_host.argtypes = tau
    psy = [ip, dp, ip, dp, ip, dp, ip, dp, ip, dp]
    L.agnpy_b200_proton_synch_sed.argtypes = psy + [vp]
    L.agnpy_b200_proton_synch_sed_host.argtypes = psy
    lr = [ip, dp, ip, dp, ip, dp, ip, dp, ip, dp, dp]
    L.agnpy_b200_ssc_loss_rate.argtypes = lr + [vp]
    L.agnpy_b200_ssc_loss_rate_host.argtypes = lr
    tsy = [ip, dp, ip, dp, ip, dp, ip, dp, dp, ip, C.c_double, ip, dp]
    L.agnpy_b200_tau_synch.argtypes = tsy + [vp]
    L.agnpy_b200_tau_synch_host.argtypes = tsy
    bb = [ip, ip, dp, dp, ip, dp, ip, ip, dp]
    L.agnpy_b200_bb_sed.argtypes = bb + [vp]
    L.agnpy_b200_bb_sed_host.argtypes = bb
    tu = [ip, ip, dp, dp, dp, ip, ip, dp]
    L.agnpy_b200_target_u.argtypes = tu + [vp]
    L.agnpy_b200_target_u_host.argtypes = tu
    ebl = [ip, dp, dp, ip, dp, ip, dp, ip, dp, ip, dp]
    L.agnpy_b200_ebl_absorption.argtypes = ebl + [vp]
    L.agnpy_b200_ebl_absorption_host.argtypes = ebl
    tl = [dp, dp, C.c_longlong, ip, ip, dp]
    L.agnpy_b200_trapz_loglog.argtypes = tl + [vp]
    L.agnpy_b200_trapz_loglog_host.argtypes = tl
    L.agnpy_b200_strict_reference.argtypes = [C.c_int]
    L.agnpy_b200_kernel_timing.argtypes = [C.c_int]
    L.agnpy_b200_kernel_timing_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    for name in EXPORTS[3:]:
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().agnpy_b200_last_error().decode()
        raise AgnpyB200Error(f"{what} failed (code {rc}): {msg}")


def strict_reference(enable=None):
    """Literal-reference switch of include/agnpy_b200.h (SSA attenuation, sigma(s)): set it (True /
    False) or just query (None); returns the previous setting.  Also usable as
    `with strict_reference_mode(): ...`."""
    return bool(lib().agnpy_b200_strict_reference(-1 if enable is None else int(bool(enable))))


class strict_reference_mode:
    """context manager: evaluate the reference's literal formulas inside the block"""

    def __init__(self, enable=True):
        self.enable = enable

    def __enter__(self):
        self.prev = strict_reference(self.enable)
        return self

    def __exit__(self, *exc):
        strict_reference(self.prev)
        return False


def launch_count():
    return int(lib().agnpy_b200_launch_count())


def kernel_timing(enable):
    check(lib().agnpy_b200_kernel_timing(int(bool(enable))), "kernel_timing")


def kernel_timing_read():
    """(total ms, launches) of the dominant kernels timed since kernel_timing(True)"""
    ms, n = C.c_double(0.0), C.c_longlong(0)
    check(lib().agnpy_b200_kernel_timing_read(C.byref(ms), C.byref(n)), "kernel_timing_read")
    return ms.value, n.value


def fp64_peak(ms=200.0, stream=None):
    """measured FP64 FMA peak of the current device in FLOP/s"""
    out = C.c_double(0.0)
    check(lib().agnpy_b200_fp64_peak(float(ms), C.byref(out), stream), "fp64_peak")
    return out.value


# ---- helpers to pass numpy arrays / torch tensors ------------------------------------------
def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def hptr(a):
    """host pointer of a contiguous float64 / model ndarray (or None)"""
    return None if a is None else a.ctypes.data


def dptr(t):
    """device pointer of a contiguous CUDA torch tensor (or None)"""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise AgnpyB200Error("device entry points need contiguous CUDA tensors")
    return t.data_ptr()


def make_models(n):
    return np.zeros(n, dtype=MODEL_DTYPE)
