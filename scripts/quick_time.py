"""First timings on the GPU box (development aid, not the bench)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch, ctypes as C
from agnpy_b200 import _lib, _core
from tests import cases

torch.cuda.set_device(0)
L = _lib.lib()
print("fp64 peak TFLOP/s:", _lib.fp64_peak(300.0) / 1e12)

def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()

def models_dev(models):
    return torch.from_numpy(models.view(np.float64).reshape(len(models), -1).copy()).cuda()

def time_ec(c, n_models, integ=0, reps=10):
    models = _lib.make_models(n_models)
    for i in range(n_models):
        _core.model_row(models, i, z=c["z"], d_L=c["d_L"], delta_D=c["delta_D"], mu_s=c["mu_s"], R_b=c["R_b"],
                        ne=list(c["ne_args"]) + [0] * (8 - len(c["ne_args"])),
                        tgt=[c["L_disk"], c["xi_line"], c["epsilon_line"], c["R_line"], c["r"] * (1 + 0.01 * i)])
    dm = models_dev(models)
    nu, gamma, mu, phi = dev(c["nu"]), dev(c["gamma"]), dev(c["mu"]), dev(c["phi"])
    out = torch.empty(n_models, nu.numel(), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def run():
        rc = L.agnpy_b200_ec_sed(_lib.EC_BLR, n_models, dm.data_ptr(), _lib.NE_BROKEN_POWERLAW, nu.data_ptr(), nu.numel(),
                                 gamma.data_ptr(), gamma.numel(), mu.data_ptr(), mu.numel(), phi.data_ptr(), phi.numel(),
                                 None, integ, out.data_ptr(), st)
        _lib.check(rc, "ec")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pts = n_models * nu.numel() * gamma.numel() * mu.numel() * phi.numel()
    print(f"EC-BLR integ={integ} models={n_models} nu={nu.numel()} gamma={gamma.numel()}: {ms:.3f} ms/call, "
          f"{pts / ms / 1e6:.1f} Gpts/s, {n_models / ms * 1e3:.1f} SED/s")
    return out

c = cases.case_c2()
time_ec(c, 1)
time_ec(c, 16)
time_ec(c, 64)
time_ec(c, 1, integ=1, reps=3)
c4 = cases.case_c2(n_nu=1000, n_gamma=1000, n_mu=200, n_phi=100)
time_ec(c4, 1, reps=2)
