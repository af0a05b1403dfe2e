"""Batched fit-level model: the components the reference's fit wrappers add up per
parameter vector (agnpy/fit/sherpa_wrapper.py:51-75, 228-325; gammapy_wrapper.py:161-181,
284-357), evaluated for a whole batch of parameter vectors -- an MCMC ensemble, a likelihood
scan, the trial points of an optimiser -- in one pass, sharded over the GPUs of a box.

Parameter vectors follow the sherpa wrapper's flat order:
  SSC scenario        [*spectral, z, delta_D, log10_B, t_var]
  EC BLR+DT scenario  [*spectral, z, delta_D, log10_B, t_var, mu_s, log10_r, log10_L_disk, M_BH,
                       m_dot, R_in, R_out, xi_line, lambda_line, R_line, xi_dt, T_dt, R_dt]
with the spectral block in `fit/core.py:92-113` order and log10 scaling
(`_scale_spectral_parameters`, sherpa_wrapper.py:29-48): log10_k first, log10_gamma_min /
log10_gamma_max last, break / pivot / cutoff Lorentz factors in log10.
As in the wrappers, Synch and SSC integrate over utils/math.py:6 `gamma_e_to_integrate`
(200 points) and the EC terms over logspace(1, 9, 300) (sherpa_wrapper.py:25-26).

The luminosity distance is an input (`d_L`, cm, one per vector): the reference derives it from z
with astropy's cosmology per call (sherpa_wrapper.py:63), which is outside the hot path.
The EC scenario also adds the two thermal terms of the wrappers: the normalised multi-T
black body of the disk and the normalised black body of the torus (sherpa_wrapper.py:283-288).
"""
import numpy as np

from . import _lib, constants as const, grids
from .batch import SedBatch, sharded_evaluate  # GraphedSum is imported lazily in graphed()

# position (from the end of the spectral block) of the log10-scaled Lorentz factors besides
# gamma_min / gamma_max, sherpa_wrapper.py:41-48
_LOG10_EXTRA = {"PowerLaw": (), "BrokenPowerLaw": (-3,), "LogParabola": (-3,),
                "ExpCutoffPowerLaw": (-3,), "ExpCutoffBrokenPowerLaw": (-3, -4)}
_N_SPECTRAL = {"PowerLaw": 4, "BrokenPowerLaw": 6, "LogParabola": 6, "ExpCutoffPowerLaw": 5,
               "ExpCutoffBrokenPowerLaw": 7}
gamma_to_integrate_ec = np.logspace(1, 9, 300)  # sherpa_wrapper.py:25-26


def scale_spectral_parameters(spectral, ne_kind):
    """vectorised `_scale_spectral_parameters` (sherpa_wrapper.py:29-48): [B, n] -> [B, n]"""
    out = np.array(spectral, dtype=np.float64, copy=True)
    out[:, 0] = 10 ** out[:, 0]
    out[:, -2] = 10 ** out[:, -2]
    out[:, -1] = 10 ** out[:, -1]
    for pos in _LOG10_EXTRA[ne_kind]:
        out[:, pos] = 10 ** out[:, pos]
    return out


def build_models(pars, d_L, ne_kind, scenario):
    """flat fit parameter vectors [B, n_par] -> dict of MODEL_DTYPE arrays per component"""
    pars = np.atleast_2d(np.asarray(pars, dtype=np.float64))
    ns = _N_SPECTRAL[ne_kind]
    n_expected = ns + (4 if scenario == "ssc" else 17)
    if pars.shape[1] != n_expected:
        raise ValueError(f"{scenario} scenario with {ne_kind} takes {n_expected} parameters, "
                         f"got {pars.shape[1]}")
    B = pars.shape[0]
    d_L = np.broadcast_to(np.asarray(d_L, dtype=np.float64), (B,))
    spectral = scale_spectral_parameters(pars[:, :ns], ne_kind)
    z, delta_D, log10_B, t_var = (pars[:, ns + i] for i in range(4))
    base = _lib.make_models(B)
    base["z"], base["d_L"], base["delta_D"] = z, d_L, delta_D
    base["B"] = 10 ** log10_B
    base["R_b"] = (const.c * t_var * delta_D) / (1 + z)  # sherpa_wrapper.py:64
    base["mu_s"] = 1.0
    base["ne"][:, :ns] = spectral
    out = {"synch": base, "ssc": base}
    if scenario == "ssc":
        return out
    (mu_s, log10_r, log10_L_disk, M_BH, m_dot, R_in, R_out, xi_line, lambda_line, R_line, xi_dt,
     T_dt, R_dt) = (pars[:, ns + 4 + i] for i in range(13))
    r, L_disk = 10 ** log10_r, 10 ** log10_L_disk
    epsilon_line = const.h * const.c / (lambda_line * 1e-8) / const.mec2  # :270-272 (Angstrom)
    epsilon_dt = 2.7 * (const.k_B * T_dt / const.mec2)                    # :275
    blr = base.copy()
    blr["mu_s"] = mu_s
    blr["tgt"][:, 0], blr["tgt"][:, 1], blr["tgt"][:, 2] = L_disk, xi_line, epsilon_line
    blr["tgt"][:, 3], blr["tgt"][:, 4] = R_line, r
    dt = base.copy()
    dt["mu_s"] = mu_s
    dt["tgt"][:, 0], dt["tgt"][:, 1], dt["tgt"][:, 2] = L_disk, xi_dt, epsilon_dt
    dt["tgt"][:, 3], dt["tgt"][:, 4] = R_dt, r
    bb_disk = base.copy()   # SSDisk.evaluate_multi_T_bb_norm_sed(nu, z, L_disk, M_BH, m_dot, R_in, R_out, d_L)
    bb_disk["mu_s"] = 1.0   # the wrappers use the default mu_s = 1 for the disk SED
    for k, v in enumerate((M_BH, m_dot, R_in, R_out, L_disk)):
        bb_disk["tgt"][:, k] = v
    bb_dt = base.copy()     # RingDustTorus.evaluate_bb_norm_sed(nu, z, xi_dt * L_disk, T_dt, R_dt, d_L)
    for k, v in enumerate((T_dt, R_dt, xi_dt * L_disk)):
        bb_dt["tgt"][:, k] = v
    out.update(ec_blr=blr, ec_dt=dt, bb_disk=bb_disk, bb_dt=bb_dt)
    return out


class FitSedBatch:
    """nu F_nu of the non-thermal components for batches of fit parameter vectors.

        model = FitSedBatch(nu_hz, "BrokenPowerLaw", scenario="ec_blr_dt")
        sed = model(pars, d_L)        # [B, n_nu] erg cm-2 s-1, summed over components

    With torch.distributed initialised every rank evaluates its shard of the batch and ONE
    all-gather joins the summed SEDs."""

    def __init__(self, nu, ne_kind="BrokenPowerLaw", scenario="ec_blr_dt", ssa=False,
                 integrator="trapz", device=None, group=None, ebl=None):
        if scenario not in ("ssc", "ec_blr_dt"):
            raise ValueError("scenario must be 'ssc' or 'ec_blr_dt'")
        self.ne_kind, self.scenario, self.group = ne_kind, scenario, group
        nu = np.ravel(np.asarray(nu, dtype=np.float64))
        kw = dict(device=device, integrator=integrator)
        # the wrappers ignore `ssa` for the EC components
        self.plans = {
            "synch": SedBatch("synch", nu, ne_kind, gamma=grids.gamma_e_to_integrate, ssa=ssa, **kw),
            "ssc": SedBatch("ssc", nu, ne_kind, gamma=grids.gamma_e_to_integrate, ssa=ssa, **kw),
        }
        if scenario == "ec_blr_dt":
            self.plans["ec_blr"] = SedBatch("ec_blr", nu, ne_kind, gamma=gamma_to_integrate_ec, **kw)
            self.plans["ec_dt"] = SedBatch("ec_dt", nu, ne_kind, gamma=gamma_to_integrate_ec, **kw)
            self.plans["bb_disk"] = SedBatch("bb_disk_norm", nu, ne_kind, device=device)
            self.plans["bb_dt"] = SedBatch("bb_dt_norm", nu, ne_kind, device=device)
        self.n_nu = nu.size
        self.device = self.plans["synch"].device
        self.torch = self.plans["synch"].torch
        # optional EBL attenuation (agnpy_b200.ebl.EBL): table resident on the device, applied to
        # the summed SEDs of each shard in place with the per-vector redshifts
        self.ebl = None
        if ebl is not None:
            dev = lambda a: self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
            self.ebl = dict(e=dev(ebl.energy_ref), z=dev(ebl.z_ref), v=dev(ebl.values_ref.T))

    def _local_components(self, models, lo, hi):
        return {name: plan._local(models[name][lo:hi]).clone() for name, plan in self.plans.items()}

    def components(self, pars, d_L):
        """per-component SEDs of the whole batch on this rank (no sharding), as numpy arrays"""
        models = build_models(pars, d_L, self.ne_kind, self.scenario)
        n = len(models["synch"])
        with self.torch.cuda.device(self.device):
            out = {k: v.cpu().numpy() for k, v in self._local_components(models, 0, n).items()}
        return out

    def evaluate_device(self, pars, d_L):
        pars = np.atleast_2d(np.asarray(pars, dtype=np.float64))
        d_L = np.broadcast_to(np.asarray(d_L, dtype=np.float64), (pars.shape[0],))
        index = np.arange(pars.shape[0])

        def local(idx):
            # only this rank's shard of the parameter table is turned into model rows
            lo, hi = int(idx[0]), int(idx[-1]) + 1
            models = build_models(pars[lo:hi], d_L[lo:hi], self.ne_kind, self.scenario)
            lo, hi = 0, hi - lo
            total = None
            for name, plan in self.plans.items():
                part = plan._local(models[name][lo:hi])
                total = part.clone() if total is None else total.add_(part)
            if self.ebl is not None:
                z_dev = self.torch.from_numpy(np.ascontiguousarray(models["synch"]["z"][lo:hi])).to(self.device)
                t, nu_dev = self.ebl, self.plans["synch"].nu
                st = self.torch.cuda.current_stream(self.device).cuda_stream
                _lib.check(_lib.lib().agnpy_b200_ebl_absorption(
                    hi - lo, z_dev.data_ptr(), nu_dev.data_ptr(), self.n_nu, t["e"].data_ptr(),
                    t["e"].numel(), t["z"].data_ptr(), t["z"].numel(), t["v"].data_ptr(), 1,
                    total.data_ptr(), st), "agnpy_b200_ebl_absorption")
            return total
        return sharded_evaluate(local, index, self.n_nu, self.group, self.device)

    def __call__(self, pars, d_L):
        with self.torch.cuda.device(self.device):
            out = self.evaluate_device(pars, d_L)
            host = out.cpu()
        return host.numpy()

    def graphed(self, n_vectors):
        """the same model for a FIXED number of parameter vectors on this GPU, as one CUDA graph
        (agnpy_b200.batch.GraphedSum): `g = model.graphed(32); sed = g(pars, d_L)` with pars of
        shape [32, n_par].  Meant for the small batches of an optimiser or a sampler step, which are
        launch-bound; results are bit-identical to `model(pars, d_L)`.  Not sharded."""
        from .batch import GraphedSum
        graph = GraphedSum(self.plans, n_vectors, self.ebl)

        def call(pars, d_L):
            return graph(build_models(pars, d_L, self.ne_kind, self.scenario))
        call.graph = graph
        return call
