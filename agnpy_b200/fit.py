"""Batched fit-level model: the components the reference's fit wrappers add up per
parameter vector (agnpy/fit/sherpa_wrapper.py:51-75, 228-325; gammapy_wrapper.py:161-181,
284-357), evaluated for a whole batch of parameter vectors -- an MCMC ensemble, a likelihood
scan, the trial points of an optimiser -- in one pass, sharded over the GPUs of a box.

Parameter vectors follow the sherpa wrapper's flat order (`parameter_names`):
  "ssc"        [*spectral, z, delta_D, log10_B, t_var]
  "ec_blr"     [..., mu_s, log10_r, log10_L_disk, M_BH, m_dot, R_in, R_out, xi_line, lambda_line, R_line]
  "ec_dt"      [..., mu_s, log10_r, log10_L_disk, M_BH, m_dot, R_in, R_out, xi_dt, T_dt, R_dt]
  "ec_blr_dt"  [..., mu_s, log10_r, log10_L_disk, M_BH, m_dot, R_in, R_out, xi_line, lambda_line,
                R_line, xi_dt, T_dt, R_dt]
with the spectral block in `fit/core.py:92-113` order and log10 scaling
(`_scale_spectral_parameters`, sherpa_wrapper.py:29-48): log10_k first, log10_gamma_min /
log10_gamma_max last, break / pivot / cutoff Lorentz factors in log10.
As in the wrappers, Synch and SSC integrate over utils/math.py:6 `gamma_e_to_integrate`
(200 points) and the EC terms over logspace(1, 9, 300) (sherpa_wrapper.py:25-26).

The luminosity distance is derived from each vector's z like the wrappers do on every evaluation
(`Distance(z=z)`, sherpa_wrapper.py:63) -- host arithmetic in agnpy_b200.cosmology, cached per
redshift -- unless `d_L` (cm) is passed.  The EC scenarios add the thermal terms of the wrappers: the
normalised multi-T black body of the disk and, with a torus, its normalised black body
(sherpa_wrapper.py:131-133, 205-210, 283-288).  The seed spectrum of the SSC term is recomputed inside the
SSC launch sequence for every batch (200 frequencies, ~11 us per call for any batch size): there is
nothing to cache across the per-instrument calls of a gammapy fit that would be cheaper than that.
"""
import numpy as np

from . import _lib, constants as const, cosmology, grids
from .batch import (GatherBuffers, SedBatch, _host_gather_call, _world,  # GraphedSum: lazily in graphed()
                    sharded_evaluate)

# position (from the end of the spectral block) of the log10-scaled Lorentz factors besides
# gamma_min / gamma_max, sherpa_wrapper.py:41-48
_LOG10_EXTRA = {"PowerLaw": (), "BrokenPowerLaw": (-3,), "LogParabola": (-3,),
                "ExpCutoffPowerLaw": (-3,), "ExpCutoffBrokenPowerLaw": (-3, -4)}
_N_SPECTRAL = {"PowerLaw": 4, "BrokenPowerLaw": 6, "LogParabola": 6, "ExpCutoffPowerLaw": 5,
               "ExpCutoffBrokenPowerLaw": 7}
gamma_to_integrate_ec = np.logspace(1, 9, 300)  # sherpa_wrapper.py:25-26
_EV_ERG = 1.602176634e-12


def scale_spectral_parameters(spectral, ne_kind):
    """vectorised `_scale_spectral_parameters` (sherpa_wrapper.py:29-48): [B, n] -> [B, n]"""
    out = np.array(spectral, dtype=np.float64, copy=True)
    out[:, 0] = 10 ** out[:, 0]
    out[:, -2] = 10 ** out[:, -2]
    out[:, -1] = 10 ** out[:, -1]
    for pos in _LOG10_EXTRA[ne_kind]:
        out[:, pos] = 10 ** out[:, pos]
    return out


SCENARIOS = ("ssc", "ec_blr", "ec_dt", "ec_blr_dt")
# names of the flat parameter vector after the spectral block, in the order the sherpa wrappers unpack
# them (sherpa_wrapper.py:55, 83-98, 157-172, 233-251) = the order gammapy's Parameters are listed in
# (gammapy_wrapper.py:216-239, fit/core.py:128-236)
_REGION_SSC = ("z", "delta_D", "log10_B", "t_var")
_REGION_EC = _REGION_SSC + ("mu_s", "log10_r")
_DISK = ("log10_L_disk", "M_BH", "m_dot", "R_in", "R_out")
_BLR = ("xi_line", "lambda_line", "R_line")
_DT = ("xi_dt", "T_dt", "R_dt")
_SPECTRAL_NAMES = {
    "PowerLaw": ("log10_k", "p", "log10_gamma_min", "log10_gamma_max"),
    "BrokenPowerLaw": ("log10_k", "p1", "p2", "log10_gamma_b", "log10_gamma_min", "log10_gamma_max"),
    "LogParabola": ("log10_k", "p", "q", "log10_gamma_0", "log10_gamma_min", "log10_gamma_max"),
    "ExpCutoffPowerLaw": ("log10_k", "p", "log10_gamma_c", "log10_gamma_min", "log10_gamma_max"),
    "ExpCutoffBrokenPowerLaw": ("log10_k", "p1", "p2", "log10_gamma_c", "log10_gamma_b",
                                "log10_gamma_min", "log10_gamma_max"),
}


def parameter_names(ne_kind, scenario):
    """names of the flat parameter vector of `scenario`, spectral block first (fit/core.py:89)"""
    if scenario not in SCENARIOS:
        raise ValueError(f"scenario must be one of {SCENARIOS}")
    names = _SPECTRAL_NAMES[ne_kind]
    if scenario == "ssc":
        return names + _REGION_SSC
    names = names + _REGION_EC + _DISK
    if "blr" in scenario:
        names = names + _BLR
    if "dt" in scenario:
        names = names + _DT
    return names


def scenario_of(targets):
    """gammapy's `targets` list (["blr"], ["dt"], ["blr", "dt"]; gammapy_wrapper.py:188-200) or None
    for the SSC model -> scenario name"""
    if not targets:
        return "ssc"
    t = set(targets)
    if not t <= {"blr", "dt"}:
        raise ValueError("targets must be a subset of ['blr', 'dt']")
    return "ec_" + "_".join(x for x in ("blr", "dt") if x in t)


_HOST_CHUNK = 4096   # parameter vectors per host -> device hand-over of a large batch


def build_models(pars, d_L, ne_kind, scenario):
    """flat fit parameter vectors [B, n_par] -> dict of MODEL_DTYPE arrays per component.
    `d_L` (cm; scalar or one per vector) may be None: it is then derived from each vector's z as the
    wrappers do on every evaluation (`Distance(z=z)`, sherpa_wrapper.py:63), cached per redshift."""
    pars = np.atleast_2d(np.asarray(pars, dtype=np.float64))
    names = parameter_names(ne_kind, scenario)
    ns = _N_SPECTRAL[ne_kind]
    if pars.shape[1] != len(names):
        raise ValueError(f"{scenario} scenario with {ne_kind} takes {len(names)} parameters "
                         f"{names}, got {pars.shape[1]}")
    B = pars.shape[0]
    col = {n: pars[:, i] for i, n in enumerate(names)}
    z, delta_D = col["z"], col["delta_D"]
    if d_L is None:
        d_L = cosmology.luminosity_distance(z)
    d_L = np.broadcast_to(np.asarray(d_L, dtype=np.float64), (B,))
    spectral = scale_spectral_parameters(pars[:, :ns], ne_kind)
    base = _lib.make_models(B)
    base["z"], base["d_L"], base["delta_D"] = z, d_L, delta_D
    base["B"] = 10 ** col["log10_B"]
    base["R_b"] = (const.c * col["t_var"] * delta_D) / (1 + z)  # sherpa_wrapper.py:64
    base["mu_s"] = 1.0
    base["ne"][:, :ns] = spectral
    out = {"synch": base, "ssc": base}
    if scenario == "ssc":
        return out
    mu_s, r, L_disk = col["mu_s"], 10 ** col["log10_r"], 10 ** col["log10_L_disk"]
    bb_disk = base.copy()   # SSDisk.evaluate_multi_T_bb_norm_sed(nu, z, L_disk, M_BH, m_dot, R_in, R_out, d_L)
    bb_disk["mu_s"] = 1.0   # the wrappers use the default mu_s = 1 for the disk SED
    for k, v in enumerate((col["M_BH"], col["m_dot"], col["R_in"], col["R_out"], L_disk)):
        bb_disk["tgt"][:, k] = v
    out["bb_disk"] = bb_disk
    if "blr" in scenario:
        # sherpa_wrapper.py:115-120 (lambda_line in Angstrom)
        epsilon_line = const.h * const.c / (col["lambda_line"] * 1e-8) / const.mec2
        blr = base.copy()
        blr["mu_s"] = mu_s
        blr["tgt"][:, 0], blr["tgt"][:, 1], blr["tgt"][:, 2] = L_disk, col["xi_line"], epsilon_line
        blr["tgt"][:, 3], blr["tgt"][:, 4] = col["R_line"], r
        out["ec_blr"] = blr
    if "dt" in scenario:
        epsilon_dt = 2.7 * (const.k_B * col["T_dt"] / const.mec2)     # sherpa_wrapper.py:194
        dt = base.copy()
        dt["mu_s"] = mu_s
        dt["tgt"][:, 0], dt["tgt"][:, 1], dt["tgt"][:, 2] = L_disk, col["xi_dt"], epsilon_dt
        dt["tgt"][:, 3], dt["tgt"][:, 4] = col["R_dt"], r
        bb_dt = base.copy()     # RingDustTorus.evaluate_bb_norm_sed(nu, z, xi_dt * L_disk, T_dt, R_dt, d_L)
        for k, v in enumerate((col["T_dt"], col["R_dt"], col["xi_dt"] * L_disk)):
            bb_dt["tgt"][:, k] = v
        out.update(ec_dt=dt, bb_dt=bb_dt)
    return out


class FitSedBatch:
    """nu F_nu of all components of a fit scenario for batches of fit parameter vectors.

        model = FitSedBatch(nu_hz, "BrokenPowerLaw", scenario="ec_blr_dt")
        sed = model(pars)             # [B, n_nu] erg cm-2 s-1, summed over components;
                                      # d_L from each vector's z (Planck18), as the wrappers do
        sed = model(pars, d_L)        # or with explicit luminosity distance(s) in cm

    scenario: "ssc", "ec_blr", "ec_dt", "ec_blr_dt" = the four `_evaluate_sed_*_scenario` functions of
    agnpy/fit/sherpa_wrapper.py:51-325; `parameter_names(ne_kind, scenario)` lists the vector layout.
    `gammapy(energy_eV, **kwargs)` is the same model in gammapy's calling convention.
    With torch.distributed initialised every rank evaluates its shard of the batch and ONE
    all-gather joins the summed SEDs."""

    def __init__(self, nu, ne_kind="BrokenPowerLaw", scenario="ec_blr_dt", ssa=False,
                 integrator="trapz", device=None, group=None, ebl=None):
        if scenario not in SCENARIOS:
            raise ValueError(f"scenario must be one of {SCENARIOS}")
        self.ne_kind, self.scenario, self.group = ne_kind, scenario, group
        nu = np.ravel(np.asarray(nu, dtype=np.float64))
        self.nu_hz = nu
        kw = dict(device=device, integrator=integrator)
        # the wrappers ignore `ssa` for the EC components
        self.plans = {
            "synch": SedBatch("synch", nu, ne_kind, gamma=grids.gamma_e_to_integrate, ssa=ssa, **kw),
            "ssc": SedBatch("ssc", nu, ne_kind, gamma=grids.gamma_e_to_integrate, ssa=ssa, **kw),
        }
        if scenario != "ssc":
            self.plans["bb_disk"] = SedBatch("bb_disk_norm", nu, ne_kind, device=device)
        if "blr" in scenario:
            self.plans["ec_blr"] = SedBatch("ec_blr", nu, ne_kind, gamma=gamma_to_integrate_ec, **kw)
        if "dt" in scenario:
            self.plans["ec_dt"] = SedBatch("ec_dt", nu, ne_kind, gamma=gamma_to_integrate_ec, **kw)
            self.plans["bb_dt"] = SedBatch("bb_dt_norm", nu, ne_kind, device=device)
        self.n_nu = nu.size
        first = self.plans["synch"]
        self.device, self.torch = first.device, first.torch
        self._gather = GatherBuffers()
        self._total = None          # device accumulator of the component sum, grown on demand
        # optional EBL attenuation (agnpy_b200.ebl.EBL): table resident on the device, applied to
        # the summed SEDs of each shard in place with the per-vector redshifts
        self.ebl = None
        if ebl is not None:
            dev = lambda a: self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
            self.ebl = dict(e=dev(ebl.energy_ref), z=dev(ebl.z_ref), v=dev(ebl.values_ref.T))

    @property
    def parameter_names(self):
        return parameter_names(self.ne_kind, self.scenario)

    def _local_components(self, models, lo, hi):
        return {name: plan._local(models[name][lo:hi]).clone() for name, plan in self.plans.items()}

    def components(self, pars, d_L=None):
        """per-component SEDs of the whole batch on this rank (no sharding), as numpy arrays"""
        models = build_models(pars, d_L, self.ne_kind, self.scenario)
        n = len(models["synch"])
        with self.torch.cuda.device(self.device):
            out = {k: v.cpu().numpy() for k, v in self._local_components(models, 0, n).items()}
        return out

    def _local_sum(self, pars, d_L):
        """closure: index range of this rank's shard -> summed SEDs of those vectors on the device"""
        B = pars.shape[0]
        if d_L is not None:
            d_L = np.broadcast_to(np.asarray(d_L, dtype=np.float64), (B,))

        def local(idx):
            # only this rank's shard of the parameter table is turned into model rows, _HOST_CHUNK
            # vectors at a time: the kernel launches are asynchronous, so the host prepares the rows
            # of the next chunk (log10 scalings, d_L(z), ...) while the device works on this one
            lo, hi = int(idx[0]), int(idx[-1]) + 1
            n = hi - lo
            if self._total is None or self._total.shape[0] < n:
                self._total = self.torch.empty(n, self.n_nu, dtype=self.torch.float64,
                                               device=self.device)
            for a in range(lo, hi, _HOST_CHUNK):
                b = min(a + _HOST_CHUNK, hi)
                models = build_models(pars[a:b], None if d_L is None else d_L[a:b], self.ne_kind,
                                      self.scenario)
                total = self._total[a - lo:b - lo]
                for k, (name, plan) in enumerate(self.plans.items()):
                    part = plan._local(models[name])
                    total.copy_(part) if k == 0 else total.add_(part)
                if self.ebl is not None:
                    z_dev = self.torch.from_numpy(np.ascontiguousarray(models["synch"]["z"])).to(self.device)
                    t, nu_dev = self.ebl, self.plans["synch"].nu
                    st = self.torch.cuda.current_stream(self.device).cuda_stream
                    _lib.check(_lib.lib().agnpy_b200_ebl_absorption(
                        b - a, z_dev.data_ptr(), nu_dev.data_ptr(), self.n_nu, t["e"].data_ptr(),
                        t["e"].numel(), t["z"].data_ptr(), t["z"].numel(), t["v"].data_ptr(), 1,
                        total.data_ptr(), st), "agnpy_b200_ebl_absorption")
            total = self._total[:n]
            return total
        return local

    def evaluate_device(self, pars, d_L=None, gather="all"):
        """summed SEDs on the device, [B, n_nu] (a buffer of this object: valid until the next call)"""
        pars = np.atleast_2d(np.asarray(pars, dtype=np.float64))
        index = np.arange(pars.shape[0])
        return sharded_evaluate(self._local_sum(pars, d_L), index, self.n_nu, self.group, self.device,
                                gather, self._gather)

    def __call__(self, pars, d_L=None, gather="all", copy=True):
        """[B, n_par] -> numpy [B, n_nu] (erg cm-2 s-1).  `gather` / `copy` as in SedBatch.__call__:
        the result goes device -> pinned host memory in one asynchronous copy; copy=False returns the
        pinned block itself (valid until the next call)."""
        with self.torch.cuda.device(self.device):
            if gather == "host" and _world(self.group) > 1:
                pars = np.atleast_2d(np.asarray(pars, dtype=np.float64))
                local = self._local_sum(pars, d_L)
                return _host_gather_call(self, local, np.arange(pars.shape[0]), self.n_nu, copy)
            out = self.evaluate_device(pars, d_L, "all" if gather == "host" else gather)
            if out is None:
                self.torch.cuda.current_stream(self.device).synchronize()
                return None
            return self.plans["synch"]._to_host(out, copy)

    # ---- the reference's two calling conventions ----------------------------------------------
    def sherpa(self, x_eV, pars, d_L=None):
        """`calc(pars, x)` of the sherpa models (sherpa_wrapper.py:376-378, 465-472) for a batch:
        energies `x_eV` must be the grid the model was built on (nu = x / h)"""
        self._check_grid(np.asarray(x_eV, dtype=np.float64) * _EV_ERG / const.h)
        return self(pars, d_L)

    def gammapy(self, energy_eV, d_L=None, **kwargs):
        """`SpectralModel.evaluate(energy, **kwargs)` of the gammapy wrappers (gammapy_wrapper.py:161-181,
        284-357) for a batch: every model parameter arrives by name (scalar or array of B values, plain
        numbers in the wrapper's units: t_var s, M_BH g, m_dot g/s, radii cm, lambda_line Angstrom,
        T_dt K; a `norm` entry is ignored, as in the reference).  Returns dN/dE = sed / E^2 in
        1 / (cm2 eV s), shape [B, n_energy]."""
        energy_eV = np.ravel(np.asarray(energy_eV, dtype=np.float64))
        self._check_grid(energy_eV * _EV_ERG / const.h)
        names = self.parameter_names
        missing = [n for n in names if n not in kwargs]
        if missing:
            raise TypeError(f"missing model parameters: {missing}")
        cols = np.broadcast_arrays(*[np.asarray(getattr(kwargs[n], "value", kwargs[n]), dtype=np.float64)
                                     for n in names])
        pars = np.stack([np.ravel(c) for c in cols], axis=1)
        sed = self(pars, d_L)
        return sed / (energy_eV**2 * _EV_ERG)

    def _check_grid(self, nu_hz):
        if nu_hz.shape != self.nu_hz.shape or not np.allclose(nu_hz, self.nu_hz, rtol=1e-12, atol=0):
            raise ValueError("the energies differ from the grid this FitSedBatch was built on; the grids "
                             "live on the device for the life of the plan -- build one model per grid")

    def graphed(self, n_vectors):
        """the same model for a FIXED number of parameter vectors on this GPU, as one CUDA graph
        (agnpy_b200.batch.GraphedSum): `g = model.graphed(32); sed = g(pars)` with pars of
        shape [32, n_par].  Meant for the small batches of an optimiser or a sampler step, which are
        launch-bound; results are bit-identical to `model(pars)`.  Not sharded."""
        from .batch import GraphedSum
        graph = GraphedSum(self.plans, n_vectors, self.ebl)

        def call(pars, d_L=None):
            return graph(build_models(pars, d_L, self.ne_kind, self.scenario))
        call.graph = graph
        return call
