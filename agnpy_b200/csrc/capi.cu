// extern "C" entry points declared in include/agnpy_b200.h: argument checks, workspace
// carving, batching of the model axis, and the host-pointer flavours (H2D, run, D2H, sync).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: a no-op unless a profiler is attached

namespace agb {

constexpr size_t kChunkBytes = (size_t)1 << 30;  // scratch budget per batch chunk

static bool bad_kind(int k) { return k < AGNPY_NE_POWERLAW || k > AGNPY_NE_TABLE; }
static bool bad_integ(int i) {
    i &= ~AGNPY_INTEG_FOLD_PHI;
    return i < AGNPY_INTEG_TRAPZ || i > AGNPY_INTEG_TRAPZ_PREFIX;
}
static const double* row_at(const double* table, int m0, int n_gamma) {
    return table ? table + (size_t)m0 * n_gamma : nullptr;  // tables are [n_models][n_gamma]
}
static int plain_integ(int i) {
    i &= ~AGNPY_INTEG_FOLD_PHI;
    return i == AGNPY_INTEG_TRAPZ_PREFIX ? AGNPY_INTEG_TRAPZ : i;
}

// NVTX range around every C-ABI entry point (named like the entry point): lets ncu / nsys filter and
// group the kernels of one call (ncu --nvtx --nvtx-include "agnpy_b200_ec_sed/").
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define AGB_RANGE() NvtxRange nvtx_range_(__func__)

static int chunk_models(size_t bytes_per_model, int n_models) {
    size_t budget = kChunkBytes;
    if (const char* e = getenv("AGNPY_B200_CHUNK_BYTES")) {  // test knob: force small chunks
        long long v = atoll(e);
        if (v > 0) budget = (size_t)v;
    }
    size_t c = budget / std::max<size_t>(bytes_per_model, 1);
    c = std::max<size_t>(1, std::min<size_t>(c, 65535));
    return (int)std::min<size_t>(c, (size_t)n_models);
}

// simple bump allocator over one device buffer for the *_host flavours
struct Arena {
    char* base;
    size_t off = 0;
    explicit Arena(void* p) : base((char*)p) {}
    template <class T>
    T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = (T*)(base + off);
        off += n * sizeof(T);
        return p;
    }
};
static size_t pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

static cudaStream_t host_stream() {
    // the *_host entry points run on one private non-blocking stream per device
    static thread_local cudaStream_t streams[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 0;
    if (!streams[dev]) cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking);
    return streams[dev];
}

static int need_device() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(AGNPY_ERR_CUDA, "no CUDA device: agnpy_b200 has no CPU fallback");
    }
    return AGNPY_OK;
}

}  // namespace agb

using namespace agb;

extern "C" {

// ------------------------------------------------------------------------------------------
int agnpy_b200_synch_sed(int n_models, const agnpy_model_t* models, int ne_kind,
                         const double* nu, int n_nu, const double* gamma, int n_gamma,
                         const double* ne_table, const double* ssa_table, int ssa, int integ,
                         double* sed, double* tau_ssa, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || !models || !nu || !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "synch_sed: bad sizes or null pointer");
    if (bad_kind(ne_kind) || bad_integ(integ)) return fail(AGNPY_ERR_ARG, "synch_sed: bad ne_kind/integ");
    if (n_nu > 65535) return fail(AGNPY_ERR_ARG, "synch_sed: n_nu > 65535");
    if (ne_kind == AGNPY_NE_TABLE && (!ne_table || (ssa && !ssa_table)))
        return fail(AGNPY_ERR_ARG, "synch_sed: AGNPY_NE_TABLE needs n_e (and SSA) tables");
    integ = plain_integ(integ);
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = 3 * (size_t)n_gamma * sizeof(double);
    int chunk = chunk_models(per, n_models);
    double* tab = (double*)workspace(per * chunk, stream);
    if (!tab) return fail(AGNPY_ERR_NOMEM, "synch_sed: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        launch_prep_gamma(models + m0, nm, ne_kind, gamma, n_gamma, row_at(ne_table, m0, n_gamma),
                          row_at(ssa_table, m0, n_gamma), 0, tab, stream);
        timing_begin(stream);
        int rc_s = launch_synch(models + m0, nm, nu, n_nu, tab, n_gamma, ssa, integ, 0,
                                sed + (size_t)m0 * n_nu,
                                tau_ssa ? tau_ssa + (size_t)m0 * n_nu : nullptr, nullptr, stream);
        timing_end(stream);
        if (rc_s) return rc_s;
    }
    return check_last("synch_sed");
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_ssc_sed(int n_models, const agnpy_model_t* models, int ne_kind,
                       const double* nu, int n_nu, const double* nu_seed, int n_seed,
                       const double* gamma, int n_gamma, const double* ne_table,
                       const double* ssa_table, int ssa, int integ, double* sed, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_seed < 2 || n_gamma < 2 || !models || !nu || !nu_seed ||
        !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "ssc_sed: bad sizes or null pointer");
    if (bad_kind(ne_kind) || bad_integ(integ)) return fail(AGNPY_ERR_ARG, "ssc_sed: bad ne_kind/integ");
    if (n_nu > 65535 || n_seed > 65535) return fail(AGNPY_ERR_ARG, "ssc_sed: n_nu/n_seed > 65535");
    // nu_seed must be strictly ascending and positive (the np.trapz kernel searches it as a monotone
    // grid); device memory cannot be inspected here cheaply, the host flavour and the Python callers
    // validate it (agnpy_b200_ssc_sed_host, agnpy_b200.grids.as_grid)
    if (ne_kind == AGNPY_NE_TABLE && (!ne_table || (ssa && !ssa_table)))
        return fail(AGNPY_ERR_ARG, "ssc_sed: AGNPY_NE_TABLE needs n_e (and SSA) tables");
    integ = plain_integ(integ);
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = (3 * (size_t)n_gamma + 2 * (size_t)n_seed + ssc_scratch_doubles(n_gamma, n_seed)) *
                 sizeof(double);
    int chunk = chunk_models(per, n_models);
    double* ws = (double*)workspace(per * chunk, stream);
    if (!ws) return fail(AGNPY_ERR_NOMEM, "ssc_sed: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        double* scratch = ws;  // first: 32-byte aligned rows of the np.trapz path
        double* tab = scratch + (size_t)nm * ssc_scratch_doubles(n_gamma, n_seed);
        double* U = tab + (size_t)nm * 3 * n_gamma;
        double* es = U + (size_t)nm * n_seed;
        launch_prep_gamma(models + m0, nm, ne_kind, gamma, n_gamma, row_at(ne_table, m0, n_gamma),
                          row_at(ssa_table, m0, n_gamma), 0, tab, stream);
        // seed synchrotron spectrum at nu_seed (synchrotron_self_compton.py:129-142)
        if (int rc = launch_synch(models + m0, nm, nu_seed, n_seed, tab, n_gamma, ssa, integ, 1, U,
                                  nullptr, es, stream))
            return rc;
        int rc_ssc = launch_ssc(models + m0, nm, nu, n_nu, tab, n_gamma, U, es, n_seed, integ,
                                sed + (size_t)m0 * n_nu, scratch, stream);
        if (rc_ssc) return rc_ssc;
    }
    return check_last("ssc_sed");
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_proton_synch_sed(int n_models, const agnpy_model_t* models, int ne_kind,
                                const double* nu, int n_nu, const double* gamma, int n_gamma,
                                const double* ne_table, int integ, double* sed, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || !models || !nu || !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "proton_synch_sed: bad sizes or null pointer");
    if (bad_kind(ne_kind) || bad_integ(integ)) return fail(AGNPY_ERR_ARG, "proton_synch_sed: bad ne_kind/integ");
    if (n_nu > 65535) return fail(AGNPY_ERR_ARG, "proton_synch_sed: n_nu > 65535");
    if (ne_kind == AGNPY_NE_TABLE && (!ne_table))
        return fail(AGNPY_ERR_ARG, "proton_synch_sed: AGNPY_NE_TABLE needs an n_e table");
    integ = plain_integ(integ);
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = 3 * (size_t)n_gamma * sizeof(double);
    int chunk = chunk_models(per, n_models);
    double* tab = (double*)workspace(per * chunk, stream);
    if (!tab) return fail(AGNPY_ERR_NOMEM, "proton_synch_sed: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        launch_prep_gamma(models + m0, nm, ne_kind, gamma, n_gamma, row_at(ne_table, m0, n_gamma), nullptr, 0,
                          tab, stream);
        timing_begin(stream);
        int rc = launch_synch(models + m0, nm, nu, n_nu, tab, n_gamma, 0, integ, 0,
                              sed + (size_t)m0 * n_nu, nullptr, nullptr, stream, 0, 1);
        timing_end(stream);
        if (rc) return rc;
    }
    return check_last("proton_synch_sed");
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_ssc_loss_rate(int n_models, const agnpy_model_t* models, int ne_kind,
                             const double* gamma_eval, int n_eval, const double* nu_seed, int n_seed,
                             const double* gamma, int n_gamma, const double* ne_table, double* out,
                             void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_eval < 1 || n_seed < 2 || n_gamma < 2 || !models || !gamma_eval || !nu_seed ||
        !gamma || !out)
        return fail(AGNPY_ERR_ARG, "ssc_loss_rate: bad sizes or null pointer");
    if (bad_kind(ne_kind)) return fail(AGNPY_ERR_ARG, "ssc_loss_rate: bad ne_kind");
    if (n_eval > 65535 || n_seed > 65535) return fail(AGNPY_ERR_ARG, "ssc_loss_rate: grid too large");
    if (ne_kind == AGNPY_NE_TABLE && (!ne_table))
        return fail(AGNPY_ERR_ARG, "ssc_loss_rate: AGNPY_NE_TABLE needs an n_e table");
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = (3 * (size_t)n_gamma + 2 * (size_t)n_seed) * sizeof(double);
    int chunk = chunk_models(per, n_models);
    double* ws = (double*)workspace(per * chunk, stream);
    if (!ws) return fail(AGNPY_ERR_NOMEM, "ssc_loss_rate: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        double* tab = ws;
        double* U = tab + (size_t)nm * 3 * n_gamma;
        double* es = U + (size_t)nm * n_seed;
        launch_prep_gamma(models + m0, nm, ne_kind, gamma, n_gamma, row_at(ne_table, m0, n_gamma), nullptr, 0,
                          tab, stream);
        if (int rc = launch_synch(models + m0, nm, nu_seed, n_seed, tab, n_gamma, 0, AGNPY_INTEG_TRAPZ, 1,
                                  U, nullptr, es, stream))
            return rc;
        launch_ssc_loss_rate(nm, gamma_eval, n_eval, U, es, n_seed, out + (size_t)m0 * n_eval, stream);
    }
    return check_last("ssc_loss_rate");
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_ec_sed(int variant, int n_models, const agnpy_model_t* models, int ne_kind,
                      const double* nu, int n_nu, const double* gamma, int n_gamma,
                      const double* mu, int n_mu, const double* phi, int n_phi,
                      const double* ne_table, int integ, double* sed, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (variant < AGNPY_EC_ISO_MONO || variant > AGNPY_EC_DT) return fail(AGNPY_ERR_ARG, "ec_sed: bad variant");
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || !models || !nu || !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "ec_sed: bad sizes or null pointer");
    if (bad_kind(ne_kind) || bad_integ(integ)) return fail(AGNPY_ERR_ARG, "ec_sed: bad ne_kind/integ");
    if (n_nu > 65535) return fail(AGNPY_ERR_ARG, "ec_sed: n_nu > 65535");
    bool need_phi = variant != AGNPY_EC_PS_BEHIND_JET;
    bool need_mu_grid = variant == AGNPY_EC_ISO_MONO || variant == AGNPY_EC_BLR;
    bool need_mu_size = need_mu_grid || variant == AGNPY_EC_SS_DISK;
    if (need_phi && (!phi || n_phi < 2)) return fail(AGNPY_ERR_ARG, "ec_sed: phi grid required");
    if (need_mu_grid && !mu) return fail(AGNPY_ERR_ARG, "ec_sed: mu grid required");
    if (need_mu_size && n_mu < 2) return fail(AGNPY_ERR_ARG, "ec_sed: n_mu < 2");
    if (ne_kind == AGNPY_NE_TABLE && (!ne_table))
        return fail(AGNPY_ERR_ARG, "ec_sed: AGNPY_NE_TABLE needs an n_e table");
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = ec_workspace_bytes(variant, 1, n_nu, n_gamma, n_mu, n_phi, integ);
    int chunk = chunk_models(per, n_models);
    double* ws = (double*)workspace(per * chunk, stream);
    if (!ws) return fail(AGNPY_ERR_NOMEM, "ec_sed: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        if (int rc = run_ec(variant, models + m0, nm, ne_kind, nu, n_nu, gamma, n_gamma, mu, n_mu,
                            phi, n_phi, row_at(ne_table, m0, n_gamma), integ, sed + (size_t)m0 * n_nu, ws,
                            stream))
            return rc;
    }
    return AGNPY_OK;
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_tau(int variant, int n_models, const agnpy_model_t* models, const double* nu,
                   int n_nu, const double* mu, int n_a, const double* phi, int n_phi,
                   const double* l_unit, int n_l, double* tau, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (variant < AGNPY_TAU_SS_DISK || variant > AGNPY_TAU_DT_MU_S) return fail(AGNPY_ERR_ARG, "tau: bad variant");
    const bool is_ps = variant == AGNPY_TAU_PS_BEHIND_BLOB || variant == AGNPY_TAU_PS_BEHIND_BLOB_MU_S;
    if (n_models < 1 || n_nu < 1 || !models || !nu || !tau)
        return fail(AGNPY_ERR_ARG, "tau: bad sizes or null pointer");
    if (!is_ps && (n_phi < 2 || !phi)) return fail(AGNPY_ERR_ARG, "tau: phi grid required");
    if (variant != AGNPY_TAU_PS_BEHIND_BLOB && (n_l < 2 || !l_unit)) return fail(AGNPY_ERR_ARG, "tau: path grid required");
    if ((variant == AGNPY_TAU_BLR || variant == AGNPY_TAU_BLR_MU_S) && (!mu || n_a < 2))
        return fail(AGNPY_ERR_ARG, "tau: mu grid required");
    if (variant == AGNPY_TAU_SS_DISK && n_a < 2) return fail(AGNPY_ERR_ARG, "tau: R_tilde_size < 2");
    if ((n_nu + 3) / 4 > 65535) return fail(AGNPY_ERR_ARG, "tau: n_nu too large");
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = tau_workspace_bytes(variant, 1, n_nu, n_a, n_phi, n_l);
    int chunk = chunk_models(per, n_models);
    double* ws = (double*)workspace(per * chunk, stream);
    if (!ws) return fail(AGNPY_ERR_NOMEM, "tau: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        if (int rc = run_tau(variant, models + m0, nm, nu, n_nu, mu, n_a, phi, n_phi, l_unit, n_l,
                             tau + (size_t)m0 * n_nu, ws, stream))
            return rc;
    }
    return AGNPY_OK;
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_tau_synch(int n_models, const agnpy_model_t* models, int ne_kind, const double* nu,
                         int n_nu, const double* gamma, int n_gamma, const double* ne_table,
                         const double* ssa_table, int n_s, double delta_margin_low, int integ,
                         double* tau, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || n_s < 2 || !models || !nu || !gamma || !tau)
        return fail(AGNPY_ERR_ARG, "tau_synch: bad sizes or null pointer");
    if (bad_kind(ne_kind) || bad_integ(integ)) return fail(AGNPY_ERR_ARG, "tau_synch: bad ne_kind/integ");
    if (n_s > 65535 || (n_nu + 3) / 4 > 65535) return fail(AGNPY_ERR_ARG, "tau_synch: grid too large");
    if (ne_kind == AGNPY_NE_TABLE && (!ne_table || !ssa_table))
        return fail(AGNPY_ERR_ARG, "tau_synch: AGNPY_NE_TABLE needs both tables");
    integ = plain_integ(integ);
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = tau_synch_workspace_bytes(1, n_nu, n_gamma, n_s);
    int chunk = chunk_models(per, n_models);
    double* ws = (double*)workspace(per * chunk, stream);
    if (!ws) return fail(AGNPY_ERR_NOMEM, "tau_synch: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        if (int rc = run_tau_synch(models + m0, nm, ne_kind, nu, n_nu, gamma, n_gamma,
                                   row_at(ne_table, m0, n_gamma), row_at(ssa_table, m0, n_gamma),
                                   n_s, delta_margin_low, integ, tau + (size_t)m0 * n_nu, ws, stream))
            return rc;
    }
    return AGNPY_OK;
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_trapz_loglog(const double* y, const double* x, long long n_rows, int n_x, int intervals,
                            double* out, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (n_rows < 1 || n_x < 2 || !y || !x || !out)
        return fail(AGNPY_ERR_ARG, "trapz_loglog: bad sizes or null pointer");
    return run_trapz_loglog(y, x, n_rows, n_x, intervals, out, (cudaStream_t)stream_);
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_bb_sed(int variant, int n_models, const agnpy_model_t* models, const double* nu,
                      int n_nu, const double* nu_norm, int n_norm, int n_R, double* sed,
                      void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (variant < AGNPY_BB_DISK || variant > AGNPY_BB_DT_NORM) return fail(AGNPY_ERR_ARG, "bb_sed: bad variant");
    if (n_models < 1 || n_nu < 1 || !models || !nu || !sed)
        return fail(AGNPY_ERR_ARG, "bb_sed: bad sizes or null pointer");
    if (variant <= AGNPY_BB_DISK_NORM && n_R < 2) return fail(AGNPY_ERR_ARG, "bb_sed: n_R < 2");
    if (variant == AGNPY_BB_DISK_NORM && (!nu_norm || n_norm < 2))
        return fail(AGNPY_ERR_ARG, "bb_sed: normalisation frequencies required");
    return run_bb(variant, models, n_models, nu, n_nu, nu_norm, n_norm, n_R > 0 ? n_R : 2, sed,
                  (cudaStream_t)stream_);
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_target_u(int variant, int n_models, const agnpy_model_t* models, const double* Gamma,
                        const double* r, int n_r, int n_mu, double* u, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (variant != AGNPY_EC_PS_BEHIND_JET && variant != AGNPY_EC_SS_DISK && variant != AGNPY_EC_BLR &&
        variant != AGNPY_EC_DT)
        return fail(AGNPY_ERR_ARG, "target_u: bad variant");
    if (n_models < 1 || n_r < 1 || !models || !r || !u)
        return fail(AGNPY_ERR_ARG, "target_u: bad sizes or null pointer");
    if ((variant == AGNPY_EC_SS_DISK || variant == AGNPY_EC_BLR) && n_mu < 2)
        return fail(AGNPY_ERR_ARG, "target_u: n_mu < 2");
    if (n_models > 65535) return fail(AGNPY_ERR_ARG, "target_u: n_models > 65535");
    return run_target_u(variant, models, n_models, Gamma, r, n_r, n_mu, u, (cudaStream_t)stream_);
}

int agnpy_b200_ebl_absorption(int n_models, const double* z, const double* nu, int n_nu,
                              const double* energy_ref, int n_e, const double* z_ref, int n_z,
                              const double* values, int multiply, double* out, void* stream_) {
    AGB_RANGE();
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_e < 2 || n_z < 2 || !z || !nu || !energy_ref || !z_ref ||
        !values || !out)
        return fail(AGNPY_ERR_ARG, "ebl_absorption: bad sizes or null pointer");
    if (n_models > 65535) return fail(AGNPY_ERR_ARG, "ebl_absorption: n_models > 65535");
    return run_ebl(n_models, z, nu, n_nu, energy_ref, n_e, z_ref, n_z, values, multiply, out,
                   (cudaStream_t)stream_);
}

}  // extern "C"

// ==========================================================================================
// host-pointer flavours.  All inputs are packed into one pinned staging block and go to the
// device with ONE cudaMemcpyAsync; outputs come back with one more (small calls are latency
// bound: a dozen pageable copies would cost more than the kernels).  Blocks above 8 MiB skip
// the staging copy and are transferred straight from / to the caller's memory.
// ==========================================================================================
namespace agb {

static void* cached_alloc(size_t bytes, bool pinned) {
    static thread_local void* ptr[2][64] = {};
    static thread_local size_t size[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    const int w = pinned ? 1 : 0;
    if (size[w][dev] >= bytes) return ptr[w][dev];
    if (ptr[w][dev]) {
        cudaDeviceSynchronize();
        if (pinned) cudaFreeHost(ptr[w][dev]); else cudaFree(ptr[w][dev]);
        ptr[w][dev] = nullptr;
        size[w][dev] = 0;
    }
    size_t want = bytes + bytes / 4 + (1u << 16);
    cudaError_t e = pinned ? cudaHostAlloc(&ptr[w][dev], want, cudaHostAllocDefault)
                           : cudaMalloc(&ptr[w][dev], want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ptr[w][dev] = nullptr;
        return nullptr;
    }
    size[w][dev] = want;
    return ptr[w][dev];
}

struct HostIO {
    static constexpr size_t kStageLimit = (size_t)8 << 20;
    struct Buf { const void* h_in; void* h_out; size_t bytes, off; };
    std::vector<Buf> bufs;
    size_t in_end = 0, total = 0;
    char* dbase = nullptr;
    char* pbase = nullptr;
    bool staged = false;
    cudaStream_t st = 0;

    int in(const void* h, size_t bytes) {  // register an input (h may be NULL -> no buffer)
        bufs.push_back({h, nullptr, h ? bytes : 0, 0});
        return (int)bufs.size() - 1;
    }
    int out(void* h, size_t bytes) {
        bufs.push_back({nullptr, h, h ? bytes : 0, 0});
        return (int)bufs.size() - 1;
    }
    // inputs must all be registered before outputs
    int commit() {
        size_t off = 0;
        for (auto& b : bufs) {
            if (b.h_out && in_end == 0) in_end = off;
            b.off = off;
            off += pad(b.bytes);
        }
        if (in_end == 0) in_end = off;
        total = off + 256;
        st = host_stream();
        dbase = (char*)cached_alloc(total, false);
        if (!dbase) return fail(AGNPY_ERR_NOMEM, "host call: device memory");
        staged = total <= kStageLimit;
        if (staged) {
            pbase = (char*)cached_alloc(total, true);
            if (!pbase) staged = false;
        }
        return AGNPY_OK;
    }
    template <class T>
    T* dev(int idx) const {
        const Buf& b = bufs[idx];
        return b.bytes ? (T*)(dbase + b.off) : nullptr;
    }
    int upload() {
        if (staged) {
            for (auto& b : bufs)
                if (b.h_in && b.bytes) memcpy(pbase + b.off, b.h_in, b.bytes);
            if (in_end && cudaMemcpyAsync(dbase, pbase, in_end, cudaMemcpyHostToDevice, st) != cudaSuccess)
                return check_last("H2D copy");
        } else {
            for (auto& b : bufs)
                if (b.h_in && b.bytes &&
                    cudaMemcpyAsync(dbase + b.off, b.h_in, b.bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
                    return check_last("H2D copy");
        }
        return AGNPY_OK;
    }
    int download() {
        size_t out_bytes = total - 256 - in_end;
        if (staged) {
            if (out_bytes && cudaMemcpyAsync(pbase + in_end, dbase + in_end, out_bytes,
                                             cudaMemcpyDeviceToHost, st) != cudaSuccess)
                return check_last("D2H copy");
            if (cudaStreamSynchronize(st) != cudaSuccess) return check_last("host call");
            for (auto& b : bufs)
                if (b.h_out && b.bytes) memcpy(b.h_out, pbase + b.off, b.bytes);
        } else {
            for (auto& b : bufs)
                if (b.h_out && b.bytes &&
                    cudaMemcpyAsync(b.h_out, dbase + b.off, b.bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                    return check_last("D2H copy");
            if (cudaStreamSynchronize(st) != cudaSuccess) return check_last("host call");
        }
        return AGNPY_OK;
    }
};

}  // namespace agb

extern "C" {

int agnpy_b200_synch_sed_host(int n_models, const agnpy_model_t* models, int ne_kind,
                              const double* nu, int n_nu, const double* gamma, int n_gamma,
                              const double* ne_table, const double* ssa_table, int ssa,
                              int integ, double* sed, double* tau_ssa) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || !models || !nu || !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "synch_sed_host: bad sizes or null pointer");
    size_t out_n = (size_t)n_models * n_nu;
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_nu = io.in(nu, 8 * (size_t)n_nu);
    int i_g = io.in(gamma, 8 * (size_t)n_gamma), i_ne = io.in(ne_table, 8 * (size_t)n_gamma * n_models);
    int i_sa = io.in(ssa_table, 8 * (size_t)n_gamma * n_models);
    int o_sed = io.out(sed, 8 * out_n), o_tau = io.out(tau_ssa, 8 * out_n);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_synch_sed(n_models, io.dev<agnpy_model_t>(i_m), ne_kind, io.dev<double>(i_nu),
                                      n_nu, io.dev<double>(i_g), n_gamma, io.dev<double>(i_ne),
                                      io.dev<double>(i_sa), ssa, integ, io.dev<double>(o_sed),
                                      io.dev<double>(o_tau), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_ssc_sed_host(int n_models, const agnpy_model_t* models, int ne_kind,
                            const double* nu, int n_nu, const double* nu_seed, int n_seed,
                            const double* gamma, int n_gamma, const double* ne_table,
                            const double* ssa_table, int ssa, int integ, double* sed) {
    if (int rc = need_device()) return rc;
    for (int j = 1; nu_seed && j < n_seed; ++j)
        if (!(nu_seed[j] > nu_seed[j - 1]) || !(nu_seed[0] > 0.0))
            return fail(AGNPY_ERR_ARG, "ssc_sed_host: nu_seed must be strictly ascending and positive");
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || n_seed < 2 || !models || !nu || !nu_seed ||
        !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "ssc_sed_host: bad sizes or null pointer");
    size_t out_n = (size_t)n_models * n_nu;
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_nu = io.in(nu, 8 * (size_t)n_nu);
    int i_sd = io.in(nu_seed, 8 * (size_t)n_seed), i_g = io.in(gamma, 8 * (size_t)n_gamma);
    int i_ne = io.in(ne_table, 8 * (size_t)n_gamma * n_models), i_sa = io.in(ssa_table, 8 * (size_t)n_gamma * n_models);
    int o_sed = io.out(sed, 8 * out_n);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_ssc_sed(n_models, io.dev<agnpy_model_t>(i_m), ne_kind, io.dev<double>(i_nu),
                                    n_nu, io.dev<double>(i_sd), n_seed, io.dev<double>(i_g), n_gamma,
                                    io.dev<double>(i_ne), io.dev<double>(i_sa), ssa, integ,
                                    io.dev<double>(o_sed), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_ec_sed_host(int variant, int n_models, const agnpy_model_t* models, int ne_kind,
                           const double* nu, int n_nu, const double* gamma, int n_gamma,
                           const double* mu, int n_mu, const double* phi, int n_phi,
                           const double* ne_table, int integ, double* sed) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || !models || !nu || !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "ec_sed_host: bad sizes or null pointer");
    size_t out_n = (size_t)n_models * n_nu;
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_nu = io.in(nu, 8 * (size_t)n_nu);
    int i_g = io.in(gamma, 8 * (size_t)n_gamma);
    int i_mu = io.in(mu, 8 * (size_t)(n_mu > 0 ? n_mu : 0)), i_ph = io.in(phi, 8 * (size_t)(n_phi > 0 ? n_phi : 0));
    int i_ne = io.in(ne_table, 8 * (size_t)n_gamma * n_models);
    int o_sed = io.out(sed, 8 * out_n);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_ec_sed(variant, n_models, io.dev<agnpy_model_t>(i_m), ne_kind,
                                   io.dev<double>(i_nu), n_nu, io.dev<double>(i_g), n_gamma,
                                   io.dev<double>(i_mu), n_mu, io.dev<double>(i_ph), n_phi,
                                   io.dev<double>(i_ne), integ, io.dev<double>(o_sed), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_tau_host(int variant, int n_models, const agnpy_model_t* models,
                        const double* nu, int n_nu, const double* mu, int n_a,
                        const double* phi, int n_phi, const double* l_unit, int n_l,
                        double* tau) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || !models || !nu || !tau)
        return fail(AGNPY_ERR_ARG, "tau_host: bad sizes or null pointer");
    size_t out_n = (size_t)n_models * n_nu;
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_nu = io.in(nu, 8 * (size_t)n_nu);
    int i_mu = io.in(mu, 8 * (size_t)(n_a > 0 ? n_a : 0)), i_ph = io.in(phi, 8 * (size_t)(n_phi > 0 ? n_phi : 0));
    int i_l = io.in(l_unit, 8 * (size_t)(n_l > 0 ? n_l : 0));
    int o_tau = io.out(tau, 8 * out_n);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_tau(variant, n_models, io.dev<agnpy_model_t>(i_m), io.dev<double>(i_nu), n_nu,
                                io.dev<double>(i_mu), n_a, io.dev<double>(i_ph), n_phi,
                                io.dev<double>(i_l), n_l, io.dev<double>(o_tau), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_trapz_loglog_host(const double* y, const double* x, long long n_rows, int n_x,
                                 int intervals, double* out) {
    if (int rc = need_device()) return rc;
    if (n_rows < 1 || n_x < 2 || !y || !x || !out)
        return fail(AGNPY_ERR_ARG, "trapz_loglog_host: bad sizes or null pointer");
    size_t out_n = (size_t)n_rows * (intervals ? (size_t)(n_x - 1) : 1);
    HostIO io;
    int i_y = io.in(y, 8 * (size_t)n_rows * n_x), i_x = io.in(x, 8 * (size_t)n_x);
    int o = io.out(out, 8 * out_n);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_trapz_loglog(io.dev<double>(i_y), io.dev<double>(i_x), n_rows, n_x, intervals,
                                         io.dev<double>(o), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_bb_sed_host(int variant, int n_models, const agnpy_model_t* models,
                           const double* nu, int n_nu, const double* nu_norm, int n_norm, int n_R,
                           double* sed) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || !models || !nu || !sed)
        return fail(AGNPY_ERR_ARG, "bb_sed_host: bad sizes or null pointer");
    size_t out_n = (size_t)n_models * n_nu;
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_nu = io.in(nu, 8 * (size_t)n_nu);
    int i_nn = io.in(nu_norm, 8 * (size_t)(n_norm > 0 ? n_norm : 0));
    int o_sed = io.out(sed, 8 * out_n);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_bb_sed(variant, n_models, io.dev<agnpy_model_t>(i_m), io.dev<double>(i_nu),
                                   n_nu, io.dev<double>(i_nn), n_norm, n_R, io.dev<double>(o_sed), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_tau_synch_host(int n_models, const agnpy_model_t* models, int ne_kind,
                              const double* nu, int n_nu, const double* gamma, int n_gamma,
                              const double* ne_table, const double* ssa_table, int n_s,
                              double delta_margin_low, int integ, double* tau) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || !models || !nu || !gamma || !tau)
        return fail(AGNPY_ERR_ARG, "tau_synch_host: bad sizes or null pointer");
    size_t out_n = (size_t)n_models * n_nu;
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_nu = io.in(nu, 8 * (size_t)n_nu);
    int i_g = io.in(gamma, 8 * (size_t)n_gamma), i_ne = io.in(ne_table, 8 * (size_t)n_gamma * n_models);
    int i_sa = io.in(ssa_table, 8 * (size_t)n_gamma * n_models);
    int o_tau = io.out(tau, 8 * out_n);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_tau_synch(n_models, io.dev<agnpy_model_t>(i_m), ne_kind, io.dev<double>(i_nu),
                                      n_nu, io.dev<double>(i_g), n_gamma, io.dev<double>(i_ne),
                                      io.dev<double>(i_sa), n_s, delta_margin_low, integ,
                                      io.dev<double>(o_tau), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_proton_synch_sed_host(int n_models, const agnpy_model_t* models, int ne_kind,
                                     const double* nu, int n_nu, const double* gamma, int n_gamma,
                                     const double* ne_table, int integ, double* sed) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || !models || !nu || !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "proton_synch_sed_host: bad sizes or null pointer");
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_nu = io.in(nu, 8 * (size_t)n_nu);
    int i_g = io.in(gamma, 8 * (size_t)n_gamma), i_ne = io.in(ne_table, 8 * (size_t)n_gamma * n_models);
    int o_sed = io.out(sed, 8 * (size_t)n_models * n_nu);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_proton_synch_sed(n_models, io.dev<agnpy_model_t>(i_m), ne_kind,
                                             io.dev<double>(i_nu), n_nu, io.dev<double>(i_g), n_gamma,
                                             io.dev<double>(i_ne), integ, io.dev<double>(o_sed), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_ssc_loss_rate_host(int n_models, const agnpy_model_t* models, int ne_kind,
                                  const double* gamma_eval, int n_eval, const double* nu_seed,
                                  int n_seed, const double* gamma, int n_gamma,
                                  const double* ne_table, double* out) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_eval < 1 || n_seed < 2 || n_gamma < 2 || !models || !gamma_eval || !nu_seed ||
        !gamma || !out)
        return fail(AGNPY_ERR_ARG, "ssc_loss_rate_host: bad sizes or null pointer");
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_ge = io.in(gamma_eval, 8 * (size_t)n_eval);
    int i_sd = io.in(nu_seed, 8 * (size_t)n_seed), i_g = io.in(gamma, 8 * (size_t)n_gamma);
    int i_ne = io.in(ne_table, 8 * (size_t)n_gamma * n_models);
    int o = io.out(out, 8 * (size_t)n_models * n_eval);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_ssc_loss_rate(n_models, io.dev<agnpy_model_t>(i_m), ne_kind, io.dev<double>(i_ge),
                                          n_eval, io.dev<double>(i_sd), n_seed, io.dev<double>(i_g), n_gamma,
                                          io.dev<double>(i_ne), io.dev<double>(o), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_target_u_host(int variant, int n_models, const agnpy_model_t* models,
                             const double* Gamma, const double* r, int n_r, int n_mu, double* u) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_r < 1 || !models || !r || !u)
        return fail(AGNPY_ERR_ARG, "target_u_host: bad sizes or null pointer");
    HostIO io;
    int i_m = io.in(models, sizeof(agnpy_model_t) * n_models), i_G = io.in(Gamma, 8 * (size_t)n_models);
    int i_r = io.in(r, 8 * (size_t)n_r);
    int o = io.out(u, 8 * (size_t)n_models * n_r);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (int rc = agnpy_b200_target_u(variant, n_models, io.dev<agnpy_model_t>(i_m), io.dev<double>(i_G),
                                     io.dev<double>(i_r), n_r, n_mu, io.dev<double>(o), io.st))
        return rc;
    return io.download();
}

int agnpy_b200_ebl_absorption_host(int n_models, const double* z, const double* nu, int n_nu,
                                   const double* energy_ref, int n_e, const double* z_ref, int n_z,
                                   const double* values, int multiply, double* out) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_e < 2 || n_z < 2 || !z || !nu || !energy_ref || !z_ref ||
        !values || !out)
        return fail(AGNPY_ERR_ARG, "ebl_absorption_host: bad sizes or null pointer");
    const size_t out_bytes = 8 * (size_t)n_models * n_nu;
    HostIO io;
    int i_z = io.in(z, 8 * (size_t)n_models), i_nu = io.in(nu, 8 * (size_t)n_nu);
    int i_e = io.in(energy_ref, 8 * (size_t)n_e), i_zr = io.in(z_ref, 8 * (size_t)n_z);
    int i_v = io.in(values, 8 * (size_t)n_e * n_z);
    int i_sed = io.in(multiply ? out : nullptr, out_bytes);  // in-place flavour: the SEDs go up too
    int o = io.out(out, out_bytes);
    if (int rc = io.commit()) return rc;
    if (int rc = io.upload()) return rc;
    if (multiply && cudaMemcpyAsync(io.dev<double>(o), io.dev<double>(i_sed), out_bytes,
                                    cudaMemcpyDeviceToDevice, io.st) != cudaSuccess)
        return check_last("ebl_absorption_host: D2D copy");
    if (int rc = agnpy_b200_ebl_absorption(n_models, io.dev<double>(i_z), io.dev<double>(i_nu), n_nu,
                                           io.dev<double>(i_e), n_e, io.dev<double>(i_zr), n_z,
                                           io.dev<double>(i_v), multiply, io.dev<double>(o), io.st))
        return rc;
    return io.download();
}

}  // extern "C"
