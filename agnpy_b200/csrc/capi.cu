// extern "C" entry points declared in include/agnpy_b200.h: argument checks, workspace
// carving, batching of the model axis, and the host-pointer flavours (H2D, run, D2H, sync).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace agb {

constexpr size_t kChunkBytes = (size_t)1 << 30;  // scratch budget per batch chunk

static bool bad_kind(int k) { return k < AGNPY_NE_POWERLAW || k > AGNPY_NE_TABLE; }
static bool bad_integ(int i) { return i != AGNPY_INTEG_TRAPZ && i != AGNPY_INTEG_LOGLOG; }

static int chunk_models(size_t bytes_per_model, int n_models) {
    size_t c = kChunkBytes / std::max<size_t>(bytes_per_model, 1);
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
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || !models || !nu || !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "synch_sed: bad sizes or null pointer");
    if (bad_kind(ne_kind) || bad_integ(integ)) return fail(AGNPY_ERR_ARG, "synch_sed: bad ne_kind/integ");
    if (n_nu > 65535) return fail(AGNPY_ERR_ARG, "synch_sed: n_nu > 65535");
    if (ne_kind == AGNPY_NE_TABLE && (n_models != 1 || !ne_table || (ssa && !ssa_table)))
        return fail(AGNPY_ERR_ARG, "synch_sed: AGNPY_NE_TABLE needs n_models == 1 and tables");
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = 3 * (size_t)n_gamma * sizeof(double);
    int chunk = chunk_models(per, n_models);
    double* tab = (double*)workspace(per * chunk, stream);
    if (!tab) return fail(AGNPY_ERR_NOMEM, "synch_sed: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        launch_prep_gamma(models + m0, nm, ne_kind, gamma, n_gamma, ne_table, ssa_table, 0, tab, stream);
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
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_seed < 2 || n_gamma < 2 || !models || !nu || !nu_seed ||
        !gamma || !sed)
        return fail(AGNPY_ERR_ARG, "ssc_sed: bad sizes or null pointer");
    if (bad_kind(ne_kind) || bad_integ(integ)) return fail(AGNPY_ERR_ARG, "ssc_sed: bad ne_kind/integ");
    if (n_nu > 65535 || n_seed > 65535) return fail(AGNPY_ERR_ARG, "ssc_sed: n_nu/n_seed > 65535");
    if (ne_kind == AGNPY_NE_TABLE && (n_models != 1 || !ne_table || (ssa && !ssa_table)))
        return fail(AGNPY_ERR_ARG, "ssc_sed: AGNPY_NE_TABLE needs n_models == 1 and tables");
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = (3 * (size_t)n_gamma + 2 * (size_t)n_seed) * sizeof(double);
    int chunk = chunk_models(per, n_models);
    double* ws = (double*)workspace(per * chunk, stream);
    if (!ws) return fail(AGNPY_ERR_NOMEM, "ssc_sed: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        double* tab = ws;
        double* U = tab + (size_t)nm * 3 * n_gamma;
        double* es = U + (size_t)nm * n_seed;
        launch_prep_gamma(models + m0, nm, ne_kind, gamma, n_gamma, ne_table, ssa_table, 0, tab, stream);
        // seed synchrotron spectrum at nu_seed (synchrotron_self_compton.py:129-142)
        if (int rc = launch_synch(models + m0, nm, nu_seed, n_seed, tab, n_gamma, ssa, integ, 1, U,
                                  nullptr, es, stream))
            return rc;
        timing_begin(stream);
        int rc_ssc = launch_ssc(models + m0, nm, nu, n_nu, tab, n_gamma, U, es, n_seed, integ,
                                sed + (size_t)m0 * n_nu, stream);
        timing_end(stream);
        if (rc_ssc) return rc_ssc;
    }
    return check_last("ssc_sed");
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_ec_sed(int variant, int n_models, const agnpy_model_t* models, int ne_kind,
                      const double* nu, int n_nu, const double* gamma, int n_gamma,
                      const double* mu, int n_mu, const double* phi, int n_phi,
                      const double* ne_table, int integ, double* sed, void* stream_) {
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
    if (ne_kind == AGNPY_NE_TABLE && (n_models != 1 || !ne_table))
        return fail(AGNPY_ERR_ARG, "ec_sed: AGNPY_NE_TABLE needs n_models == 1 and a table");
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t per = ec_workspace_bytes(variant, 1, n_nu, n_gamma, n_mu, n_phi);
    int chunk = chunk_models(per, n_models);
    double* ws = (double*)workspace(per * chunk, stream);
    if (!ws) return fail(AGNPY_ERR_NOMEM, "ec_sed: workspace");
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        int nm = std::min(chunk, n_models - m0);
        if (int rc = run_ec(variant, models + m0, nm, ne_kind, nu, n_nu, gamma, n_gamma, mu, n_mu,
                            phi, n_phi, ne_table, integ, sed + (size_t)m0 * n_nu, ws, stream))
            return rc;
    }
    return AGNPY_OK;
}

// ------------------------------------------------------------------------------------------
int agnpy_b200_tau(int variant, int n_models, const agnpy_model_t* models, const double* nu,
                   int n_nu, const double* mu, int n_a, const double* phi, int n_phi,
                   const double* l_unit, int n_l, double* tau, void* stream_) {
    if (int rc = need_device()) return rc;
    if (variant < AGNPY_TAU_SS_DISK || variant > AGNPY_TAU_DT) return fail(AGNPY_ERR_ARG, "tau: bad variant");
    if (n_models < 1 || n_nu < 1 || n_phi < 2 || n_l < 2 || !models || !nu || !phi || !l_unit || !tau)
        return fail(AGNPY_ERR_ARG, "tau: bad sizes or null pointer");
    if (variant == AGNPY_TAU_BLR && (!mu || n_a < 2)) return fail(AGNPY_ERR_ARG, "tau: mu grid required");
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

// ==========================================================================================
// host-pointer flavours: one device arena per call holds inputs and outputs; the kernel
// scratch comes from the stream workspace as in the device flavours.
// ==========================================================================================
#define H2D(dst, src, n, T)                                                                     \
    if ((src) && cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(T), cudaMemcpyHostToDevice, st) \
                     != cudaSuccess)                                                            \
        return check_last("H2D copy");
#define D2H(dst, src, n, T)                                                                     \
    if ((dst) && cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(T), cudaMemcpyDeviceToHost, st) \
                     != cudaSuccess)                                                            \
        return check_last("D2H copy");

static void* io_arena(size_t bytes) {
    // separate cache from the kernel workspace: keyed by a sentinel stream value
    static thread_local void* ptr[64] = {};
    static thread_local size_t size[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    if (size[dev] >= bytes) return ptr[dev];
    if (ptr[dev]) {
        cudaDeviceSynchronize();
        cudaFree(ptr[dev]);
        ptr[dev] = nullptr;
        size[dev] = 0;
    }
    size_t want = bytes + bytes / 4 + (1u << 16);
    if (cudaMalloc(&ptr[dev], want) != cudaSuccess) {
        cudaGetLastError();
        ptr[dev] = nullptr;
        return nullptr;
    }
    size[dev] = want;
    return ptr[dev];
}

int agnpy_b200_synch_sed_host(int n_models, const agnpy_model_t* models, int ne_kind,
                              const double* nu, int n_nu, const double* gamma, int n_gamma,
                              const double* ne_table, const double* ssa_table, int ssa,
                              int integ, double* sed, double* tau_ssa) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2) return fail(AGNPY_ERR_ARG, "synch_sed_host: bad sizes");
    cudaStream_t st = host_stream();
    size_t out_n = (size_t)n_models * n_nu;
    size_t bytes = pad(sizeof(agnpy_model_t) * n_models) + pad(8 * (size_t)n_nu) +
                   3 * pad(8 * (size_t)n_gamma) + 2 * pad(8 * out_n) + 4096;
    void* base = io_arena(bytes);
    if (!base) return fail(AGNPY_ERR_NOMEM, "synch_sed_host: device memory");
    Arena a(base);
    agnpy_model_t* d_models = a.take<agnpy_model_t>(n_models);
    double* d_nu = a.take<double>(n_nu);
    double* d_gamma = a.take<double>(n_gamma);
    double* d_net = ne_table ? a.take<double>(n_gamma) : nullptr;
    double* d_sat = ssa_table ? a.take<double>(n_gamma) : nullptr;
    double* d_sed = a.take<double>(out_n);
    double* d_tau = tau_ssa ? a.take<double>(out_n) : nullptr;
    H2D(d_models, models, n_models, agnpy_model_t);
    H2D(d_nu, nu, n_nu, double);
    H2D(d_gamma, gamma, n_gamma, double);
    H2D(d_net, ne_table, n_gamma, double);
    H2D(d_sat, ssa_table, n_gamma, double);
    if (int rc = agnpy_b200_synch_sed(n_models, d_models, ne_kind, d_nu, n_nu, d_gamma, n_gamma,
                                      d_net, d_sat, ssa, integ, d_sed, d_tau, st))
        return rc;
    D2H(sed, d_sed, out_n, double);
    D2H(tau_ssa, d_tau, out_n, double);
    if (cudaStreamSynchronize(st) != cudaSuccess) return check_last("synch_sed_host");
    return AGNPY_OK;
}

int agnpy_b200_ssc_sed_host(int n_models, const agnpy_model_t* models, int ne_kind,
                            const double* nu, int n_nu, const double* nu_seed, int n_seed,
                            const double* gamma, int n_gamma, const double* ne_table,
                            const double* ssa_table, int ssa, int integ, double* sed) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2 || n_seed < 2)
        return fail(AGNPY_ERR_ARG, "ssc_sed_host: bad sizes");
    cudaStream_t st = host_stream();
    size_t out_n = (size_t)n_models * n_nu;
    size_t bytes = pad(sizeof(agnpy_model_t) * n_models) + pad(8 * (size_t)n_nu) +
                   pad(8 * (size_t)n_seed) + 3 * pad(8 * (size_t)n_gamma) + pad(8 * out_n) + 4096;
    void* base = io_arena(bytes);
    if (!base) return fail(AGNPY_ERR_NOMEM, "ssc_sed_host: device memory");
    Arena a(base);
    agnpy_model_t* d_models = a.take<agnpy_model_t>(n_models);
    double* d_nu = a.take<double>(n_nu);
    double* d_seed = a.take<double>(n_seed);
    double* d_gamma = a.take<double>(n_gamma);
    double* d_net = ne_table ? a.take<double>(n_gamma) : nullptr;
    double* d_sat = ssa_table ? a.take<double>(n_gamma) : nullptr;
    double* d_sed = a.take<double>(out_n);
    H2D(d_models, models, n_models, agnpy_model_t);
    H2D(d_nu, nu, n_nu, double);
    H2D(d_seed, nu_seed, n_seed, double);
    H2D(d_gamma, gamma, n_gamma, double);
    H2D(d_net, ne_table, n_gamma, double);
    H2D(d_sat, ssa_table, n_gamma, double);
    if (int rc = agnpy_b200_ssc_sed(n_models, d_models, ne_kind, d_nu, n_nu, d_seed, n_seed, d_gamma,
                                    n_gamma, d_net, d_sat, ssa, integ, d_sed, st))
        return rc;
    D2H(sed, d_sed, out_n, double);
    if (cudaStreamSynchronize(st) != cudaSuccess) return check_last("ssc_sed_host");
    return AGNPY_OK;
}

int agnpy_b200_ec_sed_host(int variant, int n_models, const agnpy_model_t* models, int ne_kind,
                           const double* nu, int n_nu, const double* gamma, int n_gamma,
                           const double* mu, int n_mu, const double* phi, int n_phi,
                           const double* ne_table, int integ, double* sed) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_gamma < 2) return fail(AGNPY_ERR_ARG, "ec_sed_host: bad sizes");
    cudaStream_t st = host_stream();
    size_t out_n = (size_t)n_models * n_nu;
    size_t bytes = pad(sizeof(agnpy_model_t) * n_models) + pad(8 * (size_t)n_nu) +
                   2 * pad(8 * (size_t)n_gamma) + pad(8 * (size_t)std::max(n_mu, 1)) +
                   pad(8 * (size_t)std::max(n_phi, 1)) + pad(8 * out_n) + 4096;
    void* base = io_arena(bytes);
    if (!base) return fail(AGNPY_ERR_NOMEM, "ec_sed_host: device memory");
    Arena a(base);
    agnpy_model_t* d_models = a.take<agnpy_model_t>(n_models);
    double* d_nu = a.take<double>(n_nu);
    double* d_gamma = a.take<double>(n_gamma);
    double* d_mu = mu ? a.take<double>(n_mu) : nullptr;
    double* d_phi = phi ? a.take<double>(n_phi) : nullptr;
    double* d_net = ne_table ? a.take<double>(n_gamma) : nullptr;
    double* d_sed = a.take<double>(out_n);
    H2D(d_models, models, n_models, agnpy_model_t);
    H2D(d_nu, nu, n_nu, double);
    H2D(d_gamma, gamma, n_gamma, double);
    H2D(d_mu, mu, n_mu, double);
    H2D(d_phi, phi, n_phi, double);
    H2D(d_net, ne_table, n_gamma, double);
    if (int rc = agnpy_b200_ec_sed(variant, n_models, d_models, ne_kind, d_nu, n_nu, d_gamma, n_gamma,
                                   d_mu, n_mu, d_phi, n_phi, d_net, integ, d_sed, st))
        return rc;
    D2H(sed, d_sed, out_n, double);
    if (cudaStreamSynchronize(st) != cudaSuccess) return check_last("ec_sed_host");
    return AGNPY_OK;
}

int agnpy_b200_tau_host(int variant, int n_models, const agnpy_model_t* models,
                        const double* nu, int n_nu, const double* mu, int n_a,
                        const double* phi, int n_phi, const double* l_unit, int n_l,
                        double* tau) {
    if (int rc = need_device()) return rc;
    if (n_models < 1 || n_nu < 1 || n_phi < 2 || n_l < 2) return fail(AGNPY_ERR_ARG, "tau_host: bad sizes");
    cudaStream_t st = host_stream();
    size_t out_n = (size_t)n_models * n_nu;
    size_t bytes = pad(sizeof(agnpy_model_t) * n_models) + pad(8 * (size_t)n_nu) +
                   pad(8 * (size_t)std::max(n_a, 1)) + pad(8 * (size_t)n_phi) + pad(8 * (size_t)n_l) +
                   pad(8 * out_n) + 4096;
    void* base = io_arena(bytes);
    if (!base) return fail(AGNPY_ERR_NOMEM, "tau_host: device memory");
    Arena a(base);
    agnpy_model_t* d_models = a.take<agnpy_model_t>(n_models);
    double* d_nu = a.take<double>(n_nu);
    double* d_mu = mu ? a.take<double>(n_a) : nullptr;
    double* d_phi = a.take<double>(n_phi);
    double* d_l = a.take<double>(n_l);
    double* d_tau = a.take<double>(out_n);
    H2D(d_models, models, n_models, agnpy_model_t);
    H2D(d_nu, nu, n_nu, double);
    H2D(d_mu, mu, n_a, double);
    H2D(d_phi, phi, n_phi, double);
    H2D(d_l, l_unit, n_l, double);
    if (int rc = agnpy_b200_tau(variant, n_models, d_models, d_nu, n_nu, d_mu, n_a, d_phi, n_phi, d_l,
                                n_l, d_tau, st))
        return rc;
    D2H(tau, d_tau, out_n, double);
    if (cudaStreamSynchronize(st) != cudaSuccess) return check_last("tau_host");
    return AGNPY_OK;
}

}  // extern "C"
