// K6: thermal black-body SEDs the fit wrappers add to the non-thermal components
// (SURVEY.md 8(f) rank 1).  Reference arithmetic replaced:
//   agnpy/targets/targets.py:325-330 (SSDisk.evaluate_T), 350-415 (multi-T disk BB and its
//   normalisation), 593-610 (RingDustTorus BB and its normalisation);
//   Planck's law = astropy.modeling.physical_models.BlackBody.evaluate(nu, T, scale=1):
//   B_nu = 2 h nu^3 / (c^2 expm1(h nu / (k_B T))).
// One block per model; 1-D work (100 disk radii x (n_nu + n_norm) frequencies).
#include "common.cuh"
#include "kernels.cuh"

namespace agb {

constexpr double kKb = 1.380649e-16;
constexpr double kSigmaSB = 5.670374419e-5;  // astropy CODATA 2018 literal (5.670374419e-8 SI)

__device__ __forceinline__ double planck_nu(double nu, double T) {
    double log_boltz = kH * nu / (kKb * T);
    double boltzm1 = expm1(log_boltz);
    return 2.0 * kH * (nu * nu * nu) / ((kC * kC) * boltzm1);
}

__device__ __forceinline__ double lin_pt(double start, double stop, int n, int i) {
    if (n == 1) return start;
    if (i == n - 1) return stop;
    double step = (stop - start) / (double)(n - 1);
    return (double)i * step + start;
}

// F_nu of the multi-temperature disk at observed frequency nu (targets.py:381-390)
__device__ __forceinline__ double disk_F_nu(double nu, double z, double d_L, const double* s_R,
                                            const double* s_T, const double* s_w, int n_R) {
    double nu_prime = nu * (1.0 + z);
    double acc = 0.0;
    for (int k = 0; k < n_R; ++k) {
        double integrand = s_R[k] / (d_L * d_L) * planck_nu(nu_prime, s_T[k]);
        acc = fma(s_w[k], integrand, acc);
    }
    return 2.0 * kPi * acc;
}

__global__ void __launch_bounds__(256) bb_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ nu_norm, int n_norm, int n_R, double* __restrict__ sed) {
    extern __shared__ double sh[];
    __shared__ double red[32];
    __shared__ double s_norm;
    const int m = blockIdx.x, tid = threadIdx.x;
    const agnpy_model_t& M = models[m];
    const double* g = M.tgt;
    double* out = sed + (size_t)m * n_nu;
    if (variant == AGNPY_BB_DT || variant == AGNPY_BB_DT_NORM) {
        const double T_dt = g[0], R_dt = g[1];
        double ratio = R_dt / M.d_L;
        double prefactor = kPi * (ratio * ratio);  // targets.py:597
        double norm = 1.0;
        if (variant == AGNPY_BB_DT_NORM) {          // :606-609
            double L_tot = 4.0 * kPi * (R_dt * R_dt) * kSigmaSB * (T_dt * T_dt * T_dt * T_dt);
            norm = g[2] / L_tot;
        }
        for (int i = tid; i < n_nu; i += blockDim.x) {
            double v = prefactor * nu[i] * planck_nu(nu[i] * (1.0 + M.z), T_dt);
            out[i] = (variant == AGNPY_BB_DT_NORM) ? norm * v : v;
        }
        return;
    }
    // Shakura-Sunyaev disk: R = linspace(R_in, R_out, n_R), T(R) of targets.py:325-330
    const double M_BH = g[0], m_dot = g[1], R_in = g[2], R_out = g[3];
    double* s_R = sh;
    double* s_T = s_R + n_R;
    double* s_w = s_T + n_R;
    for (int k = tid; k < n_R; k += blockDim.x) {
        double R = lin_pt(R_in, R_out, n_R, k);
        double phi = 1.0 - sqrt(R_in / R);
        double val = (3.0 * kG * M_BH * m_dot * phi) / (8.0 * kPi * (R * R * R) * kSigmaSB);
        s_R[k] = R;
        s_T[k] = pow(val, 1.0 / 4.0);
        double lo = lin_pt(R_in, R_out, n_R, k > 0 ? k - 1 : 0);
        double hi = lin_pt(R_in, R_out, n_R, k < n_R - 1 ? k + 1 : n_R - 1);
        s_w[k] = n_R < 2 ? 0.0 : (hi - lo) * 0.5;
    }
    __syncthreads();
    double norm = 1.0;
    if (variant == AGNPY_BB_DISK_NORM) {  // targets.py:402-412
        double part = 0.0;
        for (int i = tid; i < n_norm; i += blockDim.x) {
            double nui = nu_norm[i];
            double s = M.mu_s * (nui * disk_F_nu(nui, M.z, M.d_L, s_R, s_T, s_w, n_R));
            part = fma(trapz_weight(nu_norm, i, n_norm), s / nui, part);
        }
        part = block_sum(part, red);
        if (tid == 0) {
            double A = 4.0 * kPi * (M.d_L * M.d_L);
            double L = 2.0 * (part * A);
            s_norm = g[4] / L;
        }
        __syncthreads();
        norm = s_norm;
    }
    for (int i = tid; i < n_nu; i += blockDim.x) {
        double v = M.mu_s * (nu[i] * disk_F_nu(nu[i], M.z, M.d_L, s_R, s_T, s_w, n_R));
        out[i] = (variant == AGNPY_BB_DISK_NORM) ? norm * v : v;
    }
}

int run_bb(int variant, const agnpy_model_t* models, int n_models, const double* nu, int n_nu,
           const double* nu_norm, int n_norm, int n_R, double* sed, cudaStream_t stream) {
    size_t smem = 3 * (size_t)n_R * sizeof(double);
    if (smem > 48 * 1024) return fail(AGNPY_ERR_ARG, "bb_sed: too many disk radii");
    timing_begin(stream);
    bb_kernel<<<n_models, 256, smem, stream>>>(variant, models, nu, n_nu, nu_norm, n_norm, n_R, sed);
    timing_end(stream);
    note_launch();
    return check_last("bb_sed");
}

}  // namespace agb
