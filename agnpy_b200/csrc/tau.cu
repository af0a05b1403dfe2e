// K4: gamma-gamma opacity kernels (ss_disk, blr, dt).
//
// Reference arithmetic replaced:
//   agnpy/absorption/targets_absorption.py:34-42 (sigma), 233-264, 330-353, 518-532
//   agnpy/utils/geometry.py:5-55, agnpy/targets/targets.py:307-315
//
// tau[nu] = pref * sum_t W_t sigma(s_t(nu)),   s = eps_1 * p_t * (1 - cos psi_t) / 2
// where t runs over the (mu|R_tilde, phi, l) grid.  A table kernel evaluates everything that
// does not depend on nu (geometry, np.trapz weights, target energy) once per model; the main
// kernel streams the table (L2-resident, coalesced) and evaluates sigma for TN frequencies
// per table entry in registers.  s is formed with the reference's operation order so the
// threshold mask s < 1 is bit-exact; 1/s comes from a tabulated reciprocal (one DMUL), and
// ln((1+b)/(1-b)) is evaluated as ln((1+b)^2 s), so a point costs one sqrt and one log and
// no division.
#include "common.cuh"
#include "kernels.cuh"

namespace agb {

__device__ __forceinline__ double cos_psi_t(double mu_s, double mu, double phi) {
    double t1 = mu * mu_s;
    double t2 = sqrt(1.0 - mu * mu) * sqrt(1.0 - mu_s * mu_s);
    return t1 + t2 * cos(phi);
}
__device__ __forceinline__ double lin_at(double start, double stop, int n, int i) {
    if (n == 1) return start;
    if (i == n - 1) return stop;
    double step = (stop - start) / (double)(n - 1);
    return (double)i * step + start;
}
// targets_absorption.py:333-338: l = logspace(0,5,n)*r, nudged off R_line (np.isclose, rtol 1e-4)
__device__ __forceinline__ double blr_l_at(const double* l_unit, int i, double r, double R_line) {
    double l = l_unit[i] * r;
    if (fabs(l - R_line) <= 1e-8 + 1.0e-4 * fabs(R_line)) l += 1.0e-4 * R_line;
    return l;
}
__device__ __forceinline__ double disk_eps(double L_disk, double M_BH, double eta, double R_tilde) {
    double M_8 = M_BH / (1e8 * kMsun);
    double L_Edd = 1.26 * 1e46 * M_8;
    double l_Edd = L_disk / L_Edd;
    double xi = pow(l_Edd / (M_8 * eta), 1.0 / 4.0);
    return 2.7 * 1e-4 * xi * pow(R_tilde, -3.0 / 4.0);
}

// table[m][0][t] = p_t (target energy)   table[m][1][t] = (1 - cos psi)/2
// table[m][2][t] = 1/(p_t (1-cos psi)/2)  table[m][3][t] = W_t
__global__ void __launch_bounds__(128) tau_table_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ mu, int n_a,
    const double* __restrict__ phi, int n_phi, const double* __restrict__ l_unit, int n_l,
    int n_t, double* __restrict__ table) {
    int m = blockIdx.y;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_t) return;
    const agnpy_model_t& M = models[m];
    const double* g = M.tgt;
    int il = t % n_l;
    int iphi = (t / n_l) % n_phi;
    int ia = t / (n_l * n_phi);
    double wphi = trapz_weight(phi, iphi, n_phi);
    double p, omc, W;
    if (variant == AGNPY_TAU_DT) {  // targets_absorption.py:520-530
        double R_dt = g[3], r = g[4];
        double l = l_unit[il] * r;
        double llo = l_unit[il > 0 ? il - 1 : 0] * r, lhi = l_unit[il < n_l - 1 ? il + 1 : n_l - 1] * r;
        double wl = n_l < 2 ? 0.0 : (lhi - llo) * 0.5;
        double x = sqrt(R_dt * R_dt + l * l);
        double mu_l = l / x;
        omc = 1.0 - cos_psi_t(M.mu_s, mu_l, phi[iphi]);
        p = g[2];
        W = wphi * wl * (omc / (x * x));
    } else if (variant == AGNPY_TAU_BLR) {  // :333-349
        double R_line = g[3], r = g[4];
        double l = blr_l_at(l_unit, il, r, R_line);
        double llo = blr_l_at(l_unit, il > 0 ? il - 1 : 0, r, R_line);
        double lhi = blr_l_at(l_unit, il < n_l - 1 ? il + 1 : n_l - 1, r, R_line);
        double wl = n_l < 2 ? 0.0 : (lhi - llo) * 0.5;
        double mu_i = mu[ia];
        double x = sqrt(R_line * R_line + l * l - 2.0 * l * R_line * mu_i);
        double ratio = R_line / x;
        double ms = sqrt(1.0 - ratio * ratio * (1.0 - mu_i * mu_i));
        ms = ms * ((((l > mu_i * R_line) ? 1.0 : 0.0) - 0.5) * 2.0);
        omc = 1.0 - cos_psi_t(M.mu_s, ms, phi[iphi]);
        p = g[2];
        W = trapz_weight(mu, ia, n_a) * wphi * wl * (omc / (x * x));
    } else {  // AGNPY_TAU_SS_DISK :233-262
        double M_BH = g[0], L_disk = g[1], eta = g[2];
        double R_g = kG * M_BH / (kC * kC);
        double r_t = g[5] / R_g, Rin_t = g[3] / R_g, Rout_t = g[4] / R_g;
        double R = lin_at(Rin_t, Rout_t, n_a, ia);
        double Rlo = lin_at(Rin_t, Rout_t, n_a, ia > 0 ? ia - 1 : 0);
        double Rhi = lin_at(Rin_t, Rout_t, n_a, ia < n_a - 1 ? ia + 1 : n_a - 1);
        double wR = n_a < 2 ? 0.0 : (Rhi - Rlo) * 0.5;
        double l = l_unit[il] * r_t;
        double llo = l_unit[il > 0 ? il - 1 : 0] * r_t, lhi = l_unit[il < n_l - 1 ? il + 1 : n_l - 1] * r_t;
        double wl = n_l < 2 ? 0.0 : (lhi - llo) * 0.5;
        p = disk_eps(L_disk, M_BH, eta, R);
        double phi_disk = 1.0 - sqrt(Rin_t / R);
        double q = 1.0 + (R * R) / (l * l);
        double mu_l = pow(q, -1.0 / 2.0);
        omc = 1.0 - cos_psi_t(M.mu_s, mu_l, phi[iphi]);
        W = wR * wphi * wl * (1.0 / (l * l) / (R * R) / pow(q, 3.0 / 2.0) * phi_disk / p * omc);
    }
    double* row = table + (size_t)m * 4 * n_t;
    double homc = omc / 2.0;
    row[t] = p;
    row[n_t + t] = homc;
    row[2 * (size_t)n_t + t] = 1.0 / (p * homc);
    row[3 * (size_t)n_t + t] = W;
}

// sigma(s) / (3/16 sigma_T), targets_absorption.py:34-42, with 1-beta^2 = 1/s
__device__ __forceinline__ double sigma_shape(double s, double inv_s) {
    double b2 = 1.0 - inv_s;
    double beta = sqrt(b2);
    double opb = 1.0 + beta;
    double L = log(opb * opb * s);  // = ln((1+beta)/(1-beta))
    double t1 = (3.0 - b2 * b2) * L;
    double t2 = (beta + beta) * (2.0 - b2);
    double v = inv_s * (t1 - t2);
    return (s < 1.0) ? 0.0 : v;
}

constexpr int kTauTN = 4;        // frequencies per thread
constexpr int kTauSlice = 4096;  // table entries per block

// grid = (n_slices, ceil(n_nu/TN), n_models); part[m][i][slice]
__global__ void __launch_bounds__(256) tau_main_kernel(
    const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ table, int n_t, double* __restrict__ part) {
    __shared__ double red[32];
    const int slice = blockIdx.x, m = blockIdx.z, S = gridDim.x;
    const int i0 = blockIdx.y * kTauTN;
    const agnpy_model_t& M = models[m];
    const double* t_p = table + (size_t)m * 4 * n_t;
    const double* t_h = t_p + n_t;
    const double* t_ic = t_h + n_t;
    const double* t_W = t_ic + n_t;
    double e1[kTauTN], ie1[kTauTN], acc[kTauTN];
#pragma unroll
    for (int q = 0; q < kTauTN; ++q) {
        int i = min(i0 + q, n_nu - 1);
        e1[q] = nu_to_eps(nu[i], M.z, 1.0);  // targets_absorption.py:331: no Doppler factor
        ie1[q] = 1.0 / e1[q];
        acc[q] = 0.0;
    }
    const int t_end = min(n_t, (slice + 1) * kTauSlice);
    for (int t = slice * kTauSlice + threadIdx.x; t < t_end; t += blockDim.x) {
        double p = t_p[t], h = t_h[t], ic = t_ic[t], W = t_W[t];
#pragma unroll
        for (int q = 0; q < kTauTN; ++q) {
            double s = e1[q] * p * h;  // == eps_1 * eps * (1 - cos psi) / 2, same rounding
            // below the pair-production threshold sigma is exactly 0 (targets_absorption.py:41):
            // skip the sqrt/log when no lane of the warp is above it (NaN s keeps the slow path)
            if (__all_sync(__activemask(), s < 1.0 && W == W && fabs(W) <= kFmax)) continue;
            double v = sigma_shape(s, ic * ie1[q]);
            acc[q] = fma(W, v, acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < kTauTN; ++q) {
        double v = block_sum(acc[q], red);
        if (threadIdx.x == 0 && i0 + q < n_nu) part[((size_t)m * n_nu + i0 + q) * S + slice] = v;
    }
}

__global__ void __launch_bounds__(128) tau_finish_kernel(
    int variant, const agnpy_model_t* __restrict__ models, int n_nu,
    const double* __restrict__ part, int S, double* __restrict__ tau) {
    int m = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nu) return;
    const agnpy_model_t& M = models[m];
    const double* g = M.tgt;
    const double* p = part + ((size_t)m * n_nu + i) * S;
    double integral = 0.0;
    for (int s = 0; s < S; ++s) integral += p[s];
    integral = integral * (3.0 / 16.0 * kSigmaT);
    double c3 = kC * kC * kC;
    double pref;
    if (variant == AGNPY_TAU_BLR)  // :350-352
        pref = (g[0] * g[1]) / ((4.0 * kPi) * (4.0 * kPi) * g[2] * kMe * c3);
    else if (variant == AGNPY_TAU_DT)  // :531
        pref = (g[0] * g[1]) / (8.0 * (kPi * kPi) * g[2] * kMe * c3);
    else {  // :263
        double R_g = kG * g[0] / (kC * kC);
        pref = 3.0 * g[1] / ((4.0 * kPi) * (4.0 * kPi) * g[2] * kMe * c3 * R_g);
    }
    tau[(size_t)m * n_nu + i] = pref * integral;
}

static int tau_n_t(int variant, int n_a, int n_phi, int n_l) {
    return (variant == AGNPY_TAU_DT ? 1 : n_a) * n_phi * n_l;
}

size_t tau_workspace_bytes(int variant, int n_models, int n_nu, int n_a, int n_phi, int n_l) {
    size_t n_t = tau_n_t(variant, n_a, n_phi, n_l);
    size_t S = (n_t + kTauSlice - 1) / kTauSlice;
    return (4 * n_t + (size_t)n_nu * S) * sizeof(double) * (size_t)n_models;
}

int run_tau(int variant, const agnpy_model_t* models, int n_models, const double* nu, int n_nu,
            const double* mu, int n_a, const double* phi, int n_phi, const double* l_unit, int n_l,
            double* tau, double* ws, cudaStream_t stream) {
    int n_t = tau_n_t(variant, n_a, n_phi, n_l);
    int S = (n_t + kTauSlice - 1) / kTauSlice;
    double* table = ws;
    double* part = table + (size_t)n_models * 4 * n_t;
    {
        dim3 grid((n_t + 127) / 128, n_models);
        tau_table_kernel<<<grid, 128, 0, stream>>>(variant, models, mu, n_a, phi, n_phi, l_unit,
                                                   n_l, n_t, table);
        note_launch();
    }
    {
        dim3 grid(S, (n_nu + kTauTN - 1) / kTauTN, n_models);
        timing_begin(stream);
        tau_main_kernel<<<grid, 256, 0, stream>>>(models, nu, n_nu, table, n_t, part);
        timing_end(stream);
        note_launch();
    }
    {
        dim3 grid((n_nu + 127) / 128, n_models);
        tau_finish_kernel<<<grid, 128, 0, stream>>>(variant, models, n_nu, part, S, tau);
        note_launch();
    }
    return check_last("tau");
}

}  // namespace agb
