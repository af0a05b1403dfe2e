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

// ---- off-axis geometry, utils/geometry.py:58-183, in the reference's operation order ----------
__device__ __forceinline__ double x_ring_mu_s(double R, double r, double phi_re, double u, double mu_s) {
    double sin_s = sqrt(1.0 - mu_s * mu_s);
    double x2 = u * u + R * R + r * r - 2.0 * u * R * sin_s * cos(phi_re) + 2.0 * r * u * mu_s;
    return sqrt(x2);
}
__device__ __forceinline__ double x_shell_mu_s(double R, double r, double phi_re, double mu_re, double u,
                                               double mu_s) {
    double sin_re = sqrt(1.0 - mu_re * mu_re);
    double sin_s = sqrt(1.0 - mu_s * mu_s);
    double x2 = R * R + u * u + r * r - 2.0 * R * u * sin_re * cos(phi_re) * sin_s -
                2.0 * R * mu_re * (r + u * mu_s) + 2.0 * r * u * mu_s;
    return sqrt(x2);
}

// targets_absorption.py:405-415: the off-axis BLR path grid uu = l_unit * r, nudged off the
// shell crossing (np.isclose, rtol 1e-4) and re-sorted when anything was nudged.
// path[m][i], one block per model, rank sort (n_l is ~100).
__global__ void __launch_bounds__(128) tau_blr_path_kernel(
    const agnpy_model_t* __restrict__ models, const double* __restrict__ l_unit, int n_l,
    double* __restrict__ path) {
    extern __shared__ double s_u[];
    __shared__ int s_any;
    const int m = blockIdx.x;
    const agnpy_model_t& M = models[m];
    const double R_line = M.tgt[3], r = M.tgt[4], mu_s = M.mu_s;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_l; i += blockDim.x) {
        double u = l_unit[i] * r;
        double x_cross = sqrt(r * r + u * u + 2.0 * u * r * mu_s);
        if (fabs(x_cross - R_line) <= 1e-8 + 1.0e-4 * fabs(R_line)) {
            u += 1.0e-4 * R_line;
            s_any = 1;
        }
        s_u[i] = u;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_l; i += blockDim.x) {
        double u = s_u[i];
        int rank = i;
        if (s_any) {  // stable rank = position after np.sort
            rank = 0;
            for (int j = 0; j < n_l; ++j) rank += (s_u[j] < u) || (s_u[j] == u && j < i);
        }
        path[(size_t)m * n_l + rank] = u;
    }
}

// table[m][0][t] = p_t (target energy)   table[m][1][t] = (1 - cos psi)/2
// table[m][2][t] = 1/(p_t (1-cos psi)/2)  table[m][3][t] = W_t
// one (a, phi, l) entry of the on-axis variants (dt, blr, ss_disk): target energy p, 1 - cos psi
// and the weight W = np.trapz weights x geometry, in the reference's operation order
__device__ __forceinline__ void tau_onaxis_entry(
    int variant, const agnpy_model_t& M, const double* __restrict__ mu, int n_a,
    const double* __restrict__ phi, int n_phi, const double* __restrict__ l_unit, int n_l, int ia,
    int iphi, int il, double& p, double& omc, double& W) {
    const double* g = M.tgt;
    const double wphi = trapz_weight(phi, iphi, n_phi);
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
}

// phi[i] + phi[j] = 2 pi m for an integer m: cos phi takes the same value at both points
__device__ __forceinline__ bool phi_mirror(const double* __restrict__ phi, int i, int j) {
    const double s = phi[i] + phi[j];
    const double m = rint(s / (2.0 * kPi));
    const double tol = 1e-13 * fmax(1.0, fmax(fabs(phi[i]), fabs(phi[j])));
    return fabs(s - 2.0 * kPi * m) <= tol;
}

__global__ void __launch_bounds__(128) tau_table_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ mu, int n_a,
    const double* __restrict__ phi, int n_phi, const double* __restrict__ l_unit, int n_l,
    int n_t, const double* __restrict__ path, double* __restrict__ table, int* __restrict__ n_live) {
    int m = blockIdx.y;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    // is the whole phi grid mirror-symmetric?  (decided per block: every thread needs the answer
    // to place its entry in the compacted table)
    __shared__ int s_all_mirror;
    if (threadIdx.x == 0) s_all_mirror = 1;
    __syncthreads();
    if (n_phi >= 2) {
        for (int i = threadIdx.x; 2 * i < n_phi - 1; i += blockDim.x)
            if (!phi_mirror(phi, i, n_phi - 1 - i)) s_all_mirror = 0;
    }
    __syncthreads();
    const bool all_mirror = s_all_mirror != 0 && n_phi >= 2;
    if (t >= n_t) return;
    const agnpy_model_t& M = models[m];
    const double* g = M.tgt;
    int il = t % n_l;
    int iphi = (t / n_l) % n_phi;
    int ia = t / (n_l * n_phi);
    double wphi = (variant == AGNPY_TAU_PS_BEHIND_BLOB || variant == AGNPY_TAU_PS_BEHIND_BLOB_MU_S)
                      ? 1.0 : trapz_weight(phi, iphi, n_phi);
    double p, omc, W;
    int n_keep = n_phi;   // phi indices kept by the folding of the on-axis variants
    bool dead = false;    // this entry is represented by its mirror image / by phi index 0
    if (variant == AGNPY_TAU_PS_BEHIND_BLOB) {  // targets_absorption.py:114-118 (closed form)
        p = g[0];
        omc = 1.0 - M.mu_s;
        W = omc / g[2];
    } else if (variant == AGNPY_TAU_PS_BEHIND_BLOB_MU_S) {  // :155-171
        double r = g[2], mu_s = M.mu_s;
        double u = l_unit[il] * r;
        double ulo = l_unit[il > 0 ? il - 1 : 0] * r, uhi = l_unit[il < n_l - 1 ? il + 1 : n_l - 1] * r;
        double wl = n_l < 2 ? 0.0 : (uhi - ulo) * 0.5;
        double x = sqrt(r * r + u * u + 2.0 * u * r * mu_s);
        double mu_l = (r + u * mu_s) / x;
        omc = 1.0 - cos_psi_t(mu_s, mu_l, 0.0);
        p = g[0];
        W = wl * (omc / (x * x));
    } else if (variant == AGNPY_TAU_DT_MU_S) {  // :578-594, utils/geometry.py:58-118
        double R_dt = g[3], r = g[4], mu_s = M.mu_s;
        double u = l_unit[il] * r;
        double ulo = l_unit[il > 0 ? il - 1 : 0] * r, uhi = l_unit[il < n_l - 1 ? il + 1 : n_l - 1] * r;
        double wl = n_l < 2 ? 0.0 : (uhi - ulo) * 0.5;
        double phi_re = phi[iphi];
        double sin_s = sqrt(1.0 - mu_s * mu_s);
        double x = x_ring_mu_s(R_dt, r, phi_re, u, mu_s);
        double dx = u * sin_s - R_dt * cos(phi_re);
        double dy = 0.0 - R_dt * sin(phi_re);
        double dz = r + u * mu_s;
        omc = 1.0 - cos_psi_t(mu_s, dz / x, atan2(dy, dx));
        p = g[2];
        W = wphi * wl * (omc / (x * x));
    } else if (variant == AGNPY_TAU_BLR_MU_S) {  // :416-433, utils/geometry.py:121-183
        double R_line = g[3], r = g[4], mu_s = M.mu_s;
        const double* uu = path + (size_t)m * n_l;
        double u = uu[il];
        double wl = n_l < 2 ? 0.0 : (uu[il < n_l - 1 ? il + 1 : n_l - 1] - uu[il > 0 ? il - 1 : 0]) * 0.5;
        double mu_re = mu[ia], phi_re = phi[iphi];
        double sin_s = sqrt(1.0 - mu_s * mu_s), sin_re = sqrt(1.0 - mu_re * mu_re);
        double x = x_shell_mu_s(R_line, r, phi_re, mu_re, u, mu_s);
        double dx = u * sin_s - R_line * sin_re * cos(phi_re);
        double dy = -R_line * sin_re * sin(phi_re);
        double dz = r + u * mu_s - R_line * mu_re;
        omc = 1.0 - cos_psi_t(mu_s, dz / x, atan2(dy, dx));
        p = g[2];
        W = trapz_weight(mu, ia, n_a) * wphi * wl * (omc / (x * x));
    } else {
        // on-axis variants (dt, blr, ss_disk): phi enters only through cos phi in cos psi
        // (utils/geometry.py:5-12), so table entries that differ only by a mirror image of phi carry
        // the same p and 1 - cos psi: they are merged into one entry with the sum of the weights
        // and the mirror image is not stored.  With mu_s = 1 exactly, sqrt(1 - mu_s^2) = 0 and cos psi does not depend on phi
        // at all: one entry per (a, l) carries the whole phi integral.
        tau_onaxis_entry(variant, M, mu, n_a, phi, n_phi, l_unit, n_l, ia, iphi, il, p, omc, W);
        const int ip = n_phi - 1 - iphi;
        if (M.mu_s == 1.0 && n_phi >= 2) {
            n_keep = 1;  // phi index 0 carries the whole phi integral
            if (iphi == 0) {
                double wsum = 0.0;
                for (int i = 0; i < n_phi; ++i) wsum += trapz_weight(phi, i, n_phi);
                W = W / wphi * wsum;
            } else {
                dead = true;
            }
        } else if (all_mirror) {
            n_keep = (n_phi + 1) / 2;  // phi indices 0 .. ceil(n/2)-1, each with its mirror image
            if (iphi < n_keep) {
                if (ip != iphi) {
                    double p2, omc2, W2;
                    tau_onaxis_entry(variant, M, mu, n_a, phi, n_phi, l_unit, n_l, ia, ip, il, p2, omc2, W2);
                    W = W + W2;
                }
            } else {
                dead = true;
            }
        }
    }
    // compacted table: only the kept phi indices are stored, t' = (ia * n_keep + iphi) * n_l + il,
    // and the main kernel walks n_live[m] entries.  (Models of one call may differ: mu_s is theirs.)
    const int live_total = (n_keep == n_phi) ? n_t : n_t / n_phi * n_keep;
    if (t == 0) n_live[m] = live_total;
    if (dead) return;
    if (n_keep != n_phi) t = (ia * n_keep + iphi) * n_l + il;
    double* row = table + (size_t)m * 4 * n_t;
    double homc = omc / 2.0;
    row[t] = p;
    row[n_t + t] = homc;
    row[2 * (size_t)n_t + t] = 1.0 / (p * homc);
    row[3 * (size_t)n_t + t] = W;
}

// sigma(s) / (3/16 sigma_T), targets_absorption.py:34-42 in the reference's literal form (strict
// switch): beta = sqrt(1 - 1/s), (1 - beta^2) [(3 - beta^4) ln((1+beta)/(1-beta)) - 2 beta (2 - beta^2)]
__device__ __forceinline__ double sigma_shape_strict(double s) {
    double beta = sqrt(1.0 - 1.0 / s);
    double b2 = beta * beta;
    double t1 = (3.0 - b2 * b2) * clipped_log((1.0 + beta) / (1.0 - beta));
    double t2 = -2.0 * beta * (2.0 - b2);
    double v = (1.0 - b2) * (t1 + t2);
    return (s < 1.0) ? 0.0 : v;
}
// default: the same function with 1-beta^2 = 1/s and ln((1+beta)/(1-beta)) = ln((1+beta)^2 s), which
// do not lose the ~2e-16 s the reference's 1 - fl(beta)^2 loses
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

constexpr int kTauOnSynch = 100;  // internal variant id of the finish kernel
constexpr int kTauTN = 4;        // frequencies per thread
constexpr int kTauSlice = 4096;  // table entries per block

// grid = (n_slices, ceil(n_nu/TN), n_models); part[m][i][slice]
__global__ void __launch_bounds__(256) tau_main_kernel(
    const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ table, int n_t, const int* __restrict__ n_live,
    double* __restrict__ part, int blob_frame) {
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
        // targets_absorption.py:331: no Doppler factor for the external fields; :647 the blob
        // frame for the synchrotron field
        e1[q] = nu_to_eps(nu[i], M.z, (blob_frame & 1) ? M.delta_D : 1.0);
        ie1[q] = 1.0 / e1[q];
        acc[q] = 0.0;
    }
    const int live = n_live ? n_live[m] : n_t;  // live entries of this model (the stride stays n_t)
    if (slice * kTauSlice >= live) {             // a slice behind the compacted table: nothing to add
        if (threadIdx.x < kTauTN && i0 + threadIdx.x < n_nu)
            part[((size_t)m * n_nu + i0 + threadIdx.x) * S + slice] = 0.0;
        return;
    }
    AGB_CHECK(live >= 0 && live <= n_t);
    const int t_end = min(live, (slice + 1) * kTauSlice);
    for (int t = slice * kTauSlice + threadIdx.x; t < t_end; t += blockDim.x) {
        // the weight first: a warp of retired (folded phi axis) or weightless entries costs one
        // 8-byte load per entry instead of four
        const double W = t_W[t];
        if (__all_sync(__activemask(), W == 0.0)) continue;
        const double p = t_p[t], h = t_h[t], ic = t_ic[t];
#pragma unroll
        for (int q = 0; q < kTauTN; ++q) {
            double s = e1[q] * p * h;  // == eps_1 * eps * (1 - cos psi) / 2, same rounding
            // below the pair-production threshold sigma is exactly 0 (targets_absorption.py:41):
            // skip the sqrt/log when no lane of the warp is above it (NaN s keeps the slow path)
            if (__all_sync(__activemask(), s < 1.0 && W == W && fabs(W) <= kFmax)) continue;
            double v = (blob_frame & kStrictBit) ? sigma_shape_strict(s) : sigma_shape(s, ic * ie1[q]);
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
    if (variant == AGNPY_TAU_BLR || variant == AGNPY_TAU_BLR_MU_S)  // :350-352, 433-435
        pref = (g[0] * g[1]) / ((4.0 * kPi) * (4.0 * kPi) * g[2] * kMe * c3);
    else if (variant == AGNPY_TAU_DT || variant == AGNPY_TAU_DT_MU_S)  // :531, 596
        pref = (g[0] * g[1]) / (8.0 * (kPi * kPi) * g[2] * kMe * c3);
    else if (variant == AGNPY_TAU_PS_BEHIND_BLOB || variant == AGNPY_TAU_PS_BEHIND_BLOB_MU_S)  // :118, 172
        pref = g[1] / (4.0 * kPi * g[0] * kMe * c3);
    else if (variant == kTauOnSynch)  // :694
        pref = 2.0 * M.R_b;
    else {  // :263
        double R_g = kG * g[0] / (kC * kC);
        pref = 3.0 * g[1] / ((4.0 * kPi) * (4.0 * kPi) * g[2] * kMe * c3 * R_g);
    }
    tau[(size_t)m * n_nu + i] = pref * integral;
}

// ---- opacity on the blob's own synchrotron photons, targets_absorption.py:629-694 --------------
// frequencies of the synchrotron field, one grid per model: logspace between the delta-approximation
// peaks of gamma_min and gamma_max (synchrotron.py:27-32) with margins, blob frame -> observer
__global__ void __launch_bounds__(128) tau_synch_nu_kernel(
    const agnpy_model_t* __restrict__ models, int ne_kind, int n_s, double margin_low,
    double* __restrict__ nu_s_obs) {
    const int m = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_s) return;
    const agnpy_model_t& M = models[m];
    double gmin, gmax;
    if (ne_kind == AGNPY_NE_TABLE) { gmin = M.ne[1]; gmax = M.ne[2]; }
    else ne_bounds(ne_kind, M.ne, gmin, gmax);
    double peak = kE * M.B / (2.0 * kPi * kMe * kC);
    double lo = log10(peak * (gmin * gmin) * margin_low), hi = log10(peak * (gmax * gmax) * 1.0e2);
    double y = (n_s == 1) ? lo : (j == n_s - 1 ? hi : (double)j * ((hi - lo) / (double)(n_s - 1)) + lo);
    double nu_s = pow(10.0, y);
    nu_s_obs[(size_t)m * n_s + j] = nu_s * M.delta_D / (1.0 + M.z);
}

// table entries of the tau kernel from the photon density: s = eps eps_1 / 2 (:690)
__global__ void __launch_bounds__(128) tau_synch_table_kernel(
    const double* __restrict__ n_synch, const double* __restrict__ eps, int n_s,
    double* __restrict__ table) {
    const int m = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_s) return;
    const double* e = eps + (size_t)m * n_s;
    double* row = table + (size_t)m * 4 * n_s;
    row[j] = e[j];
    row[n_s + j] = 0.5;
    row[2 * (size_t)n_s + j] = 1.0 / (e[j] * 0.5);
    row[3 * (size_t)n_s + j] = trapz_weight(e, j, n_s) * n_synch[(size_t)m * n_s + j];
}

size_t tau_synch_workspace_bytes(int n_models, int n_nu, int n_gamma, int n_s) {
    size_t S = ((size_t)n_s + kTauSlice - 1) / kTauSlice;
    return (3 * (size_t)n_gamma + 7 * (size_t)n_s + (size_t)n_nu * S) * sizeof(double) * (size_t)n_models;
}

int run_tau_synch(const agnpy_model_t* models, int n_models, int ne_kind, const double* nu, int n_nu,
                  const double* gamma, int n_gamma, const double* ne_table, const double* ssa_table,
                  int n_s, double margin_low, int integ, double* tau, double* ws, cudaStream_t stream) {
    const size_t M = (size_t)n_models;
    int S = (n_s + kTauSlice - 1) / kTauSlice;
    double* tab = ws;
    double* nu_s = tab + M * 3 * n_gamma;
    double* nsyn = nu_s + M * n_s;
    double* eps = nsyn + M * n_s;
    double* table = eps + M * n_s;
    double* part = table + M * 4 * n_s;
    launch_prep_gamma(models, n_models, ne_kind, gamma, n_gamma, ne_table, ssa_table, 0, tab, stream);
    {
        dim3 grid((n_s + 127) / 128, n_models);
        tau_synch_nu_kernel<<<grid, 128, 0, stream>>>(models, ne_kind, n_s, margin_low, nu_s);
        note_launch();
    }
    // self-absorbed synchrotron SED at those frequencies -> photon density (:668-687)
    if (int rc = launch_synch(models, n_models, nu_s, n_s, tab, n_gamma, 1, integ, 2, nsyn, nullptr, eps,
                              stream, n_s))
        return rc;
    {
        dim3 grid((n_s + 127) / 128, n_models);
        tau_synch_table_kernel<<<grid, 128, 0, stream>>>(nsyn, eps, n_s, table);
        note_launch();
    }
    {
        dim3 grid(S, (n_nu + kTauTN - 1) / kTauTN, n_models);
        timing_begin(stream);
        tau_main_kernel<<<grid, 256, 0, stream>>>(models, nu, n_nu, table, n_s, nullptr, part,
                                                   1 | (strict_reference() ? kStrictBit : 0));
        timing_end(stream);
        note_launch();
    }
    {
        dim3 grid((n_nu + 127) / 128, n_models);
        tau_finish_kernel<<<grid, 128, 0, stream>>>(kTauOnSynch, models, n_nu, part, S, tau);
        note_launch();
    }
    return check_last("tau_synch");
}

static int tau_n_t(int variant, int n_a, int n_phi, int n_l) {
    if (variant == AGNPY_TAU_PS_BEHIND_BLOB) return 1;
    if (variant == AGNPY_TAU_PS_BEHIND_BLOB_MU_S) return n_l;
    if (variant == AGNPY_TAU_DT || variant == AGNPY_TAU_DT_MU_S) return n_phi * n_l;
    return n_a * n_phi * n_l;
}

size_t tau_workspace_bytes(int variant, int n_models, int n_nu, int n_a, int n_phi, int n_l) {
    size_t n_t = tau_n_t(variant, n_a, n_phi, n_l);
    size_t S = (n_t + kTauSlice - 1) / kTauSlice;
    return (4 * n_t + (size_t)n_nu * S + (size_t)n_l + 1) * sizeof(double) * (size_t)n_models;
}

int run_tau(int variant, const agnpy_model_t* models, int n_models, const double* nu, int n_nu,
            const double* mu, int n_a, const double* phi, int n_phi, const double* l_unit, int n_l,
            double* tau, double* ws, cudaStream_t stream) {
    int n_t = tau_n_t(variant, n_a, n_phi, n_l);
    int S = (n_t + kTauSlice - 1) / kTauSlice;
    double* table = ws;
    double* part = table + (size_t)n_models * 4 * n_t;
    double* path = part + (size_t)n_models * n_nu * S;
    int* n_live = reinterpret_cast<int*>(path + (size_t)n_models * n_l);  // live table entries per model
    int nl_eff = n_l, nphi_eff = n_phi;
    if (variant == AGNPY_TAU_PS_BEHIND_BLOB) { nl_eff = 1; nphi_eff = 1; }
    if (variant == AGNPY_TAU_PS_BEHIND_BLOB_MU_S) nphi_eff = 1;
    if (variant == AGNPY_TAU_BLR_MU_S) {
        tau_blr_path_kernel<<<n_models, 128, (size_t)n_l * sizeof(double), stream>>>(models, l_unit, n_l, path);
        note_launch();
    }
    {
        dim3 grid((n_t + 127) / 128, n_models);
        tau_table_kernel<<<grid, 128, 0, stream>>>(variant, models, mu, n_a, phi, nphi_eff, l_unit,
                                                   nl_eff, n_t, path, table, n_live);
        note_launch();
    }
    {
        dim3 grid(S, (n_nu + kTauTN - 1) / kTauTN, n_models);
        timing_begin(stream);
        tau_main_kernel<<<grid, 256, 0, stream>>>(models, nu, n_nu, table, n_t, n_live, part,
                                                   strict_reference() ? kStrictBit : 0);
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
