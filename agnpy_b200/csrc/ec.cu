// K3: external-Compton SED kernels (iso_mono, ps_behind_jet, ss_disk, blr, dt).
//
// Reference arithmetic replaced:
//   agnpy/compton/external_compton.py:122-142, 221-240, 333-373, 463-490, 578-600
//   agnpy/compton/kernels.py:36-68 (get_gamma_min, compton_kernel)
//   agnpy/utils/geometry.py:5-55 (cos_psi, x_re_shell, mu_star_shell, x_re_ring)
//   agnpy/targets/targets.py:289-323 (SSDisk statics)
//
// The reference broadcasts a full (gamma, mu, phi, nu) array (8 GB at the default grids).
// Here a block owns one output frequency and a slice of (mu,phi) "pairs"; each thread keeps
// J pairs in registers and walks the gamma axis once, reading three per-(gamma,nu) values
// that the block tabulated in shared memory.  With
//     y = 1 - eps_s/gamma,  A = y + 1/y,  G = eps_s/(gamma^2 y),  b = 1/(eps (1 - cos psi))
// the kernel of kernels.py:61-65 is   K = y + 1/y - 2t + t^2 = A + t (t - 2),  t = G b,
// i.e. 4 FP64 instructions per integrand point (DMUL, DADD, DFMA, DFMA-accumulate) and no
// division: the two divisions of the nominal count are hoisted to the (gamma,nu) table and
// to the per-pair table.  The kinematic mask gamma >= gamma_min(nu, pair) is evaluated with
// the reference's own operation order (kernels.py:36-40) and turned into a start index of
// the gamma loop by a binary search of the (ascending) gamma grid.
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace agb {

// ---- geometry, in the reference's operation order -------------------------------------------
__device__ __forceinline__ double cos_psi(double mu_s, double mu, double phi) {  // geometry.py:5-12
    double t1 = mu * mu_s;
    double t2 = sqrt(1.0 - mu * mu) * sqrt(1.0 - mu_s * mu_s);
    double t3 = cos(phi);
    return t1 + t2 * t3;
}
__device__ __forceinline__ double x_re_shell(double mu, double R, double r) {  // :15-28
    return sqrt(R * R + r * r - 2.0 * r * R * mu);
}
__device__ __forceinline__ double mu_star_shell(double mu, double R, double r) {  // :31-50
    double ratio = R / x_re_shell(mu, R, r);
    double addend = ratio * ratio * (1.0 - mu * mu);
    double ms = sqrt(1.0 - addend);
    double sign = (((r > mu * R) ? 1.0 : 0.0) - 0.5) * 2.0;
    return ms * sign;
}
// numpy.linspace(start, stop, n)[i]: arange*step + start, last point forced to stop
__device__ __forceinline__ double linspace_at(double start, double stop, int n, int i) {
    if (n == 1) return start;
    if (i == n - 1) return stop;
    double step = (stop - start) / (double)(n - 1);
    return (double)i * step + start;
}
// targets.py:307-315 epsilon(R_tilde)
__device__ __forceinline__ double disk_epsilon(double L_disk, double M_BH, double eta, double R_tilde) {
    double M_8 = M_BH / (1e8 * kMsun);
    double L_Edd = 1.26 * 1e46 * M_8;
    double l_Edd = L_disk / L_Edd;
    double xi = pow(l_Edd / (M_8 * eta), 1.0 / 4.0);
    return 2.7 * 1e-4 * xi * pow(R_tilde, -3.0 / 4.0);
}

// ---------------------------------------------------------------------------------------------
// per-model pair table:  ang[m][0][j] = eps_j      target energy seen by pair j
//                        ang[m][1][j] = 1 - cos psi_j
//                        ang[m][2][j] = b_j = 1/(eps_j (1 - cos psi_j))
//                        ang[m][3][j] = W_j        np.trapz weights x geometric factor
//                        ang[m][4][j] = c_j = 2 eps_j (1 - cos psi_j)   (mask threshold on G)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ec_pairs_role(
    int variant, const agnpy_model_t& M, const double* __restrict__ mu, int n_mu,
    const double* __restrict__ phi, int n_phi, int n_pairs, double* __restrict__ row, int j) {
    if (j >= n_pairs) return;
    const double* t = M.tgt;
    double eps, cpsi, W;
    if (variant == AGNPY_EC_PS_BEHIND_JET) {  // external_compton.py:226: mu = 1, phi = 0
        eps = t[0];
        cpsi = cos_psi(M.mu_s, 1.0, 0.0);
        W = 1.0;
    } else if (variant == AGNPY_EC_DT) {  // :583-589
        double x_re = sqrt(t[3] * t[3] + t[4] * t[4]);
        double mu_dt = t[4] / x_re;
        eps = t[2];
        cpsi = cos_psi(M.mu_s, mu_dt, phi[j]);
        W = trapz_weight(phi, j, n_phi);
    } else {
        int imu = j / n_phi, iphi = j - imu * n_phi;
        double wphi = trapz_weight(phi, iphi, n_phi);
        if (variant == AGNPY_EC_ISO_MONO) {  // :127-132
            eps = t[0];
            cpsi = cos_psi(M.mu_s, mu[imu], phi[iphi]);
            W = trapz_weight(mu, imu, n_mu) * wphi;
        } else if (variant == AGNPY_EC_BLR) {  // :469-475
            double R_line = t[3], r = t[4];
            double x = x_re_shell(mu[imu], R_line, r);
            double mstar = mu_star_shell(mu[imu], R_line, r);
            eps = t[2];
            cpsi = cos_psi(M.mu_s, mstar, phi[iphi]);
            W = trapz_weight(mu, imu, n_mu) * wphi * (1.0 / (x * x));
        } else {  // AGNPY_EC_SS_DISK :335-360, targets.py:289-323
            double M_BH = t[0], L_disk = t[1], eta = t[2];
            double R_g = kG * M_BH / (kC * kC);
            double r_t = t[5] / R_g, Rin_t = t[3] / R_g, Rout_t = t[4] / R_g;
            double mu_min = 1.0 / sqrt(1.0 + (Rout_t / r_t) * (Rout_t / r_t));
            double mu_max = 1.0 / sqrt(1.0 + (Rin_t / r_t) * (Rin_t / r_t));
            double mu_i = linspace_at(mu_min, mu_max, n_mu, imu);
            double mu_lo = linspace_at(mu_min, mu_max, n_mu, imu > 0 ? imu - 1 : 0);
            double mu_hi = linspace_at(mu_min, mu_max, n_mu, imu < n_mu - 1 ? imu + 1 : n_mu - 1);
            double wmu = n_mu < 2 ? 0.0 : (mu_hi - mu_lo) * 0.5;
            double mm2 = pow(mu_i, -2.0) - 1.0;
            double R_tilde = r_t * sqrt(mm2);
            eps = disk_epsilon(L_disk, M_BH, eta, R_tilde);
            double phi_disk = 1.0 - sqrt(Rin_t / R_tilde);
            cpsi = cos_psi(M.mu_s, mu_i, phi[iphi]);
            W = wmu * wphi * (phi_disk / (eps * eps) / mu_i / pow(mm2, 3.0 / 2.0));
        }
    }
    double omc = 1.0 - cpsi;
    row[j] = eps;
    row[n_pairs + j] = omc;
    row[2 * n_pairs + j] = 1.0 / (eps * omc);
    row[3 * n_pairs + j] = W;
    row[4 * n_pairs + j] = 2.0 * (eps * omc);
}

// ---------------------------------------------------------------------------------------------
// gamma tables of one model (nu independent), K5 fused in: N_e = V_b n_e(snap(gamma/delta_D))
// (external_compton.py:125-126; the integration grid itself stays unsnapped)
//         gt[m][0][k] = gamma_k
//         gt[m][1][k] = trapz: w_k N_e/gamma^2   loglog: f_k = N_e/gamma^2
//         gt[m][2][k] = ln f_k                    gt[m][3][k] = 1/(ln g_k - ln g_{k+1})
//         gt[m][4][k] = ln(g_{k+1}/g_k)           hull[m] = {klo, khi}: non-zero range of n_e
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ec_gamma_role(
    const agnpy_model_t& M, int ne_kind, const double* __restrict__ gamma, int n_gamma,
    const double* __restrict__ ne_table, int integ, double* __restrict__ row,
    int* __restrict__ hull2, int* s_hull) {
    if (threadIdx.x == 0) { s_hull[0] = n_gamma; s_hull[1] = -1; }
    __syncthreads();
    double gmin, gmax;
    if (ne_kind == AGNPY_NE_TABLE) { gmin = M.ne[1]; gmax = M.ne[2]; }
    else ne_bounds(ne_kind, M.ne, gmin, gmax);
    const double V_b = 4.0 / 3.0 * kPi * (M.R_b * M.R_b * M.R_b);
    const double d = M.delta_D, g0 = gamma[0] / d, gN = gamma[n_gamma - 1] / d;
    for (int k = threadIdx.x; k < n_gamma; k += blockDim.x) {
        double g = gamma[k];
        double ne = (ne_kind == AGNPY_NE_TABLE)
                        ? ne_table[k]
                        : ne_eval(ne_kind, M.ne, snapped(g / d, k, n_gamma, g0, gN, gmin, gmax));
        double f = (V_b * ne) / (g * g);
        row[k] = g;
        if (integ == AGNPY_INTEG_TRAPZ) {
            row[n_gamma + k] = f * trapz_weight(gamma, k, n_gamma);
        } else {
            row[n_gamma + k] = f;
            row[2 * n_gamma + k] = clipped_log(f);
            if (k < n_gamma - 1) {
                double gu = gamma[k + 1];
                row[3 * n_gamma + k] = 1.0 / (clipped_log(g) - clipped_log(gu));
                row[4 * n_gamma + k] = clipped_log(gu / g);
            }
        }
        if (f != 0.0) { atomicMin(&s_hull[0], k); atomicMax(&s_hull[1], k); }
    }
    __syncthreads();
    if (threadIdx.x == 0) { hull2[0] = s_hull[0]; hull2[1] = s_hull[1]; }
}

// per-(gamma, nu) table of one model:  nt[m][i][k] = {A, G},
//   y = 1 - eps_s/gamma,  A = y + 1/y,  G = eps_s/(gamma^2 y)  (+inf where gamma <= eps_s)
__device__ __forceinline__ void ec_nu_role(const agnpy_model_t& M, const double* __restrict__ nu,
                                           const double* __restrict__ gamma, int n_gamma,
                                           double2* __restrict__ nt_row, int i, int k) {
    if (k >= n_gamma) return;
    // EC uses nu_to_epsilon_prime(nu, z) without the Doppler factor (external_compton.py:123)
    const double eps_s = nu_to_eps(nu[i], M.z, 1.0);
    const double g = gamma[k];
    const double y = 1.0 - eps_s / g;
    double2 v;
    v.x = y + 1.0 / y;
    v.y = (y > 0.0) ? eps_s / ((g * g) * y) : INFINITY;
    nt_row[(size_t)i * n_gamma + k] = v;
}

// One launch builds all three tables; the block role follows from blockIdx.x:
//   [0, 1)                     gamma tables + hull
//   [1, 1 + nB)                pair table, 128 pairs per block
//   [1 + nB, 1 + nB + nC*n_nu) {A, G} table, 128 gammas of one frequency per block
__global__ void __launch_bounds__(128) ec_tables_kernel(
    int variant, const agnpy_model_t* __restrict__ models, int ne_kind,
    const double* __restrict__ nu, int n_nu, const double* __restrict__ gamma, int n_gamma,
    const double* __restrict__ mu, int n_mu, const double* __restrict__ phi, int n_phi,
    const double* __restrict__ ne_table, int integ, int n_pairs, double* __restrict__ gt,
    int* __restrict__ hull, double* __restrict__ ang, double2* __restrict__ nt) {
    __shared__ int s_hull[2];
    const int m = blockIdx.y;
    const agnpy_model_t& M = models[m];
    const int nB = (n_pairs + 127) / 128, nC = (n_gamma + 127) / 128;
    int bx = blockIdx.x;
    if (bx == 0) {
        ec_gamma_role(M, ne_kind, gamma, n_gamma, ne_table ? ne_table + (size_t)m * n_gamma : nullptr,
                      integ, gt + (size_t)m * 5 * n_gamma,
                      hull + 2 * m, s_hull);
    } else if (bx < 1 + nB) {
        ec_pairs_role(variant, M, mu, n_mu, phi, n_phi, n_pairs, ang + (size_t)m * 5 * n_pairs,
                      (bx - 1) * 128 + threadIdx.x);
    } else {
        bx -= 1 + nB;
        ec_nu_role(M, nu, gamma, n_gamma, nt + (size_t)m * n_nu * n_gamma, bx / nC,
                   (bx % nC) * 128 + threadIdx.x);
    }
}

// ---------------------------------------------------------------------------------------------
// main kernel.  grid = (Sg <= S slice groups, n_nu, n_models), block = T threads, J (mu,phi)
// pairs per thread.  The block copies the per-gamma values of its (nu, model) -- produced by
// the two table kernels above -- into shared memory once, then every warp walks its slices of
// pairs independently (no barrier after the copy): per slice it locates the first unmasked
// gamma index of each pair, runs the gamma loop from shared-memory broadcasts, reduces with
// shuffles and stores its own partial:
// part[m][i][slice][warp] = sum over that warp's pairs of W_j * integral_gamma.
//
// Kinematic mask.  In exact arithmetic gamma >= gamma_min(eps_s, eps, psi)  <=>  t = G b <= 2, so
// the first unmasked index is the first k with G_k <= c, c = 2 eps (1 - cos psi): a binary search
// of the descending G table, no division or square root per (nu, pair).  The J searches of a
// thread run interleaved and branch-free (fixed trip count) so their shared-memory latencies
// overlap.  numpy decides the mask on fl(gamma_min); the two can only disagree when G_k is
// within rounding of c, so whenever a neighbour of the found index lies within a 1e-6 relative
// band of c (or y_k < 1e-8, where G_k itself is ill-conditioned) the index is recomputed from
// gamma_min in the reference's own operation order (kernels.py:36-40).  The band is ~1e9 ulp
// wide and is hit by ~1e-5 of the lookups.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ int first_unmasked_exact(const double* __restrict__ gm, int n_gamma,
                                                 double eps, double eps_s, double omc) {
    double root = sqrt(1.0 + 2.0 / (eps * eps_s * omc));
    double gmin = eps_s / 2.0 * (1.0 + root);
    int lo = 0, hi = n_gamma;  // first k with gamma_k >= gmin (NaN -> n_gamma)
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (gm[mid] >= gmin) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <int INTEG, int J>
__global__ void __launch_bounds__(256) ec_main_kernel(
    const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ gt, const int* __restrict__ hull, int n_gamma, int search_iters,
    const double2* __restrict__ nt, const double* __restrict__ ang, int n_pairs, int S,
    double* __restrict__ part) {
    extern __shared__ double sh[];
    const int i = blockIdx.y, m = blockIdx.z;
    const int T = blockDim.x, tid = threadIdx.x, W_per_block = T >> 5;
    const double* g_gm = gt + (size_t)m * 5 * n_gamma;
    const double* g_f = g_gm + n_gamma;
    const double2* AG = nt + ((size_t)m * n_nu + i) * n_gamma;
    double* s_f = sh;
    double* s_A = s_f + n_gamma;
    double* s_G = s_A + n_gamma;
    double* s_gm = s_G + n_gamma;    // loglog only
    double* s_lf = s_gm + n_gamma;   // loglog only
    double* s_idl = s_lf + n_gamma;  // loglog only
    double* s_lr = s_idl + n_gamma;  // loglog only
    for (int k = tid; k < n_gamma; k += T) {
        double2 ag = AG[k];
        s_f[k] = g_f[k];
        s_A[k] = ag.x;
        s_G[k] = ag.y;
        if (INTEG == AGNPY_INTEG_LOGLOG) {
            s_gm[k] = g_gm[k];
            s_lf[k] = g_gm[2 * n_gamma + k];
            s_idl[k] = g_gm[3 * n_gamma + k];
            s_lr[k] = g_gm[4 * n_gamma + k];
        }
    }
    const int klo = hull[2 * m], khi = hull[2 * m + 1];  // non-zero range of n_e
    __syncthreads();

    const double* a_eps = ang + (size_t)m * 5 * n_pairs;
    const double* a_omc = a_eps + n_pairs;
    const double* a_b = a_omc + n_pairs;
    const double* a_W = a_b + n_pairs;
    const double* a_c = a_W + n_pairs;
    for (int slice = blockIdx.x; slice < S; slice += gridDim.x) {
        double b[J], W[J], c[J], acc[J];
        int lo[J], hi[J];
#pragma unroll
        for (int jj = 0; jj < J; ++jj) {
            int j = slice * (T * J) + jj * T + tid;
            bool valid = j < n_pairs;
            b[jj] = valid ? a_b[j] : 0.0;
            W[jj] = valid ? a_W[j] : 0.0;
            c[jj] = valid ? a_c[j] : -1.0;  // c <= 0: never unmasked
            acc[jj] = 0.0;
            lo[jj] = 0;
            hi[jj] = n_gamma;
        }
        // J interleaved branch-free binary searches: first k with G_k <= c
        for (int it = 0; it < search_iters; ++it) {
#pragma unroll
            for (int jj = 0; jj < J; ++jj) {
                int mid = (lo[jj] + hi[jj]) >> 1;
                bool open = lo[jj] < hi[jj];
                bool le = s_G[min(mid, n_gamma - 1)] <= c[jj];
                hi[jj] = (open && le) ? mid : hi[jj];
                lo[jj] = (open && !le) ? mid + 1 : lo[jj];
            }
        }
        int kmin = INT_MAX, kmax = 0;
#pragma unroll
        for (int jj = 0; jj < J; ++jj) {
            int k = lo[jj];
            const double cj = c[jj], band = 1e-6 * cj;
            int ka = min(k, n_gamma - 1), kb = max(k - 1, 0);
            double Ga = s_G[ka], Aa = s_A[ka], Gb = s_G[kb], Ab = s_A[kb];
            bool doubt = (k < n_gamma && (fabs(Ga - cj) <= band || Aa > 1e8)) ||
                         (k > 0 && (fabs(Gb - cj) <= band || (Ab > 1e8 && Gb < kFmax)));
            doubt = (doubt || !(cj < kFmax)) && cj > 0.0;
            if (doubt) {
                int j = slice * (T * J) + jj * T + tid;
                k = first_unmasked_exact(g_gm, n_gamma, a_eps[j], nu_to_eps(nu[i], models[m].z, 1.0),
                                         a_omc[j]);
            }
            lo[jj] = k;
            if (cj > 0.0) { kmin = min(kmin, k); kmax = max(kmax, k); }
        }
        double thread_sum = 0.0;
        if (khi >= 0) {
            if (INTEG == AGNPY_INTEG_TRAPZ) {
                kmin = __reduce_min_sync(0xffffffffu, kmin);
                kmax = __reduce_max_sync(0xffffffffu, kmax);
                int kbeg = max(kmin, klo);
                int kmid = min(max(kmax, kbeg), khi + 1);
                // phase 1: some lanes still below their gamma_min -> predicated accumulate
                for (int k = kbeg; k < kmid; ++k) {
                    double wf = s_f[k], A = s_A[k], G = s_G[k];
#pragma unroll
                    for (int jj = 0; jj < J; ++jj) {
                        double t = G * b[jj];
                        double K = fma(t, t - 2.0, A);
                        if (k >= lo[jj]) acc[jj] = fma(wf, K, acc[jj]);
                    }
                }
                // phase 2: every valid lane is above its gamma_min -> 4 FP64 instructions / point
                for (int k = kmid; k <= khi; ++k) {
                    double wf = s_f[k], A = s_A[k], G = s_G[k];
#pragma unroll
                    for (int jj = 0; jj < J; ++jj) {
                        double t = G * b[jj];
                        double K = fma(t, t - 2.0, A);
                        acc[jj] = fma(wf, K, acc[jj]);
                    }
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    int ks = max(lo[jj], klo);
                    double y_prev = 0.0, ly_prev = 0.0, a = 0.0;
                    for (int k = ks; k <= khi; ++k) {
                        double f = s_f[k];
                        double y = 0.0, ly = 0.0;
                        if (f != 0.0) {
                            double t = s_G[k] * b[jj];
                            double K = fma(t, t - 2.0, s_A[k]);
                            y = f * K;
                            ly = clipped_log(y);
                        }
                        if (k > ks)
                            a += loglog_interval_pre(s_gm[k - 1], s_gm[k], s_idl[k - 1], s_lr[k - 1],
                                                     y_prev, y, ly_prev, ly);
                        y_prev = y;
                        ly_prev = ly;
                    }
                    acc[jj] = a;
                }
            }
#pragma unroll
            for (int jj = 0; jj < J; ++jj)
                if (c[jj] > 0.0) thread_sum = fma(W[jj], acc[jj], thread_sum);
        }
        thread_sum = warp_sum(thread_sum);
        if ((tid & 31) == 0)
            part[(((size_t)m * n_nu + i) * S + slice) * W_per_block + (tid >> 5)] = thread_sum;
    }
}

// ---------------------------------------------------------------------------------------------
// Opt-in FP32 integrand (AGNPY_INTEG_TRAPZ_FP32): the same kernel with the point loop in single
// precision -- K = fmaf(t, t-2, A), acc = fmaf(wf, K, acc) -- while everything that decides
// *which* points contribute stays in FP64: the mask index search runs on the double G table with
// the same guard band / exact fallback, the pair weights, the (mu,phi) sum and the prefactor
// are double.  wf is scaled by its per-model maximum so it fits the float range.  Accuracy is
// that of ~100 float FMAs per gamma integral (measured rtol ~1e-6, tests/test_gpu_prefix.py);
// it is never selected implicitly.
// ---------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(256) ec_main_fp32_kernel(
    const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ gt, const int* __restrict__ hull, int n_gamma, int search_iters,
    const double2* __restrict__ nt, const double* __restrict__ ang, int n_pairs, int S,
    double* __restrict__ part) {
    extern __shared__ double sh[];
    __shared__ double red[32];
    __shared__ double s_scale;
    const int i = blockIdx.y, m = blockIdx.z;
    const int T = blockDim.x, tid = threadIdx.x, W_per_block = T >> 5;
    const double* g_gm = gt + (size_t)m * 5 * n_gamma;
    const double* g_f = g_gm + n_gamma;
    const double2* AG = nt + ((size_t)m * n_nu + i) * n_gamma;
    double* s_G = sh;                        // double: mask search
    double* s_A = s_G + n_gamma;             // double: guard band
    float* s_ff = reinterpret_cast<float*>(s_A + n_gamma);
    float* s_Af = s_ff + n_gamma;
    float* s_Gf = s_Af + n_gamma;
    double mx = 0.0;
    for (int k = tid; k < n_gamma; k += T) mx = fmax(mx, fabs(g_f[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int w = 0; w < W_per_block; ++w) v = fmax(v, red[w]);
        s_scale = (v > 0.0 && v < kFmax) ? v : 1.0;
    }
    __syncthreads();
    const double scale = s_scale;
    for (int k = tid; k < n_gamma; k += T) {
        double2 ag = AG[k];
        s_G[k] = ag.y;
        s_A[k] = ag.x;
        s_ff[k] = (float)(g_f[k] / scale);
        s_Af[k] = (float)ag.x;
        s_Gf[k] = (float)ag.y;
    }
    const int klo = hull[2 * m], khi = hull[2 * m + 1];
    __syncthreads();

    const double* a_eps = ang + (size_t)m * 5 * n_pairs;
    const double* a_omc = a_eps + n_pairs;
    const double* a_b = a_omc + n_pairs;
    const double* a_W = a_b + n_pairs;
    const double* a_c = a_W + n_pairs;
    for (int slice = blockIdx.x; slice < S; slice += gridDim.x) {
        float bf[J], acc[J];
        double W[J], c[J];
        int lo[J], hi[J];
#pragma unroll
        for (int jj = 0; jj < J; ++jj) {
            int j = slice * (T * J) + jj * T + tid;
            bool valid = j < n_pairs;
            bf[jj] = valid ? (float)a_b[j] : 0.f;
            W[jj] = valid ? a_W[j] : 0.0;
            c[jj] = valid ? a_c[j] : -1.0;
            acc[jj] = 0.f;
            lo[jj] = 0;
            hi[jj] = n_gamma;
        }
        for (int it = 0; it < search_iters; ++it) {
#pragma unroll
            for (int jj = 0; jj < J; ++jj) {
                int mid = (lo[jj] + hi[jj]) >> 1;
                bool open = lo[jj] < hi[jj];
                bool le = s_G[min(mid, n_gamma - 1)] <= c[jj];
                hi[jj] = (open && le) ? mid : hi[jj];
                lo[jj] = (open && !le) ? mid + 1 : lo[jj];
            }
        }
        int kmin = INT_MAX, kmax = 0;
#pragma unroll
        for (int jj = 0; jj < J; ++jj) {
            int k = lo[jj];
            const double cj = c[jj], band = 1e-6 * cj;
            int ka = min(k, n_gamma - 1), kb = max(k - 1, 0);
            double Ga = s_G[ka], Aa = s_A[ka], Gb = s_G[kb], Ab = s_A[kb];
            bool doubt = (k < n_gamma && (fabs(Ga - cj) <= band || Aa > 1e8)) ||
                         (k > 0 && (fabs(Gb - cj) <= band || (Ab > 1e8 && Gb < kFmax)));
            doubt = (doubt || !(cj < kFmax)) && cj > 0.0;
            if (doubt) {
                int j = slice * (T * J) + jj * T + tid;
                k = first_unmasked_exact(g_gm, n_gamma, a_eps[j], nu_to_eps(nu[i], models[m].z, 1.0),
                                         a_omc[j]);
            }
            lo[jj] = k;
            if (cj > 0.0) { kmin = min(kmin, k); kmax = max(kmax, k); }
        }
        double thread_sum = 0.0;
        if (khi >= 0) {
            kmin = __reduce_min_sync(0xffffffffu, kmin);
            kmax = __reduce_max_sync(0xffffffffu, kmax);
            int kbeg = max(kmin, klo);
            int kmid = min(max(kmax, kbeg), khi + 1);
            for (int k = kbeg; k < kmid; ++k) {
                float wf = s_ff[k], A = s_Af[k], G = s_Gf[k];
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    float t = G * bf[jj];
                    float K = fmaf(t, t - 2.f, A);
                    if (k >= lo[jj]) acc[jj] = fmaf(wf, K, acc[jj]);
                }
            }
            for (int k = kmid; k <= khi; ++k) {
                float wf = s_ff[k], A = s_Af[k], G = s_Gf[k];
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    float t = G * bf[jj];
                    float K = fmaf(t, t - 2.f, A);
                    acc[jj] = fmaf(wf, K, acc[jj]);
                }
            }
#pragma unroll
            for (int jj = 0; jj < J; ++jj)
                if (c[jj] > 0.0) thread_sum = fma(W[jj], (double)acc[jj], thread_sum);
        }
        thread_sum = warp_sum(thread_sum);
        if ((tid & 31) == 0)
            part[(((size_t)m * n_nu + i) * S + slice) * W_per_block + (tid >> 5)] = thread_sum * scale;
    }
}

// prefactor of each evaluate_sed_flux_* (external_compton.py:133-142, 229-240, 361-373, 476-490,
// 590-600), in the reference's operation order
__device__ __forceinline__ double ec_prefactor(int variant, const agnpy_model_t& M, double nu_i) {
    const double* t = M.tgt;
    double eps_s = nu_to_eps(nu_i, M.z, 1.0);
    double d3 = M.delta_D * M.delta_D * M.delta_D;
    double es2 = eps_s * eps_s, dL2 = M.d_L * M.d_L;
    double pi2 = kPi * kPi, pi3 = kPi * kPi * kPi;
    double num, den;
    switch (variant) {
        case AGNPY_EC_ISO_MONO:
            num = 3.0 * kC * kSigmaT * t[1] * es2 * d3;
            den = 128.0 * pi2 * dL2 * (t[0] * t[0]);
            break;
        case AGNPY_EC_PS_BEHIND_JET:
            num = 3.0 * kSigmaT * t[1] * es2 * d3;
            den = 128.0 * pi2 * dL2 * (t[2] * t[2]) * (t[0] * t[0]);
            break;
        case AGNPY_EC_SS_DISK: {
            double m_dot = t[1] / (t[2] * (kC * kC));
            num = 9.0 * kSigmaT * kG * t[0] * m_dot * es2 * d3;
            den = 512.0 * pi3 * dL2 * (t[5] * t[5] * t[5]);
            break;
        }
        case AGNPY_EC_BLR:
            num = 3.0 * kSigmaT * t[1] * t[0] * es2 * d3;
            den = 512.0 * pi3 * dL2 * (t[2] * t[2]);
            break;
        default: {  // AGNPY_EC_DT
            double x_re = sqrt(t[3] * t[3] + t[4] * t[4]);
            num = 3.0 * kSigmaT * t[1] * t[0] * es2 * d3;
            den = 256.0 * pi3 * dL2 * (x_re * x_re) * (t[2] * t[2]);
            break;
        }
    }
    return num / den;
}

// sum the partials in a fixed order and apply the prefactor
__global__ void __launch_bounds__(128) ec_finish_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ part, int S, double* __restrict__ sed) {
    int m = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nu) return;
    const double* p = part + ((size_t)m * n_nu + i) * S;
    double integral = 0.0;
    for (int s = 0; s < S; ++s) integral += p[s];
    sed[(size_t)m * n_nu + i] = ec_prefactor(variant, models[m], nu[i]) * integral;
}

// ---------------------------------------------------------------------------------------------
// Opt-in fast path for the np.trapz integrator (AGNPY_INTEG_TRAPZ_PREFIX).  With np.trapz the
// gamma integral is a fixed weighted sum and the kernel is a quadratic in b:
//     sum_{k>=k0} wf_k K_k = SP[k0] + b^2 SQ[k0] - 2 b SR[k0],
//     P_k = wf_k A_k,  Q_k = wf_k G_k^2,  R_k = wf_k G_k,   S*[k] = suffix sums over k..khi,
// so a (nu, model) needs one suffix scan of three length-n_gamma arrays and, per (mu,phi) pair,
// the mask index k0 plus four FP64 instructions: O(n_gamma + n_pairs log n_gamma) instead of
// O(n_gamma n_pairs).  Same integral, same grid, same masks (k0 is found exactly as in the direct
// kernel); only the association order of the gamma sum differs (~1e-15 relative).
// grid = (n_nu, n_models), 256 threads; the prefactor is applied in-kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ec_prefix_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ gt, const int* __restrict__ hull, int n_gamma, int search_iters,
    const double2* __restrict__ nt, const double* __restrict__ ang, int n_pairs,
    double* __restrict__ sed) {
    extern __shared__ double sh[];
    __shared__ double red[32];
    const int i = blockIdx.x, m = blockIdx.y;
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* g_gm = gt + (size_t)m * 5 * n_gamma;
    const double* g_f = g_gm + n_gamma;
    const double2* AG = nt + ((size_t)m * n_nu + i) * n_gamma;
    const int klo = hull[2 * m], khi = hull[2 * m + 1];
    double* s_G = sh;                  // [n]
    double* s_A = s_G + n_gamma;       // [n]
    double* s_S = s_A + n_gamma;       // [3][n+1] suffix sums of P, Q, R
    const int n1 = n_gamma + 1;
    for (int k = tid; k < n_gamma; k += T) {
        double2 ag = AG[k];
        double wf = g_f[k];
        bool live = k >= klo && k <= khi && ag.y < kFmax;
        double R = live ? wf * ag.y : 0.0;
        s_G[k] = ag.y;
        s_A[k] = ag.x;
        s_S[k] = live ? wf * ag.x : 0.0;
        s_S[n1 + k] = R * (live ? ag.y : 0.0);
        s_S[2 * n1 + k] = R;
    }
    if (tid < 3) s_S[tid * n1 + n_gamma] = 0.0;
    __syncthreads();
    // suffix scan: warp w < 3 scans array w; lane l owns a contiguous chunk, highest k first
    if (warp < 3) {
        double* a = s_S + warp * n1;
        const int chunk = (n_gamma + 31) / 32;
        const int hi = n_gamma - lane * chunk, lo = max(hi - chunk, 0);  // lane 0 owns the top chunk
        double run = 0.0;
        for (int k = hi - 1; k >= lo; --k) { run += a[k]; a[k] = run; }
        if (hi <= 0) run = 0.0;
        // exclusive prefix over lanes (lane 0 = top chunk): offset = sum of totals of lanes < l
        double incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        double offset = incl - run;
        for (int k = hi - 1; k >= lo; --k) a[k] += offset;
    }
    __syncthreads();
    const double* SP = s_S;
    const double* SQ = s_S + n1;
    const double* SR = s_S + 2 * n1;
    const double* a_eps = ang + (size_t)m * 5 * n_pairs;
    const double* a_omc = a_eps + n_pairs;
    const double* a_b = a_omc + n_pairs;
    const double* a_W = a_b + n_pairs;
    const double* a_c = a_W + n_pairs;
    double acc = 0.0;
    for (int j = tid; j < n_pairs; j += T) {
        const double cj = a_c[j], b = a_b[j], W = a_W[j];
        if (!(cj > 0.0)) continue;
        int lo = 0, hi = n_gamma;
        for (int it = 0; it < search_iters; ++it) {
            int mid = (lo + hi) >> 1;
            bool open = lo < hi;
            bool le = s_G[min(mid, n_gamma - 1)] <= cj;
            hi = (open && le) ? mid : hi;
            lo = (open && !le) ? mid + 1 : lo;
        }
        int k = lo;
        const double band = 1e-6 * cj;
        int ka = min(k, n_gamma - 1), kb = max(k - 1, 0);
        double Ga = s_G[ka], Aa = s_A[ka], Gb = s_G[kb], Ab = s_A[kb];
        bool doubt = (k < n_gamma && (fabs(Ga - cj) <= band || Aa > 1e8)) ||
                     (k > 0 && (fabs(Gb - cj) <= band || (Ab > 1e8 && Gb < kFmax))) || !(cj < kFmax);
        if (doubt)
            k = first_unmasked_exact(g_gm, n_gamma, a_eps[j], nu_to_eps(nu[i], models[m].z, 1.0), a_omc[j]);
        k = max(k, klo);
        if (k > khi || khi < 0) continue;
        double I = fma(b, fma(b, SQ[k], -2.0 * SR[k]), SP[k]);
        acc = fma(W, I, acc);
    }
    acc = block_sum(acc, red);
    if (tid == 0) sed[(size_t)m * n_nu + i] = ec_prefactor(variant, models[m], nu[i]) * acc;
}

// ---------------------------------------------------------------------------------------------
static int ec_n_pairs(int variant, int n_mu, int n_phi) {
    if (variant == AGNPY_EC_PS_BEHIND_JET) return 1;
    if (variant == AGNPY_EC_DT) return n_phi;
    return n_mu * n_phi;
}

struct EcPlan {
    int n_pairs, T, J, S;
    int W() const { return T / 32; }  // partials per slice
};
static EcPlan ec_plan(int variant, int n_mu, int n_phi, int integ) {
    EcPlan p;
    p.n_pairs = ec_n_pairs(variant, n_mu, n_phi);
    p.T = p.n_pairs <= 32 ? 32 : p.n_pairs <= 64 ? 64 : 128;
    p.J = 1;
    // measured on B200 (scripts/tune_ec.py): 128 threads x 8 pairs is the fastest tile at the
    // default 5000 pairs; smaller pair counts keep more, smaller tiles
    if (integ == AGNPY_INTEG_TRAPZ) p.J = p.n_pairs >= 2048 ? 8 : p.n_pairs >= 512 ? 4 : p.n_pairs >= 256 ? 2 : 1;
    if (const char* e = getenv("AGNPY_B200_EC_PLAN")) {  // development knob: "T,J"
        int t = 0, j = 0;
        if (sscanf(e, "%d,%d", &t, &j) == 2 && t >= 32 && t <= 256 && t % 32 == 0 &&
            (j == 1 || j == 2 || j == 4 || j == 8) && (integ == AGNPY_INTEG_TRAPZ || j == 1)) {
            p.T = t;
            p.J = j;
        }
    }
    p.S = (p.n_pairs + p.T * p.J - 1) / (p.T * p.J);
    return p;
}

// doubles of scratch per model
static size_t ec_words(const EcPlan& p, int n_nu, int n_gamma) {
    return 5 * (size_t)n_gamma + 2      // gamma tables + hull (2 ints in one double slot, padded)
           + 5 * (size_t)p.n_pairs      // pair table
           + 2 * (size_t)n_nu * n_gamma // {A,G}
           + (size_t)n_nu * p.S * p.W();
}

size_t ec_workspace_bytes(int variant, int n_models, int n_nu, int n_gamma, int n_mu, int n_phi) {
    EcPlan a = ec_plan(variant, n_mu, n_phi, AGNPY_INTEG_LOGLOG);  // J=1 -> largest S
    EcPlan b = ec_plan(variant, n_mu, n_phi, AGNPY_INTEG_TRAPZ);
    size_t per = ec_words(a, n_nu, n_gamma);
    size_t per_b = ec_words(b, n_nu, n_gamma);
    return (per > per_b ? per : per_b) * sizeof(double) * (size_t)n_models + 256;
}

template <int INTEG, int J>
static void launch_main(dim3 grid, int T, size_t smem, cudaStream_t stream,
                        const agnpy_model_t* models, const double* nu, int n_nu, const double* gt,
                        const int* hull, int n_gamma, int iters, const double2* nt,
                        const double* ang, int n_pairs, int S, double* part) {
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(ec_main_kernel<INTEG, J>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    ec_main_kernel<INTEG, J><<<grid, T, smem, stream>>>(models, nu, n_nu, gt, hull, n_gamma, iters,
                                                        nt, ang, n_pairs, S, part);
}

int run_ec(int variant, const agnpy_model_t* models, int n_models, int ne_kind, const double* nu,
           int n_nu, const double* gamma, int n_gamma, const double* mu, int n_mu,
           const double* phi, int n_phi, const double* ne_table, int integ, double* sed,
           double* ws, cudaStream_t stream) {
    const bool prefix = integ == AGNPY_INTEG_TRAPZ_PREFIX;
    const bool fp32 = integ == AGNPY_INTEG_TRAPZ_FP32;
    if (prefix || fp32) integ = AGNPY_INTEG_TRAPZ;  // same tables (trapz weights folded into wf)
    EcPlan p = ec_plan(variant, n_mu, n_phi, integ);
    const size_t M = (size_t)n_models;
    // carve: {A,G} table first (16-byte aligned: ws comes from cudaMalloc), then the rest
    double2* nt = (double2*)ws;
    double* gt = (double*)(nt + M * n_nu * n_gamma);
    double* ang = gt + M * 5 * n_gamma;
    double* part = ang + M * 5 * p.n_pairs;
    int* hull = (int*)(part + M * n_nu * p.S * p.W());
    {
        int nB = (p.n_pairs + 127) / 128, nC = (n_gamma + 127) / 128;
        dim3 grid(1 + nB + nC * n_nu, n_models);
        ec_tables_kernel<<<grid, 128, 0, stream>>>(variant, models, ne_kind, nu, n_nu, gamma, n_gamma,
                                                   mu, n_mu, phi, n_phi, ne_table, integ, p.n_pairs,
                                                   gt, hull, ang, nt);
        note_launch();
    }
    if (prefix) {
        int iters = 1;
        while ((1 << iters) < n_gamma + 1) ++iters;
        size_t smem = (5 * (size_t)n_gamma + 3) * sizeof(double);
        if (smem > 200 * 1024) return fail(AGNPY_ERR_ARG, "n_gamma too large for the EC prefix kernel");
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(ec_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(n_nu, n_models);
        timing_begin(stream);
        ec_prefix_kernel<<<grid, 256, smem, stream>>>(variant, models, nu, n_nu, gt, hull, n_gamma, iters,
                                                      nt, ang, p.n_pairs, sed);
        timing_end(stream);
        note_launch();
        return check_last("ec_sed (prefix)");
    }
    size_t smem = (size_t)n_gamma * sizeof(double) * (integ == AGNPY_INTEG_TRAPZ ? 3 : 7);
    if (smem > 200 * 1024) return fail(AGNPY_ERR_ARG, "n_gamma too large for the EC kernel");
    // slice groups per (nu, model): a block walks all its slices once the grid fills the GPU
    // many times over (amortises the table copy), else one block per slice (latency)
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int Sg = p.S;
    while (Sg > 1 && (long long)n_nu * n_models * ((Sg + 1) / 2) >= 32LL * sms) Sg = (Sg + 1) / 2;
    if (const char* e = getenv("AGNPY_B200_EC_SG")) {  // development knob
        int v = atoi(e);
        if (v >= 1) Sg = v < p.S ? v : p.S;
    }
    int iters = 1;
    while ((1 << iters) < n_gamma + 1) ++iters;
    dim3 grid(Sg, n_nu, n_models);
    timing_begin(stream);
#define EC_LAUNCH(I, JJ) launch_main<I, JJ>(grid, p.T, smem, stream, models, nu, n_nu, gt, hull, \
                                            n_gamma, iters, nt, ang, p.n_pairs, p.S, part)
    if (fp32) {
        size_t smem32 = (size_t)n_gamma * (2 * sizeof(double) + 3 * sizeof(float));
#define EC_LAUNCH32(JJ)                                                                            \
    do {                                                                                           \
        if (smem32 > 48 * 1024)                                                                    \
            cudaFuncSetAttribute(ec_main_fp32_kernel<JJ>,                                          \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);        \
        ec_main_fp32_kernel<JJ><<<grid, p.T, smem32, stream>>>(models, nu, n_nu, gt, hull, n_gamma, \
                                                               iters, nt, ang, p.n_pairs, p.S, part); \
    } while (0)
        if (p.J == 8) EC_LAUNCH32(8);
        else if (p.J == 4) EC_LAUNCH32(4);
        else if (p.J == 2) EC_LAUNCH32(2);
        else EC_LAUNCH32(1);
#undef EC_LAUNCH32
    } else if (integ == AGNPY_INTEG_LOGLOG) EC_LAUNCH(AGNPY_INTEG_LOGLOG, 1);
    else if (p.J == 8) EC_LAUNCH(AGNPY_INTEG_TRAPZ, 8);
    else if (p.J == 4) EC_LAUNCH(AGNPY_INTEG_TRAPZ, 4);
    else if (p.J == 2) EC_LAUNCH(AGNPY_INTEG_TRAPZ, 2);
    else EC_LAUNCH(AGNPY_INTEG_TRAPZ, 1);
#undef EC_LAUNCH
    timing_end(stream);
    note_launch();
    {
        dim3 g2((n_nu + 127) / 128, n_models);
        ec_finish_kernel<<<g2, 128, 0, stream>>>(variant, models, nu, n_nu, part, p.S * p.W(), sed);
        note_launch();
    }
    return check_last("ec_sed");
}

}  // namespace agb
