// K3: external-Compton SED kernels (iso_mono, ps_behind_jet, ss_disk, blr, dt).
//
// Reference arithmetic replaced:
//   agnpy/compton/external_compton.py:122-142, 221-240, 333-373, 463-490, 578-600
//   agnpy/compton/kernels.py:36-68 (get_gamma_min, compton_kernel)
//   agnpy/utils/geometry.py:5-55 (cos_psi, x_re_shell, mu_star_shell, x_re_ring)
//   agnpy/targets/targets.py:289-323 (SSDisk statics)
//
// The reference broadcasts a full (gamma, mu, phi, nu) array (8 GB at the default grids).  Here
// every thread keeps a few (mu,phi) "pairs" in registers and walks the gamma axis, reading
// per-(gamma,nu) values tabulated once per (nu, model).  With
//     y = 1 - eps_s/gamma,  A = y + 1/y,  G = eps_s/(gamma^2 y),  b = 1/(eps (1 - cos psi))
// the kernel of kernels.py:61-65 is   K = y + 1/y - 2t + t^2 = A + t (t - 2),  t = G b:
// no division per point (the two of the nominal count are hoisted into the tables), and the
// kinematic mask gamma >= gamma_min(nu, pair) becomes a start index of the gamma loop (first k
// with G_k <= c = 2 eps (1 - cos psi)), re-decided with the reference's own operation order
// (kernels.py:36-40) whenever c lies within a guard band of a table entry.
//
// Map of this file:
//   ec_tables_kernel / ec_tables2_kernel / ec_rows_merge_kernel
//                                          gamma table, pair table (sorted by c, optionally with
//                                          the phi axis folded), {A,G} table, polynomial rows
//   ec_main_poly_kernel<J,T>               np.trapz, >= 512 pairs: 2 DFMA per point (the default)
//   ec_main_kernel<INTEG,J>                trapz_loglog, and np.trapz with few pairs (4 instr/pt)
//   ec_moments_kernel / ec_moment_kernel   opt-in: np.trapz through prefix moments of the pair table
//   ec_finish_kernel                       fixed-order sum of the warp partials x prefactor
//   run_ec                                 plan, scratch carving, launches
#include <limits.h>
#include <algorithm>
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
//                        ang[m][5][j], ang[m][6][j] = 1 - cos psi and W of the mirror partner
//                                                     (folded phi axis, see below; W = 0: none)
// Folding (AGNPY_INTEG_FOLD_PHI): when the phi grid is mirror-symmetric, phi_i + phi_{n-1-i} =
// 2 pi m, cos phi -- hence cos psi, b and c -- takes the same value at both points of a mirror
// pair, up to the last bit of cos().  The table then holds one entry per mirror pair with the
// weight W_i + W_partner in column 3, so the main kernel evaluates every distinct integrand point
// once (half the points of the default grid).  The partner's own 1 - cos psi and weight are kept
// so that the one decision that is sensitive to the last bit -- on which side of a table entry
// G_k the pair's threshold c falls -- is still taken for each partner separately with the
// reference's formula whenever c is inside the guard band (ec_main_poly_kernel).
// ---------------------------------------------------------------------------------------------
constexpr int kAngCols = 7, kRawCols = 6;  // raw column 5: sorted 64-bit keys of the segments
__device__ __forceinline__ void ec_pair_values(
    int variant, const agnpy_model_t& M, const double* __restrict__ mu, int n_mu,
    const double* __restrict__ phi, int n_phi, int imu, int iphi, double& eps, double& omc,
    double& W) {
    const double* t = M.tgt;
    double cpsi;
    const int j = iphi;  // ring torus: the pair index is the phi index
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
    omc = 1.0 - cpsi;
}

// unsorted pair values of one model: raw[m][0][j] = eps_j, [1][j] = 1 - cos psi_j, [2][j] = W_j
// (folded: W_j + W_partner), [3][j], [4][j] = 1 - cos psi and W of the mirror partner.
// n_phi_eff = n_phi (plain) or ceil(n_phi / 2) (folded): pair j = (imu, iphi < n_phi_eff).
__device__ __forceinline__ void ec_pairs_role(
    int variant, const agnpy_model_t& M, const double* __restrict__ mu, int n_mu,
    const double* __restrict__ phi, int n_phi, int n_phi_eff, int n_pairs, double* __restrict__ raw,
    int j) {
    if (j >= n_pairs) return;
    int imu = 0, iphi = j;
    if (variant != AGNPY_EC_PS_BEHIND_JET && variant != AGNPY_EC_DT) {
        imu = j / n_phi_eff;
        iphi = j - imu * n_phi_eff;
    }
    double eps, omc, W, omc_p = 0.0, W_p = 0.0;
    ec_pair_values(variant, M, mu, n_mu, phi, n_phi, imu, iphi, eps, omc, W);
    const int iphi_p = n_phi - 1 - iphi;
    if (n_phi_eff != n_phi && iphi_p != iphi) {
        double eps_p;
        ec_pair_values(variant, M, mu, n_mu, phi, n_phi, imu, iphi_p, eps_p, omc_p, W_p);
        W = W + W_p;
    }
    raw[j] = eps;
    raw[n_pairs + j] = omc;
    raw[2 * n_pairs + j] = W;
    raw[3 * n_pairs + j] = omc_p;
    raw[4 * n_pairs + j] = W_p;
}

// One 1024-thread block takes one segment of kPairSeg consecutive pairs, one pair per thread,
// and stores it sorted by descending c_j (pairs that can never contribute, c_j <= 0 or NaN,
// last): consecutive table entries then have nearly equal first-unmasked gamma indices, which
// is what lets a warp of the main kernel share one short index window.  The order of the
// (mu,phi) sum is free (it is a plain weighted sum).  The sort is a bitonic network on one
// packed 64-bit key per thread -- high word of c_j, then the pair index, so it is a total,
// deterministic order -- exchanged by warp shuffles below stride 32 and through shared memory
// above.
constexpr int kPairSeg = 1024;
// one sorted pair -> its row of the table
__device__ __forceinline__ void ec_pair_store(const double* __restrict__ raw, int n_pairs,
                                              double* __restrict__ row, int src, int j) {
    const double e = raw[src], o = raw[n_pairs + src], w = raw[2 * n_pairs + src];
    row[j] = e;
    row[n_pairs + j] = o;
    row[2 * n_pairs + j] = 1.0 / (e * o);
    row[3 * n_pairs + j] = w;
    row[4 * n_pairs + j] = 2.0 * (e * o);
    row[5 * n_pairs + j] = raw[3 * n_pairs + src];
    row[6 * n_pairs + j] = raw[4 * n_pairs + src];
}
// `keys` == nullptr (one segment): the sorted rows are written directly; else the sorted keys of
// this segment are stored for the merge role of ec_rows_merge_kernel, which interleaves the segments.
__device__ __forceinline__ void ec_pairs_sort_role(const double* __restrict__ raw, int n_pairs,
                                                   double* __restrict__ row,
                                                   unsigned long long* __restrict__ keys, int seg,
                                                   unsigned long long* s_x) {
    const int t = threadIdx.x, j0 = seg * kPairSeg, cnt = min(kPairSeg, n_pairs - j0);
    unsigned long long v = 0ull;  // padding sorts behind every real pair
    if (t < cnt) {
        const double c = 2.0 * (raw[j0 + t] * raw[n_pairs + j0 + t]);
        const unsigned hi = (c > 0.0) ? (unsigned)(__double_as_longlong(c) >> 32) : 0u;
        v = ((unsigned long long)hi << 32) | (unsigned)(~(unsigned)(j0 + t));  // unique over the model
    }
    for (int size = 2; size <= kPairSeg; size <<= 1) {
        const bool desc = (t & size) == 0;
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            unsigned long long other;
            if (stride < 32) {
                other = __shfl_xor_sync(0xffffffffu, v, stride);
            } else {
                __syncthreads();
                s_x[t] = v;
                __syncthreads();
                other = s_x[t ^ stride];
            }
            const bool keep_max = ((t & stride) == 0) == desc;
            v = keep_max ? max(v, other) : min(v, other);
        }
    }
    if (t < cnt) {
        if (keys) keys[j0 + t] = v;
        else {
            AGB_CHECK((int)(~(unsigned)(v & 0xffffffffull)) >= 0 && (int)(~(unsigned)(v & 0xffffffffull)) < n_pairs);
            ec_pair_store(raw, n_pairs, row, (int)(~(unsigned)(v & 0xffffffffull)), j0 + t);
        }
    }
}

// More than one segment: the global position of a sorted key is its position in its own segment
// plus the number of larger keys in every other segment (keys are unique), found by binary search.
// A globally sorted table halves the spread of first-unmasked indices inside a warp of the main
// kernel (a 1024-pair segment covers the whole range of thresholds, 1/n_seg of the table does not).
__device__ __forceinline__ void ec_pairs_merge_role(const double* __restrict__ raw, int n_pairs, int n_seg,
                                                    double* __restrict__ ang, int m, int j) {
    if (j >= n_pairs) return;
    const double* r = raw + (size_t)m * kRawCols * n_pairs;
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(r + 5 * (size_t)n_pairs);
    const unsigned long long key = keys[j];
    const int own = j / kPairSeg;
    int rank = j - own * kPairSeg;
    for (int seg = 0; seg < n_seg; ++seg) {
        if (seg == own) continue;
        const int j0 = seg * kPairSeg, cnt = min(kPairSeg, n_pairs - j0);
        int lo = 0, hi = cnt;  // first position whose key is smaller (descending order)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[j0 + mid] > key) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    AGB_CHECK(rank >= 0 && rank < n_pairs);
    AGB_CHECK((int)(~(unsigned)(key & 0xffffffffull)) >= 0 && (int)(~(unsigned)(key & 0xffffffffull)) < n_pairs);
    ec_pair_store(r, n_pairs, ang + (size_t)m * kAngCols * n_pairs, (int)(~(unsigned)(key & 0xffffffffull)), rank);
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

// First table launch (128 threads); the block role follows from blockIdx.x:
//   [0, 1)                     gamma tables + hull
//   [1, 1 + nB)                unsorted pair values, 128 pairs per block
//   [1 + nB, 1 + nB + nC*n_nu) {A, G} table, 128 gammas of one frequency per block: trapz_loglog only
//                              (the np.trapz paths form A and G in their row role / main kernel)
__global__ void __launch_bounds__(128) ec_tables_kernel(
    int variant, const agnpy_model_t* __restrict__ models, int ne_kind,
    const double* __restrict__ nu, int n_nu, const double* __restrict__ gamma, int n_gamma,
    const double* __restrict__ mu, int n_mu, const double* __restrict__ phi, int n_phi,
    const double* __restrict__ ne_table, int integ, int n_pairs, double* __restrict__ gt,
    int* __restrict__ hull, double* __restrict__ raw, double2* __restrict__ nt) {
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
        const int n_phi_eff =
            (variant == AGNPY_EC_PS_BEHIND_JET || variant == AGNPY_EC_DT) ? n_phi : n_pairs / n_mu;
        ec_pairs_role(variant, M, mu, n_mu, phi, n_phi, n_phi_eff, n_pairs,
                      raw + (size_t)m * kRawCols * n_pairs, (bx - 1) * 128 + threadIdx.x);
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

// trapz_loglog over gamma for one pair (utils/math.py:55-119), from its first unmasked index ks:
// y_k = f_k K_k(b), z_k = gamma_k y_k, carried sequentially.  An interval takes the logarithm-free
// form of common.cuh (loglog_interval_z); where neighbouring samples differ by more than a factor
// 2.1, or a sample is below the smallest normal double (the reference clips its logarithm there),
// it takes the logarithmic form with ln y = ln f + ln K (ln f tabulated; K >= 1 where unmasked).
__device__ __forceinline__ double ec_loglog_gamma(int ks, int khi, double b, const double* s_f,
                                                  const double* s_lf, const double* s_A,
                                                  const double* s_G, const double* s_gm,
                                                  const double* s_idl, const double* s_lr) {
    double y_prev = 0.0, z_prev = 0.0, K_prev = 1.0, a = 0.0;
    for (int k = ks; k <= khi; ++k) {
        const double f = s_f[k];
        double y = 0.0, z = 0.0, K = 1.0;
        if (f != 0.0) {
            const double t = s_G[k] * b;
            K = fma(t, t - 2.0, s_A[k]);
            y = f * K;
            z = s_gm[k] * y;
        }
        if (k > ks && y != 0.0 && y_prev != 0.0) {  // a zero end point switches the interval off (:112-117)
            bool ok;
            double v = loglog_interval_z(z_prev, z, s_lr[k - 1], ok);
            if (!ok || y < kTiny || y_prev < kTiny) {
                const double ly = (y >= kTiny) ? s_lf[k] + log(K) : clipped_log(y);
                const double lyp = (y_prev >= kTiny) ? s_lf[k - 1] + log(K_prev) : clipped_log(y_prev);
                v = loglog_interval_pre(s_gm[k - 1], s_gm[k], s_idl[k - 1], s_lr[k - 1], y_prev, y, lyp, ly);
            }
            a += v;
        }
        y_prev = y;
        z_prev = z;
        K_prev = K;
    }
    return a;
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
    // small tables on the np.trapz path come without the {A, G} table (nt == nullptr): with a few dozen
    // pairs per (nu, model) it is cheaper to form the two values here than to stream them through HBM
    const double eps_s_blk = nu_to_eps(nu[i], models[m].z, 1.0);
    for (int k = tid; k < n_gamma; k += T) {
        double2 ag;
        if (nt != nullptr) {
            ag = AG[k];
        } else {  // the arithmetic of ec_nu_role
            const double g = g_gm[k];
            const double y = 1.0 - eps_s_blk / g;
            ag.x = y + 1.0 / y;
            ag.y = (y > 0.0) ? eps_s_blk / ((g * g) * y) : INFINITY;
        }
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

    const double* a_eps = ang + (size_t)m * kAngCols * n_pairs;
    const double* a_omc = a_eps + n_pairs;
    const double* a_b = a_omc + n_pairs;
    const double* a_W = a_b + n_pairs;
    const double* a_c = a_W + n_pairs;
    const double* a_omc_p = a_c + n_pairs;  // mirror partner of a folded pair (W_p = 0: none)
    const double* a_W_p = a_omc_p + n_pairs;
    for (int slice = blockIdx.x; slice < S; slice += gridDim.x) {
        double b[J], W[J], c[J], acc[J];
        int lo[J], hi[J], lo_p[J];
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
            lo_p[jj] = -1;
            if (doubt) {
                int j = slice * (T * J) + jj * T + tid;
                const double eps_s = nu_to_eps(nu[i], models[m].z, 1.0);
                k = first_unmasked_exact(g_gm, n_gamma, a_eps[j], eps_s, a_omc[j]);
                if (INTEG == AGNPY_INTEG_LOGLOG && a_W_p[j] != 0.0) {
                    const int kp = first_unmasked_exact(g_gm, n_gamma, a_eps[j], eps_s, a_omc_p[j]);
                    if (kp != k) lo_p[jj] = kp;
                }
            }
            lo[jj] = k;
            AGB_CHECK(k >= 0 && k <= n_gamma);
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
                    acc[jj] = ec_loglog_gamma(max(lo[jj], klo), khi, b[jj], s_f, s_lf, s_A, s_G, s_gm, s_idl,
                                              s_lr);
                    // folded phi axis: a mirror partner whose own decision (inside the guard band)
                    // starts at another index gets its own gamma integral
                    if (lo_p[jj] >= 0) {
                        const double a_p = ec_loglog_gamma(max(lo_p[jj], klo), khi, b[jj], s_f, s_lf, s_A, s_G,
                                                           s_gm, s_idl, s_lr);
                        const int j = slice * (T * J) + jj * T + tid;
                        const double W_p = a_W_p[j];
                        // W acc stands for (W - W_p) acc + W_p a_p
                        acc[jj] = acc[jj] + (W_p / W[jj]) * (a_p - acc[jj]);
                    }
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
// np.trapz main path: the integrand as a polynomial in b, frequency-incremental masks.
//
// With the trapz weight folded into wf_k = w_k N_e(gamma_k)/gamma_k^2 the summand of the gamma
// integral is
//     wf_k K_k(b) = wf_k (A_k + t (t - 2)),  t = G_k b
//                 = P_k + Q_k b^2 + R_k b,   P_k = wf_k A_k,  Q_k = wf_k G_k^2,  R_k = -2 wf_k G_k,
// so each integrand point costs two DFMAs, acc = fma(Q_k, b^2, acc); acc = fma(R_k, b, acc): one
// warp-uniform operand each, which the register file feeds at 85 % of the FP64 issue rate where
// the three-register Horner form fma(fma(Q,b,R),b,acc) stops at 67 % (scripts/micro/fp64_mix.cu).
// The b-independent term sums to SP[k0] = sum_{k >= k0} P_k, a suffix sum per (nu, model); every
// pair adds SP at its own first unmasked index.  No cancellation is introduced: on the unmasked
// range 0 < t <= 2, so Q b^2 + R b = wf (t^2 - 2t) lies in [-wf, 0] while P = wf A >= 2 wf.
//
// ec_tables2_kernel writes one row per (nu, model): {Q,R}[n] (16-byte pairs), G[n], SP[n+1]
// and the index range where y = 1 - eps_s/gamma < 1e-8 (G ill-conditioned).
// ---------------------------------------------------------------------------------------------
// row of one (nu, model): {Q,R}[n + kRowPad] (the pad entries are zero: the unrolled gamma loop may run
// up to kRowPad - 1 steps past the last electron index), G[n], SP[n + 1], {ill_lo, ill_hi}; even word count
constexpr int kRowPad = 4;
__host__ __device__ __forceinline__ int ec_row_words(int n_gamma) {
    return (2 * (n_gamma + kRowPad) + 2 * n_gamma + 2 + 1) & ~1;
}

// Polynomial rows of the np.trapz path, one warp per (nu, model) row (they depend on the gamma table
// only, so they share the third table launch with the merge of the sorted pair segments, see
// ec_rows_merge_kernel).
// Row role: lane l owns gamma indices l, l+32, ...: it evaluates y, A, G (kernels.py:61-65 in the
// operation order of ec_nu_role), writes {Q,R} and G, and the warp builds the suffix sums of
// P = wf A from the top of the grid down (shuffle scan per group of 32 plus a running carry).
constexpr int kTab2Threads = 1024;
__device__ __forceinline__ void ec_poly_row_role(
    const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ gamma, int n_gamma, const double* __restrict__ gt,
    const int* __restrict__ hull, double* __restrict__ rows, int i, int m) {
    const int lane = threadIdx.x & 31;
    if (i >= n_nu) return;
    const double* g_f = gt + (size_t)m * 5 * n_gamma + n_gamma;
    double* row = rows + ((size_t)m * n_nu + i) * ec_row_words(n_gamma);
    double2* r_QR = reinterpret_cast<double2*>(row);
    double* r_G = row + 2 * (n_gamma + kRowPad);
    double* r_SP = r_G + n_gamma;
    const int klo = hull[2 * m], khi = hull[2 * m + 1];
    // EC uses nu_to_epsilon_prime(nu, z) without the Doppler factor (external_compton.py:123)
    const double eps_s = nu_to_eps(nu[i], models[m].z, 1.0);
    if (lane < kRowPad) r_QR[n_gamma + lane] = make_double2(0.0, 0.0);
    double carry = 0.0;
    int ill_lo = n_gamma, ill_hi = 0;
    // groups of 32 indices from the top of the grid down, four groups per trip: their divisions and
    // shuffle scans are independent and overlap; only the carry runs through them in order
    constexpr int kB = 4;
    for (int base = ((n_gamma - 1) >> 5) << 5; base >= 0; base -= 32 * kB) {
        double incl[kB];
#pragma unroll
        for (int q = 0; q < kB; ++q) {  // straight-line on a clamped index; only the stores are predicated
            const int k = base - 32 * q + lane;
            const bool in = k >= 0 && k < n_gamma;
            const int kk = min(max(k, 0), n_gamma - 1);
            const double g = gamma[kk];
            const double y = 1.0 - eps_s / g;
            const double A = y + 1.0 / y;
            const double G = (y > 0.0) ? eps_s / ((g * g) * y) : INFINITY;
            const double wf = g_f[kk];
            const bool live = in && kk >= klo && kk <= khi && G < kFmax;
            const double wG = live ? wf * G : 0.0;  // exactly 0 outside the electron range
            if (in) {
                r_QR[kk] = make_double2(wG * G, -2.0 * wG);
                r_G[kk] = G;
                if (A > 1e8) { ill_lo = min(ill_lo, kk); ill_hi = max(ill_hi, kk + 1); }
            }
            incl[q] = live ? wf * A : 0.0;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {  // suffix sums over the lanes of each group
#pragma unroll
            for (int q = 0; q < kB; ++q) {
                const double v = __shfl_down_sync(0xffffffffu, incl[q], o);
                if (lane + o < 32) incl[q] += v;
            }
        }
#pragma unroll
        for (int q = 0; q < kB; ++q) {
            const int k = base - 32 * q + lane;
            if (k >= 0 && k < n_gamma) r_SP[k] = incl[q] + carry;
            carry += __shfl_sync(0xffffffffu, incl[q], 0);
        }
    }
    ill_lo = __reduce_min_sync(0xffffffffu, ill_lo);
    ill_hi = __reduce_max_sync(0xffffffffu, ill_hi);
    if (lane == 0) {
        r_SP[n_gamma] = 0.0;
        *reinterpret_cast<int2*>(r_SP + n_gamma + 1) = make_int2(ill_lo, ill_hi);
    }
}

// Second table launch (1024 threads): one sorted segment of the pair table per block (from the unsorted
// values).
__global__ void __launch_bounds__(kTab2Threads) ec_tables2_kernel(const double* __restrict__ raw, int n_pairs,
                                                                  int n_seg, double* __restrict__ ang) {
    __shared__ unsigned long long s_x[kPairSeg];
    const int m = blockIdx.y;
    const double* r = raw + (size_t)m * kRawCols * n_pairs;
    // raw column 5 holds the keys; const_cast: it is scratch of this launch sequence
    unsigned long long* keys = n_seg > 1
        ? reinterpret_cast<unsigned long long*>(const_cast<double*>(r) + 5 * (size_t)n_pairs) : nullptr;
    ec_pairs_sort_role(r, n_pairs, ang + (size_t)m * kAngCols * n_pairs, keys, blockIdx.x, s_x);
}

// Third table launch (256 threads); the block role follows from blockIdx.x:
//   [0, n_rb)             np.trapz polynomial path: rows, 8 per block (one warp each); first in the
//                         grid: they run longest
//   [n_rb, n_rb + n_mb)   more than one sorted segment: global rank of every pair (merge)
constexpr int kRowWarps = 8;
__global__ void __launch_bounds__(32 * kRowWarps) ec_rows_merge_kernel(
    const double* __restrict__ raw, int n_pairs, int n_seg, double* __restrict__ ang,
    const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ gamma, int n_gamma, const double* __restrict__ gt,
    const int* __restrict__ hull, double* __restrict__ rows, int n_rb) {
    const int m = blockIdx.y;
    if ((int)blockIdx.x < n_rb)
        ec_poly_row_role(models, nu, n_nu, gamma, n_gamma, gt, hull, rows,
                         (int)blockIdx.x * kRowWarps + (threadIdx.x >> 5), m);
    else
        ec_pairs_merge_role(raw, n_pairs, n_seg, ang, m, ((int)blockIdx.x - n_rb) * (int)blockDim.x + threadIdx.x);
}

// 1-D bulk copy (TMA engine) global -> shared, completion counted on an mbarrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes,
                                              unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// the reference's own decision (kernels.py:36-40) for one (nu, pair)
__device__ __noinline__ int ec_first_unmasked_ref(const double* __restrict__ gm, int n_gamma,
                                                  const double* __restrict__ a_eps,
                                                  const double* __restrict__ a_omc, int j,
                                                  double nu_i, double z) {
    return first_unmasked_exact(gm, n_gamma, a_eps[j], nu_to_eps(nu_i, z, 1.0), a_omc[j]);
}
// 64-bit signed max / min over the warp in two 32-bit REDUX steps (high words, then the low words of
// the lanes that hold the winning high word) instead of five 64-bit shuffle rounds
__device__ __forceinline__ long long warp_max_ll(long long v) {
    const unsigned long long w = (unsigned long long)v ^ 0x8000000000000000ull;  // signed -> unsigned order
    const unsigned h = (unsigned)(w >> 32);
    const unsigned hi = __reduce_max_sync(0xffffffffu, h);
    const unsigned lo = __reduce_max_sync(0xffffffffu, h == hi ? (unsigned)w : 0u);
    return (long long)((((unsigned long long)hi << 32) | lo) ^ 0x8000000000000000ull);
}
__device__ __forceinline__ long long warp_min_ll(long long v) {
    const unsigned long long w = (unsigned long long)v ^ 0x8000000000000000ull;
    const unsigned h = (unsigned)(w >> 32);
    const unsigned hi = __reduce_min_sync(0xffffffffu, h);
    const unsigned lo = __reduce_min_sync(0xffffffffu, h == hi ? (unsigned)w : 0xffffffffu);
    return (long long)((((unsigned long long)hi << 32) | lo) ^ 0x8000000000000000ull);
}

// ---------------------------------------------------------------------------------------------
// ec_main_poly_kernel<J, T>: grid = (S pair slices, nu chunks, models), T threads x J consecutive pairs
// of the globally c-sorted table (so the first unmasked indices of a warp lie within a few entries
// of each other).  A block keeps its pairs in registers and walks the L_nu frequencies of its
// chunk; their rows arrive by TMA bulk copies (cp.async.bulk + one mbarrier per row) and the warps
// run through the chunk independently, no block barrier after the prologue.  Per (warp, nu):
//   * index window [wlo, whi] = first entries with G_k <= the warp's largest / smallest threshold,
//     by lane-parallel probes: lane l tests entry k_prev - 1 + l of the new row, one ballot gives the
//     index (G_k grows with nu, the indices move up by an entry or two per frequency step); the
//     first frequency of a chunk uses a 32-way search (2 round trips at 200 gammas).  Comparisons run
//     on the bit patterns of the positive doubles (integer pipe, not FP64);
//   * one pass over the entries [wlo-1, whi]: per pair the index count lo = wlo + #{G_k > c} (sign
//     bit of a high-word difference) and the guard band (running unsigned minimum of the same
//     difference): an entry within 3 units of the high word of c (>= 1.4e-6 relative) next to the
//     index, an ill-conditioned G in the window or a non-finite c -> that pair is decided with the
//     reference's own formula (kernels.py:36-40), mirror partners of a folded phi axis separately;
//   * phase 1 (the ragged start, window-many steps, predicated), phase 2 (2 DFMA per point, 4 gamma
//     steps per trip, zero-padded rows instead of a remainder loop), then W_j (acc_j + SP[lo_j]);
//     pair weights live in registers, dead pairs carry W = b = 0 and need no special casing;
//   * the partial sums of the chunk's frequencies are parked in shared memory and reduced once per
//     block (8 values per lane + 2 shuffles) instead of once per frequency.
// Why it looks like this (ncu, profiles/r1f_* -> r2_final_kernels_full.txt): a non-FP64 instruction
// costs the sub-partition a dispatch slot the DFMA stream cannot use (scripts/micro/fp64_mix.cu), and
// round 1 executed 1500 instructions per (warp, nu) for 800 DFMAs, 545 of them outside the gamma
// loop with 48 % of the stall samples.  Round 2: global sort of the pair table (windows of ~2 instead
// of ~5 entries, 11 instead of 39 wasted DFMAs in the ragged start), the passes above, suffix sums as
// accumulator seeds, and 5 pairs per thread at 128 registers / 4 blocks per SM (40 DFMAs for 8 other
// instructions per trip): FP64 pipe 57 % -> 68.5 % active, 2.545 -> 2.056 ms per 256 models.  Tried
// and measured slower (profiles/r2_tune_ec.txt): 8 and 10 pairs per thread, a second accumulator
// chain per pair, 8 gamma steps per trip, blocks of 160-256 threads with more frequencies each, 6
// blocks per SM at 80 registers, and a ring of row buffers with full / empty mbarriers feeding 25-200
// frequencies per block (the prologue it amortises is latency other blocks already hide; uneven
// chunks cost more).
// ---------------------------------------------------------------------------------------------
// first k in [0, n] with G_k <= c (G descending), all lanes cooperating: 32 probes per round
__device__ __forceinline__ int ec_warp_first_le(const long long* __restrict__ s_Gi, int n, long long c,
                                                int lane) {
    int lo = 0, len = n;
    while (len > 0) {
        const int s = (len + 31) >> 5;
        const int k = lo + lane * s;
        AGB_CHECK(lo >= 0 && lo + len <= n);
        const bool gt = k < lo + len && s_Gi[k] > c;
        const int cnt = __popc(__ballot_sync(0xffffffffu, gt));
        if (cnt == 0) return lo;
        const int end = lo + len;
        lo += (cnt - 1) * s + 1;  // the last probe that was still > c, plus one
        if (s == 1) return lo;
        len = min(s - 1, end - lo);
    }
    return lo;
}
// The same indices for the next frequency's table, given the previous ones (they move up by an entry or
// two per frequency step); anything the 16 probes of a half-warp cannot confirm -> full search.
// both ends of the window in one probe: lanes 0-15 test entries wlo - 1 + l against cmax, lanes 16-31
// entries whi - 1 + l against cmin; one load, one compare and one ballot per frequency step
__device__ __forceinline__ void ec_warp_step_le2(const long long* __restrict__ s_Gi, int n, long long cmax,
                                                 long long cmin, int& wlo, int& whi, int lane) {
    const bool upper = lane >= 16;
    const int k = (upper ? whi : wlo) - 1 + (lane & 15);
    const bool gt = k < 0 || (k < n && s_Gi[k] > (upper ? cmin : cmax));
    const unsigned bal = __ballot_sync(0xffffffffu, gt);
    const unsigned b0 = bal & 0xffffu, b1 = bal >> 16;
    const int r0 = __ffs(~b0) - 1, r1 = __ffs(~b1) - 1;  // runs of set bits from the first lane (16: all set)
    const bool ok0 = (b0 & 1u) && r0 < 16, ok1 = (b1 & 1u) && r1 < 16;
    wlo = ok0 ? wlo - 1 + r0 : ec_warp_first_le(s_Gi, n, cmax, lane);
    whi = ok1 ? whi - 1 + r1 : ec_warp_first_le(s_Gi, n, cmin, lane);
}

template <int J, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) ec_main_poly_kernel(
    const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ gt, const int* __restrict__ hull, int n_gamma,
    const double* __restrict__ rows, const double* __restrict__ ang, int n_pairs, int S, int L_nu,
    double* __restrict__ part) {
    extern __shared__ __align__(16) double sh[];  // L_nu rows | L_nu x (T+1) partial sums | L_nu mbarriers
    constexpr int W_per_block = T >> 5;
    const int slice = blockIdx.x, m = blockIdx.z;
    const int i0 = blockIdx.y * L_nu, i1 = min(i0 + L_nu, n_nu);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_words = ec_row_words(n_gamma);
    double* s_sum = sh + (size_t)L_nu * row_words;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(s_sum + (size_t)L_nu * (T + 1));
    if (tid == 0) {
        for (int r = 0; r < i1 - i0; ++r) mbar_init(bars + r, 1);
        mbar_init_fence();
    }
    __syncthreads();
    const double* g_rows = rows + (size_t)m * n_nu * row_words;
    if (tid == 0) {  // all rows of this block's frequency chunk, one bulk copy each
        const unsigned row_bytes = (unsigned)row_words * sizeof(double);
        for (int r = 0; r < i1 - i0; ++r) {
            mbar_expect_tx(bars + r, row_bytes);
            bulk_copy_g2s(sh + (size_t)r * row_words, g_rows + (size_t)(i0 + r) * row_words, row_bytes,
                          bars + r);
        }
    }
    const double* g_gm = gt + (size_t)m * 5 * n_gamma;
    const double* a_eps = ang + (size_t)m * kAngCols * n_pairs;
    const double* a_omc = a_eps + n_pairs;
    const double* a_b = a_omc + n_pairs;
    const double* a_W = a_b + n_pairs;
    const double* a_c = a_W + n_pairs;
    const double* a_omc_p = a_c + n_pairs;  // mirror partner of a folded pair (W_p = 0: none)
    const double* a_W_p = a_omc_p + n_pairs;
    const int klo = hull[2 * m], khi = hull[2 * m + 1];  // non-zero range of n_e
    const int jbase = (slice * T + tid) * J;

    double b[J], b2[J], W[J];
    int hc[J];  // high word of c; INT_MAX marks a pair that never contributes (W = b = 0)
    unsigned always_doubt = 0;  // non-finite c: always decided by the reference's own formula
    long long cmax = -1, cmin = LLONG_MAX;  // bits of c over the live pairs of the warp
#pragma unroll
    for (int jj = 0; jj < J; ++jj) {
        const int j = jbase + jj;
        const double c = j < n_pairs ? a_c[j] : -1.0;
        const bool live = c > 0.0;  // false for NaN as well
        b[jj] = live ? a_b[j] : 0.0;
        b2[jj] = b[jj] * b[jj];
        W[jj] = live ? a_W[j] : 0.0;
        const long long ci = live ? __double_as_longlong(c) : -1LL;
        hc[jj] = live ? (int)(ci >> 32) : INT_MAX;
        if (live) {
            cmax = max(cmax, ci);
            cmin = min(cmin, ci);
            if (hc[jj] >= 0x7ff00000) always_doubt |= 1u << jj;
        }
    }
    cmax = warp_max_ll(cmax);
    cmin = warp_min_ll(cmin);
    const bool warp_live = cmax > 0 && khi >= 0;
    int wlo = 0, whi = 0;  // first k with G_k <= cmax / <= cmin for the current frequency
    for (int i = i0; i < i1; ++i) {
        const int r = i - i0;
        // warps run through the chunk independently: the only wait is for row i to have landed
        mbar_wait(bars + r, 0);
        const double* s_row = sh + (size_t)r * row_words;
        const double2* s_QR = reinterpret_cast<const double2*>(s_row);
        const long long* s_Gi = reinterpret_cast<const long long*>(s_row + 2 * (n_gamma + kRowPad));
        const int* s_Gh = reinterpret_cast<const int*>(s_Gi);  // [2k+1] = high word
        const double* s_SP = s_row + 2 * (n_gamma + kRowPad) + n_gamma;
        double thread_sum = 0.0;
        if (warp_live) {
            // warp-uniform index window [wlo, whi]: every live pair's first unmasked index lies in it
            if (r == 0) {
                wlo = ec_warp_first_le(s_Gi, n_gamma, cmax, lane);
                whi = ec_warp_first_le(s_Gi, n_gamma, cmin, lane);
            } else {
                ec_warp_step_le2(s_Gi, n_gamma, cmax, cmin, wlo, whi, lane);
            }
            // one pass over the window entries: per pair lo = wlo + #{k in [wlo, whi) : G_k > c}
            // (G descending: a prefix count, on high words; a tie there is inside the guard band), and
            // the guard band itself: an entry within 3 units of the high word of c (>= 1.4e-6
            // relative) next to the index -> decide that pair with the reference's own formula
            AGB_CHECK(wlo >= 0 && wlo <= whi && whi <= n_gamma);
            int lo[J];
            unsigned near[J];  // min over the entries of (high word of c) - (high word of G_k) + 3, unsigned
            unsigned doubt = always_doubt;
#pragma unroll
            for (int jj = 0; jj < J; ++jj) { lo[jj] = wlo; near[jj] = 0xffffffffu; }
            if (wlo >= 1) {  // the entry below the window: guard band only
                const unsigned gh = (unsigned)s_Gh[2 * (wlo - 1) + 1];
#pragma unroll
                for (int jj = 0; jj < J; ++jj) near[jj] = min(near[jj], (unsigned)hc[jj] - gh + 3u);
            }
            for (int k = wlo; k < whi; ++k) {  // inside: count (sign bit of the difference) and band
                const unsigned gh = (unsigned)s_Gh[2 * k + 1];
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    const unsigned e = (unsigned)hc[jj] - gh;  // negative: G_k > c, still masked
                    lo[jj] += (int)(e >> 31);
                    near[jj] = min(near[jj], e + 3u);
                }
            }
            if (whi < n_gamma) {  // the entry at the upper end: guard band only
                const unsigned gh = (unsigned)s_Gh[2 * whi + 1];
#pragma unroll
                for (int jj = 0; jj < J; ++jj) near[jj] = min(near[jj], (unsigned)hc[jj] - gh + 3u);
            }
#pragma unroll
            for (int jj = 0; jj < J; ++jj)
                if (near[jj] <= 6u) doubt |= 1u << jj;
            const int2 ill = *reinterpret_cast<const int2*>(s_SP + n_gamma + 1);
            if (max(wlo - 1, 0) < ill.y && min(whi, n_gamma - 1) >= ill.x) {  // ill-conditioned G in the window
#pragma unroll
                for (int jj = 0; jj < J; ++jj)
                    if (hc[jj] != INT_MAX) doubt |= 1u << jj;
            }
            int kbeg = wlo, kmid = whi;  // lo[jj] lies in [wlo, whi] unless re-decided below
            double fold_corr = 0.0;      // see below; almost always 0
            if (__any_sync(0xffffffffu, doubt != 0)) {  // rare
                int tmin = INT_MAX, tmax = 0;
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    if (doubt & (1u << jj)) {
                        const int j = jbase + jj;
                        lo[jj] = ec_first_unmasked_ref(g_gm, n_gamma, a_eps, a_omc, j, nu[i], models[m].z);
                        // folded pair: the mirror partner takes its own decision.  If it starts
                        // at another index, the shared evaluation (weight W + W_p from lo) is off
                        // by W_p times the integrand steps between the two indices.
                        const double W_p = a_W_p[j];
                        if (W_p != 0.0) {
                            const int lo_p = ec_first_unmasked_ref(g_gm, n_gamma, a_eps, a_omc_p, j, nu[i],
                                                                   models[m].z);
                            const int ka = max(min(lo[jj], lo_p), klo), kb = min(max(lo[jj], lo_p), khi + 1);
                            double d = 0.0;
                            for (int k = ka; k < kb; ++k) {
                                const double2 qr = s_QR[k];
                                d += (s_SP[k] - s_SP[k + 1]) + fma(qr.y, b[jj], qr.x * b2[jj]);
                            }
                            fold_corr += (lo_p < lo[jj]) ? W_p * d : -(W_p * d);
                        }
                    }
                    if (hc[jj] != INT_MAX) { tmin = min(tmin, lo[jj]); tmax = max(tmax, lo[jj]); }
                }
                kbeg = __reduce_min_sync(0xffffffffu, tmin);
                kmid = __reduce_max_sync(0xffffffffu, tmax);
            }
            kbeg = max(kbeg, klo);
            kmid = min(max(kmid, kbeg), khi + 1);
            AGB_CHECK(kbeg >= 0 && kmid >= 0 && kmid <= khi + 1 && khi < n_gamma);  // kbeg > kmid: empty ranges
#pragma unroll
            for (int jj = 0; jj < J; ++jj) AGB_CHECK(lo[jj] >= 0 && lo[jj] <= n_gamma);
            // the b-independent suffix sum at each pair's own first unmasked index seeds its accumulator
            // (both are exactly 0 past the last electron index): the loads are in flight during phase 1
            double acc[J];
#pragma unroll
            for (int jj = 0; jj < J; ++jj) acc[jj] = s_SP[min(max(lo[jj], klo), n_gamma)];
            // phase 1 (a few steps: the pairs of a warp are neighbours in c): lanes join as the
            // loop passes their first unmasked index
            for (int k = kbeg; k < kmid; ++k) {
                const double2 qr = s_QR[k];
#pragma unroll
                for (int jj = 0; jj < J; ++jj)
                    if (k >= lo[jj]) acc[jj] = fma(qr.y, b[jj], fma(qr.x, b2[jj], acc[jj]));
            }
            // phase 2: every live lane is above its gamma_min -> 2 DFMA per integrand point;
            // four gamma steps per trip and no remainder loop: the row carries kRowPad zero entries
            // after the last index, so running past khi adds exact zeros
            for (int k = kmid; k <= khi; k += 4) {
                AGB_CHECK(k >= 0 && k + 3 < n_gamma + kRowPad);
                const double2 q0 = s_QR[k], q1 = s_QR[k + 1], q2 = s_QR[k + 2], q3 = s_QR[k + 3];
#define EC_STEP(q)                                                                        \
    _Pragma("unroll") for (int jj = 0; jj < J; ++jj) acc[jj] = fma(q.x, b2[jj], acc[jj]); \
    _Pragma("unroll") for (int jj = 0; jj < J; ++jj) acc[jj] = fma(q.y, b[jj], acc[jj]);
                EC_STEP(q0) EC_STEP(q1) EC_STEP(q2) EC_STEP(q3)
#undef EC_STEP
            }
            // dead pairs have W = 0
#pragma unroll
            for (int jj = 0; jj < J; ++jj) thread_sum = fma(W[jj], acc[jj], thread_sum);
            thread_sum += fold_corr;
        }
        AGB_CHECK(r >= 0 && r < L_nu);
        s_sum[(size_t)r * (T + 1) + tid] = thread_sum;
    }
    // per warp: sum its 32 columns of every frequency row.  Lane l takes row l % 8 (+8, +16, ...),
    // a quarter of the columns each, then two shuffle steps join the quarters.
    __syncwarp();
    const int n_rows = i1 - i0;
    for (int r0 = 0; r0 < n_rows; r0 += 8) {
        const int r = r0 + (lane & 7), q = lane >> 3;
        double v = 0.0;
        if (r < n_rows) {
            const double* col = s_sum + (size_t)r * (T + 1) + warp * 32 + q * 8;
#pragma unroll
            for (int x = 0; x < 8; ++x) v += col[x];
        }
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (q == 0 && r < n_rows)
            part[(((size_t)m * n_nu + i0 + r) * S + slice) * W_per_block + warp] = v;
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
// Opt-in fast path for the np.trapz integrator (AGNPY_INTEG_TRAPZ_PREFIX): the (mu, phi) sum taken
// *before* the gamma sum.  With np.trapz the integrand of a (nu, model) is a fixed weighted double
// sum, and each point is a polynomial in the pair variable b:
//     sum_j W_j sum_{k >= lo_j} (P_k + Q_k b_j^2 + R_k b_j)
//   = sum_k [ P_k M0(n_k) + Q_k M2(n_k) + R_k M1(n_k) ],      n_k = #{ j : pair j unmasked at gamma_k },
//     M0(n) = sum_{j<n} W_j,  M1(n) = sum_{j<n} W_j b_j,  M2(n) = sum_{j<n} W_j b_j^2
// over the pairs sorted by descending threshold c_j: the unmasked pairs at gamma_k are exactly a
// prefix of that order (pair j is unmasked iff G_k <= c_j).  The three prefix moments do not depend on
// the frequency, so a (nu, model) costs one binary search per gamma point and segment instead of
// one DFMA pair per (gamma, mu, phi) point: O(n_gamma log n_pairs) instead of O(n_gamma n_pairs).
// Same integral, same grids, same masks -- the guard band is the one of the direct kernel: pairs
// whose threshold lies within 3 units of the high word of G_k (>= 1.4e-6 relative) are taken out
// of the prefix and decided one by one with the reference's own formula (kernels.py:36-40), mirror
// partners of a folded phi axis separately -- only the association order of the sums differs
// (measured <= 1e-15 from the direct kernel on the BASELINE configs, tests/test_gpu_prefix.py).
//
// ec_moments_kernel: grid = (segments, models), 1024 threads: exclusive prefix sums of W, W b, W b^2
//                    over each sorted 1024-pair segment + the high words of c (dead pairs: INT_MIN)
// ec_moment_kernel:  grid = (nu / 8, models), 256 threads over the (gamma, nu) items of 8 frequencies:
//                    y, A, G, P, Q, R of kernels.py:61-65 in the operation order of the table
//                    kernels, the search (sorted high words staged in shared memory), the band, a
//                    fixed-order sum per frequency and the prefactor.  No per-(nu, model) rows are
//                    written or read.
// ---------------------------------------------------------------------------------------------
constexpr int kMomStride = kPairSeg + 1;
__global__ void __launch_bounds__(kPairSeg) ec_moments_kernel(const double* __restrict__ ang, int n_pairs,
                                                              int n_seg, double* __restrict__ mom,
                                                              int* __restrict__ hcs) {
    __shared__ double s_w[3][32];
    const int m = blockIdx.y, seg = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const double* row = ang + (size_t)m * kAngCols * n_pairs;
    const int j = seg * kPairSeg + t;
    const bool in = j < n_pairs;
    const double c = in ? row[4 * (size_t)n_pairs + j] : -1.0;
    const bool live = c > 0.0;  // false for NaN
    const double b = live ? row[2 * (size_t)n_pairs + j] : 0.0, W = live ? row[3 * (size_t)n_pairs + j] : 0.0;
    if (in) hcs[(size_t)m * n_pairs + j] = live ? (int)(__double_as_longlong(c) >> 32) : INT_MIN;
    double v[3] = {W, W * b, (W * b) * b};
    double incl[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double x = v[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        incl[q] = x;
        if (lane == 31) s_w[q][warp] = x;
    }
    __syncthreads();
    if (warp < 3) {  // exclusive scan of the 32 warp totals of quantity `warp`
        double tot = s_w[warp][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double y = __shfl_up_sync(0xffffffffu, tot, o);
            if (lane >= o) tot += y;
        }
        // exclusive = the inclusive value of the lane below (never `tot - own`: the moments of one
        // warp can exceed everything before it by 20 decades -- pairs seen almost tail-on have b ~ 1e14)
        const double excl = __shfl_up_sync(0xffffffffu, tot, 1);
        s_w[warp][lane] = lane ? excl : 0.0;
    }
    __syncthreads();
    double* out = mom + ((size_t)m * n_seg + seg) * 3 * kMomStride;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        out[q * kMomStride + t + 1] = incl[q] + s_w[q][warp];
        if (t == 0) out[q * kMomStride] = 0.0;
    }
}

constexpr int kMomNu = 8;  // frequencies per block, at most (one warp each in the final sums)
__global__ void __launch_bounds__(256) ec_moment_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ gamma, const double* __restrict__ gt, const int* __restrict__ hull,
    int n_gamma, const double* __restrict__ ang, int n_pairs, int n_seg, const double* __restrict__ mom,
    const int* __restrict__ hcs, int hc_in_smem, int L, double* __restrict__ sed) {
    extern __shared__ double sh_mom[];  // L x n_k item values, then (optionally) the sorted high words
    const int i0 = blockIdx.x * L, m = blockIdx.y, tid = threadIdx.x;
    const int n_loc = min(L, n_nu - i0);
    const agnpy_model_t& M = models[m];
    const double* g_f = gt + (size_t)m * 5 * n_gamma + n_gamma;  // w_k N_e / gamma_k^2
    const int klo = hull[2 * m], khi = hull[2 * m + 1];
    const int n_k = khi >= klo ? khi - klo + 1 : 0;
    const double* a_eps = ang + (size_t)m * kAngCols * n_pairs;
    const double* a_omc = a_eps + n_pairs;
    const double* a_b = a_omc + n_pairs;
    const double* a_W = a_b + n_pairs;
    const double* a_omc_p = a_W + 2 * (size_t)n_pairs;
    const double* a_W_p = a_omc_p + n_pairs;
    const double* mm = mom + (size_t)m * n_seg * 3 * kMomStride;
    double* s_val = sh_mom;
    const int* hc = hcs + (size_t)m * n_pairs;
    if (hc_in_smem) {
        int* s_hc = reinterpret_cast<int*>(sh_mom + (size_t)L * (n_gamma + 1));
        for (int j = tid; j < n_pairs; j += blockDim.x) s_hc[j] = hc[j];
        hc = s_hc;
        __syncthreads();
    }
    // one item per (gamma index, local frequency)
    for (int item = tid; item < n_k * n_loc; item += blockDim.x) {
        const int q = item / n_k, k = klo + (item - q * n_k);
        // EC uses nu_to_epsilon_prime(nu, z) without the Doppler factor (external_compton.py:123)
        const double eps_s = nu_to_eps(nu[i0 + q], M.z, 1.0);
        const double g = gamma[k];
        const double y = 1.0 - eps_s / g;
        double val = 0.0;
        const double A = y + 1.0 / y;
        const double G = eps_s / ((g * g) * y);
        if (y > 0.0 && G < kFmax) {  // else gamma <= eps_s: G = +inf, no pair is unmasked
            const double wf = g_f[k];
            const double wG = wf * G;
            const double P = wf * A, Q = wG * G, R = -2.0 * wG;
            const bool ill = A > 1e8;  // y < 1e-8: G itself is ill-conditioned, decide every pair exactly
            const int hG = (int)(__double_as_longlong(G) >> 32);
            double M0 = 0.0, M1 = 0.0, M2 = 0.0, fix = 0.0;
            for (int seg = 0; seg < n_seg; ++seg) {
                const int j0 = seg * kPairSeg, cnt = min(kPairSeg, n_pairs - j0);
                const int* h = hc + j0;
                // first position whose high word is <= hG + 3: everything before it is surely unmasked
                int lo = 0, hi = ill ? 0 : cnt;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (h[mid] > hG + 3) lo = mid + 1; else hi = mid;
                }
                const double* ms = mm + (size_t)seg * 3 * kMomStride;
                AGB_CHECK(lo >= 0 && lo <= cnt && cnt <= kPairSeg);
                M0 += ms[lo];
                M1 += ms[kMomStride + lo];
                M2 += ms[2 * kMomStride + lo];
                // guard band (or every live pair of an ill-conditioned entry): the reference's decision
                for (int p = lo; p < cnt && (ill ? h[p] != INT_MIN : h[p] >= hG - 3); ++p) {
                    const int j = j0 + p;
                    const double e = a_eps[j], b = a_b[j], W = a_W[j], W_p = a_W_p[j];
                    const double f = P + fma(Q, b * b, R * b);
                    double root = sqrt(1.0 + 2.0 / (e * eps_s * a_omc[j]));
                    if (g >= eps_s / 2.0 * (1.0 + root)) fix += (W - W_p) * f;
                    if (W_p != 0.0) {
                        root = sqrt(1.0 + 2.0 / (e * eps_s * a_omc_p[j]));
                        if (g >= eps_s / 2.0 * (1.0 + root)) fix += W_p * f;
                    }
                }
            }
            val = fma(Q, M2, fma(R, M1, P * M0)) + fix;
        }
        s_val[(size_t)q * (n_gamma + 1) + (k - klo)] = val;
    }
    __syncthreads();
    // warp q sums the items of local frequency q in a fixed order: lanes stride the gamma axis,
    // then the shuffle tree
    const int lane = tid & 31, warp = tid >> 5;
    for (int q = warp; q < n_loc; q += blockDim.x >> 5) {
        double acc = 0.0;
        for (int x = lane; x < n_k; x += 32) acc += s_val[(size_t)q * (n_gamma + 1) + x];
        acc = warp_sum(acc);
        if (lane == 0) sed[(size_t)m * n_nu + i0 + q] = ec_prefactor(variant, M, nu[i0 + q]) * acc;
    }
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
    if (integ == AGNPY_INTEG_TRAPZ) {
        p.J = p.n_pairs >= 512 ? 4 : p.n_pairs >= 256 ? 2 : 1;
        // polynomial kernel (>= 512 pairs): 128 threads x 5 (or 4) consecutive sorted pairs, whichever
        // pads the table less (2500 folded pairs at the default grids -> 4 slices of 640, 2 % padding).
        // Measured on B200 (scripts/tune_ec.py, profiles/r2_tune_ec.txt): 128 x 5 at 128 registers
        // (4 blocks per SM, 7 frequencies per block) 2.056 ms per 256 models, 64 x 8 2.076, 128 x 4 at
        // 96 registers 2.102; larger blocks (160-256 threads) and 10 pairs per thread lose
        if (p.n_pairs >= 512) {
            p.T = 128;
            const int pad5 = (p.n_pairs + 639) / 640 * 640, pad4 = (p.n_pairs + 511) / 512 * 512;
            p.J = pad5 <= pad4 ? 5 : 4;
        }
    }
    if (const char* e = getenv("AGNPY_B200_EC_PLAN")) {  // development knob: "T,J"
        int t = 0, j = 0;
        if (sscanf(e, "%d,%d", &t, &j) == 2 && t >= 32 && t <= 256 && t % 32 == 0 &&
            (j == 1 || j == 2 || j == 4 || j == 5 || j == 8) && (integ == AGNPY_INTEG_TRAPZ || j == 1)) {
            p.T = t;
            p.J = j;
        }
    }
    p.S = (p.n_pairs + p.T * p.J - 1) / (p.T * p.J);
    return p;
}

// Which kernels a call runs, and the scratch it needs per model (doubles).  One place decides both, so
// that the chunking of large batches (capi.cu) sees the real footprint of the path taken: 1.5 MB per
// model for the polynomial rows at the fit-wrapper grids, 30 KB for a ring torus.
struct EcPath {
    EcPlan p;
    int integ;       // AGNPY_INTEG_TRAPZ / _LOGLOG (flags stripped, prefix mapped to trapz)
    bool prefix;     // prefix-moment evaluation of the np.trapz integral
    bool v5;         // polynomial main kernel
    bool fused_ag;   // small np.trapz table: {A, G} formed inside the main kernel, no table in HBM
    size_t w_nt, w_rows, w_gt, w_ang, w_raw, w_part;  // words per model of each scratch area
    size_t words() const { return w_nt + w_rows + w_gt + w_ang + w_raw + w_part + 2; }
};
static EcPath ec_path(int variant, int n_nu, int n_gamma, int n_mu, int n_phi, int integ_flags) {
    EcPath q;
    const bool fold_req =
        (integ_flags & AGNPY_INTEG_FOLD_PHI) != 0 && getenv("AGNPY_B200_EC_NOFOLD") == nullptr;
    int integ = integ_flags & ~AGNPY_INTEG_FOLD_PHI;
    q.prefix = integ == AGNPY_INTEG_TRAPZ_PREFIX;
    if (q.prefix) integ = AGNPY_INTEG_TRAPZ;  // same tables (trapz weights folded into wf)
    q.integ = integ;
    EcPlan p = ec_plan(variant, n_mu, n_phi, integ);
    // mirror-symmetric phi grid (the caller vouches for it): one table entry per mirror pair, if
    // the folded (mu, phi) table still feeds the polynomial kernel
    if (fold_req && n_phi >= 2 &&
        (variant == AGNPY_EC_ISO_MONO || variant == AGNPY_EC_SS_DISK || variant == AGNPY_EC_BLR)) {
        EcPlan pf = ec_plan(variant, n_mu, (n_phi + 1) / 2, integ);
        // np.trapz: only where the folded table still feeds the polynomial kernel (the small-table
        // kernel does not carry the partner bookkeeping); trapz_loglog and the prefix path: always
        if (q.prefix || integ == AGNPY_INTEG_LOGLOG || pf.n_pairs >= 512) p = pf;
    }
    const bool use_v4 = getenv("AGNPY_B200_EC_V4") != nullptr;  // A/B knob: 4-instruction body
    if (use_v4 && p.J == 5) {  // the small-table kernel is instantiated for 1, 2, 4 and 8 pairs per thread
        p.J = 4;
        p.S = (p.n_pairs + p.T * p.J - 1) / (p.T * p.J);
    }
    // the polynomial path pays off once a (nu, model) has a few hundred pairs; the ring / point
    // source variants (<= n_phi pairs) stay on the small-table kernel
    q.v5 = integ == AGNPY_INTEG_TRAPZ && !q.prefix && !use_v4 && p.n_pairs >= 512;
    // np.trapz with a small pair table (ring torus, point source, coarse grids): no {A, G} table in HBM
    q.fused_ag = integ == AGNPY_INTEG_TRAPZ && !q.prefix && !q.v5 && p.S <= 4;
    q.p = p;
    const size_t n_seg = (p.n_pairs + kPairSeg - 1) / kPairSeg;
    q.w_nt = (q.v5 || q.prefix || q.fused_ag) ? 0 : 2 * (size_t)n_nu * n_gamma;
    q.w_rows = q.v5 ? (size_t)n_nu * ec_row_words(n_gamma)
               : q.prefix ? ((n_seg * 3 * kMomStride + p.n_pairs / 2 + 2 + 1) & ~(size_t)1) : 0;
    q.w_gt = 5 * (size_t)n_gamma;
    q.w_ang = kAngCols * (size_t)p.n_pairs;
    q.w_raw = kRawCols * (size_t)p.n_pairs;
    q.w_part = q.prefix ? 0 : (size_t)n_nu * p.S * p.W();
    return q;
}

size_t ec_workspace_bytes(int variant, int n_models, int n_nu, int n_gamma, int n_mu, int n_phi, int integ) {
    return ec_path(variant, n_nu, n_gamma, n_mu, n_phi, integ).words() * sizeof(double) * (size_t)n_models + 256;
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
    const EcPath q = ec_path(variant, n_nu, n_gamma, n_mu, n_phi, integ);
    const EcPlan p = q.p;
    const bool prefix = q.prefix, v5 = q.v5, fused_ag = q.fused_ag;
    integ = q.integ;
    const size_t M = (size_t)n_models;
    // carve (every area has an even word count except the last ones: nt and rows stay 16-byte aligned,
    // ws comes from cudaMalloc)
    double2* nt = (double2*)ws;
    double* rows = ws + M * q.w_nt;
    double* gt = rows + M * q.w_rows;
    double* ang = gt + M * q.w_gt;
    double* raw = ang + M * q.w_ang;
    double* part = raw + M * q.w_raw;
    int* hull = (int*)(part + M * q.w_part);
    {
        int nB = (p.n_pairs + 127) / 128, nC = (n_gamma + 127) / 128;
        dim3 grid(1 + nB + ((v5 || prefix || fused_ag) ? 0 : nC * n_nu), n_models);
        ec_tables_kernel<<<grid, 128, 0, stream>>>(variant, models, ne_kind, nu, n_nu, gamma, n_gamma,
                                                   mu, n_mu, phi, n_phi, ne_table, integ, p.n_pairs,
                                                   gt, hull, raw, nt);
        note_launch();
        // sorted segments of the pair table
        const int n_seg = (p.n_pairs + kPairSeg - 1) / kPairSeg;
        dim3 grid2(n_seg, n_models);
        ec_tables2_kernel<<<grid2, kTab2Threads, 0, stream>>>(raw, p.n_pairs, n_seg, ang);
        note_launch();
        // polynomial rows of the np.trapz path + (more than one segment) the global pair order
        const int n_rb = v5 ? (n_nu + kRowWarps - 1) / kRowWarps : 0;
        const int n_mb = n_seg > 1 ? (p.n_pairs + 32 * kRowWarps - 1) / (32 * kRowWarps) : 0;
        if (n_rb + n_mb > 0) {
            dim3 grid3(n_rb + n_mb, n_models);
            ec_rows_merge_kernel<<<grid3, 32 * kRowWarps, 0, stream>>>(raw, p.n_pairs, n_seg, ang, models, nu,
                                                                       n_nu, gamma, n_gamma, gt, hull, rows, n_rb);
            note_launch();
        }
    }
    if (prefix) {
        // prefix moments over the sorted segments (the `rows` area is free on this path), then one
        // block per (nu, model)
        const int n_seg = (p.n_pairs + kPairSeg - 1) / kPairSeg;
        int sms_prefix = 148;
        {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms_prefix, cudaDevAttrMultiProcessorCount, dev);
        }
        double* mom = rows;
        int* hcs = (int*)(mom + M * n_seg * 3 * kMomStride);
        dim3 gm(n_seg, n_models);
        ec_moments_kernel<<<gm, kPairSeg, 0, stream>>>(ang, p.n_pairs, n_seg, mom, hcs);
        note_launch();
        // frequencies per block: 8 once that still leaves a few waves of blocks, fewer for small calls
        int L = kMomNu;
        while (L > 1 && (long long)((n_nu + L - 1) / L) * n_models < 4LL * sms_prefix) L >>= 1;
        dim3 grid((n_nu + L - 1) / L, n_models);
        size_t smem = (size_t)L * (n_gamma + 1) * sizeof(double);
        const size_t hc_bytes = (size_t)p.n_pairs * sizeof(int);
        const int hc_in_smem = smem + hc_bytes <= 96 * 1024;  // the sorted high words, if they fit
        if (hc_in_smem) smem += hc_bytes;
        if (smem > 200 * 1024) return fail(AGNPY_ERR_ARG, "n_gamma too large for the EC prefix kernel");
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(ec_moment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        timing_begin(stream);
        ec_moment_kernel<<<grid, 256, smem, stream>>>(variant, models, nu, n_nu, gamma, gt, hull, n_gamma,
                                                      ang, p.n_pairs, n_seg, mom, hcs, hc_in_smem, L, sed);
        timing_end(stream);
        note_launch();
        return check_last("ec_sed (prefix)");
    }
    size_t smem = v5 ? (size_t)ec_row_words(n_gamma) * sizeof(double) + 8
                     : (size_t)n_gamma * sizeof(double) * (integ == AGNPY_INTEG_TRAPZ ? 3 : 7);
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
    if (v5) {
        const int row_words = ec_row_words(n_gamma);
        // frequencies per block: as many as keep >= 8 waves of blocks in flight, at most 8
        int L_nu = 16;
        while (L_nu > 1 && (long long)p.S * ((n_nu + L_nu - 1) / L_nu) * n_models < 8LL * 5 * sms) L_nu >>= 1;
        if (const char* e = getenv("AGNPY_B200_EC_LNU")) {  // development knob
            int v = atoi(e);
            if (v >= 1) L_nu = v;
        }
        const size_t row_bytes = (size_t)row_words * sizeof(double);
        // shared memory per frequency: the row, its parked partial sums and its barrier; blocks per SM
        // as the register budget allows (128 threads: 4 at 5 pairs per thread / 128 registers, 5 at 4 pairs /
        // 96 registers, 3 at 8 pairs; 64 threads x 8 pairs: 6)
        const size_t per_nu = row_bytes + 8 + (size_t)(p.T + 1) * sizeof(double);
        const int blocks_per_sm = p.T >= 128 ? (p.J == 8 ? 3 : p.J == 5 ? 4 : 5) : 6;
        const size_t smem_cap = (size_t)(227 * 1024) / blocks_per_sm - 1024;
        if (L_nu * per_nu > smem_cap) L_nu = (int)std::max<size_t>(1, smem_cap / per_nu);
        dim3 grid(p.S, (n_nu + L_nu - 1) / L_nu, n_models);
        const size_t smem6 = L_nu * per_nu;
        timing_begin(stream);
#define EC_LAUNCH_K(KERNEL)                                                                         \
    do {                                                                                            \
        if (smem6 > 48 * 1024)                                                                      \
            cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6);  \
        KERNEL<<<grid, p.T, smem6, stream>>>(models, nu, n_nu, gt, hull, n_gamma, rows, ang,        \
                                             p.n_pairs, p.S, L_nu, part);                           \
    } while (0)
        if (p.T == 128 && p.J == 5) EC_LAUNCH_K((ec_main_poly_kernel<5, 128, 4>));
        else if (p.T == 128 && p.J == 4) EC_LAUNCH_K((ec_main_poly_kernel<4, 128, 5>));
        else if (p.T == 64 && p.J == 8) EC_LAUNCH_K((ec_main_poly_kernel<8, 64, 6>));
        else if (p.T == 128 && p.J == 8) EC_LAUNCH_K((ec_main_poly_kernel<8, 128, 3>));
        else return fail(AGNPY_ERR_ARG, "ec_sed: unsupported AGNPY_B200_EC_PLAN for the polynomial kernel");
#undef EC_LAUNCH_K
        timing_end(stream);
        note_launch();
        dim3 g2((n_nu + 127) / 128, n_models);
        ec_finish_kernel<<<g2, 128, 0, stream>>>(variant, models, nu, n_nu, part, p.S * p.W(), sed);
        note_launch();
        return check_last("ec_sed");
    }
    dim3 grid(Sg, n_nu, n_models);
    timing_begin(stream);
    const double2* nt_arg = fused_ag ? nullptr : nt;
#define EC_LAUNCH(I, JJ) launch_main<I, JJ>(grid, p.T, smem, stream, models, nu, n_nu, gt, hull, \
                                            n_gamma, iters, nt_arg, ang, p.n_pairs, p.S, part)
    if (integ == AGNPY_INTEG_LOGLOG) EC_LAUNCH(AGNPY_INTEG_LOGLOG, 1);
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
