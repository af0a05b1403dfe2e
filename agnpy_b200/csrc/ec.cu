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
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ec_pairs_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ mu,
    int n_mu, const double* __restrict__ phi, int n_phi, int n_pairs, double* __restrict__ ang) {
    int m = blockIdx.y;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pairs) return;
    const agnpy_model_t& M = models[m];
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
    double* row = ang + (size_t)m * 4 * n_pairs;
    row[j] = eps;
    row[n_pairs + j] = omc;
    row[2 * n_pairs + j] = 1.0 / (eps * omc);
    row[3 * n_pairs + j] = W;
}

// ---------------------------------------------------------------------------------------------
// main kernel.  grid = (S slices, n_nu, n_models), block = T threads, J pairs per thread.
// shared (trapz):  gamma_k | wf_k = w_k N_e/gamma^2 | A_k | G_k
// shared (loglog): gamma_k | f_k = N_e/gamma^2 | ln f_k | A_k | G_k | 1/(ln g_k - ln g_k+1) | ln(g_k+1/g_k)
// part[m][i][slice] = sum over the slice's pairs of W_j * integral_gamma
// ---------------------------------------------------------------------------------------------
template <int INTEG, int J>
__global__ void __launch_bounds__(256) ec_main_kernel(
    const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ tab, int n_gamma, const double* __restrict__ ang, int n_pairs,
    double* __restrict__ part) {
    extern __shared__ double sh[];
    __shared__ double red[32];
    __shared__ int s_klo, s_khi;
    const int slice = blockIdx.x, i = blockIdx.y, m = blockIdx.z, S = gridDim.x;
    const int T = blockDim.x, tid = threadIdx.x;
    const agnpy_model_t& M = models[m];
    const double* gs = tab + (size_t)m * 3 * n_gamma;
    const double* Ne = gs + n_gamma;
    // EC uses nu_to_epsilon_prime(nu, z) without the Doppler factor (external_compton.py:123)
    const double eps_s = nu_to_eps(nu[i], M.z, 1.0);
    double* s_gm = sh;
    double* s_f = s_gm + n_gamma;
    double* s_A = s_f + n_gamma;
    double* s_G = s_A + n_gamma;
    double* s_lf = s_G + n_gamma;    // loglog only
    double* s_idl = s_lf + n_gamma;  // loglog only
    double* s_lr = s_idl + n_gamma;  // loglog only
    if (tid == 0) { s_klo = n_gamma; s_khi = -1; }
    __syncthreads();
    for (int k = tid; k < n_gamma; k += T) {
        double g = gs[k];
        double f = Ne[k] / (g * g);
        double y = 1.0 - eps_s / g;
        s_gm[k] = g;
        s_A[k] = y + 1.0 / y;
        s_G[k] = eps_s / ((g * g) * y);
        if (INTEG == AGNPY_INTEG_TRAPZ) {
            s_f[k] = f * trapz_weight(gs, k, n_gamma);
        } else {
            s_f[k] = f;
            s_lf[k] = clipped_log(f);
            if (k < n_gamma - 1) {
                double gu = gs[k + 1];
                s_idl[k] = 1.0 / (clipped_log(g) - clipped_log(gu));
                s_lr[k] = clipped_log(gu / g);
            }
        }
        if (f != 0.0) { atomicMin(&s_klo, k); atomicMax(&s_khi, k); }
    }
    __syncthreads();
    const int klo = s_klo, khi = s_khi;  // hull of the non-zero electron distribution

    const double* a_eps = ang + (size_t)m * 4 * n_pairs;
    const double* a_omc = a_eps + n_pairs;
    const double* a_b = a_omc + n_pairs;
    const double* a_W = a_b + n_pairs;
    double b[J], W[J], acc[J];
    int k0[J];
    int kmin = INT_MAX, kmax = 0;
#pragma unroll
    for (int jj = 0; jj < J; ++jj) {
        int j = slice * (T * J) + jj * T + tid;
        bool valid = j < n_pairs;
        acc[jj] = 0.0;
        b[jj] = 0.0;
        W[jj] = 0.0;
        k0[jj] = n_gamma;
        if (valid) {
            b[jj] = a_b[j];
            W[jj] = a_W[j];
            // kernels.py:36-40 in the reference's operation order
            double root = sqrt(1.0 + 2.0 / (a_eps[j] * eps_s * a_omc[j]));
            double gmin = eps_s / 2.0 * (1.0 + root);
            int lo = 0, hi = n_gamma;  // first k with gamma_k >= gmin (NaN -> n_gamma)
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (s_gm[mid] >= gmin) hi = mid; else lo = mid + 1;
            }
            k0[jj] = lo;
            kmin = min(kmin, lo);
            kmax = max(kmax, lo);
        }
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
                    if (k >= k0[jj]) acc[jj] = fma(wf, K, acc[jj]);
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
                int ks = max(k0[jj], klo);
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
        for (int jj = 0; jj < J; ++jj) {
            int j = slice * (T * J) + jj * T + tid;
            if (j < n_pairs) thread_sum = fma(W[jj], acc[jj], thread_sum);
        }
    }
    thread_sum = block_sum(thread_sum, red);
    if (tid == 0) part[((size_t)m * n_nu + i) * S + slice] = thread_sum;
}

// sum the slices in a fixed order and apply the prefactor of each evaluate_sed_flux_*
__global__ void __launch_bounds__(128) ec_finish_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ part, int S, double* __restrict__ sed) {
    int m = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nu) return;
    const agnpy_model_t& M = models[m];
    const double* t = M.tgt;
    const double* p = part + ((size_t)m * n_nu + i) * S;
    double integral = 0.0;
    for (int s = 0; s < S; ++s) integral += p[s];
    double eps_s = nu_to_eps(nu[i], M.z, 1.0);
    double d3 = M.delta_D * M.delta_D * M.delta_D;
    double es2 = eps_s * eps_s, dL2 = M.d_L * M.d_L;
    double pi2 = kPi * kPi, pi3 = kPi * kPi * kPi;
    double num, den;
    switch (variant) {
        case AGNPY_EC_ISO_MONO:  // external_compton.py:133-142
            num = 3.0 * kC * kSigmaT * t[1] * es2 * d3;
            den = 128.0 * pi2 * dL2 * (t[0] * t[0]);
            break;
        case AGNPY_EC_PS_BEHIND_JET:  // :229-240
            num = 3.0 * kSigmaT * t[1] * es2 * d3;
            den = 128.0 * pi2 * dL2 * (t[2] * t[2]) * (t[0] * t[0]);
            break;
        case AGNPY_EC_SS_DISK: {  // :361-373
            double m_dot = t[1] / (t[2] * (kC * kC));
            num = 9.0 * kSigmaT * kG * t[0] * m_dot * es2 * d3;
            den = 512.0 * pi3 * dL2 * (t[5] * t[5] * t[5]);
            break;
        }
        case AGNPY_EC_BLR:  // :476-490
            num = 3.0 * kSigmaT * t[1] * t[0] * es2 * d3;
            den = 512.0 * pi3 * dL2 * (t[2] * t[2]);
            break;
        default: {  // AGNPY_EC_DT :590-600
            double x_re = sqrt(t[3] * t[3] + t[4] * t[4]);
            num = 3.0 * kSigmaT * t[1] * t[0] * es2 * d3;
            den = 256.0 * pi3 * dL2 * (x_re * x_re) * (t[2] * t[2]);
            break;
        }
    }
    sed[(size_t)m * n_nu + i] = num / den * integral;
}

// ---------------------------------------------------------------------------------------------
static int ec_n_pairs(int variant, int n_mu, int n_phi) {
    if (variant == AGNPY_EC_PS_BEHIND_JET) return 1;
    if (variant == AGNPY_EC_DT) return n_phi;
    return n_mu * n_phi;
}

struct EcPlan {
    int n_pairs, T, J, S;
};
static EcPlan ec_plan(int variant, int n_mu, int n_phi, int integ) {
    EcPlan p;
    p.n_pairs = ec_n_pairs(variant, n_mu, n_phi);
    p.T = p.n_pairs <= 32 ? 32 : p.n_pairs <= 64 ? 64 : p.n_pairs <= 128 ? 128 : 256;
    p.J = 1;
    if (integ == AGNPY_INTEG_TRAPZ) p.J = p.n_pairs >= 2048 ? 4 : p.n_pairs >= 512 ? 2 : 1;
    p.S = (p.n_pairs + p.T * p.J - 1) / (p.T * p.J);
    return p;
}

size_t ec_workspace_bytes(int variant, int n_models, int n_nu, int n_gamma, int n_mu, int n_phi) {
    EcPlan p = ec_plan(variant, n_mu, n_phi, AGNPY_INTEG_LOGLOG);  // J=1 -> largest S
    size_t per = 3 * (size_t)n_gamma + 4 * (size_t)p.n_pairs + (size_t)n_nu * p.S;
    return per * sizeof(double) * (size_t)n_models;
}

template <int INTEG, int J>
static void launch_main(dim3 grid, int T, size_t smem, cudaStream_t stream,
                        const agnpy_model_t* models, const double* nu, int n_nu, const double* tab,
                        int n_gamma, const double* ang, int n_pairs, double* part) {
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(ec_main_kernel<INTEG, J>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    ec_main_kernel<INTEG, J><<<grid, T, smem, stream>>>(models, nu, n_nu, tab, n_gamma, ang,
                                                        n_pairs, part);
}

int run_ec(int variant, const agnpy_model_t* models, int n_models, int ne_kind, const double* nu,
           int n_nu, const double* gamma, int n_gamma, const double* mu, int n_mu,
           const double* phi, int n_phi, const double* ne_table, int integ, double* sed,
           double* ws, cudaStream_t stream) {
    EcPlan p = ec_plan(variant, n_mu, n_phi, integ);
    double* tab = ws;
    double* ang = tab + (size_t)n_models * 3 * n_gamma;
    double* part = ang + (size_t)n_models * 4 * p.n_pairs;
    launch_prep_gamma(models, n_models, ne_kind, gamma, n_gamma, ne_table, nullptr, 1, tab, stream);
    {
        dim3 grid((p.n_pairs + 127) / 128, n_models);
        ec_pairs_kernel<<<grid, 128, 0, stream>>>(variant, models, mu, n_mu, phi, n_phi, p.n_pairs, ang);
        note_launch();
    }
    size_t smem = (size_t)n_gamma * sizeof(double) * (integ == AGNPY_INTEG_TRAPZ ? 4 : 7);
    if (smem > 200 * 1024) return fail(AGNPY_ERR_ARG, "n_gamma too large for the EC kernel");
    dim3 grid(p.S, n_nu, n_models);
    timing_begin(stream);
    if (integ == AGNPY_INTEG_LOGLOG)
        launch_main<AGNPY_INTEG_LOGLOG, 1>(grid, p.T, smem, stream, models, nu, n_nu, tab, n_gamma, ang, p.n_pairs, part);
    else if (p.J == 4)
        launch_main<AGNPY_INTEG_TRAPZ, 4>(grid, p.T, smem, stream, models, nu, n_nu, tab, n_gamma, ang, p.n_pairs, part);
    else if (p.J == 2)
        launch_main<AGNPY_INTEG_TRAPZ, 2>(grid, p.T, smem, stream, models, nu, n_nu, tab, n_gamma, ang, p.n_pairs, part);
    else
        launch_main<AGNPY_INTEG_TRAPZ, 1>(grid, p.T, smem, stream, models, nu, n_nu, tab, n_gamma, ang, p.n_pairs, part);
    timing_end(stream);
    note_launch();
    {
        dim3 g2((n_nu + 127) / 128, n_models);
        ec_finish_kernel<<<g2, 128, 0, stream>>>(variant, models, nu, n_nu, part, p.S, sed);
        note_launch();
    }
    return check_last("ec_sed");
}

}  // namespace agb
