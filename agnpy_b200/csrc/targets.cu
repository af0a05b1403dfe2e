// integral energy densities of the target photon fields, u(r[, blob]), and the
// EBL absorption table lookup -- the cheap 1-D pieces that keep a fit loop entirely on the device.
//
// Reference arithmetic replaced:
//   agnpy/targets/targets.py:202-218 (PointSourceBehindJet.u), 429-466 (SSDisk.u),
//   516-543 (SphericalShellBLR.u), 620-638 (RingDustTorus.u)      [CMB.u is two flops: host side]
//   agnpy/absorption/ebl_absorption.py:59-73 (scipy RegularGridInterpolator, linear, NaN outside)
#include "common.cuh"
#include "kernels.cuh"

namespace agb {

// numpy.linspace(start, stop, n)[i]
__device__ __forceinline__ double linspace_i(double start, double stop, int n, int i) {
    if (n == 1) return start;
    if (i == n - 1) return stop;
    double step = (stop - start) / (double)(n - 1);
    return (double)i * step + start;
}

// one thread per (r, model); the mu integrals (50 / 100 points, np.trapz) run in registers
__global__ void __launch_bounds__(128) target_u_kernel(
    int variant, const agnpy_model_t* __restrict__ models, const double* __restrict__ Gamma,
    const double* __restrict__ r_grid, int n_r, int n_mu, double* __restrict__ u_out) {
    const int m = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_r) return;
    const double* t = models[m].tgt;
    const double r = r_grid[i];
    const bool comoving = Gamma != nullptr;
    const double G = comoving ? Gamma[m] : 1.0;
    const double Beta = comoving ? sqrt(1.0 - 1.0 / (G * G)) : 0.0;  // blob.py:134-136
    double u;
    if (variant == AGNPY_EC_PS_BEHIND_JET) {  // tgt = {eps_0, L_0}
        u = t[1] / (4.0 * kPi * kC * (r * r));
        if (comoving) u = u / ((G * G) * ((1.0 + Beta) * (1.0 + Beta)));
    } else if (variant == AGNPY_EC_DT) {  // tgt = {L_disk, xi_dt, eps_dt, R_dt}
        const double x2 = t[3] * t[3] + r * r;
        const double x = sqrt(x2);
        u = t[1] * t[0] / (4.0 * kPi * kC * x2);
        if (comoving) {
            const double mu = r / x;
            const double f = G * (1.0 - Beta * mu);
            u = u * (f * f);
        }
    } else if (variant == AGNPY_EC_BLR) {  // tgt = {L_disk, xi_line, eps_line, R_line}; mu = linspace(-1,1,50)
        const double R = t[3];
        const double prefactor = t[1] * t[0] / (8.0 * kPi * kC);
        double integral = 0.0, y_prev = 0.0, mu_prev = 0.0;
        for (int k = 0; k < n_mu; ++k) {
            const double mu = linspace_i(-1.0, 1.0, n_mu, k);
            const double x2 = r * r + R * R - 2.0 * r * R * mu;
            double y = 1.0 / x2;
            if (comoving) {
                const double mu_star = sqrt(1.0 - (R * R) / x2 * (1.0 - mu * mu));
                const double f = 1.0 - Beta * mu_star;
                y = y * ((G * G) * (f * f));
            }
            if (k > 0) integral += (mu - mu_prev) * (y + y_prev) / 2.0;
            y_prev = y;
            mu_prev = mu;
        }
        u = prefactor * integral;
    } else {  // AGNPY_EC_SS_DISK  tgt = {M_BH, L_disk, eta, R_in, R_out}; mu grid of targets.py:289-296
        const double M_BH = t[0], L_disk = t[1], eta = t[2];
        const double R_g = kG * M_BH / (kC * kC);
        const double M_8 = M_BH / (1e8 * kMsun);
        const double L_Edd = 1.26 * 1e46 * M_8;
        const double l_Edd = L_disk / L_Edd;
        const double Rin_t = t[3] / R_g, Rout_t = t[4] / R_g, r_t = r / R_g;
        const double mu_min = 1.0 / sqrt(1.0 + (Rout_t / r_t) * (Rout_t / r_t));
        const double mu_max = 1.0 / sqrt(1.0 + (Rin_t / r_t) * (Rin_t / r_t));
        const double prefactor = 3.0 * l_Edd * L_Edd * R_g / (8.0 * kPi * kC * eta * (r * r * r));
        double integral = 0.0, y_prev = 0.0, mu_prev = 0.0;
        for (int k = 0; k < n_mu; ++k) {
            const double mu = linspace_i(mu_min, mu_max, n_mu, k);
            const double mm2 = pow(mu, -2.0) - 1.0;
            const double R_tilde = r_t * sqrt(mm2);
            const double phi_disk = 1.0 - sqrt(Rin_t / R_tilde);
            double y = 1.0 / mu * pow(mm2, -3.0 / 2.0) * phi_disk;
            if (comoving) {
                const double mu_p = (mu - Beta) / (1.0 - Beta * mu);
                const double a = 1.0 - Beta * mu, b = 1.0 + Beta * mu_p;
                y = y * (1.0 / (pow(G, 6.0) * (a * a) * ((b * b) * (b * b))));
            }
            if (k > 0) integral += (mu - mu_prev) * (y + y_prev) / 2.0;
            y_prev = y;
            mu_prev = mu;
        }
        u = prefactor * integral;
    }
    u_out[(size_t)m * n_r + i] = u;
}

int run_target_u(int variant, const agnpy_model_t* models, int n_models, const double* Gamma,
                 const double* r, int n_r, int n_mu, double* u, cudaStream_t stream) {
    dim3 grid((n_r + 127) / 128, n_models);
    target_u_kernel<<<grid, 128, 0, stream>>>(variant, models, Gamma, r, n_r, n_mu, u);
    note_launch();
    return check_last("target_u");
}

// ---------------------------------------------------------------------------------------------
// EBL absorption: bilinear interpolation of values[n_e][n_z] at (E = h nu [keV], z_m), NaN outside
// the table (bounds_error=False, fill_value=nan).  Interval search as scipy's find_indices:
// i = searchsorted(grid, x) - 1 clipped to [0, n-2]; weights (x - g_i)/(g_{i+1} - g_i).
// multiply != 0: out[m][i] *= absorption (apply to a batch of SEDs in place).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int grid_interval(const double* __restrict__ g, int n, double x) {
    int lo = 0, hi = n;  // first index with g[idx] >= x  (numpy.searchsorted side='left')
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (g[mid] < x) lo = mid + 1; else hi = mid;
    }
    int i = lo - 1;
    return i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
}

__global__ void __launch_bounds__(128) ebl_kernel(
    const double* __restrict__ z_models, const double* __restrict__ nu, int n_nu,
    const double* __restrict__ e_ref, int n_e, const double* __restrict__ z_ref, int n_z,
    const double* __restrict__ values, int multiply, double* __restrict__ out) {
    const int m = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nu) return;
    const double E = kH * nu[i] / 1.602176634e-9;  // keV
    const double z = z_models[m];
    double a;
    if (!(E >= e_ref[0] && E <= e_ref[n_e - 1] && z >= z_ref[0] && z <= z_ref[n_z - 1])) {
        a = nan("");
    } else {
        const int ie = grid_interval(e_ref, n_e, E), iz = grid_interval(z_ref, n_z, z);
        const double ye = (E - e_ref[ie]) / (e_ref[ie + 1] - e_ref[ie]);
        const double yz = (z - z_ref[iz]) / (z_ref[iz + 1] - z_ref[iz]);
        const double v00 = values[(size_t)ie * n_z + iz], v01 = values[(size_t)ie * n_z + iz + 1];
        const double v10 = values[(size_t)(ie + 1) * n_z + iz], v11 = values[(size_t)(ie + 1) * n_z + iz + 1];
        a = v00 * (1.0 - ye) * (1.0 - yz) + v01 * (1.0 - ye) * yz + v10 * ye * (1.0 - yz) + v11 * ye * yz;
    }
    const size_t o = (size_t)m * n_nu + i;
    out[o] = multiply ? out[o] * a : a;
}

int run_ebl(int n_models, const double* z, const double* nu, int n_nu, const double* e_ref, int n_e,
            const double* z_ref, int n_z, const double* values, int multiply, double* out,
            cudaStream_t stream) {
    dim3 grid((n_nu + 127) / 128, n_models);
    ebl_kernel<<<grid, 128, 0, stream>>>(z, nu, n_nu, e_ref, n_e, z_ref, n_z, values, multiply, out);
    note_launch();
    return check_last("ebl_absorption");
}

}  // namespace agb
