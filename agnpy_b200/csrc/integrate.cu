// K6 as a callable: the reference's trapz_loglog(y, x, axis) for user arrays
// (agnpy/utils/math.py:55-119; tested directly in agnpy/utils/tests/test_utils.py:30-39).
//
// y is [n_rows][n_x] (the host side moves `axis` last), x is [n_x]; every row is integrated with the
// reference's per-interval formula in its literal form -- here nothing is known about y, so the
// power is evaluated like the reference does, np.power(x_up / x_low, m), not through the
// end-point identity the fused kernels use:
//     m   = (ln y_lo - ln y_up) / (ln x_lo - ln x_up)                       :104   (clipped logs, :49-52)
//     I_k = y_lo / (m + 1) (x_up (x_up / x_lo)^m - x_lo)   if |m + 1| > 1e-10  :106-108
//         = x_lo y_lo ln(x_up / x_lo)                      otherwise            :109
//     I_k = 0 where y_lo == 0, y_up == 0 or x_lo == x_up                        :112-117
// and summed in index order like np.add.reduce over a strided axis (:119).
// One warp per row: lanes take intervals k = lane, lane + 32, ...; with `intervals` the per-interval
// values are stored instead of the sum.  Row sums add the 32 lane partials in a fixed shuffle tree
// (differences to numpy's sequential order are of rounding size; rows of <= 33 points are exact).
#include "common.cuh"
#include "kernels.cuh"

namespace agb {

__device__ __forceinline__ double loglog_interval_literal(double x_lo, double x_up, double y_lo,
                                                          double y_up) {
    if (y_lo == 0.0 || y_up == 0.0 || x_lo == x_up) return 0.0;
    double m = (clipped_log(y_lo) - clipped_log(y_up)) / (clipped_log(x_lo) - clipped_log(x_up));
    if (fabs(m + 1.0) > 1e-10) return y_lo / (m + 1.0) * (x_up * pow(x_up / x_lo, m) - x_lo);
    return x_lo * y_lo * clipped_log(x_up / x_lo);
}

__global__ void __launch_bounds__(128) trapz_loglog_kernel(const double* __restrict__ y,
                                                           const double* __restrict__ x, long long n_rows,
                                                           int n_x, int intervals,
                                                           double* __restrict__ out) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const double* yr = y + row * n_x;
    double acc = 0.0;
    for (int k = lane; k < n_x - 1; k += 32) {
        double v = loglog_interval_literal(x[k], x[k + 1], yr[k], yr[k + 1]);
        if (intervals) out[row * (n_x - 1) + k] = v;
        else acc += v;
    }
    if (!intervals) {
        acc = warp_sum(acc);
        if (lane == 0) out[row] = acc;
    }
}

int run_trapz_loglog(const double* y, const double* x, long long n_rows, int n_x, int intervals,
                     double* out, cudaStream_t stream) {
    const int warps = 4;
    const long long blocks = (n_rows + warps - 1) / warps;
    if (blocks > 2147483647LL) return fail(AGNPY_ERR_ARG, "trapz_loglog: too many rows");
    trapz_loglog_kernel<<<(unsigned)blocks, 32 * warps, 0, stream>>>(y, x, n_rows, n_x, intervals, out);
    note_launch();
    return check_last("trapz_loglog");
}

}  // namespace agb
