// Development microbenchmark: FP64 pipe behaviour at low occupancy on B200 (how many warps in the
// EC loop body does an SM need to saturate the pipe; DFMA dependent-issue latency).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int J>
__global__ void __launch_bounds__(128) body(double* out, const double2* tabs, int n, int reps, long long* cyc) {
    extern __shared__ double2 qr[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) qr[i] = tabs[i];
    __syncthreads();
    double acc[J], bb[J], b2[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { acc[j] = threadIdx.x * 1e-3 + j; bb[j] = 1.0 + 1e-3 * (threadIdx.x + j); b2[j] = bb[j] * bb[j]; }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        double2 v = qr[0];
        for (int kk = 0; kk < n; ++kk) {
            double2 nx = qr[min(kk + 1, n - 1)];
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = fma(v.x, b2[j], acc[j]);
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = fma(v.y, bb[j], acc[j]);
            v = nx;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) s += acc[j];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void chain(double* out, int n, double a, double b, long long* cyc) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
    long long t1 = clock64();
    if (x == 123.456) out[0] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int J>
void run(int blocks_per_sm, int threads) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int n = 200, reps = 400;
    double* out; double2* tabs; long long* cyc;
    cudaMalloc(&out, 64); cudaMalloc(&tabs, n * 16); cudaMalloc(&cyc, 8); cudaMemset(tabs, 0, n * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        body<J><<<sms * blocks_per_sm, threads, n * 16>>>(out, tabs, n, reps, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double warps_per_sm = blocks_per_sm * threads / 32.0;
    double dfma_per_warp = (double)reps * n * 2 * J;
    double cycles = best * 1e-3 * 1.965e9;                       // whole-kernel time in SM clocks
    double need = warps_per_sm * dfma_per_warp * 0.5;            // 64 FP64 lanes per SM per clock
    printf("J=%d blocks/SM=%d threads=%3d (%4.1f warps/SM): kernel %8.0f cyc (block 0 alone: %8lld) -> pipe %.1f%%\n",
           J, blocks_per_sm, threads, warps_per_sm, cycles, h, need / cycles * 100);
    cudaFree(out); cudaFree(tabs); cudaFree(cyc);
}

int main() {
    double* out; long long* cyc; cudaMalloc(&out, 64); cudaMalloc(&cyc, 8);
    chain<<<1, 32>>>(out, 10000, 0.999, 1e-9, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA chain, 1 warp: %.2f cycles per DFMA\n", h / 40000.0);
    run<8>(1, 32); run<8>(1, 64); run<8>(1, 128); run<8>(2, 128); run<8>(3, 128); run<8>(4, 128);
    run<4>(1, 32); run<4>(1, 128); run<4>(2, 128); run<4>(4, 128); run<4>(8, 128);
    return 0;
}
