// Development microbenchmark: attainable FP64 issue rate for the EC inner-loop instruction mix.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int J>
__global__ void __launch_bounds__(256) k(double* out, const double* tabs, int n, int reps, double a, double b) {
    extern __shared__ double sh[];
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) sh[i] = tabs[i];
    __syncthreads();
    double acc[J], bb[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { acc[j] = threadIdx.x * 1e-3 + j; bb[j] = 1.0 + 1e-3 * (threadIdx.x + j); }
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {  // uniform-operand DFMA chains (the library's peak probe)
            for (int kk = 0; kk < n; ++kk) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    acc[j] = fma(acc[j], a, b); acc[j] = fma(acc[j], a, b);
                    acc[j] = fma(acc[j], a, b); acc[j] = fma(acc[j], a, b);
                }
            }
        } else if (MODE == 1) {  // all-register DFMA chains
            for (int kk = 0; kk < n; ++kk) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    acc[j] = fma(acc[j], bb[j], bb[(j + 1) % J]); acc[j] = fma(acc[j], bb[j], bb[(j + 1) % J]);
                    acc[j] = fma(acc[j], bb[j], bb[(j + 1) % J]); acc[j] = fma(acc[j], bb[j], bb[(j + 1) % J]);
                }
            }
        } else if (MODE == 2) {  // EC loop body, tables from shared memory
            for (int kk = 0; kk < n; ++kk) {
                double wf = sh[kk], A = sh[n + kk], G = sh[2 * n + kk];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    double t = G * bb[j];
                    double K = fma(t, t - 2.0, A);
                    acc[j] = fma(wf, K, acc[j]);
                }
            }
        } else if (MODE == 3) {  // EC loop body, loop-invariant "table" values in registers
            double wf = a, A = b, G = a * b;
            for (int kk = 0; kk < n; ++kk) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    double t = G * bb[j];
                    double K = fma(t, t - 2.0, A);
                    acc[j] = fma(wf, K, acc[j]);
                }
                G += 1e-9;  // keep the loop from collapsing
            }
        } else if (MODE == 5) {  // v5 body: u = fma(Q,b,R); acc = fma(u,b,acc), {Q,R} one 16-byte smem read
            const double2* qr = reinterpret_cast<const double2*>(sh);
            for (int kk = 0; kk < n; ++kk) {
                double2 v = qr[kk];
#pragma unroll
                for (int j = 0; j < J; ++j) acc[j] = fma(fma(v.x, bb[j], v.y), bb[j], acc[j]);
            }
        } else if (MODE == 6) {  // split body: acc = fma(Q,b^2,acc); acc = fma(R,b,acc)
            const double2* qr = reinterpret_cast<const double2*>(sh);
            double b2[J];
#pragma unroll
            for (int j = 0; j < J; ++j) b2[j] = bb[j] * bb[j];
            for (int kk = 0; kk < n; ++kk) {
                double2 v = qr[kk];
#pragma unroll
                for (int j = 0; j < J; ++j) acc[j] = fma(v.x, b2[j], acc[j]);
#pragma unroll
                for (int j = 0; j < J; ++j) acc[j] = fma(v.y, bb[j], acc[j]);
            }
        } else if (MODE == 7) {  // split body, two accumulator sets
            const double2* qr = reinterpret_cast<const double2*>(sh);
            double b2[J], acc2[J];
#pragma unroll
            for (int j = 0; j < J; ++j) { b2[j] = bb[j] * bb[j]; acc2[j] = 0.0; }
            for (int kk = 0; kk < n; ++kk) {
                double2 v = qr[kk];
#pragma unroll
                for (int j = 0; j < J; ++j) { acc[j] = fma(v.x, b2[j], acc[j]); acc2[j] = fma(v.y, bb[j], acc2[j]); }
            }
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] += acc2[j];
        } else if (MODE == 4) {  // factored form: 2 DFMA per point, tables from shared memory
            for (int kk = 0; kk < n; ++kk) {
                double H = sh[kk], G = sh[2 * n + kk];
#pragma unroll
                for (int j = 0; j < J; ++j) acc[j] = fma(H, fma(G, bb[j], -2.0), acc[j]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) s += acc[j];
    if (s == 123.456) out[0] = s;
}

template <int MODE, int J>
void run(const char* name, int blocks_per_sm, int threads, double fp64_per_iter) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int n = 200, reps = 200;
    double *out, *tabs; cudaMalloc(&out, 64); cudaMalloc(&tabs, 3 * n * 8);
    cudaMemset(tabs, 0, 3 * n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * blocks_per_sm;
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        k<MODE, J><<<blocks, threads, 3 * n * 8>>>(out, tabs, n, reps, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double instr = (double)blocks * threads * reps * n * fp64_per_iter * J;  // thread-level FP64 instr
    double rate = instr / (best * 1e-3);                                       // per second
    double per_clk_smsp = rate / (sms * 4.0) / 1.965e9;                        // lanes/clk/SMSP (peak 16)
    printf("%-44s J=%d blocks/SM=%d thr=%d: %7.3f ms  %6.2f FP64 lane-instr/clk/SMSP (%.1f%% of 16)\n", name, J,
           blocks_per_sm, threads, best, per_clk_smsp, per_clk_smsp / 16 * 100);
    cudaFree(out); cudaFree(tabs);
}

int main() {
    run<0, 8>("uniform-operand DFMA chains", 8, 256, 4);
    run<1, 8>("all-register DFMA chains", 8, 256, 4);
    run<1, 4>("all-register DFMA chains", 8, 256, 4);
    run<2, 4>("EC body (DMUL,DADD,2 DFMA), smem tables", 8, 256, 4);
    run<2, 4>("EC body (DMUL,DADD,2 DFMA), smem tables", 4, 256, 4);
    run<2, 8>("EC body (DMUL,DADD,2 DFMA), smem tables", 4, 256, 4);
    run<2, 8>("EC body (DMUL,DADD,2 DFMA), smem tables", 4, 128, 4);
    run<2, 2>("EC body (DMUL,DADD,2 DFMA), smem tables", 8, 256, 4);
    run<3, 4>("EC body, register 'tables'", 8, 256, 4);
    run<3, 8>("EC body, register 'tables'", 4, 256, 4);
    run<4, 4>("factored body (2 DFMA), smem tables", 8, 256, 2);
    run<4, 8>("factored body (2 DFMA), smem tables", 4, 256, 2);
    run<5, 8>("v5 body (2 DFMA, 3 reg operands), smem", 4, 128, 2);
    run<5, 8>("v5 body (2 DFMA, 3 reg operands), smem", 7, 128, 2);
    run<5, 4>("v5 body (2 DFMA, 3 reg operands), smem", 8, 128, 2);
    run<6, 8>("split body (Q b^2 + R b), smem", 4, 128, 2);
    run<6, 8>("split body (Q b^2 + R b), smem", 7, 128, 2);
    run<6, 4>("split body (Q b^2 + R b), smem", 8, 128, 2);
    run<7, 8>("split body, 2 accumulator sets, smem", 4, 128, 2);
    run<7, 4>("split body, 2 accumulator sets, smem", 8, 128, 2);
    return 0;
}
