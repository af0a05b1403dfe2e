// Library runtime: error text, launch counter, stream-keyed workspace cache, FP64-peak probe.
#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "common.cuh"

namespace agb {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int fail(int code, const char* what) {
    g_err = what ? what : "";
    return code;
}

int check_last(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return AGNPY_OK;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return AGNPY_ERR_CUDA;
}

// literal-reference switch (see include/agnpy_b200.h): initial value from AGNPY_B200_STRICT_REFERENCE
static std::atomic<int> g_strict{-1};
bool strict_reference() {
    int v = g_strict.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("AGNPY_B200_STRICT_REFERENCE");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
        g_strict.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}

// One growing scratch buffer per (device, stream).  Work on one stream is ordered, so a
// buffer can be reused by consecutive calls on that stream; growing it synchronises the
// stream first (rare: sizes are stable inside a fit loop).  A buffer that was handed out while its
// stream was capturing is baked into a CUDA graph: it is pinned -- never freed or replaced -- and a
// request that does not fit it fails instead of growing (cudaStreamSynchronize / cudaFree are
// illegal during capture and would invalidate the graph's pointers afterwards).
struct Slot {
    void* ptr = nullptr;
    size_t size = 0;
    bool pinned = false;
};
static std::mutex g_ws_mutex;
static std::map<std::pair<int, cudaStream_t>, Slot> g_ws;

static bool stream_capturing(cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

void* workspace(size_t bytes, cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    const bool capturing = stream_capturing(stream);
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    Slot& s = g_ws[{dev, stream}];
    if (s.size >= bytes) {
        if (capturing) s.pinned = true;
        return s.ptr;
    }
    if (capturing || s.pinned) {
        g_err = capturing ? "workspace would have to grow during stream capture: run the call once "
                            "outside the capture with the same sizes first"
                          : "workspace of this stream is baked into a CUDA graph and cannot grow: use "
                            "another stream for larger calls";
        return nullptr;
    }
    if (s.ptr) {
        cudaStreamSynchronize(stream);
        cudaFree(s.ptr);
        s.ptr = nullptr;
        s.size = 0;
    }
    size_t want = bytes + bytes / 4 + (1u << 20);
    if (cudaMalloc(&s.ptr, want) != cudaSuccess) {
        cudaGetLastError();
        s.ptr = nullptr;
        return nullptr;
    }
    s.size = want;
    return s.ptr;
}

// ---- optional CUDA-event timing of the dominant kernel of a call (bench.py roofline) ---------
// One set of events per device; begin/end are no-ops on a capturing stream (events recorded into a
// graph cannot be queried) and on a device other than the one the events were made on.
struct Timing {
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool pending = false;
};
static std::mutex g_timing_mutex;
static bool g_timing = false;
static std::map<int, Timing> g_timers;
static double g_timed_ms = 0.0;
static long long g_timed_n = 0;

static void timing_collect_locked() {
    for (auto& kv : g_timers) {
        Timing& t = kv.second;
        if (!t.pending) continue;
        float ms = 0.f;
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != kv.first) cudaSetDevice(kv.first);
        if (cudaEventSynchronize(t.t1) == cudaSuccess && cudaEventElapsedTime(&ms, t.t0, t.t1) == cudaSuccess) {
            g_timed_ms += ms;
            g_timed_n += 1;
        } else {
            cudaGetLastError();
        }
        if (cur != kv.first) cudaSetDevice(cur);
        t.pending = false;
    }
}
static Timing* timer_here(cudaStream_t stream) {
    if (stream_capturing(stream)) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    Timing& t = g_timers[dev];
    if (!t.t0 && (cudaEventCreate(&t.t0) != cudaSuccess || cudaEventCreate(&t.t1) != cudaSuccess)) {
        cudaGetLastError();
        t.t0 = t.t1 = nullptr;
        return nullptr;
    }
    return &t;
}
void timing_begin(cudaStream_t stream) {
    if (!g_timing) return;
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    timing_collect_locked();
    Timing* t = timer_here(stream);
    if (t && cudaEventRecord(t->t0, stream) != cudaSuccess) cudaGetLastError();
}
void timing_end(cudaStream_t stream) {
    if (!g_timing) return;
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    Timing* t = timer_here(stream);
    if (!t) return;
    if (cudaEventRecord(t->t1, stream) == cudaSuccess) t->pending = true;
    else cudaGetLastError();
}

// ---- FP64 peak probe: 8 independent DFMA chains per thread, all SMs full -------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

}  // namespace agb

extern "C" {

const char* agnpy_b200_version(void) { return "agnpy_b200 0.2.0 (sm_100a)"; }
const char* agnpy_b200_last_error(void) { return agb::g_err.c_str(); }
long long agnpy_b200_launch_count(void) { return agb::g_launches.load(); }

int agnpy_b200_kernel_timing(int enable) {
    using namespace agb;
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    if (!enable) timing_collect_locked();
    g_timing = enable != 0;
    if (enable) {
        g_timed_ms = 0.0;
        g_timed_n = 0;
        for (auto& kv : g_timers) kv.second.pending = false;
    }
    return AGNPY_OK;
}

int agnpy_b200_kernel_timing_read(double* total_ms, long long* n_launches) {
    using namespace agb;
    std::lock_guard<std::mutex> lock(g_timing_mutex);
    timing_collect_locked();
    *total_ms = g_timed_ms;
    *n_launches = g_timed_n;
    return AGNPY_OK;
}

int agnpy_b200_strict_reference(int enable) {
    using namespace agb;
    int prev = strict_reference() ? 1 : 0;
    if (enable >= 0) g_strict.store(enable ? 1 : 0, std::memory_order_relaxed);
    return prev;
}

int agnpy_b200_fp64_peak(double ms, double* flops_per_s, void* stream_) {
    using namespace agb;
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(AGNPY_ERR_CUDA, "no CUDA device");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out = (double*)workspace(64, stream);
    if (!out) return fail(AGNPY_ERR_NOMEM, "workspace");
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 2000;
    float t = 0.f;
    double best = 0.0;
    double spent = 0.0;
    // warm-up + calibrate, then repeat until `ms` of device time has been spent
    for (int rep = 0; rep < 64; ++rep) {
        cudaEventRecord(e0, stream);
        fp64_peak_kernel<<<blocks, threads, 0, stream>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, stream);
        note_launch();
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        cudaEventElapsedTime(&t, e0, e1);
        double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        if (rep > 0) {
            double rate = flops / (t * 1e-3);
            if (rate > best) best = rate;
            spent += t;
        }
        if (t < 2.0f) iters *= 2;
        if (spent >= ms) break;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (int rc = check_last("fp64_peak")) return rc;
    *flops_per_s = best;
    return AGNPY_OK;
}

}  // extern "C"
