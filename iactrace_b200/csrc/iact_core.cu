// Library plumbing: error reporting, launch accounting, device probes used as roofline denominators.
#include "iact_common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void iact_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int iact_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return IACT_OK;
    iact_set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return IACT_ERR_CUDA;
}

void iact_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* iact_last_error(void) { return g_err; }
extern "C" int iact_version(void) { return 100; }
extern "C" long long iact_launch_count(void) { return g_launches.load(); }

extern "C" int iact_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { iact_check_cuda(e, "cudaGetDeviceCount"); return -1; }
    return n;
}

// ---------------------------------------------------------------- roofline probes
// FP32 FMA throughput: 8 independent accumulator chains per thread, all SMs saturated.
__global__ void __launch_bounds__(256) probe_fp32_kernel(int iters, float a, float b, float* sink) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
    float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) sink[0] = s;
}

// Shared-memory float atomicAdd throughput with `n_distinct` addresses per warp instruction.
__global__ void __launch_bounds__(256) probe_atoms_kernel(int iters, int n_distinct, float* sink) {
    __shared__ float h[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) h[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int idx = ((lane % n_distinct) * 33 + warp * 7) & 2047;
    for (int i = 0; i < iters; ++i) {
        atomicAdd(&h[idx], 1.0f);
        idx = (idx + 64) & 2047;
    }
    __syncthreads();
    if (h[threadIdx.x] == -1.f) sink[0] = 1.f;
}

static int time_kernel(cudaStream_t st, float* ms, void (*launch)(cudaStream_t, void*), void* ctx) {
    cudaEvent_t e0, e1;
    IACT_CUDA(cudaEventCreate(&e0));
    IACT_CUDA(cudaEventCreate(&e1));
    launch(st, ctx);                       // warm-up
    IACT_CUDA(cudaEventRecord(e0, st));
    launch(st, ctx);
    IACT_CUDA(cudaEventRecord(e1, st));
    IACT_CUDA(cudaEventSynchronize(e1));
    IACT_CUDA(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return iact_check_cuda(cudaGetLastError(), "probe");
}

struct ProbeCtx { int iters, n_distinct, blocks; float* sink; };

extern "C" int iact_probe_fp32(int iters, double* out_flops, void* stream) {
    IACT_REQUIRE(out_flops && iters > 0, "bad arguments");
    int dev = 0, sms = 0;
    IACT_CUDA(cudaGetDevice(&dev));
    IACT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ProbeCtx c{iters, 0, sms * 8, nullptr};
    IACT_CUDA(cudaMalloc(&c.sink, 16));
    float ms = 0.f;
    int rc = time_kernel((cudaStream_t)stream, &ms, [](cudaStream_t st, void* p) {
        ProbeCtx* c = (ProbeCtx*)p;
        probe_fp32_kernel<<<c->blocks, 256, 0, st>>>(c->iters, 1.0000001f, 1e-9f, c->sink);
        iact_count_launch();
    }, &c);
    cudaFree(c.sink);
    if (rc) return rc;
    *out_flops = 2.0 * 64.0 * (double)iters * 256.0 * c.blocks / (ms * 1e-3);
    return IACT_OK;
}

extern "C" int iact_probe_smem_atomics(int iters, int n_distinct, double* out_atomics, void* stream) {
    IACT_REQUIRE(out_atomics && iters > 0 && n_distinct >= 1 && n_distinct <= 32, "bad arguments");
    int dev = 0, sms = 0;
    IACT_CUDA(cudaGetDevice(&dev));
    IACT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ProbeCtx c{iters, n_distinct, sms * 8, nullptr};
    IACT_CUDA(cudaMalloc(&c.sink, 16));
    float ms = 0.f;
    int rc = time_kernel((cudaStream_t)stream, &ms, [](cudaStream_t st, void* p) {
        ProbeCtx* c = (ProbeCtx*)p;
        probe_atoms_kernel<<<c->blocks, 256, 0, st>>>(c->iters, c->n_distinct, c->sink);
        iact_count_launch();
    }, &c);
    cudaFree(c.sink);
    if (rc) return rc;
    *out_atomics = (double)iters * 256.0 * c.blocks / (ms * 1e-3);
    return IACT_OK;
}
