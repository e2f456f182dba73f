// microbench.cu -- measured denominators for the roofline fractions bench.py reports.
//
// myrio_bench_smem_bandwidth: aggregate shared-memory READ bandwidth of the device, the bound of the K6
// tile kernels (cluster.cu: every multiply-add term reads one column value from shared memory).  One
// 1024-thread CTA per SM (the K6 launch shape) streams a 64 KB shared-memory array with conflict-free
// 16-byte loads (LDS.128: a quarter-warp covers all 32 banks per phase); bytes read / elapsed time.
#include "common.cuh"

namespace myrio {

static constexpr int MB_THREADS = 1024;
static constexpr int MB_VEC = 4096;  // float4 elements = 64 KB

__global__ void __launch_bounds__(MB_THREADS) smem_read_kernel(uint32_t iters, float *__restrict__ sink) {
    extern __shared__ __align__(16) float4 buf[];
    for (int i = threadIdx.x; i < MB_VEC; i += MB_THREADS) buf[i] = make_float4((float)i, 1.0f, 2.0f, 3.0f);
    __syncthreads();
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a, d = a;
    uint32_t idx = threadIdx.x;
    for (uint32_t it = 0; it < iters; ++it) {
        // four independent loads in flight per thread; consecutive threads -> consecutive 16-byte slots
        const float4 v0 = buf[idx & (MB_VEC - 1)];
        const float4 v1 = buf[(idx + MB_THREADS) & (MB_VEC - 1)];
        const float4 v2 = buf[(idx + 2 * MB_THREADS) & (MB_VEC - 1)];
        const float4 v3 = buf[(idx + 3 * MB_THREADS) & (MB_VEC - 1)];
        a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
        b.x += v1.x; b.y += v1.y; b.z += v1.z; b.w += v1.w;
        c.x += v2.x; c.y += v2.y; c.z += v2.z; c.w += v2.w;
        d.x += v3.x; d.y += v3.y; d.z += v3.z; d.w += v3.w;
        idx += 17;  // a different slot every iteration (still one 16-byte slot per thread, conflict-free)
    }
    const float s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    if (s == -1.0f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the loads alive
}

void bench_smem_bandwidth(myrio_ctx *ctx, double *gbps, double *ms_out) {
    const size_t smem = (size_t)MB_VEC * sizeof(float4);
    MYRIO_CK(cudaFuncSetAttribute(smem_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DevBuf<float> sink((size_t)ctx->sm_count * MB_THREADS);
    const uint32_t iters = 20000;
    cudaEvent_t e0, e1;
    MYRIO_CK(cudaEventCreate(&e0));
    MYRIO_CK(cudaEventCreate(&e1));
    double best_ms = 1e30;
    for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
        MYRIO_CK(cudaEventRecord(e0, ctx->stream));
        MYRIO_LAUNCH(ctx, "mb_smem_read", smem_read_kernel, ctx->sm_count, MB_THREADS, smem, iters, sink.p);
        MYRIO_CK(cudaEventRecord(e1, ctx->stream));
        MYRIO_CK(cudaEventSynchronize(e1));
        float ms = 0;
        MYRIO_CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double bytes = (double)ctx->sm_count * MB_THREADS * (double)iters * 4.0 * 16.0;
    if (gbps) *gbps = bytes / (best_ms * 1e-3) / 1e9;
    if (ms_out) *ms_out = best_ms;
}

// ---- SM clock under load, measured on the device: cycles of %clock64 per nanosecond of %globaltimer over a
// ~0.5 ms spin of one warp, enqueued on the context's stream between two steps of a timed region.  No driver or
// NVML call is involved (an NVML query inside a timed region costs 2-36 ms on these boxes and can stall the CUDA
// calls of other ranks: profiles/r02_smi_hiccups.md).
__global__ void clock_probe_kernel(unsigned long long *__restrict__ out) {
    if (threadIdx.x != 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    long long c1 = c0;
    while (c1 - c0 < 1000000) c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[0] = (unsigned long long)(c1 - c0);
    out[1] = t1 - t0;
}

static constexpr uint32_t CLOCK_PROBE_SLOTS = 64;

void bench_clock_probe(myrio_ctx *ctx, uint32_t slot) {
    if (slot >= CLOCK_PROBE_SLOTS) fail(MYRIO_ERR_BAD_ARG, "clock probe slot %u out of range (%u)", slot, CLOCK_PROBE_SLOTS);
    if (ctx->clock_probes.n == 0) {
        ctx->clock_probes.alloc(2 * CLOCK_PROBE_SLOTS);
        MYRIO_CK(cudaMemsetAsync(ctx->clock_probes.p, 0, 2 * CLOCK_PROBE_SLOTS * sizeof(unsigned long long), ctx->stream));
    }
    MYRIO_LAUNCH(ctx, "mb_clock_probe", clock_probe_kernel, 1, 32, 0, ctx->clock_probes.p + 2 * slot);
}

void bench_clock_read(myrio_ctx *ctx, uint32_t n, double *mhz) {
    if (n > CLOCK_PROBE_SLOTS) fail(MYRIO_ERR_BAD_ARG, "at most %u clock probes", CLOCK_PROBE_SLOTS);
    std::vector<unsigned long long> h(2 * (size_t)n, 0ull);
    if (ctx->clock_probes.n && n) d2h(ctx, h.data(), ctx->clock_probes.p, 2 * (size_t)n);
    sync(ctx);
    for (uint32_t i = 0; i < n; ++i) mhz[i] = h[2 * i + 1] ? (double)h[2 * i] * 1e3 / (double)h[2 * i + 1] : 0.0;
}

}  // namespace myrio
