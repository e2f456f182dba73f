// scan.cu -- device-wide exclusive prefix sums (row offsets, CSR offsets).
// Two-level: per-block partial sums, a single-block carry scan over the block
// sums, then a final pass that re-reads the input.  HBM traffic 2 reads + 1 write.
#include "common.cuh"

namespace myrio {

static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns exclusive
// prefix, *total = block sum.  wbuf: >= 33 u64 of shared memory.
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t *wbuf, uint64_t *total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned nwarps = (blockDim.x + 31) >> 5;
    uint64_t inc = warp_incl_scan(v);
    if (lane == 31) wbuf[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < nwarps ? wbuf[lane] : 0;
        uint64_t wi = warp_incl_scan(w);
        if (lane < nwarps) wbuf[lane] = wi - w;
        if (lane == 31) wbuf[32] = wi;
    }
    __syncthreads();
    uint64_t res = wbuf[warp] + inc - v;
    *total = wbuf[32];
    __syncthreads();
    return res;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const TIn *__restrict__ in,
                                                                       uint64_t *__restrict__ bsum,
                                                                       size_t n) {
    __shared__ uint64_t wbuf[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += (uint64_t)in[idx];
    }
    uint64_t total;
    block_excl_scan(s, wbuf, &total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single block: in-place exclusive scan of bsum[0..m), writes the grand total to bsum[m]
__global__ void __launch_bounds__(1024) scan_carry_kernel(uint64_t *bsum, size_t m) {
    __shared__ uint64_t wbuf[33];
    uint64_t carry = 0;
    for (size_t base = 0; base < m; base += blockDim.x) {
        size_t idx = base + threadIdx.x;
        uint64_t v = idx < m ? bsum[idx] : 0;
        uint64_t total;
        uint64_t ex = block_excl_scan(v, wbuf, &total);
        if (idx < m) bsum[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) bsum[m] = carry;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_final_kernel(const TIn *__restrict__ in,
                                                                  const uint64_t *__restrict__ bofs,
                                                                  uint64_t *__restrict__ out,
                                                                  size_t n, size_t nblocks) {
    __shared__ uint64_t wbuf[33];
    // thread t owns SCAN_ITEMS consecutive elements so the prefix stays ordered
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + i;
        v[i] = idx < n ? (uint64_t)in[idx] : 0;
        s += v[i];
    }
    uint64_t total;
    uint64_t ex = block_excl_scan(s, wbuf, &total) + bofs[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + i;
        if (idx < n) out[idx] = ex;
        ex += v[i];
    }
    if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) out[n] = bofs[nblocks];
}

template <typename TIn>
static void exclusive_scan_impl(myrio_ctx *ctx, const TIn *in, uint64_t *out, size_t n) {
    if (n == 0) {
        MYRIO_CK(cudaMemsetAsync(out, 0, sizeof(uint64_t), ctx->stream));
        return;
    }
    size_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf<uint64_t> bsum(nblocks + 1);
    MYRIO_LAUNCH(ctx, "scan_block_sums", scan_block_sums_kernel<TIn>, (unsigned)nblocks,
                 SCAN_THREADS, 0, in, bsum.p, n);
    MYRIO_LAUNCH(ctx, "scan_carry", scan_carry_kernel, 1, 1024, 0, bsum.p, nblocks);
    MYRIO_LAUNCH(ctx, "scan_final", scan_final_kernel<TIn>, (unsigned)nblocks, SCAN_THREADS, 0, in,
                 bsum.p, out, n, nblocks);
    sync(ctx);  // bsum is freed on return
}

void exclusive_scan_u32_to_u64(myrio_ctx *ctx, const uint32_t *in, uint64_t *out, size_t n) {
    exclusive_scan_impl<uint32_t>(ctx, in, out, n);
}
void exclusive_scan_u64(myrio_ctx *ctx, const uint64_t *in, uint64_t *out, size_t n) {
    exclusive_scan_impl<uint64_t>(ctx, in, out, n);
}

}  // namespace myrio
