// radix_sort.cu -- stable LSD radix sort of (u64 key, u64 payload) pairs, 8-bit digits.
// Used by the index build (K3): postings sorted by k-mer key, leaf order preserved
// inside each key because every pass is stable and the input is in leaf order.
//
// Per pass: (1) per-CTA digit histogram, (2) exclusive scan over [digit][cta],
// (3) scatter with a stable in-CTA rank: every warp owns a contiguous slice of
// the CTA's tile and ranks its items with __match_any_sync; per-warp digit
// counters are prefix-summed across warps.
#include "common.cuh"

namespace myrio {

static constexpr int RS_THREADS = 256;
static constexpr int RS_WARPS = RS_THREADS / 32;
static constexpr int RS_ITEMS = 16;
static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096

// digit of one (key, payload) pair for this pass: key bits [shift, shift+8), or -- for the
// trailing "group" passes -- bits of (payload >> 32) / group_div (the tile of a posting)
__device__ __forceinline__ uint32_t rs_digit(uint64_t key, uint64_t val, int shift, uint32_t group_div) {
    if (group_div == 0) return (uint32_t)(key >> shift) & 255u;
    return (((uint32_t)(val >> 32)) / group_div >> shift) & 255u;
}

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ keys,
                                                             const uint64_t *__restrict__ vals, size_t n,
                                                             int shift, uint32_t group_div,
                                                             uint32_t *__restrict__ hist,
                                                             uint32_t nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        size_t idx = base + (size_t)i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[rs_digit(keys[idx], group_div ? vals[idx] : 0ull, shift, group_div)], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(
    const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals, size_t n, int shift,
    uint32_t group_div, const uint64_t *__restrict__ offs, uint32_t nblocks,
    uint64_t *__restrict__ out_keys, uint64_t *__restrict__ out_vals) {
    __shared__ uint32_t warp_cnt[RS_WARPS][256];
    __shared__ uint64_t gofs[256];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;
    gofs[threadIdx.x] = offs[(size_t)threadIdx.x * nblocks + blockIdx.x];
    __syncthreads();

    const size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)warp * (RS_ITEMS * 32);
    uint64_t k[RS_ITEMS], v[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t idx = wbase + (size_t)i * 32 + lane;
        const bool valid = idx < n;
        k[i] = valid ? keys[idx] : 0;
        v[i] = valid ? vals[idx] : 0;
        const uint32_t d = valid ? rs_digit(k[i], v[i], shift, group_div) : 0xFFFFFFFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t r = 0;
        if (valid) r = warp_cnt[warp][d] + __popc(peers & lt);
        __syncwarp();
        if (valid && (int)lane == __ffs(peers) - 1) warp_cnt[warp][d] += __popc(peers);
        __syncwarp();
        rank[i] = r;
    }
    __syncthreads();
    {  // exclusive prefix across warps for digit = threadIdx.x
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = warp_cnt[w][threadIdx.x];
            warp_cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t idx = wbase + (size_t)i * 32 + lane;
        if (idx < n) {
            const uint32_t d = rs_digit(k[i], v[i], shift, group_div);
            const uint64_t pos = gofs[d] + warp_cnt[warp][d] + rank[i];
            out_keys[pos] = k[i];
            out_vals[pos] = v[i];
        }
    }
}

// Sorts by key bits [0, key_bits) and then -- stable, so the key order survives inside a
// group -- by group = (payload >> 32) / group_div for groups in [0, n_groups) (skipped when
// group_div == 0).  Returns true if the result ended up in the tmp buffers (odd number of
// passes); the caller swaps its buffer roles.
bool radix_sort_pairs_raw(myrio_ctx *ctx, uint64_t *keys, uint64_t *vals, uint64_t *keys_tmp,
                          uint64_t *vals_tmp, size_t n, int key_bits, DevBuf<uint32_t> &hist,
                          DevBuf<uint64_t> &offs, uint32_t group_div, uint64_t n_groups) {
    if (n == 0) return false;
    const int key_passes = (key_bits + 7) / 8;
    int group_passes = 0;
    if (group_div)
        for (uint64_t g = n_groups > 0 ? n_groups - 1 : 0; g > 0; g >>= 8) ++group_passes;
    const int passes = key_passes + group_passes;
    const uint32_t nblocks = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    hist.reserve((size_t)256 * nblocks);
    offs.reserve((size_t)256 * nblocks + 1);
    uint64_t *sk = keys, *sv = vals, *dk = keys_tmp, *dv = vals_tmp;
    for (int p = 0; p < passes; ++p) {
        const bool on_group = p >= key_passes;
        const int shift = 8 * (on_group ? p - key_passes : p);
        const uint32_t gd = on_group ? group_div : 0u;
        MYRIO_LAUNCH(ctx, "k3_radix_hist", rs_hist_kernel, nblocks, RS_THREADS, 0, sk, sv, n, shift, gd,
                     hist.p, nblocks);
        exclusive_scan_u32_to_u64(ctx, hist.p, offs.p, (size_t)256 * nblocks);
        MYRIO_LAUNCH(ctx, "k3_radix_scatter", rs_scatter_kernel, nblocks, RS_THREADS, 0, sk, sv, n,
                     shift, gd, offs.p, nblocks, dk, dv);
        std::swap(sk, dk);
        std::swap(sv, dv);
    }
    return (passes & 1) != 0;
}

}  // namespace myrio
