// cluster.cu -- K6: the reads x centroids / reads x reads similarity kernels and the
// host loop of clustering::cluster (myrio-core/src/clustering.rs:55-270).
//
// At the clustering k (default 6) a k-mer vector has at most 4^k = 4096
// dimensions, so the column side of a similarity tile is kept DENSE in shared
// memory (transposed [key][column]) while every thread walks one sparse row in
// ascending key order: acc_c = acc_c + row_val * col_c[key], separate multiply and
// add.  A key missing from the column contributes +0.0, which leaves an f32
// accumulator unchanged, so the result is bit-identical to the reference's
// sparse merge-join (data/sparse.rs:317-360).  No tensor cores: the sum order is
// part of the result.
//
// File map: row slabs (coalesced interleaved rows) and pair_tile_kernel (plain tile: centroid
// columns, myrio_pairwise); the conflict-free silhouette tile (canonical-key planes, three
// bank-permuted replicas, slab_schedule_kernel, pair_tile_canon_kernel); the order-preserving row
// reductions (silsum_rows / argmax_rows); recentering; cluster_reads = the host loop of cluster(),
// optionally row-sharded over ranks; compute_cluster_centroid / fingerprints; myrio_pairwise.
// Similarity::Overlap takes the merge-join kernel of merge.cu everywhere.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>

#include "cluster.cuh"
#include "comm.cuh"
#include "index.cuh"
#include "kmer_count.cuh"
#include "merge.cuh"

namespace myrio {

// ------------------------------------------------------------------ K6 tile
//
// Rows: the sparse rows of a tile are re-packed once per call into a ROW SLAB: groups of 32
// rows (one per lane), entries interleaved four at a time -- keys[(off + j/4) * 128 + lane * 4 +
// j % 4] -- so that a warp reads 4 entries of each of its 32 rows with one coalesced 8-byte
// (u16 keys) and one 16-byte (f32 values) load per lane.  Short rows are padded with
// (key 0, +0.0): a +0.0 term leaves an f32 accumulator unchanged.
// Columns: CB columns are dense in shared memory as [dim][4]-float planes (one LDS.128 per
// plane and entry); a persistent grid of one 1024-thread CTA per SM walks (column chunk, row
// slice) tasks, every thread accumulating CB scores of its row in ascending key order.
// The tile is stored TRANSPOSED ([column][row], coalesced) for the order-preserving row
// reductions below (silhouette sums, last-maximum argmax), or row-major for myrio_pairwise.
static constexpr int PT_THREADS = 1024;
static constexpr int PT_WARPS = PT_THREADS / 32;

struct RowSlab {
    DevBuf<uint16_t> keys;
    DevBuf<float> vals;
    DevBuf<uint32_t> grp_len;  // per group: entries per row / 4 (padded)
    DevBuf<uint64_t> grp_off;  // per group: first quad of the group
    DevBuf<unsigned long long> d_nnz;
    uint32_t n_rows = 0, n_groups = 0;
    uint64_t nnz = 0;  // entries of the packed rows (without padding)
};

__global__ void __launch_bounds__(256) slab_count_kernel(const uint64_t *__restrict__ r_off,
                                                         const uint32_t *__restrict__ row_ids, uint32_t n_rows,
                                                         uint32_t n_groups, uint32_t *__restrict__ grp_len,
                                                         unsigned long long *__restrict__ nnz) {
    const uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t ri = g * 32 + lane;
    uint32_t len = 0;
    if (ri < n_rows) {
        const uint32_t r = row_ids ? row_ids[ri] : ri;
        len = (uint32_t)(r_off[r + 1] - r_off[r]);
    }
    uint32_t sum = len;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        len = max(len, __shfl_xor_sync(0xffffffffu, len, d));
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
    }
    if (lane == 0) {
        grp_len[g] = (len + 3) >> 2;
        atomicAdd(nnz, (unsigned long long)sum);
    }
}

__global__ void __launch_bounds__(256) slab_fill_kernel(const uint64_t *__restrict__ r_off,
                                                        const uint64_t *__restrict__ r_keys,
                                                        const float *__restrict__ r_vals,
                                                        const uint32_t *__restrict__ row_ids, uint32_t n_rows,
                                                        uint32_t n_groups, const uint32_t *__restrict__ grp_len,
                                                        const uint64_t *__restrict__ grp_off,
                                                        uint16_t *__restrict__ keys, float *__restrict__ vals) {
    const uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t ri = g * 32 + lane;
    uint64_t b = 0;
    uint32_t len = 0;
    if (ri < n_rows) {
        const uint32_t r = row_ids ? row_ids[ri] : ri;
        b = r_off[r];
        len = (uint32_t)(r_off[r + 1] - b);
    }
    const uint32_t quads = grp_len[g];
    const uint64_t base = grp_off[g] * 128ull + lane * 4u;
    for (uint32_t q = 0; q < quads; ++q) {
        uint16_t k4[4];
        float v4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t j = q * 4 + u;
            k4[u] = j < len ? (uint16_t)r_keys[b + j] : (uint16_t)0;
            v4[u] = j < len ? r_vals[b + j] : 0.0f;
        }
        *reinterpret_cast<uint2 *>(keys + base + (uint64_t)q * 128) =
            make_uint2((uint32_t)k4[0] | ((uint32_t)k4[1] << 16), (uint32_t)k4[2] | ((uint32_t)k4[3] << 16));
        *reinterpret_cast<float4 *>(vals + base + (uint64_t)q * 128) = make_float4(v4[0], v4[1], v4[2], v4[3]);
    }
}

static void build_row_slab(myrio_ctx *ctx, const uint64_t *r_off, const uint64_t *r_keys, const float *r_vals,
                           const uint32_t *row_ids, uint32_t n_rows, RowSlab &s) {
    s.n_rows = n_rows;
    s.n_groups = (n_rows + 31) / 32;
    if (s.n_groups == 0) return;
    s.grp_len.alloc(s.n_groups);
    s.grp_off.alloc(s.n_groups + 1);
    s.d_nnz.alloc(1);
    MYRIO_CK(cudaMemsetAsync(s.d_nnz.p, 0, sizeof(unsigned long long), ctx->stream));
    const unsigned grid = (s.n_groups + 7) / 8;
    MYRIO_LAUNCH(ctx, "k6_slab_count", slab_count_kernel, grid, 256, 0, r_off, row_ids, n_rows, s.n_groups,
                 s.grp_len.p, s.d_nnz.p);
    exclusive_scan_u32_to_u64(ctx, s.grp_len.p, s.grp_off.p, s.n_groups);
    uint64_t quads = 0;
    unsigned long long h_nnz = 0;
    d2h(ctx, &quads, s.grp_off.p + s.n_groups, 1);
    d2h(ctx, &h_nnz, s.d_nnz.p, 1);
    sync(ctx);
    s.nnz = h_nnz;
    s.keys.alloc(std::max<uint64_t>(quads, 1) * 128);
    s.vals.alloc(std::max<uint64_t>(quads, 1) * 128);
    MYRIO_LAUNCH(ctx, "k6_slab_fill", slab_fill_kernel, grid, 256, 0, r_off, r_keys, r_vals, row_ids, n_rows,
                 s.n_groups, s.grp_len.p, s.grp_off.p, s.keys.p, s.vals.p);
}

struct PairArgs {
    // rows: slab
    const uint16_t *s_keys;
    const float *s_vals;
    const uint32_t *grp_len;
    const uint64_t *grp_off;
    uint32_t n_rows, n_groups;
    // columns [c_first, c_first + n_cols): sparse (CSR batch + optional id list) or dense [..][dim]
    const uint64_t *c_off;
    const uint64_t *c_keys;
    const float *c_vals;
    const uint32_t *col_ids;
    const float *c_dense;
    uint32_t c_first, n_cols;
    uint32_t dim;
    uint32_t n_slices;  // row slices per column chunk
    int sim;
    float *out;   // transposed: out[c * ld + row]; row-major: out[row * ld + c]   (c relative to c_first)
    uint64_t ld;
};

__device__ __forceinline__ float pk_finite_or_zero(float x) { return (fabsf(x) <= 3.402823466e+38f) ? x : 0.0f; }

__device__ __forceinline__ float pk_sim(int sim, float dot) {
    if (sim == MYRIO_SIM_JACARD_TANIMOTO) return pk_finite_or_zero(__fdiv_rn(dot, __fsub_rn(2.0f, dot)));
    return dot;
}

// CB = 8: two [dim][4] planes; CB = 2 (dim > 4096): one [dim][2] plane
template <int CB, bool TRANSPOSED>
__global__ void __launch_bounds__(PT_THREADS, 1) pair_tile_kernel(PairArgs a) {
    extern __shared__ __align__(16) float col_sm[];
    constexpr int VW = CB >= 4 ? 4 : 2;        // floats per plane entry
    constexpr int NP = CB / VW;                // planes
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_chunks = (a.n_cols + CB - 1) / CB;
    const uint32_t n_tasks = n_chunks * a.n_slices;
    const uint32_t plane = a.dim * VW;
    for (uint32_t i = tid; i < plane * NP; i += PT_THREADS) col_sm[i] = 0.0f;
    uint32_t staged = 0xFFFFFFFFu;

    // slice-major task order: at any time all CTAs walk the SAME slice of the row slab (sized to stay
    // L2-resident), each against its own column chunk
    for (uint32_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const uint32_t slice = task / n_chunks, chunk = task % n_chunks;
        const uint32_t c0 = chunk * CB;
        const uint32_t nc = min((uint32_t)CB, a.n_cols - c0);
        if (chunk != staged) {
            __syncthreads();  // everyone is done with the previous chunk (and with the zero fill)
            if (a.c_dense) {
                for (uint32_t i = tid; i < a.dim * CB; i += PT_THREADS) {
                    const uint32_t c = i / a.dim, key = i % a.dim;
                    col_sm[(c / VW) * plane + key * VW + (c % VW)] =
                        c < nc ? a.c_dense[(size_t)(a.c_first + c0 + c) * a.dim + key] : 0.0f;
                }
            } else {
                if (staged != 0xFFFFFFFFu) {  // un-stage: zero the planes (16-byte stores, ~1k cycles)
                    float4 *z = reinterpret_cast<float4 *>(col_sm);
                    for (uint32_t i = tid; i < plane * NP / 4; i += PT_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    __syncthreads();
                }
                // PT_WARPS / CB warps per column, four independent loads in flight per lane
                constexpr uint32_t WPC = PT_WARPS / CB, LPC = WPC * 32;
                const uint32_t c = warp / WPC, l = (warp % WPC) * 32 + lane;
                if (c < nc) {
                    const uint32_t cr = a.col_ids ? a.col_ids[a.c_first + c0 + c] : (a.c_first + c0 + c);
                    const uint64_t b = a.c_off[cr], e = a.c_off[cr + 1];
                    float *dst = col_sm + (c / VW) * plane + (c % VW);
                    for (uint64_t j = b + l; j < e; j += 4 * LPC) {
                        uint32_t kk[4];
                        float vv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint64_t jj = j + (uint64_t)u * LPC;
                            kk[u] = jj < e ? (uint32_t)a.c_keys[jj] : 0xFFFFFFFFu;
                            vv[u] = jj < e ? a.c_vals[jj] : 0.0f;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (kk[u] != 0xFFFFFFFFu) dst[kk[u] * VW] = vv[u];
                    }
                }
            }
            staged = chunk;
            __syncthreads();
        }
        // row groups of this slice, one per warp at a time
        const uint32_t g_lo = (uint32_t)((uint64_t)a.n_groups * slice / a.n_slices);
        const uint32_t g_hi = (uint32_t)((uint64_t)a.n_groups * (slice + 1) / a.n_slices);
        for (uint32_t g = g_lo + warp; g < g_hi; g += PT_WARPS) {
            const uint32_t quads = a.grp_len[g];
            const uint64_t base = a.grp_off[g] * 128ull + lane * 4u;
            const uint2 *kp = reinterpret_cast<const uint2 *>(a.s_keys + base);
            const float4 *vp = reinterpret_cast<const float4 *>(a.s_vals + base);
            float acc[CB];
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[c] = 0.0f;
            uint2 kq = make_uint2(0, 0);
            float4 vq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (quads) {
                kq = __ldg(kp);
                vq = __ldg(vp);
            }
            for (uint32_t q = 0; q < quads; ++q) {
                const uint2 kc = kq;
                const float4 vc = vq;
                if (q + 1 < quads) {  // next quad in flight while this one is applied (32 lanes x 8 / 16 B apart)
                    kq = __ldg(kp + (size_t)(q + 1) * 32);
                    vq = __ldg(vp + (size_t)(q + 1) * 32);
                }
                const uint32_t keys4[4] = {kc.x & 0xFFFFu, kc.x >> 16, kc.y & 0xFFFFu, kc.y >> 16};
                const float vals4[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {  // ascending key order
                    const float v = vals4[u];
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        if (VW == 4) {
                            const float4 x = *reinterpret_cast<const float4 *>(col_sm + p * plane + keys4[u] * 4);
                            acc[p * 4 + 0] = __fadd_rn(acc[p * 4 + 0], __fmul_rn(v, x.x));
                            acc[p * 4 + 1] = __fadd_rn(acc[p * 4 + 1], __fmul_rn(v, x.y));
                            acc[p * 4 + 2] = __fadd_rn(acc[p * 4 + 2], __fmul_rn(v, x.z));
                            acc[p * 4 + 3] = __fadd_rn(acc[p * 4 + 3], __fmul_rn(v, x.w));
                        } else {
                            const float2 x = *reinterpret_cast<const float2 *>(col_sm + p * plane + keys4[u] * 2);
                            acc[p * 2 + 0] = __fadd_rn(acc[p * 2 + 0], __fmul_rn(v, x.x));
                            acc[p * 2 + 1] = __fadd_rn(acc[p * 2 + 1], __fmul_rn(v, x.y));
                        }
                    }
                }
            }
            const uint32_t ri = g * 32 + lane;
            if (ri < a.n_rows) {
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    if ((uint32_t)c < nc) {
                        const float s = pk_sim(a.sim, acc[c]);
                        if (TRANSPOSED)
                            a.out[(size_t)(c0 + c) * a.ld + ri] = s;
                        else
                            a.out[(size_t)ri * a.ld + c0 + c] = s;
                    }
                }
            }
        }
    }
}

// Silhouette sums (clustering.rs:323-336) over a transposed tile: one thread per row adds
// dist = 1 - sim in COLUMN ORDER (the member order of the reference's iterators), continuing
// the running sums of the previous column pass.
// Rows from swap_row on belong to the SECOND group of columns (two mutually nearest clusters scored in one
// tile): for them the columns from `split` on are the own cluster and the first `split` the neighbour's -- the
// two sums are independent sequences, so only their roles swap.
__global__ void __launch_bounds__(128) silsum_rows_kernel(const float *__restrict__ tile, uint64_t ld,
                                                          uint32_t n_rows, uint32_t c_first, uint32_t n_cols,
                                                          uint32_t split, float *__restrict__ sums,
                                                          uint32_t swap_row) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const bool swp = r >= swap_row;
    float sa = sums[(size_t)r * 2 + (swp ? 1 : 0)], sb = sums[(size_t)r * 2 + (swp ? 0 : 1)];
    uint32_t c = 0;
    for (; c + 8 <= n_cols; c += 8) {
        float s[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = __ldg(tile + (size_t)(c + u) * ld + r);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float d = __fsub_rn(1.0f, s[u]);  // Similarity::dist, similarity.rs:53-56
            if (c_first + c + u < split)
                sa = __fadd_rn(sa, d);
            else
                sb = __fadd_rn(sb, d);
        }
    }
    for (; c < n_cols; ++c) {
        const float d = __fsub_rn(1.0f, __ldg(tile + (size_t)c * ld + r));
        if (c_first + c < split)
            sa = __fadd_rn(sa, d);
        else
            sb = __fadd_rn(sb, d);
    }
    sums[(size_t)r * 2 + (swp ? 1 : 0)] = sa;
    sums[(size_t)r * 2 + (swp ? 0 : 1)] = sb;
}

// argmax over the columns of a transposed tile; max_by_key keeps the LAST maximum.  One 8-byte record per
// row (best similarity bits, best column): a row-sharded run exchanges both with ONE all-gather per iteration.
__global__ void __launch_bounds__(128) argmax_rows_kernel(const float *__restrict__ tile, uint64_t ld,
                                                          uint32_t n_rows, uint32_t n_cols,
                                                          uint2 *__restrict__ best_out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    float best = 0.0f;
    uint32_t best_idx = 0;
    for (uint32_t c = 0; c < n_cols; ++c) {
        const float s = __ldg(tile + (size_t)c * ld + r);
        if (c == 0 || s >= best) {
            best = s;
            best_idx = c;
        }
    }
    best_out[r] = make_uint2(__float_as_uint(best), best_idx);
}

static uint64_t pt_slice_bytes() {  // MYRIO_K6_SLICE_MB overrides (tuning)
    static const uint64_t v = [] {
        const char *e = getenv("MYRIO_K6_SLICE_MB");
        const long mb = e ? atol(e) : 0;
        return (uint64_t)(mb > 0 ? mb : 32) << 20;
    }();
    return v;
}

// rows x columns [c_first, c_first + n_cols) -> tile (transposed [col][ld] or row-major [row][ld])
template <bool TRANSPOSED>
static void launch_pair_tile(myrio_ctx *ctx, PairArgs a, const RowSlab &s, const char *name) {
    if (s.n_rows == 0 || a.n_cols == 0) return;
    a.s_keys = s.keys.p;
    a.s_vals = s.vals.p;
    a.grp_len = s.grp_len.p;
    a.grp_off = s.grp_off.p;
    a.n_rows = s.n_rows;
    a.n_groups = s.n_groups;
    ctx->cluster_stats[0] += (uint64_t)s.n_rows * a.n_cols;  // pair similarities
    ctx->cluster_stats[1] += s.nnz * a.n_cols;               // row entries x columns = multiply-add terms
    const uint32_t cb = a.dim <= 4096 ? 8 : 2;
    const uint32_t n_chunks = (a.n_cols + cb - 1) / cb;
    // row slices: (a) small enough that a slice of the slab stays L2-resident while every CTA streams
    // it once per column chunk, (b) enough (chunk, slice) tasks for two rounds of the persistent grid
    const uint64_t slab_bytes = (uint64_t)s.keys.n * 2 + (uint64_t)s.vals.n * 4;
    uint32_t n_slices = (uint32_t)((slab_bytes + pt_slice_bytes() - 1) / pt_slice_bytes());
    const uint32_t want = 2u * (uint32_t)ctx->sm_count;
    if ((uint64_t)n_chunks * n_slices < want) n_slices = (want + n_chunks - 1) / n_chunks;
    n_slices = std::max(1u, std::min(n_slices, (s.n_groups + PT_WARPS - 1) / PT_WARPS));
    a.n_slices = n_slices;
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)n_chunks * n_slices, (uint64_t)ctx->sm_count);
    const size_t smem = (size_t)a.dim * cb * sizeof(float);
    if (cb == 8) {
        auto kern = pair_tile_kernel<8, TRANSPOSED>;
        MYRIO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MYRIO_LAUNCH(ctx, name, kern, grid, PT_THREADS, smem, a);
    } else {
        auto kern = pair_tile_kernel<2, TRANSPOSED>;
        MYRIO_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MYRIO_LAUNCH(ctx, name, kern, grid, PT_THREADS, smem, a);
    }
}

static constexpr uint64_t PT_TILE_BUDGET = 1ull << 30;  // bytes of similarity tile kept at once

// ------------------------------------------------ K6 tile, conflict-free variant
//
// pair_tile_kernel is bound by shared-memory wavefronts: every lane reads the 16-byte column
// entry of ITS row's next key, and the 8 lanes of an LDS.128 phase collide in the 8 16-byte bank
// groups about 2.1x (profiles/r01_k6_tile_ncu_summary.json).  For vectors that are symmetric
// under reverse complement (every k-mer vector counted on both strands is: count(key) ==
// count(revcomp key), data/myrseq.rs:96-103) the columns only need their CANONICAL keys --
// 2080 of 4096 at k = 6 -- which leaves room for THREE replicas of the two [canonical key][4]
// planes, the key's slot permuted so that its bank group is octal digit 0, digit 1 or
// (digit 0 ^ digit 1) of the canonical index.  The row slab is then SCHEDULED: for every 8-lane
// phase a greedy matcher (one level of augmentation) gives each lane's next entry a replica
// whose bank group is still free in this slot; a lane that finds none waits (a +0.0 pad that
// re-reads a placed lane's address, i.e. a broadcast).  Entries of a row stay in ascending key
// order, so every accumulator sees the same f32 sequence as before.
static constexpr uint32_t PC_REPLICAS = 3;

struct CanonTable {
    DevBuf<uint16_t> idx;  // [4^k]: canonical index | 0x8000 when key > revcomp(key)
    uint32_t n_canon = 0;  // canonical keys
    uint32_t plane_rows = 0;  // n_canon rounded up to a multiple of 64
    uint32_t k = 0;
};

static uint64_t host_revcomp(uint64_t key, uint32_t k) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < k; ++i) {
        r = (r << 2) | (3u - (key & 3u));
        key >>= 2;
    }
    return r;
}

static void build_canon_table(myrio_ctx *ctx, uint32_t k, CanonTable &t) {
    const uint32_t dim = 1u << (2 * k);
    std::vector<uint16_t> h(dim);
    uint32_t n = 0;
    for (uint32_t key = 0; key < dim; ++key)
        if (key <= host_revcomp(key, k)) h[key] = (uint16_t)n++;
    for (uint32_t key = 0; key < dim; ++key) {
        const uint32_t rc = (uint32_t)host_revcomp(key, k);
        if (key > rc) h[key] = (uint16_t)(h[rc] | 0x8000u);
    }
    t.k = k;
    t.n_canon = n;
    t.plane_rows = (n + 63) / 64 * 64;
    t.idx.alloc(dim);
    h2d(ctx, t.idx.p, h.data(), dim);
    sync(ctx);  // h goes out of scope
}

__device__ __forceinline__ uint32_t pc_slot(uint32_t ci, uint32_t r) {
    const uint32_t d0 = ci & 7u, d1 = (ci >> 3) & 7u, hi = ci & ~63u;
    if (r == 0) return ci;
    if (r == 1) return hi | (d0 << 3) | d1;
    return hi | (d1 << 3) | (d0 ^ d1);
}

// One warp per group of 32 rows (lane = row; the 8 lanes of an LDS.128 phase match among
// themselves, the four phases advance in lockstep).  Per slot: three parallel rounds -- every
// waiting lane proposes the bank group of replica 0, then 1, then 2 of its next entry, the
// lowest lane wins a free group -- then one level of augmentation for the lanes still waiting
// (a placed lane holding one of their groups moves to a free alternative).  ~1.12 slots per
// entry of the longest row.  WRITE == false: only the number of slots (grp_len, in quads).
template <bool WRITE>
__global__ void __launch_bounds__(128) slab_schedule_kernel(const uint64_t *__restrict__ r_off,
                                                            const uint64_t *__restrict__ r_keys,
                                                            const float *__restrict__ r_vals,
                                                            const uint32_t *__restrict__ row_ids, uint32_t n_rows,
                                                            uint32_t n_groups, const uint16_t *__restrict__ canon,
                                                            uint32_t plane_rows, uint32_t *__restrict__ grp_len,
                                                            const uint64_t *__restrict__ grp_off,
                                                            uint16_t *__restrict__ keys, float *__restrict__ vals) {
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const uint32_t lane = threadIdx.x & 31, ph = lane >> 3;
    const unsigned phase_mask = 0xFFu << (ph * 8);
    const uint32_t ri = g * 32 + lane;
    uint64_t ptr = 0, end = 0;
    if (ri < n_rows) {
        const uint32_t r = row_ids ? row_ids[ri] : ri;
        ptr = r_off[r];
        end = r_off[r + 1];
    }
    uint64_t base = 0;
    uint32_t total = 0;
    if (WRITE) {
        base = grp_off[g] * 128ull + lane * 4u;
        total = grp_len[g] * 4u;
    }
    auto phase_or = [&](uint32_t v) {  // OR over the 8 lanes of the phase
        v |= __shfl_xor_sync(FULL, v, 1);
        v |= __shfl_xor_sync(FULL, v, 2);
        v |= __shfl_xor_sync(FULL, v, 4);
        return v;
    };
    uint32_t slot = 0;
    for (;; ++slot) {
        const bool has = ptr < end;
        if (!__any_sync(FULL, has)) break;
        const uint32_t ci = has ? (uint32_t)(canon[(uint32_t)r_keys[ptr]] & 0x7FFFu) : 0u;
        const uint32_t g0 = ci & 7u, g1 = (ci >> 3) & 7u, g2 = g0 ^ g1;
        uint32_t taken = 0;  // bank groups in use in my phase
        int choice = -1;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const uint32_t gr = r == 0 ? g0 : (r == 1 ? g1 : g2);
            const bool want = has && choice < 0 && !((taken >> gr) & 1u);
            const unsigned peers = __match_any_sync(FULL, want ? (ph * 8u + gr) : (64u + lane));
            const bool win = want && ((uint32_t)(__ffs(peers) - 1) == lane);
            if (win) choice = r;
            taken |= phase_or(win ? (1u << gr) : 0u);
        }
        // augmentation, one waiting lane of every phase at a time
        unsigned um = __ballot_sync(FULL, has && choice < 0) & phase_mask;
        while (__any_sync(FULL, um != 0u)) {
            const bool active = um != 0u;
            const uint32_t i = active ? (uint32_t)(__ffs(um) - 1) : lane;
            um &= um - 1u;
            const uint32_t i0 = __shfl_sync(FULL, g0, i), i1 = __shfl_sync(FULL, g1, i), i2 = __shfl_sync(FULL, g2, i);
            const uint32_t mine = choice == 0 ? g0 : (choice == 1 ? g1 : g2);
            int alt = -1;
#pragma unroll
            for (int r2 = 2; r2 >= 0; --r2) {
                const uint32_t gr2 = r2 == 0 ? g0 : (r2 == 1 ? g1 : g2);
                if (r2 != choice && !((taken >> gr2) & 1u)) alt = r2;
            }
            const bool cand = active && choice >= 0 && alt >= 0 && (mine == i0 || mine == i1 || mine == i2);
            const unsigned cm = __ballot_sync(FULL, cand) & phase_mask;
            const uint32_t j = cm ? (uint32_t)(__ffs(cm) - 1) : lane;
            const uint32_t jold = __shfl_sync(FULL, mine, j);
            const uint32_t altg = alt == 0 ? g0 : (alt == 1 ? g1 : g2);
            const uint32_t jnew = __shfl_sync(FULL, altg, j);
            if (cm) {
                if (lane == j) choice = alt;
                if (lane == i) choice = (jold == g0) ? 0 : ((jold == g1) ? 1 : 2);
                taken |= 1u << jnew;
            }
        }
        if (WRITE) {
            const uint32_t mine_e = choice >= 0 ? (uint32_t)choice * plane_rows + pc_slot(ci, (uint32_t)choice) : 0u;
            // waiting lanes re-read the address of a placed lane of their phase (broadcast), value +0.0
            const unsigned pm = __ballot_sync(FULL, choice >= 0) & phase_mask;
            const uint32_t src = pm ? (uint32_t)(__ffs(pm) - 1) : lane;
            const uint32_t pad_e = __shfl_sync(FULL, mine_e, src);
            const uint64_t at = base + (uint64_t)(slot >> 2) * 128 + (slot & 3u);
            keys[at] = (uint16_t)(choice >= 0 ? mine_e : pad_e);
            vals[at] = choice >= 0 ? r_vals[ptr] : 0.0f;
        }
        if (choice >= 0) ++ptr;
    }
    if (!WRITE) {
        if (lane == 0) grp_len[g] = (slot + 3) >> 2;
    } else {
        for (; slot < total; ++slot) {  // round the group up to whole quads
            const uint64_t at = base + (uint64_t)(slot >> 2) * 128 + (slot & 3u);
            keys[at] = 0;
            vals[at] = 0.0f;
        }
    }
}

static void build_row_slab_scheduled(myrio_ctx *ctx, const uint64_t *r_off, const uint64_t *r_keys,
                                     const float *r_vals, const uint32_t *row_ids, uint32_t n_rows,
                                     const CanonTable &ct, RowSlab &s) {
    s.n_rows = n_rows;
    s.n_groups = (n_rows + 31) / 32;
    if (s.n_groups == 0) return;
    s.grp_len.alloc(s.n_groups);
    s.grp_off.alloc(s.n_groups + 1);
    s.d_nnz.alloc(1);
    MYRIO_CK(cudaMemsetAsync(s.d_nnz.p, 0, sizeof(unsigned long long), ctx->stream));
    // exact entry count (for the work statistics) with the plain counter kernel; its grp_len is replaced below
    MYRIO_LAUNCH(ctx, "k6_slab_count", slab_count_kernel, (s.n_groups + 7) / 8, 256, 0, r_off, row_ids, n_rows,
                 s.n_groups, s.grp_len.p, s.d_nnz.p);
    const unsigned grid = (s.n_groups + 3) / 4;
    MYRIO_LAUNCH(ctx, "k6_slab_schedule_count", slab_schedule_kernel<false>, grid, 128, 0, r_off, r_keys, r_vals,
                 row_ids, n_rows, s.n_groups, ct.idx.p, ct.plane_rows, s.grp_len.p, nullptr, nullptr, nullptr);
    exclusive_scan_u32_to_u64(ctx, s.grp_len.p, s.grp_off.p, s.n_groups);
    uint64_t quads = 0;
    unsigned long long h_nnz = 0;
    d2h(ctx, &quads, s.grp_off.p + s.n_groups, 1);
    d2h(ctx, &h_nnz, s.d_nnz.p, 1);
    sync(ctx);
    s.nnz = h_nnz;
    s.keys.alloc(std::max<uint64_t>(quads, 1) * 128);
    s.vals.alloc(std::max<uint64_t>(quads, 1) * 128);
    MYRIO_LAUNCH(ctx, "k6_slab_schedule_fill", slab_schedule_kernel<true>, grid, 128, 0, r_off, r_keys, r_vals,
                 row_ids, n_rows, s.n_groups, ct.idx.p, ct.plane_rows, s.grp_len.p, s.grp_off.p, s.keys.p,
                 s.vals.p);
}

struct CanonArgs {
    const uint16_t *s_keys;  // scheduled slab: replica * plane_rows + permuted canonical index
    const float *s_vals;
    const uint32_t *grp_len;
    const uint64_t *grp_off;
    uint32_t n_rows, n_groups;
    const uint64_t *c_off;  // sparse, reverse-complement-symmetric columns
    const uint64_t *c_keys;
    const float *c_vals;
    const uint32_t *col_ids;
    uint32_t c_first, n_cols;
    const uint16_t *canon;
    uint32_t plane_rows;
    uint32_t n_slices;
    int sim;
    float *out;  // transposed: out[c * ld + row]
    uint64_t ld;
    // The first sym_cols columns are the rows themselves, in the same order (a cluster's members against
    // themselves): sim(i, j) == sim(j, i) bit for bit (the products commute, the key order is shared), so a
    // column chunk that lies inside that square only takes the rows at or below its first column and every
    // score is also stored at its mirror position.  Needs c_first == 0 and n_cols >= sym_cols.
    uint32_t sym_cols;
};

// shared memory: plane A (columns 0-3) = PC_REPLICAS x plane_rows float4, then plane B (columns 4-7)
__global__ void __launch_bounds__(PT_THREADS, 1) pair_tile_canon_kernel(CanonArgs a) {
    extern __shared__ __align__(16) float col_sm[];
    constexpr int CB = 8;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_chunks = (a.n_cols + CB - 1) / CB;
    const uint32_t n_tasks = n_chunks * a.n_slices;
    const uint32_t plane_f4 = PC_REPLICAS * a.plane_rows;  // float4 entries per plane
    float4 *sm4 = reinterpret_cast<float4 *>(col_sm);
    const float4 *planeA = sm4, *planeB = sm4 + plane_f4;
    uint32_t staged = 0xFFFFFFFFu;
    for (uint32_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const uint32_t slice = task / n_chunks, chunk = task % n_chunks;
        const uint32_t c0 = chunk * CB;
        const uint32_t nc = min((uint32_t)CB, a.n_cols - c0);
        const bool sym = c0 + CB <= a.sym_cols;  // the chunk lies inside the symmetric square
        uint32_t g_lo = (uint32_t)((uint64_t)a.n_groups * slice / a.n_slices);
        const uint32_t g_hi = (uint32_t)((uint64_t)a.n_groups * (slice + 1) / a.n_slices);
        if (sym) g_lo = max(g_lo, c0 >> 5);  // rows above the chunk come from the mirror stores
        if (g_lo >= g_hi) continue;           // (uniform over the CTA: nothing staged for an empty task)
        if (chunk != staged) {
            __syncthreads();
            for (uint32_t i = tid; i < 2 * plane_f4; i += PT_THREADS) sm4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncthreads();
            constexpr uint32_t WPC = PT_WARPS / CB, LPC = WPC * 32;
            const uint32_t c = warp / WPC, l = (warp % WPC) * 32 + lane;
            if (c < nc) {
                const uint32_t cr = a.col_ids ? a.col_ids[a.c_first + c0 + c] : (a.c_first + c0 + c);
                const uint64_t b = a.c_off[cr], e = a.c_off[cr + 1];
                float *dst = col_sm + (c / 4) * plane_f4 * 4 + (c % 4);
                for (uint64_t j = b + l; j < e; j += 4 * LPC) {
                    uint32_t kk[4];
                    float vv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint64_t jj = j + (uint64_t)u * LPC;
                        kk[u] = jj < e ? (uint32_t)a.canon[(uint32_t)a.c_keys[jj]] : 0x8000u;
                        vv[u] = jj < e ? a.c_vals[jj] : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (!(kk[u] & 0x8000u)) {  // the canonical member of a (key, revcomp key) pair
#pragma unroll
                            for (uint32_t r = 0; r < PC_REPLICAS; ++r)
                                dst[(r * a.plane_rows + pc_slot(kk[u], r)) * 4] = vv[u];
                        }
                }
            }
            staged = chunk;
            __syncthreads();
        }
        for (uint32_t g = g_lo + warp; g < g_hi; g += PT_WARPS) {
            const uint32_t quads = a.grp_len[g];
            const uint64_t base = a.grp_off[g] * 128ull + lane * 4u;
            const uint2 *kp = reinterpret_cast<const uint2 *>(a.s_keys + base);
            const float4 *vp = reinterpret_cast<const float4 *>(a.s_vals + base);
            float acc[CB];
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[c] = 0.0f;
            uint2 kq = make_uint2(0, 0);
            float4 vq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (quads) {
                kq = __ldg(kp);
                vq = __ldg(vp);
            }
            for (uint32_t q = 0; q < quads; ++q) {
                const uint2 kc = kq;
                const float4 vc = vq;
                if (q + 1 < quads) {
                    kq = __ldg(kp + (size_t)(q + 1) * 32);
                    vq = __ldg(vp + (size_t)(q + 1) * 32);
                }
                const uint32_t e4[4] = {kc.x & 0xFFFFu, kc.x >> 16, kc.y & 0xFFFFu, kc.y >> 16};
                const float v4[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {  // ascending key order per row
                    const float v = v4[u];
                    const float4 x = planeA[e4[u]], y = planeB[e4[u]];
                    acc[0] = __fadd_rn(acc[0], __fmul_rn(v, x.x));
                    acc[1] = __fadd_rn(acc[1], __fmul_rn(v, x.y));
                    acc[2] = __fadd_rn(acc[2], __fmul_rn(v, x.z));
                    acc[3] = __fadd_rn(acc[3], __fmul_rn(v, x.w));
                    acc[4] = __fadd_rn(acc[4], __fmul_rn(v, y.x));
                    acc[5] = __fadd_rn(acc[5], __fmul_rn(v, y.y));
                    acc[6] = __fadd_rn(acc[6], __fmul_rn(v, y.z));
                    acc[7] = __fadd_rn(acc[7], __fmul_rn(v, y.w));
                }
            }
            const uint32_t ri = g * 32 + lane;
            if (ri < a.n_rows) {
#pragma unroll
                for (int c = 0; c < CB; ++c)
                    if ((uint32_t)c < nc) a.out[(size_t)(c0 + c) * a.ld + ri] = pk_sim(a.sim, acc[c]);
                if (sym && ri >= c0 + CB && ri < a.sym_cols) {  // mirror: column ri, rows c0 .. c0+7
#pragma unroll
                    for (int c = 0; c < CB; ++c) a.out[(size_t)ri * a.ld + c0 + c] = pk_sim(a.sim, acc[c]);
                }
            }
        }
    }
}

static bool canon_path_enabled() {  // MYRIO_K6_CANON=0 keeps the plain tile kernel (comparison runs)
    static const bool v = [] {
        const char *e = getenv("MYRIO_K6_CANON");
        return !(e && e[0] == '0');
    }();
    return v;
}

static void launch_pair_tile_canon(myrio_ctx *ctx, CanonArgs a, const RowSlab &s, const CanonTable &ct,
                                   const char *name) {
    if (s.n_rows == 0 || a.n_cols == 0) return;
    a.s_keys = s.keys.p;
    a.s_vals = s.vals.p;
    a.grp_len = s.grp_len.p;
    a.grp_off = s.grp_off.p;
    a.n_rows = s.n_rows;
    a.n_groups = s.n_groups;
    a.canon = ct.idx.p;
    a.plane_rows = ct.plane_rows;
    ctx->cluster_stats[0] += (uint64_t)s.n_rows * a.n_cols;  // similarities delivered (mirrored ones included)
    if (a.sym_cols == 0) ctx->cluster_stats[1] += s.nnz * a.n_cols;  // terms: the caller counts a symmetric launch
    const uint32_t n_chunks = (a.n_cols + 7) / 8;
    const uint64_t slab_bytes = (uint64_t)s.keys.n * 2 + (uint64_t)s.vals.n * 4;
    uint32_t n_slices = (uint32_t)((slab_bytes + pt_slice_bytes() - 1) / pt_slice_bytes());
    const uint32_t want = 2u * (uint32_t)ctx->sm_count;
    if ((uint64_t)n_chunks * n_slices < want) n_slices = (want + n_chunks - 1) / n_chunks;
    n_slices = std::max(1u, std::min(n_slices, (s.n_groups + PT_WARPS - 1) / PT_WARPS));
    a.n_slices = n_slices;
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)n_chunks * n_slices, (uint64_t)ctx->sm_count);
    const size_t smem = (size_t)2 * PC_REPLICAS * ct.plane_rows * 16;
    MYRIO_CK(cudaFuncSetAttribute(pair_tile_canon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MYRIO_LAUNCH(ctx, name, pair_tile_canon_kernel, grid, PT_THREADS, smem, a);
}


// ---------------------------------------------------------------- recentering

// one warp per kept read: sums[c][key] += raw count (integer-valued, exact)
__global__ void __launch_bounds__(256) recenter_accumulate_kernel(const uint64_t *__restrict__ r_off,
                                                                  const uint64_t *__restrict__ r_keys,
                                                                  const float *__restrict__ raw,
                                                                  const uint32_t *__restrict__ kept,
                                                                  const uint32_t *__restrict__ target,
                                                                  uint32_t n_kept, uint32_t dim,
                                                                  uint32_t *__restrict__ sums,
                                                                  uint32_t *__restrict__ members) {
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_kept) return;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t r = kept[i], c = target[i];
    const uint64_t b = r_off[r], e = r_off[r + 1];
    for (uint64_t j = b + lane; j < e; j += 32)
        atomicAdd(&sums[(size_t)c * dim + (uint32_t)r_keys[j]], (uint32_t)raw[j]);
    if (lane == 0) atomicAdd(&members[c], 1u);
}

// The same sums with shared-memory privatisation (n_clusters * dim counters fit in shared memory: the reference
// default of a handful of clusters at k = 6): a persistent grid, one warp per kept read at a time, shared-memory
// atomics, then ONE global atomic per non-zero counter and CTA -- ~25x fewer global atomics on 16 k addresses.
__global__ void __launch_bounds__(256) recenter_accumulate_smem_kernel(const uint64_t *__restrict__ r_off,
                                                                       const uint64_t *__restrict__ r_keys,
                                                                       const float *__restrict__ raw,
                                                                       const uint32_t *__restrict__ kept,
                                                                       const uint32_t *__restrict__ target,
                                                                       uint32_t n_kept, uint32_t dim, uint32_t n_clusters,
                                                                       uint32_t *__restrict__ sums,
                                                                       uint32_t *__restrict__ members) {
    extern __shared__ uint32_t rc_smem[];  // [n_clusters * dim] sums | [n_clusters] members
    const uint32_t total = n_clusters * dim + n_clusters;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) rc_smem[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const uint32_t wpb = blockDim.x >> 5;
    for (uint32_t i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n_kept; i += gridDim.x * wpb) {
        const uint32_t r = kept[i], c = target[i];
        const uint64_t b = r_off[r], e = r_off[r + 1];
        for (uint64_t j = b + lane; j < e; j += 32)
            atomicAdd(&rc_smem[c * dim + (uint32_t)r_keys[j]], (uint32_t)raw[j]);
        if (lane == 0) atomicAdd(&rc_smem[n_clusters * dim + c], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_clusters * dim; i += blockDim.x)
        if (rc_smem[i]) atomicAdd(&sums[i], rc_smem[i]);
    for (uint32_t c = threadIdx.x; c < n_clusters; c += blockDim.x)
        if (rc_smem[n_clusters * dim + c]) atomicAdd(&members[c], rc_smem[n_clusters * dim + c]);
}

// one CTA per cluster: centroid = normalize_l2(sum of raw member counts)
// (clustering.rs:44-53, 288-295); empty clusters keep their centroid.  The squares are formed in parallel,
// their sum runs sequentially in key order on one thread (data/sparse.rs:366: a zero square leaves it unchanged).
__global__ void __launch_bounds__(256) recenter_finalize_kernel(const uint32_t *__restrict__ sums,
                                                                const uint32_t *__restrict__ members,
                                                                uint32_t n_clusters, uint32_t dim,
                                                                float *__restrict__ cen) {
    extern __shared__ float rf_sq[];  // [dim]
    __shared__ float s_m;
    const uint32_t c = blockIdx.x;
    if (c >= n_clusters || members[c] == 0) return;
    const uint32_t *s = sums + (size_t)c * dim;
    float *o = cen + (size_t)c * dim;
    for (uint32_t k = threadIdx.x; k < dim; k += blockDim.x) {
        const float v = (float)s[k];
        rf_sq[k] = __fmul_rn(v, v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float nrm = 0.0f;
#pragma unroll 8
        for (uint32_t k = 0; k < dim; ++k) nrm = __fadd_rn(nrm, rf_sq[k]);
        s_m = __fsqrt_rn(nrm);
    }
    __syncthreads();
    const float m = s_m;
    for (uint32_t k = threadIdx.x; k < dim; k += blockDim.x) {
        const uint32_t u = s[k];
        o[k] = u ? __fdiv_rn((float)u, m) : 0.0f;
    }
}

// ----------------------------------------------------------------- host side

static void host_normalize(std::vector<float> &v) {  // data/sparse.rs:366-383
    float s = 0.0f;
    for (float x : v) {
        const float sq = x * x;
        s = s + sq;
    }
    const float m = sqrtf(s);
    for (float &x : v) x = x / m;
}

static float host_dense_dot(const float *a, const float *b, uint32_t dim) {
    float acc = 0.0f;
    for (uint32_t k = 0; k < dim; ++k) {
        const float p = a[k] * b[k];
        acc = acc + p;
    }
    return acc;
}

static float host_sim(int sim, float dot) {
    if (sim == MYRIO_SIM_JACARD_TANIMOTO) {
        const float s = dot / (2.0f - dot);
        return std::isfinite(s) ? s : 0.0f;
    }
    return dot;
}

// similarity.rs:172-182 over dense vectors: a key absent from both contributes +0.0 to both sums
static float host_dense_overlap(const float *a, const float *b, uint32_t dim) {
    float sum_min = 0.0f, sum_max = 0.0f;
    for (uint32_t k = 0; k < dim; ++k) {
        sum_min = sum_min + fminf(a[k], b[k]);
        sum_max = sum_max + fmaxf(a[k], b[k]);
    }
    const float s = sum_min / sum_max;
    return std::isfinite(s) ? s : 0.0f;
}

static float host_dense_sim(int sim, const float *a, const float *b, uint32_t dim) {
    if (sim == MYRIO_SIM_OVERLAP) return host_dense_overlap(a, b, dim);
    return host_sim(sim, host_dense_dot(a, b, dim));
}

// dense centroid rows -> CSR with a fixed capacity of `dim` entries per row (one warp per row):
// the merge-join kernel (Overlap) needs sparse columns
__global__ void __launch_bounds__(32) dense_to_csr_kernel(const float *__restrict__ dense, uint32_t dim,
                                                          uint64_t *__restrict__ off, uint64_t *__restrict__ end,
                                                          uint64_t *__restrict__ keys, float *__restrict__ vals) {
    const uint32_t c = blockIdx.x, lane = threadIdx.x;
    const float *src = dense + (size_t)c * dim;
    const uint64_t base = (uint64_t)c * dim;
    uint32_t n = 0;
    for (uint32_t k0 = 0; k0 < dim; k0 += 32) {
        const float v = src[k0 + lane];
        const unsigned m = __ballot_sync(0xffffffffu, v != 0.0f);
        if (v != 0.0f) {
            const uint32_t pos = n + __popc(m & ((1u << lane) - 1u));
            keys[base + pos] = k0 + lane;
            vals[base + pos] = v;
        }
        n += __popc(m);
    }
    if (lane == 0) {
        off[c] = base;
        end[c] = base + n;
    }
}

static uint64_t max_key_of(myrio_ctx *ctx, const myrio_svecs *v);

// `sh` (optional): row-sharded run.  Every rank holds ALL reads and runs the same host
// loop; the similarity tiles -- reads x centroids per iteration, members x (members + neighbours) for
// the silhouette -- are computed for the rank's contiguous share of the ROWS only and the per-row
// results (best similarity + target, dots, silhouette sums) are all-gathered through sh->allgather,
// after which every rank continues with identical data.  Recentering is replicated (cheap, exact).
void cluster_reads(myrio_ctx *ctx, const myrio_seqs *reads, const myrio_cluster_params *p,
                   const myrio_svecs *init, const myrio_shard *sh, int32_t *assign, uint32_t *n_iters,
                   float *silhouette) {
    const uint32_t W = sh ? sh->world : 1, R = sh ? sh->rank : 0;
    // exchange of the per-row results: the caller's callback (host buffers), or -- allgather == NULL -- the
    // context's own NCCL communicator (myrio_comm_init), device to device on the context's stream
    const bool use_nccl = sh && W > 1 && !sh->allgather;
    if (W == 0 || R >= W) fail(MYRIO_ERR_BAD_ARG, "bad shard description");
    if (use_nccl && (!ctx->comm || ctx->comm_world != W || ctx->comm_rank != R))
        fail(MYRIO_ERR_BAD_ARG, "myrio_cluster_sharded without a callback needs myrio_comm_init with the same rank/world");
    // contiguous share of n rows: rank r owns [r * chunk, min(n, (r + 1) * chunk))
    auto share = [&](uint32_t n, uint32_t &lo, uint32_t &cnt, uint32_t &chunk) {
        chunk = (n + W - 1) / W;
        lo = std::min(n, R * chunk);
        cnt = std::min(n, lo + chunk) - lo;
    };
    // buf holds W chunks of bytes_per_rank; this rank's chunk is filled, the others arrive
    const bool timing = getenv("MYRIO_CLUSTER_TIMING") != nullptr;
    double t_exchange = 0.0;
    const auto t_begin = std::chrono::steady_clock::now();
    auto seconds_since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count();
    };
    auto exchange = [&](void *buf, uint64_t bytes_per_rank) {
        if (W == 1 || bytes_per_rank == 0) return;
        const auto t0 = std::chrono::steady_clock::now();
        const int32_t rc = sh->allgather(sh->user, buf, bytes_per_rank);
        t_exchange += seconds_since(t0);
        if (rc != 0) fail(MYRIO_ERR_BAD_ARG, "allgather callback failed (%d)", rc);
    };
    // per-row results of all ranks: this rank's `cnt` elements (device) -> W chunks of `chunk` elements (host)
    DevBuf<uint8_t> gbuf;
    auto gather_rows = [&](const void *d_local, size_t elem, uint32_t lo, uint32_t cnt, uint32_t chunk, void *host) {
        uint8_t *h = static_cast<uint8_t *>(host);
        if (use_nccl) {
            const auto t0 = std::chrono::steady_clock::now();
            gbuf.reserve((size_t)W * chunk * elem);
            if (cnt)
                MYRIO_CK(cudaMemcpyAsync(gbuf.p + (size_t)R * chunk * elem, d_local, (size_t)cnt * elem,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
            comm_allgather_inplace(ctx, gbuf.p, (uint64_t)chunk * elem);
            d2h(ctx, h, gbuf.p, (size_t)W * chunk * elem);
            sync(ctx);
            t_exchange += seconds_since(t0);
            return;
        }
        if (cnt) d2h(ctx, h + (size_t)lo * elem, static_cast<const uint8_t *>(d_local), (size_t)cnt * elem);
        sync(ctx);
        exchange(h, (uint64_t)chunk * elem);
    };
    struct TimingReport {
        bool on;
        uint32_t rank;
        const double &ex;
        std::chrono::steady_clock::time_point t0;
        std::vector<std::pair<const char *, double>> marks;
        ~TimingReport() {
            if (!on) return;
            std::fprintf(stderr, "[myrio_cluster rank %u] total %.1f ms, exchange %.1f ms", rank,
                         std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3, ex * 1e3);
            for (auto &m : marks) std::fprintf(stderr, ", %s @%.1f", m.first, m.second * 1e3);
            std::fprintf(stderr, "\n");
        }
    } report{timing, R, t_exchange, t_begin, {}};
    auto mark = [&](const char *what) {
        if (timing) report.marks.emplace_back(what, seconds_since(t_begin));
    };
    const uint64_t n_init = init ? init->n_rows : 0;
    if (n_init == 0 && p->expected_nb_of_clusters == 0)
        fail(MYRIO_ERR_BAD_ARG, "Ensure `initial_centroids` is not empty or `expected_nb_of_clusters` > 0");
    if (p->k < 2 || p->k > 32) fail(MYRIO_ERR_INVALID_K, "k=%u: only k in 2..32 is supported", p->k);
    if (p->k > 7)
        fail(MYRIO_ERR_UNSUPPORTED, "GPU clustering keeps 4^k dense columns in shared memory: k <= 7 (got %u)", p->k);
    if (p->similarity != MYRIO_SIM_COSINE && p->similarity != MYRIO_SIM_JACARD_TANIMOTO &&
        p->similarity != MYRIO_SIM_OVERLAP)
        fail(MYRIO_ERR_BAD_ARG, "unknown similarity %d", p->similarity);
    if (init && init->vtype != MYRIO_VAL_F32) fail(MYRIO_ERR_BAD_ARG, "initial centroids must be f32");
    const uint64_t N = reads->n;
    const uint32_t dim = 1u << (2 * p->k);
    const int sim = p->similarity;
    if (n_iters) *n_iters = 0;
    ctx->cluster_stats[0] = ctx->cluster_stats[1] = 0;

    // Step 1 (clustering.rs:78-90)
    std::unique_ptr<myrio_svecs> counts(count_reads(ctx, reads, p->k, p->t1));
    std::vector<uint64_t> hck(N);
    d2h(ctx, hck.data(), counts->aux_u64.p, N);
    sync(ctx);
    std::vector<uint32_t> kept;
    for (uint64_t r = 0; r < N; ++r) {
        if (silhouette) silhouette[r] = NAN;
        if (hck[r] > p->t2) {
            kept.push_back((uint32_t)r);
            assign[r] = 0;
        } else {
            assign[r] = -1;
        }
    }
    const uint32_t nk = (uint32_t)kept.size();
    if (nk == 0) fail(MYRIO_ERR_EMPTY, "no read has more than t2=%llu high-confidence k-mers", (unsigned long long)p->t2);
    std::unique_ptr<myrio_svecs> norm(svecs_normalize(ctx, counts.get()));  // :93-94
    DevBuf<uint32_t> d_kept(nk);
    h2d(ctx, d_kept.p, kept.data(), nk);

    // Step 2 (:99-129): centroids live as dense rows
    const uint64_t cmax = std::max<uint64_t>(std::max<uint64_t>(n_init, p->expected_nb_of_clusters), 1);
    DevBuf<float> cen(cmax * dim);
    MYRIO_CK(cudaMemsetAsync(cen.p, 0, cmax * dim * sizeof(float), ctx->stream));
    uint32_t nc = 0;
    std::vector<float> dense(dim);
    if (n_init) {
        std::vector<uint64_t> ro(n_init + 1), ik(init->nnz);
        std::vector<float> iv(init->nnz);
        d2h(ctx, ro.data(), init->row_off.p, n_init + 1);
        d2h(ctx, ik.data(), init->keys.p, init->nnz);
        d2h(ctx, iv.data(), init->vals_f32.p, init->nnz);
        sync(ctx);
        for (uint64_t c = 0; c < n_init; ++c) {
            std::fill(dense.begin(), dense.end(), 0.0f);
            for (uint64_t j = ro[c]; j < ro[c + 1]; ++j) {
                if (ik[j] >= dim) fail(MYRIO_ERR_BAD_ARG, "initial centroid key out of range for k=%u", p->k);
                dense[ik[j]] = iv[j];
            }
            h2d(ctx, cen.p + (size_t)nc * dim, dense.data(), dim);
            sync(ctx);
            ++nc;
        }
    }
    // Cluster::from_normalized_centroid (:281-286): clone the normalised read and normalise AGAIN
    auto push_centroid_from_read = [&](uint32_t r) {
        uint64_t be[2];
        d2h(ctx, be, norm->row_off.p + r, 2);
        sync(ctx);
        const uint64_t n = be[1] - be[0];
        std::vector<uint64_t> kk(n);
        std::vector<float> vv(n);
        d2h(ctx, kk.data(), norm->keys.p + be[0], n);
        d2h(ctx, vv.data(), norm->vals_f32.p + be[0], n);
        sync(ctx);
        host_normalize(vv);
        std::fill(dense.begin(), dense.end(), 0.0f);
        for (uint64_t j = 0; j < n; ++j) dense[kk[j]] = vv[j];
        h2d(ctx, cen.p + (size_t)nc * dim, dense.data(), dim);
        sync(ctx);
        ++nc;
    };
    uint32_t max_idx = 0;  // :106-108 max_by_key -> last maximum
    uint64_t max_hck = 0;
    for (uint32_t i = 0; i < nk; ++i)
        if (i == 0 || hck[kept[i]] >= max_hck) {
            max_hck = hck[kept[i]];
            max_idx = i;
        }
    const float max_nb_hck = (float)max_hck;
    if (nc == 0) push_centroid_from_read(kept[max_idx]);

    mark("counted+init");
    const bool merge_path = sim == MYRIO_SIM_OVERLAP;  // union-of-keys sums: merge-join kernel
    uint32_t klo, kcnt, kchunk;  // this rank's share of the kept reads
    share(nk, klo, kcnt, kchunk);
    RowSlab slab;  // this rank's kept reads (normalised), packed once for every reads x centroids pass
    if (!merge_path) build_row_slab(ctx, norm->row_off.p, norm->keys.p, norm->vals_f32.p, d_kept.p + klo, kcnt, slab);
    PairArgs pa{};
    pa.c_dense = cen.p;
    pa.dim = dim;
    pa.sim = sim;
    DevBuf<uint64_t> cs_off, cs_end, cs_keys;
    DevBuf<float> cs_vals;
    if (merge_path) {
        cs_off.alloc(cmax);
        cs_end.alloc(cmax);
        cs_keys.alloc(cmax * dim);
        cs_vals.alloc(cmax * dim);
    }
    // kept reads x current centroids -> tile (pa.out / pa.ld / pa.n_cols), transposed or row-major
    auto centroid_tile = [&](bool transposed, const char *name) {
        if (!merge_path) {
            if (transposed)
                launch_pair_tile<true>(ctx, pa, slab, name);
            else
                launch_pair_tile<false>(ctx, pa, slab, name);
            return;
        }
        MYRIO_LAUNCH(ctx, "k6_dense_to_csr", dense_to_csr_kernel, pa.n_cols, 32, 0, cen.p, dim, cs_off.p, cs_end.p,
                     cs_keys.p, cs_vals.p);
        MergeArgs m{};
        m.r_off = norm->row_off.p;
        m.r_keys = norm->keys.p;
        m.r_vals = norm->vals_f32.p;
        m.row_ids = d_kept.p + klo;
        m.n_rows = kcnt;
        m.c_off = cs_off.p;
        m.c_end = cs_end.p;
        m.c_keys = cs_keys.p;
        m.c_vals = cs_vals.p;
        m.n_cols = pa.n_cols;
        m.sim = sim;
        m.out = pa.out;
        m.ld = pa.ld;
        m.transposed = transposed;
        ctx->cluster_stats[0] += (uint64_t)kcnt * pa.n_cols;
        launch_merge_pairs(ctx, m, name);
    };

    while (nc < p->expected_nb_of_clusters) {  // :114-129
        DevBuf<float> d_dots(std::max<size_t>((size_t)kcnt * nc, 1));
        pa.n_cols = nc;
        pa.out = d_dots.p;
        pa.ld = nc;
        centroid_tile(false, "k6_pair_store");
        std::vector<float> dots((size_t)W * kchunk * nc);
        gather_rows(d_dots.p, nc * sizeof(float), klo, kcnt, kchunk, dots.data());
        uint32_t best = 0;
        float best_score = 0.0f;
        for (uint32_t i = 0; i < nk; ++i) {
            float s = 0.0f;
            for (uint32_t c = 0; c < nc; ++c) s = s + dots[(size_t)i * nc + c];
            s = s / (float)nc;
            const float h = (float)hck[kept[i]];
            const float pen = 0.8f + 0.2f * (1.0f - h / max_nb_hck);
            s = s * pen;
            if (!std::isfinite(s)) fail(MYRIO_ERR_NONFINITE, "non-finite centroid-selection score");
            if (i == 0 || s < best_score) {  // min_by_key -> first minimum
                best_score = s;
                best = i;
            }
        }
        push_centroid_from_read(kept[best]);
    }

    mark("centroids filled");
    // Step 3 (:132-169)
    DevBuf<uint2> d_best(std::max<uint32_t>(kcnt, 1));  // (best similarity bits, target) of this rank's rows
    DevBuf<uint32_t> d_target(nk), d_sums((size_t)nc * dim), d_members(nc);
    std::vector<uint2> best_target((size_t)W * kchunk);
    std::vector<float> best((size_t)W * kchunk);
    std::vector<uint32_t> target((size_t)W * kchunk);
    DevBuf<float> d_tile(std::max<size_t>((size_t)nc * kcnt, 1));  // reads x centroids similarities, [centroid][read]
    pa.n_cols = nc;
    pa.out = d_tile.p;
    pa.ld = kcnt;
    // best similarity + target of every kept read (host), target also on the device for the recentering
    auto assign_pass = [&]() {
        if (kcnt) {
            centroid_tile(true, "k6_pair_argmax");
            MYRIO_LAUNCH(ctx, "k6_argmax_rows", argmax_rows_kernel, (kcnt + 127) / 128, 128, 0, d_tile.p,
                         (uint64_t)kcnt, kcnt, nc, d_best.p);
        }
        gather_rows(d_best.p, sizeof(uint2), klo, kcnt, kchunk, best_target.data());
        for (uint32_t i = 0; i < nk; ++i) {
            std::memcpy(&best[i], &best_target[i].x, sizeof(float));
            target[i] = best_target[i].y;
        }
        h2d(ctx, d_target.p, target.data(), nk);
    };
    float previous_mean_wcsss = 1E-6f;
    float improvement_score = 3.40282347e+38f;
    uint64_t nb_iters = 0;
    while (improvement_score > p->eta_improvement && nb_iters < p->nb_iters_max) {
        assign_pass();
        // wcsss (:307-316): sum1 over members (push order = read order)
        std::vector<float> w(nc, 0.0f);
        std::vector<uint8_t> seen(nc, 0);
        for (uint32_t i = 0; i < nk; ++i) {
            const uint32_t c = target[i];
            if (!seen[c]) {
                w[c] = best[i];
                seen[c] = 1;
            } else {
                w[c] = w[c] + best[i];
            }
        }
        float sum_w = 0.0f;
        for (uint32_t c = 0; c < nc; ++c) sum_w = sum_w + w[c];
        const float current_mean_wcsss = sum_w / (float)p->expected_nb_of_clusters;
        improvement_score = (current_mean_wcsss - previous_mean_wcsss) / previous_mean_wcsss;
        previous_mean_wcsss = current_mean_wcsss;
        // recenter (:167)
        MYRIO_CK(cudaMemsetAsync(d_sums.p, 0, (size_t)nc * dim * sizeof(uint32_t), ctx->stream));
        MYRIO_CK(cudaMemsetAsync(d_members.p, 0, nc * sizeof(uint32_t), ctx->stream));
        const size_t rc_bytes = ((size_t)nc * dim + nc) * sizeof(uint32_t);
        if (rc_bytes + 2048 <= ctx->smem_optin) {
            MYRIO_CK(cudaFuncSetAttribute(recenter_accumulate_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)rc_bytes));
            MYRIO_LAUNCH(ctx, "k6_recenter_accumulate", recenter_accumulate_smem_kernel,
                         (unsigned)std::min<uint32_t>((nk + 7) / 8, ctx->sm_count), 256, rc_bytes, counts->row_off.p,
                         counts->keys.p, counts->vals_f32.p, d_kept.p, d_target.p, nk, dim, nc, d_sums.p, d_members.p);
        } else {
            MYRIO_LAUNCH(ctx, "k6_recenter_accumulate", recenter_accumulate_kernel, (nk + 7) / 8, 256, 0,
                         counts->row_off.p, counts->keys.p, counts->vals_f32.p, d_kept.p, d_target.p, nk, dim,
                         d_sums.p, d_members.p);
        }
        MYRIO_CK(cudaFuncSetAttribute(recenter_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(dim * sizeof(float))));
        MYRIO_LAUNCH(ctx, "k6_recenter_finalize", recenter_finalize_kernel, nc, 256, dim * sizeof(float), d_sums.p,
                     d_members.p, nc, dim, cen.p);
        nb_iters += 1;
    }
    if (n_iters) *n_iters = (uint32_t)nb_iters;

    mark("iterations done");
    // Step 5 / first half of step 4 (:200-233): final assignment
    assign_pass();
    sync(ctx);
    for (uint32_t i = 0; i < nk; ++i) assign[kept[i]] = (int32_t)target[i];
    if (!p->has_silhouette_trimming || nc == 1) return;

    // Step 4 (:235-267) with compute_silhouette_scores_vec (:318-365)
    std::vector<std::vector<uint32_t>> members(nc);
    for (uint32_t i = 0; i < nk; ++i) members[target[i]].push_back(kept[i]);
    std::vector<float> hcen((size_t)nc * dim);
    d2h(ctx, hcen.data(), cen.p, hcen.size());
    sync(ctx);
    const float ssdcf = p->silhouette_trimming;
    // read vectors counted on both strands are reverse-complement symmetric: canonical-key columns,
    // three bank-permuted replicas, scheduled row slabs (needs 6 * plane_rows * 16 B of shared memory)
    CanonTable canon;
    const bool canon_path = !merge_path && p->k <= 6 && canon_path_enabled();
    if (canon_path) build_canon_table(ctx, p->k, canon);
    // neighbour of every cluster: last maximum over j != c of sim(centroid_c, centroid_j) (:345-351)
    std::vector<uint32_t> nbr(nc, 0xFFFFFFFFu);
    for (uint32_t c = 0; c < nc; ++c) {
        float bs = 0.0f;
        for (uint32_t j = 0; j < nc; ++j) {
            if (j == c) continue;
            const float s = host_dense_sim(sim, &hcen[(size_t)c * dim], &hcen[(size_t)j * dim], dim);
            if (nbr[c] == 0xFFFFFFFFu || s >= bs) {
                bs = s;
                nbr[c] = j;
            }
        }
    }
    // row lengths of the normalised reads (exact term counts of the symmetric launches)
    std::vector<uint64_t> h_roff;
    const bool sym_ok = canon_path && W == 1 && !(getenv("MYRIO_K6_SYM") && getenv("MYRIO_K6_SYM")[0] == '0');
    if (sym_ok) {
        h_roff.resize(N + 1);
        d2h(ctx, h_roff.data(), norm->row_off.p, N + 1);
        sync(ctx);
    }
    uint64_t sym_budget = PT_TILE_BUDGET;  // bytes of similarity tile a symmetric launch may hold
    if (sym_ok) {
        size_t free_b = 0, total_b = 0;
        MYRIO_CK(cudaMemGetInfo(&free_b, &total_b));
        sym_budget = std::max<uint64_t>(PT_TILE_BUDGET, std::min<uint64_t>(free_b / 2, 48ull << 30));
    }

    // Silhouette sums (a, b) of `rows` against the columns `cols`: columns [0, split) are the rows' own cluster
    // (for the rows before swap_row; the other way round after it).  sym_cols > 0: the first sym_cols columns
    // ARE the rows (same reads, same order) and the whole square is held in one tile.
    auto tile_sums = [&](const uint32_t *rows_h, uint32_t n_rows, const std::vector<uint32_t> &cols, uint32_t split,
                         uint32_t swap_row, uint32_t sym_cols, DevBuf<float> &d_sums2) {
        DevBuf<uint32_t> d_rows(n_rows), d_cols(cols.size());
        h2d(ctx, d_rows.p, rows_h, n_rows);
        h2d(ctx, d_cols.p, cols.data(), cols.size());
        MYRIO_CK(cudaMemsetAsync(d_sums2.p, 0, (size_t)n_rows * 2 * sizeof(float), ctx->stream));
        RowSlab mslab;  // the rows of the tile
        if (canon_path)
            build_row_slab_scheduled(ctx, norm->row_off.p, norm->keys.p, norm->vals_f32.p, d_rows.p, n_rows, canon, mslab);
        else if (!merge_path)
            build_row_slab(ctx, norm->row_off.p, norm->keys.p, norm->vals_f32.p, d_rows.p, n_rows, mslab);
        PairArgs sa{};
        sa.c_off = norm->row_off.p;
        sa.c_keys = norm->keys.p;
        sa.c_vals = norm->vals_f32.p;
        sa.col_ids = d_cols.p;
        sa.dim = dim;
        sa.sim = sim;
        sa.ld = n_rows;
        // column passes of at most PT_TILE_BUDGET bytes of [column][member] similarities; the row
        // sums continue from pass to pass, so the f32 addition order is the member order
        const uint32_t n_cols_all = (uint32_t)cols.size();
        const uint64_t budget = sym_cols ? sym_budget : PT_TILE_BUDGET;
        uint32_t pass_cols = (uint32_t)std::min<uint64_t>(
            n_cols_all, std::max<uint64_t>(8, budget / (4ull * n_rows) / 8 * 8));
        if (sym_cols && pass_cols < sym_cols) sym_cols = 0;  // (the callers check the budget; belt and braces)
        DevBuf<float> d_sil_tile((size_t)pass_cols * n_rows);
        sa.out = d_sil_tile.p;
        for (uint32_t cf = 0; cf < n_cols_all; cf += pass_cols) {
            sa.c_first = cf;
            sa.n_cols = std::min(pass_cols, n_cols_all - cf);
            if (canon_path) {
                CanonArgs ca{};
                ca.c_off = sa.c_off;
                ca.c_keys = sa.c_keys;
                ca.c_vals = sa.c_vals;
                ca.col_ids = sa.col_ids;
                ca.c_first = cf;
                ca.n_cols = sa.n_cols;
                ca.sim = sim;
                ca.out = d_sil_tile.p;
                ca.ld = n_rows;
                ca.sym_cols = cf == 0 ? sym_cols : 0;
                if (ca.sym_cols) {
                    // terms actually formed: inside the square a 32-row group meets the column chunks that start
                    // below its last row; outside it, every column
                    const uint32_t sq = ca.sym_cols / 8 * 8;  // columns in symmetric chunks
                    uint64_t terms = 0;
                    for (uint32_t i = 0; i < n_rows; ++i) {
                        const uint64_t nnz_i = h_roff[rows_h[i] + 1] - h_roff[rows_h[i]];
                        const uint32_t inside = std::min<uint32_t>(sq, ((i >> 5) * 32 + 32) / 8 * 8);  // whole 32-row groups
                        terms += nnz_i * (inside + (sa.n_cols - sq));
                    }
                    ctx->cluster_stats[1] += terms;
                }
                launch_pair_tile_canon(ctx, ca, mslab, canon, "k6_pair_silhouette");
            } else if (!merge_path) {
                launch_pair_tile<true>(ctx, sa, mslab, "k6_pair_silhouette");
            } else {
                MergeArgs m{};
                m.r_off = norm->row_off.p;
                m.r_keys = norm->keys.p;
                m.r_vals = norm->vals_f32.p;
                m.row_ids = d_rows.p;
                m.n_rows = n_rows;
                m.c_off = norm->row_off.p;
                m.c_keys = norm->keys.p;
                m.c_vals = norm->vals_f32.p;
                m.col_ids = d_cols.p;
                m.c_first = cf;
                m.n_cols = sa.n_cols;
                m.sim = sim;
                m.out = d_sil_tile.p;
                m.ld = n_rows;
                m.transposed = true;
                ctx->cluster_stats[0] += (uint64_t)n_rows * sa.n_cols;
                launch_merge_pairs(ctx, m, "k6_pair_silhouette");
            }
            MYRIO_LAUNCH(ctx, "k6_silsum_rows", silsum_rows_kernel, (n_rows + 127) / 128, 128, 0, d_sil_tile.p,
                         (uint64_t)n_rows, n_rows, cf, sa.n_cols, split, d_sums2.p, swap_row);
        }
    };
    // silhouette scores of cluster c from its members' sums -> trimming (:244-266)
    auto finish_cluster = [&](uint32_t c, const float *sums, uint32_t nn) {
        const uint32_t nm = (uint32_t)members[c].size();
        std::vector<float> sil(nm);
        for (uint32_t t = 0; t < nm; ++t) {  // inner() :323-336
            const float a = sums[(size_t)t * 2] / (float)(nm - 1);
            const float b = sums[(size_t)t * 2 + 1] / (float)nn;
            sil[t] = (b - a) / fmaxf(a, b);
        }
        const float n = (float)nm;  // :244-249
        float sum = 0.0f;
        for (uint32_t t = 0; t < nm; ++t) sum = sum + sil[t];
        const float mean = sum / n;
        float sq = 0.0f;
        for (uint32_t t = 0; t < nm; ++t) {
            const float d = sil[t] - mean;
            sq = sq + d * d;
        }
        const float sd = sqrtf((1.0f / n) * sq);
        const float left_cutoff = mean - sd * ssdcf;
        for (uint32_t t = 0; t < nm; ++t) {
            const uint32_t r = members[c][t];
            if (silhouette) silhouette[r] = sil[t];
            if (sil[t] < left_cutoff) assign[r] = -2;  // :257-262
        }
    };
    std::vector<uint8_t> done(nc, 0);
    for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t nm = (uint32_t)members[c].size();
        if (nm == 0 || done[c]) continue;
        const uint32_t nb = nbr[c];
        const uint32_t nn = (uint32_t)members[nb].size();
        mark("silhouette cluster");
        // Two clusters that are each other's neighbour need the same similarities: ONE symmetric tile over
        // both member lists gives the four sums of both (half the work of two separate rectangles).
        if (sym_ok && nb > c && nbr[nb] == c && nn > 0 && !done[nb] &&
            (uint64_t)(nm + nn) * (nm + nn) * 4ull <= sym_budget) {
            std::vector<uint32_t> both(members[c]);
            both.insert(both.end(), members[nb].begin(), members[nb].end());
            DevBuf<float> d_sums2((size_t)(nm + nn) * 2);
            tile_sums(both.data(), nm + nn, both, nm, nm, nm + nn, d_sums2);
            std::vector<float> sums((size_t)(nm + nn) * 2);
            d2h(ctx, sums.data(), d_sums2.p, sums.size());
            sync(ctx);
            finish_cluster(c, sums.data(), nn);
            finish_cluster(nb, sums.data() + (size_t)nm * 2, nm);
            done[c] = done[nb] = 1;
            continue;
        }
        std::vector<uint32_t> cols(members[c]);
        cols.insert(cols.end(), members[nb].begin(), members[nb].end());
        uint32_t mlo, mcnt, mchunk;  // this rank's share of the cluster's members (rows of the tile)
        share(nm, mlo, mcnt, mchunk);
        std::vector<float> sums((size_t)W * mchunk * 2);
        DevBuf<float> d_sums2(std::max<size_t>((size_t)mcnt * 2, 1));
        if (mcnt) {
            // one rank: the members x members square is symmetric as well (when it fits one tile pass)
            const bool own_sym = sym_ok && (uint64_t)nm * (nm + 8) * 4ull <= sym_budget;
            tile_sums(members[c].data() + mlo, mcnt, cols, nm, 0xFFFFFFFFu, own_sym ? nm : 0, d_sums2);
        }
        gather_rows(d_sums2.p, 2 * sizeof(float), mlo, mcnt, mchunk, sums.data());
        finish_cluster(c, sums.data(), nn);
        done[c] = 1;
    }
}

// ------------------------------------------------ compute_cluster_centroid

// VT = float: rows of an f32 batch (read counts, fingerprints).  VT = uint16_t: cached leaf counts,
// converted like kmer_store_counts_to_kmer_counts (tax/mod.rs:55-58): val as f32 * rescale_factor.
template <typename VT>
__global__ void __launch_bounds__(256) gather_members_kernel(const uint64_t *__restrict__ r_off,
                                                             const uint64_t *__restrict__ r_keys,
                                                             const VT *__restrict__ r_vals,
                                                             const float *__restrict__ rescale,
                                                             const uint64_t *__restrict__ ids,
                                                             const uint64_t *__restrict__ dst_off, uint64_t m,
                                                             uint64_t *__restrict__ okeys,
                                                             uint64_t *__restrict__ ovals) {
    const uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= m) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t r = ids[i], b = r_off[r], n = r_off[r + 1] - b, d = dst_off[i];
    const float f = rescale ? rescale[r] : 1.0f;
    for (uint64_t j = lane; j < n; j += 32) {
        okeys[d + j] = r_keys[b + j];
        const float v = rescale ? __fmul_rn((float)r_vals[b + j], f) : (float)r_vals[b + j];
        ovals[d + j] = (uint64_t)__float_as_uint(v);
    }
}

__global__ void __launch_bounds__(256) cc_mark_heads_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                            uint32_t *__restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) cc_reduce_runs_kernel(const uint64_t *__restrict__ keys,
                                                             const uint64_t *__restrict__ vals, uint64_t n,
                                                             const uint32_t *__restrict__ flags,
                                                             const uint64_t *__restrict__ uidx,
                                                             uint64_t *__restrict__ ukeys,
                                                             float *__restrict__ usum) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    // iter.sum::<SFVec>() (data/sparse.rs:521-528): a fold of `acc + x` from the empty vector, i.e. per
    // key a sequential f32 sum in member order (the stable sort kept it); exact for integer counts
    const uint64_t key = keys[i];
    float s = __uint_as_float((uint32_t)vals[i]);
    for (uint64_t j = i + 1; j < n && keys[j] == key; ++j) s = __fadd_rn(s, __uint_as_float((uint32_t)vals[j]));
    const uint64_t u = uidx[i];
    ukeys[u] = key;
    usum[u] = s;
}

__global__ void cc_norm_kernel(const float *__restrict__ v, uint64_t n, float *__restrict__ magn) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float s = 0.0f;
        for (uint64_t i = 0; i < n; ++i) s = __fadd_rn(s, __fmul_rn(v[i], v[i]));
        *magn = __fsqrt_rn(s);
    }
}

__global__ void __launch_bounds__(256) cc_scale_kernel(float *__restrict__ v, uint64_t n,
                                                       const float *__restrict__ magn) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = __fdiv_rn(v[i], *magn);
}

myrio_svecs *cluster_centroid(myrio_ctx *ctx, const myrio_svecs *counts, const uint64_t *member_ids,
                              uint64_t m, const float *d_row_scale) {
    const bool leaf_counts = counts->vtype == MYRIO_VAL_U16;
    if (leaf_counts && d_row_scale) fail(MYRIO_ERR_BAD_ARG, "explicit row factors go with f32 batches");
    if (leaf_counts && counts->aux_kind != 2)
        fail(MYRIO_ERR_BAD_ARG, "u16 counts need their rescale factors (a myrio_count_leaves batch)");
    for (uint64_t i = 0; i < m; ++i)
        if (member_ids[i] >= counts->n_rows) fail(MYRIO_ERR_BAD_ARG, "member id out of range");
    auto out = std::make_unique<myrio_svecs>();
    out->n_rows = 1;
    out->vtype = MYRIO_VAL_F32;
    out->row_off.alloc(2);
    std::vector<uint64_t> ro(counts->n_rows + 1);
    d2h(ctx, ro.data(), counts->row_off.p, counts->n_rows + 1);
    sync(ctx);
    std::vector<uint64_t> dst(m + 1, 0);
    for (uint64_t i = 0; i < m; ++i) dst[i + 1] = dst[i] + (ro[member_ids[i] + 1] - ro[member_ids[i]]);
    const uint64_t total = dst[m];
    uint64_t U = 0;
    if (total) {
        DevBuf<uint64_t> d_ids(m), d_dst(m + 1), k1(total), v1(total), k2(total), v2(total), uidx(total + 1);
        DevBuf<uint32_t> hist, flags(total);
        DevBuf<uint64_t> offs;
        h2d(ctx, d_ids.p, member_ids, m);
        h2d(ctx, d_dst.p, dst.data(), m + 1);
        if (leaf_counts)
            MYRIO_LAUNCH(ctx, "cc_gather", gather_members_kernel<uint16_t>, (unsigned)((m + 7) / 8), 256, 0,
                         counts->row_off.p, counts->keys.p, counts->vals_u16.p, counts->aux_f32.p, d_ids.p, d_dst.p, m,
                         k1.p, v1.p);
        else
            MYRIO_LAUNCH(ctx, "cc_gather", gather_members_kernel<float>, (unsigned)((m + 7) / 8), 256, 0,
                         counts->row_off.p, counts->keys.p, counts->vals_f32.p, d_row_scale, d_ids.p, d_dst.p, m,
                         k1.p, v1.p);
        const bool in_tmp = radix_sort_pairs_raw(ctx, k1.p, v1.p, k2.p, v2.p, total, 64, hist, offs);
        const uint64_t *sk = in_tmp ? k2.p : k1.p, *sv = in_tmp ? v2.p : v1.p;
        const unsigned g = (unsigned)((total + 255) / 256);
        MYRIO_LAUNCH(ctx, "cc_mark_heads", cc_mark_heads_kernel, g, 256, 0, sk, total, flags.p);
        exclusive_scan_u32_to_u64(ctx, flags.p, uidx.p, total);
        d2h(ctx, &U, uidx.p + total, 1);
        sync(ctx);
        out->keys.alloc(U);
        out->vals_f32.alloc(U);
        MYRIO_LAUNCH(ctx, "cc_reduce_runs", cc_reduce_runs_kernel, g, 256, 0, sk, sv, total, flags.p, uidx.p,
                     out->keys.p, out->vals_f32.p);
        DevBuf<float> magn(1);
        MYRIO_LAUNCH(ctx, "cc_norm", cc_norm_kernel, 1, 32, 0, out->vals_f32.p, U, magn.p);
        MYRIO_LAUNCH(ctx, "cc_scale", cc_scale_kernel, (unsigned)((U + 255) / 256), 256, 0, out->vals_f32.p, U,
                     magn.p);
        sync(ctx);
    } else {
        out->keys.alloc(1);
        out->vals_f32.alloc(1);
    }
    out->nnz = U;
    const uint64_t h_ro[2] = {0, U};
    h2d(ctx, out->row_off.p, h_ro, 2);
    sync(ctx);
    return out.release();
}

// The re-weighting of the local fingerprints before the second clustering pass
// (myrio-cli/src/app/process/run.rs:243-262): score_i = simfunc(local_i, naive_query) (pre-normalised),
// adjustor_i = exp(1 + alpha * score_i), result = normalize_l2(sum_i local_i * adjustor_i).  The similarities
// and the weighted sparse sum run on the device; the n_local scalars exp() go through libm's expf on the host,
// which is what Rust's f32::exp calls.
myrio_svecs *reweight_fingerprints(myrio_ctx *ctx, const myrio_svecs *local, const myrio_svecs *naive_query,
                                   int32_t sim, float alpha) {
    if (local->vtype != MYRIO_VAL_F32 || naive_query->vtype != MYRIO_VAL_F32)
        fail(MYRIO_ERR_BAD_ARG, "fingerprints and the query are f32 (normalised) vectors");
    if (naive_query->n_rows != 1) fail(MYRIO_ERR_BAD_ARG, "the naive query is one vector (a 1-row batch)");
    const uint64_t n = local->n_rows;
    std::vector<float> score(std::max<uint64_t>(n, 1)), adjustor(std::max<uint64_t>(n, 1));
    if (n) pairwise(ctx, local, naive_query, sim, score.data());
    for (uint64_t i = 0; i < n; ++i) adjustor[i] = expf(1.0f + alpha * score[i]);  // :255
    DevBuf<float> d_adj(std::max<uint64_t>(n, 1));
    h2d(ctx, d_adj.p, adjustor.data(), n);
    std::vector<uint64_t> ids(n);
    std::iota(ids.begin(), ids.end(), 0ull);
    myrio_svecs *out = cluster_centroid(ctx, local, ids.data(), n, d_adj.p);  // :257-260 sum + into_normalized_l2
    sync(ctx);
    return out;
}

// --------------------------------------------------------------- pairwise

__global__ void __launch_bounds__(256) max_key_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                      unsigned long long *__restrict__ out) {
    uint64_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, keys[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, (unsigned long long)m);
}

static uint64_t max_key_of(myrio_ctx *ctx, const myrio_svecs *v) {
    if (v->nnz == 0) return 0;
    DevBuf<unsigned long long> d(1);
    MYRIO_CK(cudaMemsetAsync(d.p, 0, sizeof(unsigned long long), ctx->stream));
    MYRIO_LAUNCH(ctx, "max_key", max_key_kernel, (unsigned)std::min<uint64_t>(1024, (v->nnz + 255) / 256), 256, 0,
                 v->keys.p, v->nnz, d.p);
    unsigned long long h = 0;
    d2h(ctx, &h, d.p, 1);
    sync(ctx);
    return h;
}

void pairwise(myrio_ctx *ctx, const myrio_svecs *rows, const myrio_svecs *cols, int32_t sim, float *out) {
    if (rows->vtype != MYRIO_VAL_F32 || cols->vtype != MYRIO_VAL_F32)
        fail(MYRIO_ERR_BAD_ARG, "pairwise needs f32 (pre-normalised) batches");
    if (sim != MYRIO_SIM_COSINE && sim != MYRIO_SIM_JACARD_TANIMOTO && sim != MYRIO_SIM_OVERLAP)
        fail(MYRIO_ERR_BAD_ARG, "unknown similarity %d", sim);
    const uint64_t nr = rows->n_rows, ncol = cols->n_rows;
    ctx->cluster_stats[0] = ctx->cluster_stats[1] = 0;
    if (nr == 0 || ncol == 0) return;
    if (nr > 0xFFFFFFFFull || ncol > 0xFFFFFFFFull) fail(MYRIO_ERR_BAD_ARG, "too many rows/cols");
    const uint64_t mk = std::max(max_key_of(ctx, rows), max_key_of(ctx, cols));
    DevBuf<float> d_out(nr * ncol);
    if (mk < 16384 && sim != MYRIO_SIM_OVERLAP) {
        uint32_t dim = 16;
        while (dim <= mk) dim <<= 2;
        RowSlab slab;
        build_row_slab(ctx, rows->row_off.p, rows->keys.p, rows->vals_f32.p, nullptr, (uint32_t)nr, slab);
        PairArgs a{};
        a.c_off = cols->row_off.p;
        a.c_keys = cols->keys.p;
        a.c_vals = cols->vals_f32.p;
        a.n_cols = (uint32_t)ncol;
        a.dim = dim;
        a.sim = sim;
        a.out = d_out.p;
        a.ld = ncol;
        launch_pair_tile<false>(ctx, a, slab, "k6_pair_store");
    } else {
        // Overlap (sum of maxima over the union of the keys) and key spaces too large for dense
        // columns: the merge-join kernel
        MergeArgs m{};
        m.r_off = rows->row_off.p;
        m.r_keys = rows->keys.p;
        m.r_vals = rows->vals_f32.p;
        m.n_rows = (uint32_t)nr;
        m.c_off = cols->row_off.p;
        m.c_keys = cols->keys.p;
        m.c_vals = cols->vals_f32.p;
        m.n_cols = (uint32_t)ncol;
        m.sim = sim;
        m.out = d_out.p;
        m.ld = ncol;
        m.transposed = false;
        ctx->cluster_stats[0] += nr * ncol;
        launch_merge_pairs(ctx, m, "k6_pair_merge");
    }
    d2h(ctx, out, d_out.p, nr * ncol);
    sync(ctx);
}

}  // namespace myrio
