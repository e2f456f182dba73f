// kmer_count.cu -- K1 (read k-mer counting) and K2 (reference-leaf k-mer counting).
//
// One CTA per sequence.  The packed sequence is staged in shared memory with
// 16-byte loads, every window's key is produced by a funnel shift over the 2-bit
// words (forward key + reverse-complement key via brev), the keys are sorted in
// shared memory -- one MSD counting pass into 1024 order-preserving buckets, then
// every key finds its place inside its bucket (a handful of keys) by counting; a
// bitonic network is the fallback for low-complexity sequences -- and run-length
// reduced into a sorted (key, count) row with ballot-ranked compaction.  Rows go to
// an over-allocated scratch at a per-row offset known up front (2*(n-k+1) keys at
// most); compact_rows_kernel then packs them into the final CSR once the exact row
// sizes have been prefix-summed.
//
// Reference semantics restated here:
//   reads : MyrSeq::compute_kmer_counts      myrio-core/src/data/myrseq.rs:76-111
//   leaves: compute_kmer_store_counts_for_fasta_seq  myrio-core/src/tax/mod.rs:70-148
//   reduce: SparseVec::from_unsorted_singles  myrio-core/src/data/sparse.rs:229-272
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kmer_count.cuh"

namespace myrio {

static constexpr int KC_THREADS = 256;  // largest CTA; small size classes launch fewer threads (kc_threads_for)

static const uint32_t h_phred_bits[128] = {
#include "phred_table.inc"
};

__device__ float g_phred[128];

void upload_phred_table(myrio_ctx *ctx) {
    MYRIO_CK(cudaMemcpyToSymbolAsync(g_phred, h_phred_bits, sizeof(h_phred_bits), 0,
                                     cudaMemcpyHostToDevice, ctx->stream));
}

// ----------------------------------------------------------------- helpers

// key of the window starting at base i: bits [2i, 2i+2k) of the LSB-first packed words
__device__ __forceinline__ uint64_t window_key(const uint64_t *w, uint32_t i, uint32_t k) {
    const uint32_t wi = i >> 5, sh = (i & 31u) * 2u;
    uint64_t key = w[wi] >> sh;
    if (sh) key |= w[wi + 1] << (64u - sh);
    if (k < 32) key &= (1ull << (2u * k)) - 1ull;
    return key;
}

// key of the reverse complement of a window: complement (3-b == ~b per 2-bit
// group), reverse the order of the groups, realign to bit 0
__device__ __forceinline__ uint64_t revcomp_key(uint64_t key, uint32_t k) {
    uint64_t x = __brevll(~key);
    x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
    return x >> (64u - 2u * k);
}

// The (forward, reverse-complement) key pair of a window in the configured bit order.  LSB-first
// (base j at bits 2j, bio-seq's BitVec<usize, Lsb0>) is the default; MSB-first (base j at bits
// 2(k-1-j)) is the same pair with its 2-bit groups reversed: reversing the groups of kf is
// revcomp_key(~kf), and reversing those of kr = revcomp_key(kf) leaves ~kf.
__device__ __forceinline__ void order_key_pair(uint64_t &kf, uint64_t &kr, uint32_t k, uint32_t msb_first) {
    if (msb_first) {
        const uint64_t mask = k < 32 ? (1ull << (2u * k)) - 1ull : ~0ull;
        const uint64_t f = kf;
        kf = revcomp_key(~f, k);
        kr = ~f & mask;
    }
}

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one u32 per thread (blockDim.x threads, a multiple of 32); wbuf: 33 u32
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t *wbuf, uint32_t *total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan_u32(v);
    if (lane == 31) wbuf[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned nw = blockDim.x >> 5;
        uint32_t w = lane < nw ? wbuf[lane] : 0;
        uint32_t wi = warp_incl_scan_u32(w);
        if (lane < nw) wbuf[lane] = wi - w;
        if (lane == 31) wbuf[32] = wi;
    }
    __syncthreads();
    uint32_t res = wbuf[warp] + inc - v;
    *total = wbuf[32];
    __syncthreads();
    return res;
}

// Bitonic sort of S (power of two) keys, ascending; optional u16 payload moves along.
template <bool HAS_W>
__device__ __forceinline__ void bitonic_sort(uint64_t *keys, uint16_t *w, uint32_t S) {
    for (uint32_t size = 2; size <= S; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < (S >> 1); t += blockDim.x) {
                uint32_t lo = 2u * t - (t & (stride - 1u));
                uint32_t hi = lo + stride;
                bool asc = (lo & size) == 0u;
                uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == asc) {
                    keys[lo] = b;
                    keys[hi] = a;
                    if (HAS_W) {
                        uint16_t t0 = w[lo];
                        w[lo] = w[hi];
                        w[hi] = t0;
                    }
                }
            }
        }
    }
    __syncthreads();
}

// ---- sort of E u64 keys (K1 and the unambiguous path of K2).
// 1024 order-preserving buckets on the leading 10 bits of the key; the histogram is 512 packed pairs of u16
// counters (E <= 4096 here, so a half never carries into its neighbour) and doubles as the scatter cursor:
// after the scan a half holds its bucket's START, after the scatter its END (= the next bucket's start).
// Inside a bucket (E/1024 keys on average) a key's place is the number of smaller keys (ties: lower index
// first), counted by the key's own thread -- no per-thread insertion sort, no divergence beyond the bucket size.
// Returns false, `keys` untouched, if a bucket holds more than 32 keys (the caller falls back to bitonic).
static constexpr int KC_BINS = 1024;
__device__ __forceinline__ uint32_t kc_half(const uint32_t *h32, uint32_t d) {
    return (h32[d >> 1] >> ((d & 1u) * 16u)) & 0xFFFFu;
}
__device__ __forceinline__ bool bucket_rank_sort(uint64_t *keys, uint64_t *keys2, uint32_t E, uint32_t key_bits,
                                                 uint32_t *h32, uint32_t *wbuf) {
    const uint32_t tid = threadIdx.x, T = blockDim.x;
    const uint32_t WPT = (KC_BINS / 2) / T;  // packed words per thread (T is 64, 128 or 256)
    const uint32_t shift = key_bits > 10 ? key_bits - 10 : 0;
    for (uint32_t b = tid; b < KC_BINS / 2; b += T) h32[b] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < E; i += T) {
        const uint32_t d = (uint32_t)(keys[i] >> shift) & (KC_BINS - 1);
        atomicAdd(&h32[d >> 1], 1u << ((d & 1u) * 16u));
    }
    __syncthreads();
    uint32_t c = 0, cmax = 0;
    for (uint32_t j = 0; j < WPT; ++j) {
        const uint32_t w = h32[tid * WPT + j];
        c += (w & 0xFFFFu) + (w >> 16);
        cmax = max(cmax, max(w & 0xFFFFu, w >> 16));
    }
    uint32_t total;
    uint32_t run = block_excl_scan_u32(c, wbuf, &total);
    if (__syncthreads_or(cmax > 32u)) return false;
    for (uint32_t j = 0; j < WPT; ++j) {
        const uint32_t w = h32[tid * WPT + j];
        const uint32_t lo = w & 0xFFFFu, hi = w >> 16;
        h32[tid * WPT + j] = run | ((run + lo) << 16);
        run += lo + hi;
    }
    __syncthreads();
    for (uint32_t i = tid; i < E; i += T) {
        const uint64_t key = keys[i];
        const uint32_t d = (uint32_t)(key >> shift) & (KC_BINS - 1);
        const uint32_t old = atomicAdd(&h32[d >> 1], 1u << ((d & 1u) * 16u));
        keys2[(old >> ((d & 1u) * 16u)) & 0xFFFFu] = key;
    }
    __syncthreads();
    for (uint32_t i = tid; i < E; i += T) {
        const uint64_t x = keys2[i];
        const uint32_t d = (uint32_t)(x >> shift) & (KC_BINS - 1);
        const uint32_t e = kc_half(h32, d), b = d ? kc_half(h32, d - 1) : 0u;
        uint32_t rank = b;
        for (uint32_t j = b; j < e; ++j) {
            const uint64_t y = keys2[j];
            rank += (y < x || (y == x && j < i)) ? 1u : 0u;
        }
        keys[rank] = x;
    }
    __syncthreads();
    return true;
}

// Run-length reduce of the sorted keys[0..E) straight to the row's scratch, in key order: ballot-ranked
// compaction in position order (segment = 32 consecutive positions of one warp), two sweeps with a scan of
// the per-segment counts in between.  `emit(i, run) -> value` decides per run head whether the run is kept
// (value != 0) and what is stored; returns the number of runs kept.  `seg` holds round_len/32 + 8 counters
// (it aliases the sort buffer, dead by now); round_len positions are reduced per round.
template <typename VT, typename F>
__device__ __forceinline__ uint32_t reduce_runs(const uint64_t *keys, uint32_t E, uint32_t *seg, uint32_t round_len,
                                                uint32_t *s_d, uint64_t *out_keys, VT *out_vals, F value_of) {
    const uint32_t tid = threadIdx.x, T = blockDim.x, lane = tid & 31u, warp = tid >> 5, nW = T >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t done = 0;  // runs emitted by earlier rounds
    for (uint32_t r0 = 0; r0 < E; r0 += round_len) {
        const uint32_t re = min(E, r0 + round_len);
        const uint32_t iters = (re - r0 + T - 1) / T;
        for (uint32_t m = 0; m < iters; ++m) {
            const uint32_t i = r0 + m * T + tid;
            bool keep = false;
            if (i < re && (i == 0 || keys[i] != keys[i - 1])) {
                uint32_t j = i + 1;
                const uint64_t x = keys[i];
                while (j < E && keys[j] == x) ++j;
                keep = value_of(i, j - i) != (VT)0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) seg[m * nW + warp] = __popc(bal);
        }
        __syncthreads();
        if (warp == 0) {
            const uint32_t ns = iters * nW;
            uint32_t carry = 0;
            for (uint32_t c0 = 0; c0 < ns; c0 += 32) {
                const uint32_t v = c0 + lane < ns ? seg[c0 + lane] : 0u;
                const uint32_t inc = warp_incl_scan_u32(v);
                if (c0 + lane < ns) seg[c0 + lane] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) *s_d = carry;
        }
        __syncthreads();
        for (uint32_t m = 0; m < iters; ++m) {
            const uint32_t i = r0 + m * T + tid;
            VT v = (VT)0;
            uint64_t x = 0;
            if (i < re && (i == 0 || keys[i] != keys[i - 1])) {
                uint32_t j = i + 1;
                x = keys[i];
                while (j < E && keys[j] == x) ++j;
                v = value_of(i, j - i);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, v != (VT)0);
            if (v != (VT)0) {
                const uint32_t d = done + seg[m * nW + warp] + __popc(bal & lt);
                out_keys[d] = x;
                out_vals[d] = v;
            }
        }
        __syncthreads();
        done += *s_d;
        __syncthreads();
    }
    return done;
}

// bucket sort needs a second key buffer: only the small size classes carry one
__host__ __device__ inline bool has_bucket_buffer(uint32_t ecap) { return ecap <= 4096; }

__device__ __forceinline__ uint32_t pow2_ceil(uint32_t v) {
    if (v <= 2) return 2;
    return 1u << (32 - __clz(v - 1));
}

// Marks run heads of the sorted keys[0..E) and records their positions:
// headpos[idx] = index of the idx-th distinct key; returns D (number of runs).
// Each thread owns a contiguous chunk so the scan keeps the order.
__device__ __forceinline__ uint32_t find_run_heads(const uint64_t *keys, uint32_t E,
                                                   uint32_t *headpos, uint32_t *wbuf) {
    const uint32_t C = (E + blockDim.x - 1) / blockDim.x;
    const uint32_t b = min(E, threadIdx.x * C), e = min(E, b + C);
    uint32_t cnt = 0;
    for (uint32_t i = b; i < e; ++i) cnt += (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
    uint32_t D;
    uint32_t pos = block_excl_scan_u32(cnt, wbuf, &D);
    for (uint32_t i = b; i < e; ++i)
        if (i == 0 || keys[i] != keys[i - 1]) headpos[pos++] = i;
    __syncthreads();
    return D;
}

// ------------------------------------------------------------- K1: reads

struct ReadCountArgs {
    const uint64_t *words;
    const uint64_t *word_off;
    const uint64_t *base_off;
    const uint8_t *qual;
    const uint32_t *ids;  // rows handled by this launch
    uint32_t k;
    uint32_t msb_first;    // key bit order (myrio_ctx_set_kmer_bit_order)
    float cutoff;
    uint32_t ecap;         // sort capacity (power of two) >= 2*(n-k+1) for every row of the launch
    const uint64_t *eoff;  // scratch offset of every row
    uint64_t *skeys;
    float *svals;
    uint32_t *d_out;    // distinct keys per row
    uint64_t *hck_out;  // nb_hck per row
    int *err;
    uint8_t *gscratch;  // GLOBAL variant: per-CTA work area
    uint64_t gstride;
};

// shared/global work area: keys[ecap] u64
//                          | buf: keys2[ecap] u64 (small classes: bucket-sort buffer) or (ecap/2+64) f32
//                            -- the per-base confidence factors live here until the sort starts
//                          | words[ecap/64+4] u64
__host__ __device__ inline size_t read_buf_bytes(uint32_t ecap) {
    return has_bucket_buffer(ecap) ? (size_t)ecap * 8 : ((size_t)ecap / 2 + 64) * 4;
}
__host__ __device__ inline size_t read_work_bytes(uint32_t ecap) {
    return (size_t)ecap * 8 + read_buf_bytes(ecap) + ((size_t)ecap / 64 + 4) * 8;
}

template <bool GLOBAL>
__global__ void __launch_bounds__(KC_THREADS) count_reads_kernel(ReadCountArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ float s_phred[128];
    __shared__ uint32_t s_wbuf[33];
    __shared__ uint32_t s_cnt, s_d;
    __shared__ uint32_t s_h32[KC_BINS / 2];

    uint8_t *work = GLOBAL ? a.gscratch + (size_t)blockIdx.x * a.gstride : smem_raw;
    uint64_t *keys = reinterpret_cast<uint64_t *>(work);
    uint64_t *keys2 = reinterpret_cast<uint64_t *>(work + (size_t)a.ecap * 8);
    float *pfac = reinterpret_cast<float *>(keys2);  // dead before the sort writes keys2
    uint64_t *words = reinterpret_cast<uint64_t *>(work + (size_t)a.ecap * 8 + read_buf_bytes(a.ecap));

    const uint32_t tid = threadIdx.x, T = blockDim.x, lane = tid & 31u;
    const uint32_t r = a.ids[blockIdx.x];
    const uint64_t b0 = a.base_off[r];
    const uint32_t n = (uint32_t)(a.base_off[r + 1] - b0);
    const uint32_t k = a.k;

    for (uint32_t i = tid; i < 128; i += T) s_phred[i] = g_phred[i];
    if (tid == 0) s_cnt = 0;
    if (n < k) {  // kmers::<K>() yields nothing
        if (tid == 0) {
            a.d_out[r] = 0;
            a.hck_out[r] = 0;
        }
        return;
    }
    const uint32_t nb = n - k + 1;

    // packed bases: 16-byte loads (every read starts on an even word)
    {
        const uint64_t w0 = a.word_off[r];
        const uint32_t nw = (n + 31) >> 5;
        const uint4 *src = reinterpret_cast<const uint4 *>(a.words + w0);
        uint4 *dst = reinterpret_cast<uint4 *>(words);
        for (uint32_t i = tid; i < (nw + 1) / 2; i += T) dst[i] = __ldg(src + i);
        if (tid == 0) {  // window_key may touch one word past the end
            words[((nw + 1) / 2) * 2] = 0;
        }
    }
    __syncthreads();  // s_phred ready
    // quality bytes -> per-base probability of a correct call: one aligned 4-byte load per thread and round
    {
        const uint8_t *qp = a.qual + b0;
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(qp) & 3u);
        const uint32_t *q4 = reinterpret_cast<const uint32_t *>(qp - mis);
        const uint32_t nchunks = (mis + n + 3) >> 2;
        bool bad = false;
        for (uint32_t c = tid; c < nchunks; c += T) {
            const uint32_t v = __ldg(q4 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = (int)(c * 4 + j) - (int)mis;
                if (idx >= 0 && idx < (int)n) {
                    const uint32_t q = (v >> (j * 8)) & 0xFFu;
                    bad |= q >= 128u;
                    pfac[idx] = s_phred[q & 127u];
                }
            }
        }
        if (tid < 4) pfac[n + tid] = 1.0f;  // a thread's four windows read three factors past its last window
        if (bad) atomicMax(a.err, MYRIO_ERR_BAD_QUALITY);
    }
    __syncthreads();

    // windows: confidence = product of k factors (sequential f32, left to right, as the reference); a thread
    // takes FOUR consecutive windows so that a factor is loaded once and multiplied into every chain it belongs
    // to (1.0 * v == v exactly, so a chain starts with its first factor).  Every passing window emits its key
    // and its reverse complement's key.
    const uint32_t ngroups = (nb + 3) >> 2;
    for (uint32_t g0 = 0; g0 < ngroups; g0 += T) {
        const uint32_t i0 = 4u * (g0 + tid);
        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (g0 + tid < ngroups) {
            const float *pp = pfac + i0;
            if (k >= 4) {
                const float v0 = pp[0], v1 = pp[1], v2 = pp[2], v3 = pp[3];
                c[0] = __fmul_rn(__fmul_rn(__fmul_rn(v0, v1), v2), v3);
                c[1] = __fmul_rn(__fmul_rn(v1, v2), v3);
                c[2] = __fmul_rn(v2, v3);
                c[3] = v3;
#pragma unroll 4
                for (uint32_t j = 4; j < k; ++j) {
                    const float v = pp[j];
                    c[0] = __fmul_rn(c[0], v);
                    c[1] = __fmul_rn(c[1], v);
                    c[2] = __fmul_rn(c[2], v);
                    c[3] = __fmul_rn(c[3], v);
                }
                const float e0 = pp[k], e1 = pp[k + 1], e2 = pp[k + 2];
                c[1] = __fmul_rn(c[1], e0);
                c[2] = __fmul_rn(__fmul_rn(c[2], e0), e1);
                c[3] = __fmul_rn(__fmul_rn(__fmul_rn(c[3], e0), e1), e2);
            } else {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    float conf = 1.0f;
                    for (uint32_t j = 0; j < k; ++j) conf = __fmul_rn(conf, pp[w + j]);
                    c[w] = conf;
                }
            }
        }
        bool ps[4];
        uint32_t np = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            ps[w] = g0 + tid < ngroups && i0 + w < nb && c[w] > a.cutoff;
            np += ps[w] ? 1u : 0u;
        }
        const uint32_t incl = warp_incl_scan_u32(np);
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tot) {
            uint32_t slot = 0;
            if (lane == 31) slot = atomicAdd(&s_cnt, 2u * tot);
            slot = __shfl_sync(0xffffffffu, slot, 31) + 2u * (incl - np);
#pragma unroll
            for (int w = 0; w < 4; ++w)
                if (ps[w]) {
                    uint64_t kf = window_key(words, i0 + w, k), kr = revcomp_key(kf, k);
                    order_key_pair(kf, kr, k, a.msb_first);
                    keys[slot] = kf;
                    keys[slot + 1] = kr;
                    slot += 2;
                }
        }
    }
    __syncthreads();
    const uint32_t E = s_cnt;  // == nb_hck
    if (E == 0) {
        if (tid == 0) {
            a.d_out[r] = 0;
            a.hck_out[r] = 0;
        }
        return;
    }
    if (!(has_bucket_buffer(a.ecap) && bucket_rank_sort(keys, keys2, E, 2u * k, s_h32, s_wbuf))) {
        const uint32_t S = pow2_ceil(E);
        for (uint32_t i = E + tid; i < S; i += T) keys[i] = ~0ull;
        bitonic_sort<false>(keys, nullptr, S);
    }

    const uint64_t o = a.eoff[r];
    // value += 1.0 per occurrence: exact
    // (the segment counters alias the sort buffer: (ecap/2 + 64) words at least, ecap/32 + 8 needed)
    const uint32_t D = reduce_runs<float>(keys, E, reinterpret_cast<uint32_t *>(keys2), a.ecap, &s_d, a.skeys + o,
                                          a.svals + o, [](uint32_t, uint32_t run) { return (float)run; });
    if (tid == 0) {
        a.d_out[r] = D;
        a.hck_out[r] = E;
    }
}

// ------------------------------------------------------------ K2: leaves

// shared definition with oracle/myrio_oracle.c:myo_resample_pick
__device__ __forceinline__ uint32_t resample_pick(uint64_t seed, uint64_t leaf, uint32_t resample,
                                                  uint32_t pos, uint32_t n_opts) {
    uint64_t x = seed;
    x ^= leaf * 0x9E3779B97F4A7C15ull;
    x ^= ((uint64_t)resample << 32 | (uint64_t)pos) * 0xD1B54A32D192ED03ull;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (uint32_t)(((x >> 32) * (uint64_t)n_opts) >> 32);
}

// tax/mod.rs:77-97: options in ascending Dna order = descending bit of the 4-bit code
__device__ __forceinline__ uint32_t resolve_iupac(uint32_t code, uint64_t seed, uint64_t leaf,
                                                  uint32_t resample, uint32_t pos) {
    const uint32_t nopt = __popc(code);
    if (nopt == 1) return (uint32_t)__clz(code) - 28u;
    uint32_t pick = resample_pick(seed, leaf, resample, pos, nopt);
    // pick-th set bit counted from bit 3 downwards
    for (uint32_t b = 0; b < 4; ++b) {
        if (code & (8u >> b)) {
            if (pick == 0) return b;
            --pick;
        }
    }
    return 0;
}

struct LeafCountArgs {
    const uint64_t *words;  // 4-bit packed
    const uint64_t *word_off;
    const uint64_t *base_off;
    const uint32_t *ids;
    uint32_t k;
    uint32_t msb_first;  // key bit order (myrio_ctx_set_kmer_bit_order)
    uint32_t nb_resamples;
    uint64_t seed;
    uint64_t leaf_id_base;
    uint32_t ecap;  // sort capacity and max symbols of every row of the launch
    const uint64_t *eoff;
    uint64_t *skeys;
    uint16_t *svals;
    uint32_t *d_out;
    float *rescale_out;
    int *err;
    uint8_t *gscratch;
    uint64_t gstride;
};

// work area: keys[ecap] u64 | aux[ecap+64] u32 | wts[ecap] u16 | raw[ecap+16] u8 | comp[ecap+16] u8
//            | words2[ecap/32+4] u64
// The unambiguous path sorts with a second key buffer keys2[ecap] u64 that ALIASES aux..comp (8*ecap + 288
// bytes, all dead once the keys exist; aux then holds the segment counters of the run-length reduce).
__host__ __device__ inline size_t leaf_work_bytes(uint32_t ecap) {
    return (size_t)ecap * 8 + ((size_t)ecap + 64) * 4 + (size_t)ecap * 2 + 2 * ((size_t)ecap + 16) +
           ((size_t)ecap / 32 + 4) * 8 + 16;
}

template <bool GLOBAL>
__global__ void __launch_bounds__(KC_THREADS) count_leaves_kernel(LeafCountArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_wbuf[33];
    __shared__ uint32_t s_cnt, s_namb, s_over;
    __shared__ uint32_t s_sum, s_d;
    __shared__ uint32_t s_h32[KC_BINS / 2];

    uint8_t *work = GLOBAL ? a.gscratch + (size_t)blockIdx.x * a.gstride : smem_raw;
    size_t o_ = 0;
    uint64_t *keys = reinterpret_cast<uint64_t *>(work + o_);
    o_ += (size_t)a.ecap * 8;
    uint32_t *aux = reinterpret_cast<uint32_t *>(work + o_);
    o_ += ((size_t)a.ecap + 64) * 4;
    uint16_t *wts = reinterpret_cast<uint16_t *>(work + o_);
    o_ += (size_t)a.ecap * 2;
    uint8_t *raw = work + o_;
    o_ += (size_t)a.ecap + 16;
    uint8_t *comp = work + o_;
    o_ += (size_t)a.ecap + 16;
    o_ = (o_ + 15) & ~(size_t)15;
    uint64_t *words2 = reinterpret_cast<uint64_t *>(work + o_);
    uint64_t *keys2 = reinterpret_cast<uint64_t *>(aux);

    const uint32_t tid = threadIdx.x;
    const uint32_t r = a.ids[blockIdx.x];
    const uint64_t b0 = a.base_off[r];
    const uint32_t n_in = (uint32_t)(a.base_off[r + 1] - b0);
    const uint32_t k = a.k, R = a.nb_resamples;
    const uint64_t leaf = a.leaf_id_base + r;
    const uint64_t o = a.eoff[r];

    if (tid == 0) {
        s_cnt = 0;
        s_namb = 0;
        s_over = 0;
        s_sum = 0;
    }
    // Fast front end: a leaf of plain A/C/G/T (every 4-bit symbol one-hot: no gap, no ambiguity code -- the
    // normal case) goes from the 4-bit words straight to 2-bit words with word-wide bit tricks, two input
    // words (32 symbols) per thread.  Anything else takes the symbol-by-symbol path below.
    bool irregular = false;
    {
        const uint64_t *src = a.words + a.word_off[r];
        const uint32_t nw2 = (n_in + 31) >> 5;
        for (uint32_t wi = tid; wi <= nw2; wi += blockDim.x) {  // one zero word past the end for window_key
            uint64_t out = 0;
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t s0 = (2 * wi + h) * 16;
                if (s0 < n_in) {
                    const uint64_t w = __ldg(src + 2 * wi + h);
                    const uint32_t valid = min(16u, n_in - s0);
                    const uint64_t m = valid == 16 ? ~0ull : (1ull << (4 * valid)) - 1ull;
                    const uint64_t ones = 0x1111111111111111ull;
                    const uint64_t multi = ((w & (w >> 1)) & (7 * ones)) | ((w & (w >> 2)) & (3 * ones)) |
                                           ((w & (w >> 3)) & ones);
                    const uint64_t zero = ~(w | (w >> 1) | (w >> 2) | (w >> 3)) & ones;
                    irregular |= ((multi | zero) & m) != 0ull;
                    // A=8 -> 0, C=4 -> 1, G=2 -> 2, T=1 -> 3: high bit = G|T, low bit = C|T
                    const uint64_t hi = ((w >> 1) | w) & ones, lo = ((w >> 2) | w) & ones;
                    uint64_t x = ((hi << 1) | lo) & m;
                    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
                    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
                    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
                    x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
                    out |= x << (32 * h);
                }
            }
            words2[wi] = out;
        }
    }
    const bool fast = !__syncthreads_or(irregular);
    uint32_t n = n_in;
    if (!fast) {
        // unpack 4-bit symbols (16-byte loads: 32 symbols each)
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.words + a.word_off[r]);
            const uint32_t nchunks = (n_in + 31) >> 5;
            for (uint32_t c = tid; c < nchunks; c += blockDim.x) {
                uint4 v = __ldg(src + c);
                uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    uint32_t idx = c * 32 + j;
                    if (idx < n_in) raw[idx] = (uint8_t)((wv[j >> 3] >> ((j & 7) * 4)) & 15u);
                }
            }
        }
        __syncthreads();
        // drop gaps (Iupac::X == 0), tax/mod.rs:98 -- ordered compaction
        {
            const uint32_t C = (n_in + blockDim.x - 1) / blockDim.x;
            const uint32_t b = min(n_in, tid * C), e = min(n_in, b + C);
            uint32_t cnt = 0, amb = 0;
            for (uint32_t i = b; i < e; ++i) {
                cnt += raw[i] != 0;
                amb += __popc((uint32_t)raw[i]) > 1;
            }
            uint32_t pos = block_excl_scan_u32(cnt, s_wbuf, &n);
            for (uint32_t i = b; i < e; ++i)
                if (raw[i] != 0) comp[pos++] = raw[i];
            if (amb) atomicAdd(&s_namb, amb);
        }
        __syncthreads();
    }
    const uint32_t cutoff = (R + 5u) >> 3;  // tax/mod.rs:130
    if (n < k || R == 0) {                  // :99-102 / empty singles -> :136-139
        if (tid == 0) {
            a.skeys[o] = 0;
            a.svals[o] = 1;
            a.d_out[r] = 1;
            a.rescale_out[r] = 1.0f;
        }
        return;
    }
    const uint32_t nb = n - k + 1;
    const bool ambiguous = !fast && s_namb != 0;
    uint32_t E;
    if (!ambiguous) {
        // all resamples are identical: count once, scale by R (wrapping u16 like the
        // release-mode `+=` of from_unsorted_singles)
        if (!fast) {  // gaps were dropped: pack the compacted symbols
            const uint32_t nw = (n + 31) >> 5;
            for (uint32_t wi = tid; wi <= nw; wi += blockDim.x) {
                uint64_t w = 0;
                const uint32_t lim = min(32u, n > wi * 32 ? n - wi * 32 : 0u);
                for (uint32_t j = 0; j < lim; ++j)
                    w |= (uint64_t)((uint32_t)__clz((uint32_t)comp[wi * 32 + j]) - 28u) << (2 * j);
                words2[wi] = w;
            }
            __syncthreads();
        }
        for (uint32_t i = tid; i < nb; i += blockDim.x) {
            uint64_t kf = window_key(words2, i, k), kr = revcomp_key(kf, k);
            order_key_pair(kf, kr, k, a.msb_first);
            keys[2 * i] = kf;
            keys[2 * i + 1] = kr;
        }
        E = 2 * nb;
    } else {
        // prefix count of ambiguous symbols -> window i is ambiguous iff pre[i+k] > pre[i]
        uint32_t *pre = aux;
        {
            const uint32_t C = (n + blockDim.x - 1) / blockDim.x;
            const uint32_t b = min(n, tid * C), e = min(n, b + C);
            uint32_t cnt = 0;
            for (uint32_t i = b; i < e; ++i) cnt += __popc((uint32_t)comp[i]) > 1;
            uint32_t tot;
            uint32_t pos = block_excl_scan_u32(cnt, s_wbuf, &tot);
            for (uint32_t i = b; i < e; ++i) {
                pre[i] = pos;
                pos += __popc((uint32_t)comp[i]) > 1;
            }
            if (tid == blockDim.x - 1 || e == n) pre[n] = tot;
        }
        __syncthreads();
        for (uint32_t i = tid; i < nb; i += blockDim.x) {
            const bool amb = pre[i + k] != pre[i];
            const uint32_t need = amb ? 2 * R : 2;
            const uint32_t slot = atomicAdd(&s_cnt, need);
            if (slot + need > a.ecap) {
                s_over = 1;
                continue;
            }
            if (!amb) {
                uint64_t kf = 0;
                for (uint32_t j = 0; j < k; ++j)
                    kf |= (uint64_t)((uint32_t)__clz((uint32_t)comp[i + j]) - 28u) << (2 * j);
                uint64_t kr = revcomp_key(kf, k);
                order_key_pair(kf, kr, k, a.msb_first);
                keys[slot] = kf;
                wts[slot] = (uint16_t)R;
                keys[slot + 1] = kr;
                wts[slot + 1] = (uint16_t)R;
            } else {
                for (uint32_t rs = 0; rs < R; ++rs) {
                    uint64_t kf = 0;
                    for (uint32_t j = 0; j < k; ++j)
                        kf |= (uint64_t)resolve_iupac(comp[i + j], a.seed, leaf, rs, i + j) << (2 * j);
                    uint64_t kr = revcomp_key(kf, k);
                    order_key_pair(kf, kr, k, a.msb_first);
                    keys[slot + 2 * rs] = kf;
                    wts[slot + 2 * rs] = 1;
                    keys[slot + 2 * rs + 1] = kr;
                    wts[slot + 2 * rs + 1] = 1;
                }
            }
        }
        __syncthreads();
        if (s_over) {
            if (tid == 0) atomicMax(a.err, MYRIO_ERR_UNSUPPORTED);
            return;
        }
        E = s_cnt;
    }
    __syncthreads();
    if (!(!ambiguous && has_bucket_buffer(a.ecap) && bucket_rank_sort(keys, keys2, E, 2u * k, s_h32, s_wbuf))) {
        const uint32_t S = pow2_ceil(E);
        for (uint32_t i = E + tid; i < S; i += blockDim.x) {
            keys[i] = ~0ull;
            if (ambiguous) wts[i] = 0;
        }
        if (ambiguous)
            bitonic_sort<true>(keys, wts, S);
        else
            bitonic_sort<false>(keys, nullptr, S);
    }

    if (!ambiguous) {
        // value = run length x R (u16, wrapping); strip values <= cutoff (:130-134); sum of the survivors
        uint32_t sum = 0;
        const uint32_t D = reduce_runs<uint16_t>(keys, E, aux, a.ecap, &s_d, a.skeys + o, a.svals + o,
                                                 [&](uint32_t, uint32_t run) {
                                                     const uint32_t v = (run * R) & 0xFFFFu;
                                                     return (uint16_t)(v > cutoff ? v : 0u);
                                                 });
        // the survivors' sum: every thread re-reads its share of what was just written (L1-resident)
        for (uint32_t i = tid; i < D; i += blockDim.x) sum += a.svals[o + i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if ((tid & 31u) == 0 && sum) atomicAdd(&s_sum, sum);  // u32 wrapping like sum::<u32>() in release
        __syncthreads();
        if (tid == 0) {
            if (D == 0) {  // :136-139
                a.skeys[o] = 0;
                a.svals[o] = 1;
                a.d_out[r] = 1;
                a.rescale_out[r] = 1.0f;
            } else {
                a.d_out[r] = D;
                a.rescale_out[r] = __fdiv_rn((float)(2u * nb), (float)s_sum);  // :144-145
            }
        }
        return;
    }
    uint32_t *headpos = aux;
    const uint32_t Dall = find_run_heads(keys, E, headpos, s_wbuf);
    // value per run, strip values <= cutoff (:130-134), ordered compaction of the survivors
    {
        const uint32_t C = (Dall + blockDim.x - 1) / blockDim.x;
        const uint32_t b = min(Dall, tid * C), e = min(Dall, b + C);
        uint32_t kept = 0, sum = 0;
        // pass 1: count survivors of this thread's chunk
        for (uint32_t idx = b; idx < e; ++idx) {
            const uint32_t i = headpos[idx];
            const uint32_t nx = (idx + 1 < Dall) ? headpos[idx + 1] : E;
            uint32_t v;
            if (ambiguous) {
                v = 0;
                for (uint32_t j = i; j < nx; ++j) v += wts[j];
            } else {
                v = (nx - i) * R;
            }
            v &= 0xFFFFu;
            kept += v > cutoff;
        }
        uint32_t D;
        uint32_t pos = block_excl_scan_u32(kept, s_wbuf, &D);
        for (uint32_t idx = b; idx < e; ++idx) {
            const uint32_t i = headpos[idx];
            const uint32_t nx = (idx + 1 < Dall) ? headpos[idx + 1] : E;
            uint32_t v;
            if (ambiguous) {
                v = 0;
                for (uint32_t j = i; j < nx; ++j) v += wts[j];
            } else {
                v = (nx - i) * R;
            }
            v &= 0xFFFFu;
            if (v > cutoff) {
                a.skeys[o + pos] = keys[i];
                a.svals[o + pos] = (uint16_t)v;
                sum += v;
                ++pos;
            }
        }
        if (sum) atomicAdd(&s_sum, sum);  // u32 wrapping like sum::<u32>() in release
        __syncthreads();
        if (tid == 0) {
            if (D == 0) {  // :136-139
                a.skeys[o] = 0;
                a.svals[o] = 1;
                a.d_out[r] = 1;
                a.rescale_out[r] = 1.0f;
            } else {
                a.d_out[r] = D;
                a.rescale_out[r] = __fdiv_rn((float)(2u * nb), (float)s_sum);  // :144-145
            }
        }
    }
}

// ------------------------------------------------- scratch -> CSR compaction

template <typename VT>
__global__ void __launch_bounds__(256) compact_rows_kernel(const uint64_t *__restrict__ eoff,
                                                           const uint64_t *__restrict__ row_off,
                                                           const uint64_t *__restrict__ skeys,
                                                           const VT *__restrict__ svals,
                                                           uint64_t *__restrict__ keys,
                                                           VT *__restrict__ vals, uint64_t n_rows) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t src = eoff[row], dst = row_off[row];
    const uint32_t d = (uint32_t)(row_off[row + 1] - dst);
    for (uint32_t i = lane; i < d; i += 32) {
        keys[dst + i] = skeys[src + i];
        vals[dst + i] = svals[src + i];
    }
}

// ------------------------------------------------------------ host drivers

static const uint32_t kSmemClasses[] = {1024, 2048, 4096, 8192, 16384};

// threads per CTA for a sort capacity: the CTA's fixed costs (scans, barriers, idle lanes of short
// loops) scale with its warps, so a ~500-key read gets 2-4 warps, not 8.  MYRIO_KC_THREADS=a,b
// overrides the first two classes (tuning).
static unsigned kc_threads_for(uint32_t ecap) {
    static const int ov[2] = {[] { const char *e = getenv("MYRIO_KC_THREADS"); return e ? atoi(e) : 0; }(),
                              [] { const char *e = getenv("MYRIO_KC_THREADS"); const char *c = e ? strchr(e, ',') : nullptr; return c ? atoi(c + 1) : 0; }()};
    auto ok = [](int v) { return v == 64 || v == 128 || v == 256; };
    if (ecap <= 1024) return ok(ov[0]) ? ov[0] : 128;  // measured: 64 / 128 / 256 threads = 1.23 / 1.02 / 1.07 ms (K1)
    if (ecap <= 2048) return ok(ov[1]) ? ov[1] : 256;  // measured: 64 / 128 / 256 threads = 4.3 / 3.1 / 2.7 ms (K2)
    return KC_THREADS;
}

template <typename KernelT>
static void allow_smem(KernelT kernel, size_t bytes) {
    MYRIO_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

static uint32_t host_pow2_ceil(uint64_t v) {
    uint32_t p = 2;
    while (p < v) p <<= 1;
    return p;
}

myrio_svecs *count_reads(myrio_ctx *ctx, const myrio_seqs *reads, uint32_t k, float cutoff) {
    if (k < 2 || k > 32) fail(MYRIO_ERR_INVALID_K, "k=%u: only k in 2..32 is supported", k);
    if (reads->kind != 0) fail(MYRIO_ERR_BAD_ARG, "count_reads needs a reads batch");
    const uint64_t N = reads->n;
    auto *out = new myrio_svecs();
    std::unique_ptr<myrio_svecs> guard(out);
    out->n_rows = N;
    out->vtype = MYRIO_VAL_F32;
    out->aux_kind = 1;
    out->row_off.alloc(N + 1);
    out->aux_u64.alloc(std::max<uint64_t>(N, 1));
    if (N == 0) {
        MYRIO_CK(cudaMemsetAsync(out->row_off.p, 0, sizeof(uint64_t), ctx->stream));
        sync(ctx);
        return guard.release();
    }
    // scratch row offsets + size classes from the host copy of the lengths
    std::vector<uint64_t> eoff(N + 1);
    std::vector<std::vector<uint32_t>> cls(sizeof(kSmemClasses) / sizeof(uint32_t) + 1);
    uint64_t acc = 0, max_e = 0;
    for (uint64_t r = 0; r < N; ++r) {
        uint64_t n = reads->h_base_off[r + 1] - reads->h_base_off[r];
        uint64_t e = n >= k ? 2 * (n - k + 1) : 0;
        eoff[r] = acc;
        acc += e;
        max_e = std::max(max_e, e);
        size_t c = 0;
        while (c < sizeof(kSmemClasses) / sizeof(uint32_t) && e > kSmemClasses[c]) ++c;
        cls[c].push_back((uint32_t)r);
    }
    eoff[N] = acc;
    DevBuf<uint64_t> d_eoff(N + 1), skeys(std::max<uint64_t>(acc, 1));
    DevBuf<float> svals(std::max<uint64_t>(acc, 1));
    DevBuf<uint32_t> d_d(N);
    DevBuf<int> d_err(1);
    h2d(ctx, d_eoff.p, eoff.data(), N + 1);
    MYRIO_CK(cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));

    std::vector<DevBuf<uint32_t>> id_bufs;
    DevBuf<uint8_t> gscratch;
    for (size_t c = 0; c < cls.size(); ++c) {
        if (cls[c].empty()) continue;
        id_bufs.emplace_back(cls[c].size());
        h2d(ctx, id_bufs.back().p, cls[c].data(), cls[c].size());
        ReadCountArgs a{};
        a.words = reads->words.p;
        a.word_off = reads->word_off.p;
        a.base_off = reads->base_off.p;
        a.qual = reads->qual.p;
        a.ids = id_bufs.back().p;
        a.k = k;
        a.msb_first = ctx->kmer_msb_first;
        a.cutoff = cutoff;
        a.eoff = d_eoff.p;
        a.skeys = skeys.p;
        a.svals = svals.p;
        a.d_out = d_d.p;
        a.hck_out = out->aux_u64.p;
        a.err = d_err.p;
        if (c < sizeof(kSmemClasses) / sizeof(uint32_t)) {
            a.ecap = kSmemClasses[c];
            size_t smem = read_work_bytes(a.ecap);
            allow_smem(count_reads_kernel<false>, smem);
            MYRIO_LAUNCH(ctx, "k1_count_reads", count_reads_kernel<false>, (unsigned)cls[c].size(),
                         kc_threads_for(a.ecap), smem, a);
        } else {
            // sequences too long for shared memory: same algorithm on a global work area,
            // processed in waves to bound the scratch
            a.ecap = host_pow2_ceil(max_e);
            a.gstride = (read_work_bytes(a.ecap) + 255) / 256 * 256;
            const size_t wave = std::max<size_t>(1, std::min<size_t>(cls[c].size(), (1ull << 30) / a.gstride));
            gscratch.alloc(wave * a.gstride);
            a.gscratch = gscratch.p;
            for (size_t b = 0; b < cls[c].size(); b += wave) {
                a.ids = id_bufs.back().p + b;
                unsigned g = (unsigned)std::min(wave, cls[c].size() - b);
                MYRIO_LAUNCH(ctx, "k1_count_reads_long", count_reads_kernel<true>, g, KC_THREADS, 0, a);
            }
        }
    }
    exclusive_scan_u32_to_u64(ctx, d_d.p, out->row_off.p, N);
    int h_err = 0;
    uint64_t nnz = 0;
    d2h(ctx, &h_err, d_err.p, 1);
    d2h(ctx, &nnz, out->row_off.p + N, 1);
    sync(ctx);
    if (h_err == MYRIO_ERR_BAD_QUALITY)
        fail(MYRIO_ERR_BAD_QUALITY, "a quality score >= 128 indexes past the Phred table");
    if (h_err) fail(h_err, "read counting failed (%d)", h_err);
    out->nnz = nnz;
    out->keys.alloc(std::max<uint64_t>(nnz, 1));
    out->vals_f32.alloc(std::max<uint64_t>(nnz, 1));
    MYRIO_LAUNCH(ctx, "compact_rows", compact_rows_kernel<float>, (unsigned)((N + 7) / 8), 256, 0,
                 d_eoff.p, out->row_off.p, skeys.p, svals.p, out->keys.p, out->vals_f32.p, N);
    sync(ctx);
    return guard.release();
}

myrio_svecs *count_leaves(myrio_ctx *ctx, const myrio_seqs *leaves, uint32_t k, uint32_t nb_resamples,
                          uint64_t seed, uint64_t leaf_id_base) {
    if (k < 2 || k > 32) fail(MYRIO_ERR_INVALID_K, "k=%u: only k in 2..32 is supported", k);
    if (leaves->kind != 1) fail(MYRIO_ERR_BAD_ARG, "count_leaves needs a leaves batch");
    if (nb_resamples > 65535) fail(MYRIO_ERR_BAD_ARG, "nb_resamples must fit in u16");
    const uint64_t L = leaves->n;
    auto *out = new myrio_svecs();
    std::unique_ptr<myrio_svecs> guard(out);
    out->n_rows = L;
    out->vtype = MYRIO_VAL_U16;
    out->aux_kind = 2;
    out->row_off.alloc(L + 1);
    out->aux_f32.alloc(std::max<uint64_t>(L, 1));
    if (L == 0) {
        MYRIO_CK(cudaMemsetAsync(out->row_off.p, 0, sizeof(uint64_t), ctx->stream));
        sync(ctx);
        return guard.release();
    }
    // Capacity per row.  ACGT-only leaves need 2*(n-k+1) keys; a leaf with `amb`
    // ambiguity codes needs at most 2*(nb + (R-1)*min(nb, k*amb)).
    const std::vector<uint32_t> &amb = leaves->h_aux;
    std::vector<uint64_t> eoff(L + 1);
    // size classes whose work area fits in shared memory (the 16 k class of a leaf is 266 KB: leaves
    // with many ambiguity codes take the global-scratch variant)
    size_t n_cls = 0;
    while (n_cls < sizeof(kSmemClasses) / sizeof(uint32_t) &&
           leaf_work_bytes(kSmemClasses[n_cls]) + 4096 <= ctx->smem_optin)
        ++n_cls;
    std::vector<std::vector<uint32_t>> cls(n_cls + 1);
    uint64_t acc = 0, max_cap = 0;
    for (uint64_t r = 0; r < L; ++r) {
        uint64_t n = leaves->h_base_off[r + 1] - leaves->h_base_off[r];
        uint64_t nb = n >= k ? n - k + 1 : 0;
        uint64_t na = amb.empty() ? 0 : amb[r];
        uint64_t e = 2 * (nb + (uint64_t)(nb_resamples > 0 ? nb_resamples - 1 : 0) * std::min<uint64_t>(nb, (uint64_t)k * na));
        uint64_t cap = std::max<uint64_t>(std::max<uint64_t>(e, n), 1);
        eoff[r] = acc;
        acc += std::max<uint64_t>(e, 1);
        max_cap = std::max(max_cap, cap);
        size_t c = 0;
        while (c < n_cls && cap > kSmemClasses[c]) ++c;
        cls[c].push_back((uint32_t)r);
    }
    eoff[L] = acc;
    DevBuf<uint64_t> d_eoff(L + 1), skeys(acc);
    DevBuf<uint16_t> svals(acc);
    DevBuf<uint32_t> d_d(L);
    DevBuf<int> d_err(1);
    h2d(ctx, d_eoff.p, eoff.data(), L + 1);
    MYRIO_CK(cudaMemsetAsync(d_err.p, 0, sizeof(int), ctx->stream));

    std::vector<DevBuf<uint32_t>> id_bufs;
    DevBuf<uint8_t> gscratch;
    for (size_t c = 0; c < cls.size(); ++c) {
        if (cls[c].empty()) continue;
        id_bufs.emplace_back(cls[c].size());
        h2d(ctx, id_bufs.back().p, cls[c].data(), cls[c].size());
        LeafCountArgs a{};
        a.words = leaves->words.p;
        a.word_off = leaves->word_off.p;
        a.base_off = leaves->base_off.p;
        a.ids = id_bufs.back().p;
        a.k = k;
        a.msb_first = ctx->kmer_msb_first;
        a.nb_resamples = nb_resamples;
        a.seed = seed;
        a.leaf_id_base = leaf_id_base;
        a.eoff = d_eoff.p;
        a.skeys = skeys.p;
        a.svals = svals.p;
        a.d_out = d_d.p;
        a.rescale_out = out->aux_f32.p;
        a.err = d_err.p;
        if (c < n_cls) {
            a.ecap = kSmemClasses[c];
            size_t smem = leaf_work_bytes(a.ecap);
            allow_smem(count_leaves_kernel<false>, smem);
            MYRIO_LAUNCH(ctx, "k2_count_leaves", count_leaves_kernel<false>, (unsigned)cls[c].size(),
                         kc_threads_for(a.ecap), smem, a);
        } else {
            a.ecap = host_pow2_ceil(max_cap);
            a.gstride = (leaf_work_bytes(a.ecap) + 255) / 256 * 256;
            const size_t wave = std::max<size_t>(1, std::min<size_t>(cls[c].size(), (1ull << 30) / a.gstride));
            gscratch.alloc(wave * a.gstride);
            a.gscratch = gscratch.p;
            for (size_t b = 0; b < cls[c].size(); b += wave) {
                a.ids = id_bufs.back().p + b;
                unsigned g = (unsigned)std::min(wave, cls[c].size() - b);
                MYRIO_LAUNCH(ctx, "k2_count_leaves_long", count_leaves_kernel<true>, g, KC_THREADS, 0, a);
            }
        }
    }
    exclusive_scan_u32_to_u64(ctx, d_d.p, out->row_off.p, L);
    int h_err = 0;
    uint64_t nnz = 0;
    d2h(ctx, &h_err, d_err.p, 1);
    d2h(ctx, &nnz, out->row_off.p + L, 1);
    sync(ctx);
    if (h_err) fail(h_err, "leaf counting failed (%d)", h_err);
    out->nnz = nnz;
    out->keys.alloc(std::max<uint64_t>(nnz, 1));
    out->vals_u16.alloc(std::max<uint64_t>(nnz, 1));
    MYRIO_LAUNCH(ctx, "compact_rows", compact_rows_kernel<uint16_t>, (unsigned)((L + 7) / 8), 256, 0,
                 d_eoff.p, out->row_off.p, skeys.p, svals.p, out->keys.p, out->vals_u16.p, L);
    sync(ctx);
    return guard.release();
}

}  // namespace myrio
