// common.cuh -- context, error handling, device buffers, launch bookkeeping.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/myrio_b200.h"

namespace myrio {

// internal exception type (never crosses the ABI; NOT the myrio::Error of include/myrio_b200.hpp)
struct Failure {
    int32_t code;
    std::string msg;
};

void set_last_error(const std::string &msg);

[[noreturn]] inline void fail(int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Failure{code, std::string(buf)};
}

#define MYRIO_CK(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::myrio::fail(e__ == cudaErrorMemoryAllocation ? MYRIO_ERR_OOM : MYRIO_ERR_CUDA,    \
                          "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
        }                                                                                       \
    } while (0)

// Runs `f`, translating exceptions into a status + thread-local message.
template <typename F>
inline int32_t guarded(F &&f) {
    try {
        f();
        return MYRIO_OK;
    } catch (const Failure &e) {
        set_last_error(e.msg);
        return e.code;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return MYRIO_ERR_BAD_ARG;
    } catch (...) {
        set_last_error("unknown error");
        return MYRIO_ERR_BAD_ARG;
    }
}

// Stream of the context whose API call is running on this thread (set by ScopedCtx).
// Device buffers are stream-ordered allocations from the device's default memory
// pool (cudaMallocAsync): scratch of one call is recycled by the next without
// going back to the driver, and a buffer may be released while kernels that use it
// are still queued on the same stream.
cudaStream_t &alloc_stream();

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), s(o.s) {
        o.p = nullptr;
        o.n = 0;
    }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            s = o.s;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        s = alloc_stream();
        if (count) MYRIO_CK(cudaMallocAsync(&p, count * sizeof(T), s));
    }
    // grow-only scratch
    void reserve(size_t count) {
        if (count > n) alloc(count + count / 8);
    }
    void release() {
        if (p) {
            // the owning context (and its stream) may already be gone: fall back to cudaFree
            if (cudaFreeAsync(p, s) != cudaSuccess) {
                cudaGetLastError();
                cudaFree(p);
            }
        }
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

struct ProfEntry {
    std::string name;
    cudaEvent_t start, stop;
};

}  // namespace myrio

struct myrio_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    size_t smem_optin = 0;
    uint64_t launches = 0;
    uint32_t kmer_msb_first = 0;  // key bit order of K1/K2 (0 = bio-seq's LSB-first)
    int topx_bound_pruning = -1;  // myrio_ctx_set_topx_bound_pruning; -1 = the process default (MYRIO_K4_UB)
    bool prof_on = false;
    std::vector<myrio::ProfEntry> prof;
    // scoring scratch (grow-only)
    myrio::DevBuf<float> score_rows;
    myrio::DevBuf<unsigned long long> score_counters;
    myrio::DevBuf<uint64_t> topx_cand;
    myrio::DevBuf<uint32_t> topx_cand_cnt, topx_gthr;
    void *plan = nullptr;  // myrio::Plan (score.cu): hit-list scratch, grow-only
    // state between myrio_score_topx_seed_async and myrio_score_topx_sweep_async
    bool twophase_ready = false, twophase_seeded = false;
    uint64_t twophase_q = 0;
    uint32_t twophase_x = 0;
    const void *twophase_ix = nullptr;  // the index the seed pass planned against
    int32_t twophase_sim = 0;
    // multi-GPU (comm.cu): ncclComm_t of this rank, bound by myrio_comm_init
    void *comm = nullptr;
    uint32_t comm_rank = 0, comm_world = 1;
    // exchange buffers of myrio_score_topx_sharded (grow-only: a steady-state step never goes back to the allocator)
    myrio::DevBuf<uint32_t> xchg_bound;
    myrio::DevBuf<uint64_t> xchg_top, xchg_all;
    myrio::DevBuf<unsigned long long> clock_probes;  // myrio_bench_clock_probe: (cycles, ns) per slot
    uint64_t last_stats[3] = {0, 0, 0};
    uint64_t cluster_stats[2] = {0, 0};  // pair similarities, multiply-add terms of the last K6 call
    // pinned staging for small D2H reads
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
};

struct myrio_seqs {
    int kind = 0;  // 0 reads (2-bit + qual), 1 leaves (4-bit)
    uint64_t n = 0;
    myrio::DevBuf<uint64_t> words, word_off, base_off;
    myrio::DevBuf<uint8_t> qual;
    std::vector<uint64_t> h_base_off;  // host copy of the lengths (classification)
    std::vector<uint32_t> h_aux;       // leaves: number of ambiguity codes per sequence
};

struct myrio_svecs {
    uint64_t n_rows = 0, nnz = 0;
    int32_t vtype = MYRIO_VAL_F32;
    int aux_kind = 0;  // 0 none, 1 u64 (nb_hck), 2 f32 (rescale)
    myrio::DevBuf<uint64_t> row_off, keys;
    myrio::DevBuf<float> vals_f32;
    myrio::DevBuf<uint16_t> vals_u16;
    myrio::DevBuf<uint64_t> aux_u64;
    myrio::DevBuf<float> aux_f32;
};

namespace myrio {

// Makes `ctx` current for the calling thread: device + allocation stream.
struct ScopedCtx {
    int prev = 0, cur = 0;
    cudaStream_t prev_stream;
    explicit ScopedCtx(const myrio_ctx *ctx) : prev_stream(alloc_stream()) {
        cudaGetDevice(&prev);
        cur = ctx->device;
        if (prev != cur) cudaSetDevice(cur);
        alloc_stream() = ctx->stream;
    }
    ~ScopedCtx() {
        alloc_stream() = prev_stream;
        if (prev != cur) cudaSetDevice(prev);
    }
};

// Bookkeeping around every kernel launch: count it, optionally time it with
// CUDA events on the context's stream, and surface launch errors.
struct LaunchScope {
    myrio_ctx *ctx;
    bool timed;
    cudaEvent_t stop = nullptr;
    LaunchScope(myrio_ctx *c, const char *name) : ctx(c), timed(c->prof_on) {
        ctx->launches++;
        if (timed) {
            ProfEntry e;
            e.name = name;
            MYRIO_CK(cudaEventCreate(&e.start));
            MYRIO_CK(cudaEventCreate(&e.stop));
            MYRIO_CK(cudaEventRecord(e.start, ctx->stream));
            stop = e.stop;
            ctx->prof.push_back(e);
        }
    }
    void done() {
        if (timed) MYRIO_CK(cudaEventRecord(stop, ctx->stream));
        MYRIO_CK(cudaGetLastError());
    }
};

#define MYRIO_LAUNCH(ctx, name, kernel, grid, block, smem, ...)                      \
    do {                                                                             \
        ::myrio::LaunchScope ls__((ctx), (name));                                    \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);             \
        ls__.done();                                                                 \
    } while (0)

inline void sync(myrio_ctx *ctx) { MYRIO_CK(cudaStreamSynchronize(ctx->stream)); }

template <typename T>
inline void h2d(myrio_ctx *ctx, T *dst, const T *src, size_t n) {
    if (n) MYRIO_CK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
}
template <typename T>
inline void d2h(myrio_ctx *ctx, T *dst, const T *src, size_t n) {
    if (n) MYRIO_CK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
}

// --- device-wide primitives (scan.cu) ---
// out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 elements.
void exclusive_scan_u32_to_u64(myrio_ctx *ctx, const uint32_t *in, uint64_t *out, size_t n);
void exclusive_scan_u64(myrio_ctx *ctx, const uint64_t *in, uint64_t *out, size_t n);

// --- LSD radix sort of (u64 key, u64 payload) pairs, stable (radix_sort.cu) ---
// Sorts by key bits [0, key_bits).  keys/vals are overwritten with the result.
void radix_sort_pairs(myrio_ctx *ctx, uint64_t *keys, uint64_t *vals, uint64_t *keys_tmp,
                      uint64_t *vals_tmp, size_t n, int key_bits);

}  // namespace myrio
