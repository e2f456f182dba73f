// ingest.cu -- FASTQ records (ASCII sequence + quality lines) -> packed MyrSeq batch on
// the device: io::read_fastq's per-record conversion (myrio-core/src/io/mod.rs:46-48:
// Seq::<Dna>::from_str, quality = byte saturating_sub 33) and MyrSeq::pre_process
// (data/myrseq.rs:113-131: non-empty, len >= min_length, mean quality >= min_mean_qual,
// max quality <= max_qual).  The host only sees one flag per record.
#include <memory>

#include "kmer_count.cuh"

namespace myrio {

// one warp per record: validate the bases, decide keep/drop
__global__ void __launch_bounds__(256) fastq_scan_kernel(const uint8_t *__restrict__ seq,
                                                         const uint8_t *__restrict__ qual,
                                                         const uint64_t *__restrict__ base_off, uint64_t n,
                                                         uint64_t min_length, float min_mean_qual,
                                                         uint32_t max_qual, uint8_t *__restrict__ keep,
                                                         unsigned long long *__restrict__ bad_record) {
    const uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t b = base_off[r], len = base_off[r + 1] - b;
    uint32_t sum = 0, mx = 0;
    bool bad = false;
    for (uint64_t i = lane; i < len; i += 32) {
        const uint8_t c = seq[b + i];
        bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
        const uint32_t qa = qual[b + i];
        const uint32_t q = qa >= 33u ? qa - 33u : 0u;  // saturating_sub(33)
        sum += q;
        mx = max(mx, q);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        if (bad) atomicMin(bad_record, (unsigned long long)r);
        // the reference sums f32 sequentially; an integer sum converts to the same f32 while
        // every partial sum is < 2^24 (always true below 65 793 bases at Q <= 255)
        float fsum = (float)sum;
        if (len > 65536) {
            fsum = 0.0f;
            for (uint64_t i = 0; i < len; ++i) {
                const uint32_t qa = qual[b + i];
                fsum = __fadd_rn(fsum, (float)(qa >= 33u ? qa - 33u : 0u));
            }
        }
        const float mean = len ? __fdiv_rn(fsum, (float)len) : 0.0f;
        keep[r] = (len != 0 && len >= min_length && mean >= min_mean_qual && mx <= max_qual) ? 1 : 0;
    }
}

// one warp per kept record: 2-bit packing (base i at bits 2*(i%32) of word i/32) + Phred bytes
__global__ void __launch_bounds__(256) fastq_pack_kernel(const uint8_t *__restrict__ seq,
                                                         const uint8_t *__restrict__ qual,
                                                         const uint64_t *__restrict__ base_off,
                                                         const uint64_t *__restrict__ src_rec, uint64_t n_kept,
                                                         const uint64_t *__restrict__ out_word_off,
                                                         const uint64_t *__restrict__ out_base_off,
                                                         uint64_t *__restrict__ words,
                                                         uint8_t *__restrict__ out_qual) {
    const uint64_t j = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= n_kept) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t r = src_rec[j];
    const uint64_t b = base_off[r], len = base_off[r + 1] - b;
    const uint64_t ob = out_base_off[j], ow = out_word_off[j];
    const uint64_t nw = out_word_off[j + 1] - ow;
    for (uint64_t i = lane; i < len; i += 32) {
        const uint32_t qa = qual[b + i];
        out_qual[ob + i] = (uint8_t)(qa >= 33u ? qa - 33u : 0u);
    }
    // each lane builds whole words: word w covers bases [32w, 32w+32)
    for (uint64_t w = lane; w < nw; w += 32) {
        uint64_t word = 0;
        const uint64_t i0 = w * 32, i1 = min(len, i0 + 32);
        for (uint64_t i = i0; i < i1; ++i) {
            const uint8_t c = seq[b + i];
            const uint64_t code = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u;  // bio-seq Dna
            word |= code << (2 * (i - i0));
        }
        words[ow + w] = word;
    }
}

myrio_seqs *reads_from_fastq(myrio_ctx *ctx, const uint8_t *seq_ascii, const uint8_t *qual_ascii,
                             const uint64_t *base_off, uint64_t n, uint64_t min_length, float min_mean_qual,
                             uint8_t max_qual, uint8_t *keep_host) {
    if (n && (!seq_ascii || !qual_ascii || !base_off)) fail(MYRIO_ERR_BAD_ARG, "NULL input array");
    auto s = std::make_unique<myrio_seqs>();
    s->kind = 0;
    const uint64_t total = n ? base_off[n] : 0;
    std::vector<uint8_t> keep(n, 0);
    DevBuf<uint8_t> d_seq(total + 1), d_qual(total + 1), d_keep(std::max<uint64_t>(n, 1));
    DevBuf<uint64_t> d_off(n + 1);
    DevBuf<unsigned long long> d_bad(1);
    if (n) {
        h2d(ctx, d_seq.p, seq_ascii, total);
        h2d(ctx, d_qual.p, qual_ascii, total);
        h2d(ctx, d_off.p, base_off, n + 1);
        MYRIO_CK(cudaMemsetAsync(d_bad.p, 0xFF, sizeof(unsigned long long), ctx->stream));
        MYRIO_LAUNCH(ctx, "k0_fastq_scan", fastq_scan_kernel, (unsigned)((n + 7) / 8), 256, 0, d_seq.p, d_qual.p,
                     d_off.p, n, min_length, min_mean_qual, (uint32_t)max_qual, d_keep.p, d_bad.p);
        unsigned long long bad = 0;
        d2h(ctx, &bad, d_bad.p, 1);
        d2h(ctx, keep.data(), d_keep.p, n);
        sync(ctx);
        if (bad != ~0ull)
            fail(MYRIO_ERR_BAD_BASE, "record %llu holds a base other than A, C, G, T (read_fastq rejects the file)", bad);
    }
    // layout of the kept reads (host: one pass over the flags)
    std::vector<uint64_t> src_rec, word_off(1, 0);
    s->h_base_off.assign(1, 0);
    for (uint64_t r = 0; r < n; ++r) {
        if (!keep[r]) continue;
        const uint64_t len = base_off[r + 1] - base_off[r];
        src_rec.push_back(r);
        word_off.push_back(word_off.back() + ((len + 31) / 32 + 1) / 2 * 2);
        s->h_base_off.push_back(s->h_base_off.back() + len);
    }
    const uint64_t nk = src_rec.size();
    s->n = nk;
    s->words.alloc(word_off.back() + 4);
    s->word_off.alloc(nk + 1);
    s->base_off.alloc(nk + 1);
    s->qual.alloc(s->h_base_off.back() + 32);
    MYRIO_CK(cudaMemsetAsync(s->words.p + word_off.back(), 0, 4 * sizeof(uint64_t), ctx->stream));
    MYRIO_CK(cudaMemsetAsync(s->qual.p + s->h_base_off.back(), 0, 32, ctx->stream));
    h2d(ctx, s->word_off.p, word_off.data(), nk + 1);
    h2d(ctx, s->base_off.p, s->h_base_off.data(), nk + 1);
    if (nk) {
        DevBuf<uint64_t> d_src(nk);
        h2d(ctx, d_src.p, src_rec.data(), nk);
        MYRIO_LAUNCH(ctx, "k0_fastq_pack", fastq_pack_kernel, (unsigned)((nk + 7) / 8), 256, 0, d_seq.p, d_qual.p,
                     d_off.p, d_src.p, nk, s->word_off.p, s->base_off.p, s->words.p, s->qual.p);
        sync(ctx);  // src_rec / word_off (host vectors) are read by the async copies above
    } else {
        sync(ctx);
    }
    if (keep_host)
        for (uint64_t r = 0; r < n; ++r) keep_host[r] = keep[r];
    return s.release();
}

// ---------------------------------------------------------------------------------------------
// FASTA records -> packed Seq<Iupac> leaves on the device: the `Seq::from_str(&string_seq)` of
// TaxTreeStore::load_from_fasta_string (myrio-core/src/tax/store.rs:178-188).  bio-seq's Iupac codec:
// A=8 C=4 G=2 T=1, an ambiguity code is the OR of its bases (R=A|G ... N=15), '-' is the gap X=0; upper
// case only.  A record holding any other byte is reported and SKIPPED by the reference (store.rs:181-187:
// eprintln + continue, no payload is created), so it gets keep = 0 here and the kept records are
// renumbered in input order, like payload_id (store.rs:190-192).

__device__ __forceinline__ uint32_t iupac_code(uint8_t c) {
    switch (c) {
        case 'A': return 8;   case 'C': return 4;   case 'G': return 2;   case 'T': return 1;
        case 'R': return 10;  case 'Y': return 5;   case 'S': return 6;   case 'W': return 9;
        case 'K': return 3;   case 'M': return 12;  case 'B': return 7;   case 'D': return 11;
        case 'H': return 13;  case 'V': return 14;  case 'N': return 15;  case '-': return 0;
        default: return 0xFFu;
    }
}

// one warp per record: validity + number of ambiguity codes (scratch sizing of K2)
__global__ void __launch_bounds__(256) fasta_scan_kernel(const uint8_t *__restrict__ seq,
                                                         const uint64_t *__restrict__ base_off, uint64_t n,
                                                         uint8_t *__restrict__ keep, uint32_t *__restrict__ n_amb) {
    const uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t b = base_off[r], len = base_off[r + 1] - b;
    bool bad = false;
    uint32_t amb = 0;
    for (uint64_t i = lane; i < len; i += 32) {
        const uint32_t c = iupac_code(seq[b + i]);
        bad |= c == 0xFFu;
        amb += (c != 0xFFu && __popc(c) > 1) ? 1u : 0u;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) amb += __shfl_xor_sync(0xffffffffu, amb, d);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        keep[r] = bad ? 0 : 1;
        n_amb[r] = amb;
    }
}

// one warp per kept record: 4-bit packing (symbol i at bits 4*(i%16) of word i/16)
__global__ void __launch_bounds__(256) fasta_pack_kernel(const uint8_t *__restrict__ seq,
                                                         const uint64_t *__restrict__ base_off,
                                                         const uint64_t *__restrict__ src_rec, uint64_t n_kept,
                                                         const uint64_t *__restrict__ out_word_off,
                                                         uint64_t *__restrict__ words) {
    const uint64_t j = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= n_kept) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t r = src_rec[j];
    const uint64_t b = base_off[r], len = base_off[r + 1] - b;
    const uint64_t ow = out_word_off[j], nw = out_word_off[j + 1] - ow;
    for (uint64_t w = lane; w < nw; w += 32) {
        uint64_t word = 0;
        const uint64_t i0 = w * 16, i1 = min(len, i0 + 16);
        for (uint64_t i = i0; i < i1; ++i) word |= (uint64_t)(iupac_code(seq[b + i]) & 15u) << (4 * (i - i0));
        words[ow + w] = word;
    }
}

myrio_seqs *leaves_from_fasta(myrio_ctx *ctx, const uint8_t *seq_ascii, const uint64_t *base_off, uint64_t n,
                              uint8_t *keep_host) {
    if (n && (!seq_ascii || !base_off)) fail(MYRIO_ERR_BAD_ARG, "NULL input array");
    auto s = std::make_unique<myrio_seqs>();
    s->kind = 1;
    const uint64_t total = n ? base_off[n] : 0;
    std::vector<uint8_t> keep(n, 0);
    std::vector<uint32_t> amb(n, 0);
    DevBuf<uint8_t> d_seq(total + 1), d_keep(std::max<uint64_t>(n, 1));
    DevBuf<uint32_t> d_amb(std::max<uint64_t>(n, 1));
    DevBuf<uint64_t> d_off(n + 1);
    if (n) {
        h2d(ctx, d_seq.p, seq_ascii, total);
        h2d(ctx, d_off.p, base_off, n + 1);
        MYRIO_LAUNCH(ctx, "k0_fasta_scan", fasta_scan_kernel, (unsigned)((n + 7) / 8), 256, 0, d_seq.p, d_off.p, n,
                     d_keep.p, d_amb.p);
        d2h(ctx, keep.data(), d_keep.p, n);
        d2h(ctx, amb.data(), d_amb.p, n);
        sync(ctx);
    }
    std::vector<uint64_t> src_rec, word_off(1, 0);
    s->h_base_off.assign(1, 0);
    for (uint64_t r = 0; r < n; ++r) {
        if (!keep[r]) continue;
        const uint64_t len = base_off[r + 1] - base_off[r];
        src_rec.push_back(r);
        word_off.push_back(word_off.back() + ((len + 15) / 16 + 1) / 2 * 2);
        s->h_base_off.push_back(s->h_base_off.back() + len);
        s->h_aux.push_back(amb[r]);
    }
    const uint64_t nk = src_rec.size();
    s->n = nk;
    s->words.alloc(word_off.back() + 4);
    s->word_off.alloc(nk + 1);
    s->base_off.alloc(nk + 1);
    MYRIO_CK(cudaMemsetAsync(s->words.p + word_off.back(), 0, 4 * sizeof(uint64_t), ctx->stream));
    h2d(ctx, s->word_off.p, word_off.data(), nk + 1);
    h2d(ctx, s->base_off.p, s->h_base_off.data(), nk + 1);
    if (nk) {
        DevBuf<uint64_t> d_src(nk);
        h2d(ctx, d_src.p, src_rec.data(), nk);
        MYRIO_LAUNCH(ctx, "k0_fasta_pack", fasta_pack_kernel, (unsigned)((nk + 7) / 8), 256, 0, d_seq.p, d_off.p,
                     d_src.p, nk, s->word_off.p, s->words.p);
    }
    sync(ctx);  // the host vectors above are read by the async copies
    if (keep_host)
        for (uint64_t r = 0; r < n; ++r) keep_host[r] = keep[r];
    return s.release();
}

}  // namespace myrio
