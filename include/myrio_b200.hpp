// myrio_b200.hpp -- header-only C++17 host mirror of the `myrio-core` API for the hot
// path, over the C ABI of myrio_b200.h.  The reference is compiled Rust
// (myrio-core/src/prelude.rs:1-12); this is the same surface for a compiled-language
// caller: same names, argument meaning and error behaviour (an ABI error code that
// stands for a reference panic throws myrio::Panic, the others throw myrio::Error).
//
//   MyrSeqs::compute_kmer_counts          data/myrseq.rs:76-111 (batch of MyrSeq)
//   SFVecs::into_normalized_l2            data/sparse.rs:385-388
//   TaxTreeStore::compute_and_append_kmer_counts / _overwrite_   tax/store.rs:265-308
//   TaxTreeCompute::from_store_tree       tax/compute.rs:31-120 (leaf vectors -> index)
//   TaxTreeResults::from_compute_tree     tax/results.rs:68-93 (+ cut() top-x :310-329)
//   cluster / compute_cluster_centroid    clustering.rs:55-270 / :44-53
#ifndef MYRIO_B200_HPP
#define MYRIO_B200_HPP

#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "myrio_b200.h"

namespace myrio {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
};
// the reference panics in these situations (invalid k, Phred >= 128, empty clustering, ...)
struct Panic : Error {
    using Error::Error;
};

inline void check(int32_t rc) {
    if (rc == MYRIO_OK) return;
    const std::string msg = myrio_last_error();
    if (rc == MYRIO_ERR_INVALID_K || rc == MYRIO_ERR_BAD_QUALITY || rc == MYRIO_ERR_EMPTY || rc == MYRIO_ERR_NONFINITE)
        throw Panic(rc, msg);
    throw Error(rc, msg);
}

using Float = float;  // data/sparse.rs:13
enum class Similarity : int32_t { Cosine = 0, JacardTanimoto = 1, Overlap = 2 };  // similarity.rs:26-33

class Context {
  public:
    explicit Context(int device = 0) { check(myrio_ctx_create(device, &h_)); }
    ~Context() { myrio_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    myrio_ctx *get() const { return h_; }
    uint64_t launch_count() const { return myrio_ctx_launch_count(h_); }

  private:
    myrio_ctx *h_ = nullptr;
};

// One SparseVec<Float> on the host (data/sparse.rs:19-26)
struct SFVec {
    std::vector<uint64_t> keys;
    std::vector<Float> values;
    size_t count() const { return keys.size(); }
};

// A device-resident batch of SparseVec<T>
class SFVecs {
  public:
    SFVecs(const Context &ctx, myrio_svecs *h) : ctx_(&ctx), h_(h) {}
    SFVecs(SFVecs &&o) noexcept : ctx_(o.ctx_), h_(std::exchange(o.h_, nullptr)) {}
    SFVecs(const SFVecs &) = delete;
    ~SFVecs() { myrio_svecs_free(h_); }
    static SFVecs upload(const Context &ctx, const std::vector<SFVec> &rows) {
        std::vector<uint64_t> off(rows.size() + 1, 0), keys;
        std::vector<Float> vals;
        for (size_t i = 0; i < rows.size(); ++i) {
            off[i + 1] = off[i] + rows[i].count();
            keys.insert(keys.end(), rows[i].keys.begin(), rows[i].keys.end());
            vals.insert(vals.end(), rows[i].values.begin(), rows[i].values.end());
        }
        myrio_svecs *h = nullptr;
        check(myrio_svecs_upload(ctx.get(), rows.size(), off.data(), keys.data(), vals.data(), MYRIO_VAL_F32, &h));
        return SFVecs(ctx, h);
    }
    uint64_t n_rows() const {
        uint64_t n = 0;
        check(myrio_svecs_info(h_, &n, nullptr, nullptr));
        return n;
    }
    SFVecs into_normalized_l2() const {
        myrio_svecs *o = nullptr;
        check(myrio_svecs_normalize(ctx_->get(), h_, &o));
        return SFVecs(*ctx_, o);
    }
    // rows as host SFVec (f32 batches); aux = nb_hck per row when the batch carries it
    std::vector<SFVec> download(std::vector<uint64_t> *nb_hck = nullptr) const {
        uint64_t n = 0, nnz = 0;
        int32_t vt = 0;
        check(myrio_svecs_info(h_, &n, &nnz, &vt));
        if (vt != MYRIO_VAL_F32) throw Error(MYRIO_ERR_BAD_ARG, "download(): not an f32 batch");
        std::vector<uint64_t> off(n + 1), keys(nnz);
        std::vector<Float> vals(nnz);
        if (nb_hck) nb_hck->assign(n, 0);
        check(myrio_svecs_download(ctx_->get(), h_, off.data(), keys.data(), vals.data(),
                                   nb_hck ? nb_hck->data() : nullptr));
        std::vector<SFVec> rows(n);
        for (uint64_t i = 0; i < n; ++i) {
            rows[i].keys.assign(keys.begin() + off[i], keys.begin() + off[i + 1]);
            rows[i].values.assign(vals.begin() + off[i], vals.begin() + off[i + 1]);
        }
        return rows;
    }
    myrio_svecs *get() const { return h_; }
    const Context &ctx() const { return *ctx_; }

  private:
    const Context *ctx_;
    myrio_svecs *h_;
};

// A batch of MyrSeq (data/myrseq.rs:21-33): ASCII "ACGT" sequences + Phred qualities
class MyrSeqs {
  public:
    MyrSeqs(const Context &ctx, const std::vector<std::string> &seqs, const std::vector<std::vector<uint8_t>> &quals)
        : ctx_(&ctx), n_(seqs.size()) {
        std::vector<uint64_t> words, word_off(n_ + 1, 0), base_off(n_ + 1, 0);
        std::vector<uint8_t> qual;
        for (size_t r = 0; r < n_; ++r) {
            const std::string &s = seqs[r];
            if (quals[r].size() != s.size()) throw Error(MYRIO_ERR_BAD_ARG, "sequence / quality length mismatch");
            const size_t nw = ((s.size() + 31) / 32 + 1) / 2 * 2;  // even word count: 16-byte aligned reads
            words.resize(word_off[r] + nw, 0);
            for (size_t i = 0; i < s.size(); ++i) {
                uint64_t c;
                switch (s[i]) {  // bio-seq Dna: A=0 C=1 G=2 T=3
                    case 'A': c = 0; break;
                    case 'C': c = 1; break;
                    case 'G': c = 2; break;
                    case 'T': c = 3; break;
                    default: throw Error(MYRIO_ERR_BAD_ARG, "non-ACGT base (io/mod.rs:46-47 rejects the file)");
                }
                words[word_off[r] + i / 32] |= c << (2 * (i % 32));
            }
            word_off[r + 1] = word_off[r] + nw;
            base_off[r + 1] = base_off[r] + s.size();
            qual.insert(qual.end(), quals[r].begin(), quals[r].end());
        }
        check(myrio_reads_upload(ctx.get(), words.data(), word_off.data(), base_off.data(), qual.data(), n_, &h_));
    }
    ~MyrSeqs() { myrio_seqs_free(h_); }
    MyrSeqs(const MyrSeqs &) = delete;
    // (SFVec, nb_hck) for every read: the batch keeps nb_hck as its aux column
    SFVecs compute_kmer_counts(size_t k, Float cutoff) const {
        myrio_svecs *o = nullptr;
        check(myrio_count_reads(ctx_->get(), h_, (uint32_t)k, cutoff, &o));
        return SFVecs(*ctx_, o);
    }
    size_t len() const { return n_; }
    myrio_seqs *get() const { return h_; }
    const Context &ctx() const { return *ctx_; }

  private:
    const Context *ctx_;
    size_t n_;
    myrio_seqs *h_ = nullptr;
};

// Leaf sequences + cached counts of a .myrtree (tax/store.rs:28-50); IUPAC ASCII input
class TaxTreeStore {
  public:
    TaxTreeStore(const Context &ctx, const std::vector<std::string> &leaves) : ctx_(&ctx), n_(leaves.size()) {
        std::vector<uint64_t> words, word_off(n_ + 1, 0), base_off(n_ + 1, 0);
        for (size_t r = 0; r < n_; ++r) {
            const std::string &s = leaves[r];
            const size_t nw = ((s.size() + 15) / 16 + 1) / 2 * 2;
            words.resize(word_off[r] + nw, 0);
            for (size_t i = 0; i < s.size(); ++i)
                words[word_off[r] + i / 16] |= (uint64_t)iupac_code(s[i]) << (4 * (i % 16));
            word_off[r + 1] = word_off[r] + nw;
            base_off[r + 1] = base_off[r] + s.size();
        }
        check(myrio_leaves_upload(ctx.get(), words.data(), word_off.data(), base_off.data(), n_, &h_));
    }
    ~TaxTreeStore() {
        for (auto *c : kmer_store_counts_vec) myrio_svecs_free(c);
        myrio_seqs_free(h_);
    }
    TaxTreeStore(const TaxTreeStore &) = delete;
    void compute_and_append_kmer_counts(size_t k, size_t nb_resamples, uint64_t seed = 0) {
        myrio_svecs *o = nullptr;
        check(myrio_count_leaves(ctx_->get(), h_, (uint32_t)k, (uint32_t)nb_resamples, seed, 0, &o));
        kmer_store_counts_vec.push_back(o);
        k_precomputed.push_back(k);
    }
    void compute_and_overwrite_kmer_counts(size_t k, size_t nb_resamples, uint64_t seed = 0) {
        for (auto *c : kmer_store_counts_vec) myrio_svecs_free(c);
        kmer_store_counts_vec.clear();
        k_precomputed.clear();
        compute_and_append_kmer_counts(k, nb_resamples, seed);
    }
    // bio-seq Iupac: A=8 C=4 G=2 T=1, ambiguity codes are ORs, '-' (X) = 0
    static uint8_t iupac_code(char c) {
        switch (c) {
            case 'A': return 8; case 'C': return 4; case 'G': return 2; case 'T': return 1;
            case 'R': return 10; case 'Y': return 5; case 'S': return 6; case 'W': return 9;
            case 'K': return 3; case 'M': return 12; case 'B': return 7; case 'D': return 11;
            case 'H': return 13; case 'V': return 14; case 'N': return 15; case '-': return 0;
            default: throw Error(MYRIO_ERR_BAD_ARG, "not an IUPAC nucleotide code");
        }
    }
    std::vector<size_t> k_precomputed;
    std::vector<myrio_svecs *> kmer_store_counts_vec;
    size_t n_leaves() const { return n_; }
    const Context &ctx() const { return *ctx_; }

  private:
    const Context *ctx_;
    size_t n_;
    myrio_seqs *h_ = nullptr;
};

enum class CacheOptions { Enabled, Disabled };  // tax/compute.rs:22-27

class TaxTreeCompute {
  public:
    // tax/compute.rs:80-118: pick or compute the counts for search_k, normalise, index
    static TaxTreeCompute from_store_tree(TaxTreeStore &ttstore, size_t fasta_nb_resamples, size_t search_k,
                                          CacheOptions cache_opt = CacheOptions::Enabled) {
        size_t idx = 0;
        bool found = false;
        for (size_t i = 0; i < ttstore.k_precomputed.size(); ++i)
            if (ttstore.k_precomputed[i] == search_k) {
                idx = i;
                found = true;
                break;
            }
        if (!found) {
            if (cache_opt == CacheOptions::Enabled) {
                ttstore.compute_and_append_kmer_counts(search_k, fasta_nb_resamples);
                idx = ttstore.k_precomputed.size() - 1;
            } else {
                ttstore.compute_and_overwrite_kmer_counts(search_k, fasta_nb_resamples);
                idx = 0;
            }
        }
        myrio_svecs *norm = nullptr;
        check(myrio_svecs_normalize(ttstore.ctx().get(), ttstore.kmer_store_counts_vec[idx], &norm));
        myrio_index *ix = nullptr;
        const int32_t rc = myrio_index_build(ttstore.ctx().get(), norm, 0, 0, &ix);
        myrio_svecs_free(norm);
        check(rc);
        TaxTreeCompute out(ttstore.ctx(), ix, ttstore.n_leaves());
        out.counts_ = ttstore.kmer_store_counts_vec[idx];
        return out;
    }
    TaxTreeCompute(TaxTreeCompute &&o) noexcept
        : ctx_(o.ctx_), h_(std::exchange(o.h_, nullptr)), n_(o.n_), counts_(o.counts_),
          leaf_vecs_(std::exchange(o.leaf_vecs_, nullptr)) {}
    ~TaxTreeCompute() {
        myrio_index_free(h_);
        if (leaf_vecs_) myrio_svecs_free(leaf_vecs_);
    }
    myrio_index *get() const { return h_; }
    // the reference's Box<[SFVec]> of normalised leaf vectors (tax/compute.rs:100-108), materialised
    // only when a merge-join similarity (Overlap) needs them; the store must outlive this object
    const myrio_svecs *leaf_vectors() const {
        if (!leaf_vecs_) check(myrio_svecs_normalize(ctx_->get(), counts_, &leaf_vecs_));
        return leaf_vecs_;
    }
    size_t payloads_len() const { return n_; }
    const Context &ctx() const { return *ctx_; }

  private:
    TaxTreeCompute(const Context &ctx, myrio_index *h, size_t n) : ctx_(&ctx), h_(h), n_(n) {}
    const Context *ctx_;
    myrio_index *h_;
    size_t n_;
    const myrio_svecs *counts_ = nullptr;
    mutable myrio_svecs *leaf_vecs_ = nullptr;
};

// the scoring part of TaxTreeResults::from_compute_tree (tax/results.rs:82-93) and cut()
struct TaxTreeResults {
    std::vector<Float> payloads;                              // SimScore per payload_id (one query)
    std::vector<std::pair<uint32_t, Float>> best;             // cut(nb_best): (payload_id, score)
    static TaxTreeResults from_compute_tree(const SFVec &query, const TaxTreeCompute &ttcompute, Similarity similarity,
                                            size_t nb_best_analysis) {
        const Context &ctx = ttcompute.ctx();
        SFVecs q = SFVecs::upload(ctx, {query});
        TaxTreeResults r;
        r.payloads.assign(ttcompute.payloads_len(), 0.0f);
        // Overlap sums maxima over the union of the keys (similarity.rs:172-182): merge-join over the
        // leaf vectors like the reference; Cosine / JacardTanimoto run on the inverted index
        const bool direct = similarity == Similarity::Overlap;
        std::vector<Float> ts(nb_best_analysis);
        std::vector<uint32_t> ti(nb_best_analysis);
        if (direct) {
            check(myrio_score_all_direct(ctx.get(), ttcompute.leaf_vectors(), q.get(), 0, 1, (int32_t)similarity,
                                         r.payloads.data()));
            if (nb_best_analysis)
                check(myrio_score_topx_direct(ctx.get(), ttcompute.leaf_vectors(), 0, q.get(), 0, 1,
                                              (uint32_t)nb_best_analysis, (int32_t)similarity, ts.data(), ti.data()));
        } else {
            check(myrio_score_all(ctx.get(), ttcompute.get(), q.get(), 0, 1, (int32_t)similarity, r.payloads.data()));
            if (nb_best_analysis)
                check(myrio_score_topx(ctx.get(), ttcompute.get(), q.get(), 0, 1, (uint32_t)nb_best_analysis,
                                       (int32_t)similarity, ts.data(), ti.data()));
        }
        for (size_t j = 0; j < nb_best_analysis; ++j)
            if (ti[j] != 0xFFFFFFFFu) r.best.emplace_back(ti[j], ts[j]);
        return r;
    }
};

// clustering.rs:15-36
struct ClusteringParameters {
    size_t k = 6;
    Float t1 = 0.2f;
    size_t t2 = 20;
    Similarity similarity = Similarity::Cosine;
    std::vector<SFVec> intial_centroids;  // (sic)
    size_t expected_nb_of_clusters = 1;
    Float eta_improvement = 1e-4f;
    size_t nb_iters_max = 20;
    std::optional<Float> silhouette_trimming;
};

// clustering.rs:38-41 as index lists into the input batch
struct ClusteringOutput {
    std::vector<std::vector<size_t>> clusters;
    std::vector<size_t> rejected;  // t2 rejections first, then silhouette rejections
};

inline ClusteringOutput cluster(const MyrSeqs &myrseqs, const ClusteringParameters &params) {
    myrio_cluster_params p{};
    p.k = (uint32_t)params.k;
    p.t1 = params.t1;
    p.t2 = params.t2;
    p.similarity = (int32_t)params.similarity;
    p.expected_nb_of_clusters = params.expected_nb_of_clusters;
    p.eta_improvement = params.eta_improvement;
    p.nb_iters_max = params.nb_iters_max;
    p.has_silhouette_trimming = params.silhouette_trimming.has_value();
    p.silhouette_trimming = params.silhouette_trimming.value_or(0.0f);
    std::optional<SFVecs> init;
    if (!params.intial_centroids.empty()) init.emplace(SFVecs::upload(myrseqs.ctx(), params.intial_centroids));
    std::vector<int32_t> assign(myrseqs.len() + 1);
    check(myrio_cluster(myrseqs.ctx().get(), myrseqs.get(), &p, init ? init->get() : nullptr, assign.data(), nullptr,
                        nullptr));
    ClusteringOutput out;
    out.clusters.resize(std::max(params.expected_nb_of_clusters, params.intial_centroids.size()));
    std::vector<size_t> rej_sil;
    for (size_t i = 0; i < myrseqs.len(); ++i) {
        if (assign[i] == -1)
            out.rejected.push_back(i);
        else if (assign[i] == -2)
            rej_sil.push_back(i);
        else
            out.clusters[(size_t)assign[i]].push_back(i);
    }
    out.rejected.insert(out.rejected.end(), rej_sil.begin(), rej_sil.end());
    return out;
}

// clustering.rs:44-53 over rows `members` of a raw count batch
inline SFVec compute_cluster_centroid(const SFVecs &counts, const std::vector<uint64_t> &members) {
    myrio_svecs *o = nullptr;
    check(myrio_compute_cluster_centroid(counts.ctx().get(), counts.get(), members.data(), members.size(), &o));
    SFVecs c(counts.ctx(), o);
    return c.download().at(0);
}

}  // namespace myrio
#endif  // MYRIO_B200_HPP
