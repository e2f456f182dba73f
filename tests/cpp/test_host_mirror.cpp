// The reference's own unit tests for the hot path, run through the C++ host mirror
// (include/myrio_b200.hpp) on the GPU.  Built and run by tests/test_gpu_cpp_mirror.py.
#include <cmath>
#include <cstdio>
#include <limits>

#include "myrio_b200.hpp"

using namespace myrio;

static int failures = 0;
#define ASSERT_FLOAT_EQ(lhs, rhs)                                                               \
    do { /* myrio-core/src/lib.rs:11-17: |lhs - rhs| < f32::EPSILON */                          \
        if (!(std::fabs((lhs) - (rhs)) < std::numeric_limits<float>::epsilon())) {              \
            std::printf("FAIL %s:%d: %s=%.9g vs %s=%.9g\n", __FILE__, __LINE__, #lhs, (double)(lhs), #rhs, \
                        (double)(rhs));                                                         \
            ++failures;                                                                         \
        }                                                                                       \
    } while (0)
#define ASSERT_TRUE(c)                                                     \
    do {                                                                   \
        if (!(c)) {                                                        \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c);       \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

static float dot(const SFVec &a, const SFVec &b) {  // data/sparse.rs:317-360 on the host, for the checks
    float d = 0.0f;
    size_t i = 0, j = 0;
    while (i < a.count() && j < b.count()) {
        if (a.keys[i] < b.keys[j]) ++i;
        else if (a.keys[i] > b.keys[j]) ++j;
        else { d = d + a.values[i] * b.values[j]; ++i; ++j; }
    }
    return d;
}

int main() {
    Context ctx(0);

    std::fprintf(stderr, "[1] simfunc_test\n");
    {  // similarity.rs:217-232 simfunc_test: ACTG vs ACTC, k=2, cutoff f32::MIN, Q=0
        MyrSeqs seqs(ctx, {"ACTG", "ACTC"}, {{0, 0, 0, 0}, {0, 0, 0, 0}});
        SFVecs counts = seqs.compute_kmer_counts(2, std::numeric_limits<float>::lowest());
        std::vector<uint64_t> nb_hck;
        auto rows = counts.download(&nb_hck);
        ASSERT_TRUE(nb_hck[0] == 6 && nb_hck[1] == 6);
        const SFVec &a = rows[0], &b = rows[1];
        float na = 0, nb = 0;
        for (float v : a.values) na += v * v;
        for (float v : b.values) nb += v * v;
        const float cosine = dot(a, b) / (std::sqrt(na) * std::sqrt(nb));
        ASSERT_FLOAT_EQ(cosine, 2.0f / 3.0f);
        auto norm = counts.into_normalized_l2().download();
        ASSERT_FLOAT_EQ(dot(norm[0], norm[1]), cosine);  // cosine_pre_normalized(normalised) == cosine
    }
    std::fprintf(stderr, "[2] invalid k\n");
    {  // invalid k panics in the reference (myrio-proc/src/lib.rs:70)
        MyrSeqs seqs(ctx, {"ACGTACGT"}, {{40, 40, 40, 40, 40, 40, 40, 40}});
        bool panicked = false;
        try { seqs.compute_kmer_counts(33, 0.1f); } catch (const Panic &) { panicked = true; }
        ASSERT_TRUE(panicked);
    }
    std::fprintf(stderr, "[3] tree new + run\n");
    {  // tree new + run on a tiny tree: from_store_tree -> from_compute_tree -> cut
        std::vector<std::string> leaves = {
            "ACGTTGCAAGCTTGCATGCAAGTCGATCGATTAGCTAGCTAGGATCGATCGGATTACGATCGAT",
            "ACGTTGCAAGCTTGCATGCAAGTCGATCGATTAGCTAGCTAGGATCGATCGGATTACGATCGAA",
            "TTTTGGGGCCCCAAAATTTTGGGGCCCCAAAATTTTGGGGCCCCAAAATTTTGGGGCCCCAAAA",
            "ACGT-NACGT"};  // a gap and an ambiguity code; shorter than k -> sentinel ({0:1}, 1.0)
        TaxTreeStore store(ctx, leaves);
        store.compute_and_append_kmer_counts(18, 32);
        ASSERT_TRUE(store.k_precomputed.size() == 1 && store.k_precomputed[0] == 18);
        TaxTreeCompute ttc = TaxTreeCompute::from_store_tree(store, 32, 18);
        MyrSeqs reads(ctx, {leaves[1]}, {std::vector<uint8_t>(leaves[1].size(), 40)});
        auto q = reads.compute_kmer_counts(18, 0.1f).into_normalized_l2().download();
        TaxTreeResults res = TaxTreeResults::from_compute_tree(q[0], ttc, Similarity::Cosine, 3);
        ASSERT_TRUE(res.payloads.size() == 4);
        ASSERT_TRUE(res.best.size() == 3 && res.best[0].first == 1);
        ASSERT_TRUE(std::fabs(res.best[0].second - 1.0f) < 1e-5f);
        ASSERT_TRUE(res.payloads[2] == 0.0f && res.payloads[3] == 0.0f);
        ASSERT_TRUE(res.best[0].second == res.payloads[1] && res.best[1].second == res.payloads[0]);
        // Overlap (similarity.rs:172-182) takes the merge-join path over the leaf vectors:
        // a leaf scored against itself has sum_min == sum_max
        TaxTreeResults ov = TaxTreeResults::from_compute_tree(q[0], ttc, Similarity::Overlap, 3);
        ASSERT_TRUE(ov.best.size() == 3 && ov.best[0].first == 1 && ov.best[0].second == 1.0f);
        ASSERT_TRUE(ov.payloads[2] == 0.0f && ov.best[0].second == ov.payloads[1]);
    }
    std::fprintf(stderr, "[4] cluster\n");
    {  // cluster(): two obviously different read families, no initial centroids
        std::vector<std::string> seqs;
        std::vector<std::vector<uint8_t>> quals;
        const std::string fa = "ACGTTGCAAGCTTGCATGCAAGTCGATCGATTAGCTAGCTAGGATCGATCGGATTACGATCGATCCGATATCGGCTA";
        const std::string fb = "TTTTGGGGCCCCAAAATGTGTGTCACACAGAGAGCTCTCTAAAACCCCGGGGTTTTACACGTGTCAGTCAGTCAGTAA";
        for (int i = 0; i < 12; ++i) {
            std::string s = (i % 2 ? fb : fa);
            s[5 + i] = s[5 + i] == 'A' ? 'C' : 'A';
            seqs.push_back(s);
            quals.emplace_back(s.size(), 40);
        }
        MyrSeqs reads(ctx, seqs, quals);
        ClusteringParameters p;
        p.k = 6;
        p.expected_nb_of_clusters = 2;
        ClusteringOutput out = cluster(reads, p);
        ASSERT_TRUE(out.clusters.size() == 2 && out.rejected.empty());
        ASSERT_TRUE(out.clusters[0].size() == 6 && out.clusters[1].size() == 6);
        for (auto &c : out.clusters)
            for (size_t i : c) ASSERT_TRUE(i % 2 == c[0] % 2);
        SFVecs counts = reads.compute_kmer_counts(6, 0.2f);
        SFVec cen = compute_cluster_centroid(counts, {0, 2, 4});
        float n2 = 0;
        for (float v : cen.values) n2 += v * v;
        ASSERT_TRUE(std::fabs(n2 - 1.0f) < 1e-5f);
        ClusteringParameters bad;
        bad.expected_nb_of_clusters = 0;  // panics in the reference (clustering.rs:71-73)
        bool threw = false;
        try { cluster(reads, bad); } catch (const Error &) { threw = true; }
        ASSERT_TRUE(threw);
    }
    std::fprintf(stderr, "[5] done\n");
    if (failures)
        std::printf("%d FAILURES\n", failures);
    else
        std::printf("host mirror ok (launches: %llu)\n", (unsigned long long)ctx.launch_count());
    return failures ? 1 : 0;
}
