/* The Myrio hot path through the C ABI from plain C (what a cgo / Rust -sys / JNI binding does):
 * pack three reference leaves and two reads the way bio-seq stores them, count k-mers, build the
 * index, score, take the top hit.  Exit code 0 iff every read's best leaf is the leaf it came from.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -o c_abi_example -Lmyrio_b200 -lmyrio_b200 \
 *       -Wl,-rpath,$PWD/myrio_b200
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "myrio_b200.h"

#define CHECK(call)                                                                    \
    do {                                                                               \
        int32_t rc__ = (call);                                                         \
        if (rc__ != MYRIO_OK) {                                                        \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc__, myrio_last_error());   \
            return 2;                                                                  \
        }                                                                              \
    } while (0)

enum { N_LEAVES = 3, N_READS = 2, K = 8 };

static const char *LEAVES[N_LEAVES] = {
    "ACGTTGCAAGCTTGCATGCAAGTCGATCGATTAGCTAGCTAGGATCGATCGGATTACGATCGATCCGATATCGGCTAAGCTTACG",
    "TTTTGGGGCCCCAAAATGTGTGTCACACAGAGAGCTCTCTAAAACCCCGGGGTTTTACACGTGTCAGTCAGTCAGTAACCGGTTA",
    "GATTACAGATTACAGGCCTTAAGGCCTTAACTGACTGACTGAAATTTCCCGGGATATATCGCGCGTATACGCGATCGTAGCTAGC"};
/* reads: a window of leaf 1 and a window of leaf 2 */
static const int READ_SRC[N_READS] = {1, 2};
static const int READ_BEGIN[N_READS] = {10, 20};
static const int READ_LEN[N_READS] = {50, 55};

static unsigned dna2(char c) { return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u; }    /* bio-seq Dna */
static unsigned iupac4(char c) { return c == 'A' ? 8u : c == 'C' ? 4u : c == 'G' ? 2u : 1u; }  /* bio-seq Iupac */

int main(void) {
    myrio_ctx *ctx = NULL;
    CHECK(myrio_ctx_create(0, &ctx));

    /* ---- leaves: Seq<Iupac>, symbol i at bits 4*(i%16) of word i/16, every sequence on an even word */
    uint64_t lbase[N_LEAVES + 1] = {0}, lword[N_LEAVES + 1] = {0};
    for (int l = 0; l < N_LEAVES; ++l) {
        const uint64_t n = strlen(LEAVES[l]);
        lbase[l + 1] = lbase[l] + n;
        lword[l + 1] = lword[l] + ((n + 15) / 16 + 1) / 2 * 2;
    }
    uint64_t *lwords = (uint64_t *)calloc(lword[N_LEAVES], sizeof(uint64_t));
    for (int l = 0; l < N_LEAVES; ++l)
        for (uint64_t i = 0; i < lbase[l + 1] - lbase[l]; ++i)
            lwords[lword[l] + i / 16] |= (uint64_t)iupac4(LEAVES[l][i]) << (4 * (i % 16));
    myrio_seqs *leaves = NULL;
    CHECK(myrio_leaves_upload(ctx, lwords, lword, lbase, N_LEAVES, &leaves));

    /* ---- reads: Seq<Dna>, base i at bits 2*(i%32) of word i/32, + Phred bytes */
    uint64_t rbase[N_READS + 1] = {0}, rword[N_READS + 1] = {0};
    for (int r = 0; r < N_READS; ++r) {
        rbase[r + 1] = rbase[r] + (uint64_t)READ_LEN[r];
        rword[r + 1] = rword[r] + (((uint64_t)READ_LEN[r] + 31) / 32 + 1) / 2 * 2;
    }
    uint64_t *rwords = (uint64_t *)calloc(rword[N_READS], sizeof(uint64_t));
    uint8_t *qual = (uint8_t *)malloc(rbase[N_READS]);
    for (int r = 0; r < N_READS; ++r)
        for (int i = 0; i < READ_LEN[r]; ++i) {
            rwords[rword[r] + (uint64_t)i / 32] |= (uint64_t)dna2(LEAVES[READ_SRC[r]][READ_BEGIN[r] + i]) << (2 * (i % 32));
            qual[rbase[r] + (uint64_t)i] = 35;
        }
    myrio_seqs *reads = NULL;
    CHECK(myrio_reads_upload(ctx, rwords, rword, rbase, qual, N_READS, &reads));

    /* ---- tree new: cached leaf counts; run: normalise, index, count the reads, score */
    myrio_svecs *leaf_counts = NULL, *leaf_vecs = NULL, *read_counts = NULL, *queries = NULL;
    myrio_index *index = NULL;
    CHECK(myrio_count_leaves(ctx, leaves, K, 32, 0, 0, &leaf_counts));
    CHECK(myrio_svecs_normalize(ctx, leaf_counts, &leaf_vecs));
    CHECK(myrio_index_build(ctx, leaf_vecs, 0, 0, &index));
    CHECK(myrio_count_reads(ctx, reads, K, 0.1f, &read_counts));
    CHECK(myrio_svecs_normalize(ctx, read_counts, &queries));

    float scores[N_READS * N_LEAVES], top_s[N_READS];
    uint32_t top_i[N_READS];
    CHECK(myrio_score_all(ctx, index, queries, 0, N_READS, MYRIO_SIM_COSINE, scores));
    CHECK(myrio_score_topx(ctx, index, queries, 0, N_READS, 1, MYRIO_SIM_COSINE, top_s, top_i));
    int ok = 1;
    for (int r = 0; r < N_READS; ++r) {
        printf("read %d: best leaf %u (score %.4f); row = %.4f %.4f %.4f\n", r, (unsigned)top_i[r], top_s[r],
               scores[r * N_LEAVES], scores[r * N_LEAVES + 1], scores[r * N_LEAVES + 2]);
        ok = ok && top_i[r] == (uint32_t)READ_SRC[r] && top_s[r] == scores[r * N_LEAVES + READ_SRC[r]];
    }
    printf("kernel launches: %llu\n", (unsigned long long)myrio_ctx_launch_count(ctx));

    myrio_index_free(index);
    myrio_svecs_free(queries);
    myrio_svecs_free(read_counts);
    myrio_svecs_free(leaf_vecs);
    myrio_svecs_free(leaf_counts);
    myrio_seqs_free(reads);
    myrio_seqs_free(leaves);
    myrio_ctx_destroy(ctx);
    free(lwords);
    free(rwords);
    free(qual);
    return ok ? 0 : 1;
}
