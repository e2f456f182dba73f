/*
 * myrio_oracle.c -- CPU restatement of Myrio's hot path (see myrio_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the product (myrio_b200/) never links this.
 *
 * Build: gcc -O2 -std=c11 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC
 *   -ffp-contract=off is REQUIRED: Rust never contracts a*b+c into an FMA and the
 *   reference's f32 results depend on that (data/sparse.rs:347).
 *
 * All floating point is IEEE f32, operations in exactly the reference's order.
 */
#include "myrio_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ tables */

/* constants.rs:28 -- the reference stores 128 decimal literals of
 * 1 - 10^(-q/10) evaluated in f64, which rustc rounds to f32.  Recomputed here
 * from the same formula; tests/test_oracle_golden.py checks the f32 bit patterns
 * against tests/golden/phred_correct_f32.json (made from the reference literals
 * by tests/golden/make_golden.py with exact rational rounding). */
static float g_phred[128];
static int g_phred_ready = 0;

const float *myo_phred_correct_table(void) {
    if (!g_phred_ready) {
#pragma omp critical(myo_phred)
        {
            if (!g_phred_ready) {
                for (int q = 0; q < 128; ++q) {
                    double p = 1.0 - pow(10.0, -(double)q / 10.0);
                    g_phred[q] = (float)p;
                }
                g_phred_ready = 1;
            }
        }
    }
    return g_phred;
}

/* tax/store.rs:178-188: `Seq::<Iupac>::from_str(&string_seq)` of every FASTA record.  bio-seq Iupac:
 * A=8 C=4 G=2 T=1, ambiguity codes = OR of their bases, '-' = X = 0; upper case only (assumed, see
 * header).  A record with any other byte is reported and skipped (store.rs:181-187). */
long long myo_fasta_to_iupac(const uint8_t *seq_ascii, const uint64_t *base_off, size_t n, uint8_t *codes_out,
                             uint8_t *keep_out, uint64_t *kept_off_out) {
    static const char sym[] = "-TGKCYSBAWRDMHVN"; /* index = 4-bit code */
    size_t kept = 0;
    uint64_t w = 0;
    kept_off_out[0] = 0;
    for (size_t r = 0; r < n; ++r) {
        uint64_t b = base_off[r], e = base_off[r + 1];
        int ok = 1;
        for (uint64_t i = b; i < e && ok; ++i) {
            const char *p = seq_ascii[i] ? strchr(sym, (int)seq_ascii[i]) : NULL;
            if (!p) ok = 0;
        }
        keep_out[r] = (uint8_t)ok;
        if (!ok) continue;
        for (uint64_t i = b; i < e; ++i) codes_out[w++] = (uint8_t)(strchr(sym, (int)seq_ascii[i]) - sym);
        kept_off_out[++kept] = w;
    }
    return (long long)kept;
}

int myo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py --impl reference under torchrun: the launcher exports OMP_NUM_THREADS=1; the CPU arm must
 * still use every host core it can */
void myo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static int clamp_threads(int threads) {
    int m = myo_max_threads();
    if (threads <= 0 || threads > m) return m;
    return threads;
}

/* ------------------------------------------------------- FASTQ ingestion */

long long myo_fastq_preprocess(const uint8_t *seq_ascii, const uint8_t *qual_ascii, const uint64_t *base_off,
                               size_t n, size_t min_length, float min_mean_qual, uint8_t max_qual,
                               uint8_t *codes_out, uint8_t *qual_out, uint8_t *keep_out,
                               uint64_t *kept_off_out) {
    size_t kept = 0;
    uint64_t w = 0;
    kept_off_out[0] = 0;
    for (size_t r = 0; r < n; ++r) {
        size_t b = (size_t)base_off[r], len = (size_t)(base_off[r + 1] - base_off[r]);
        /* io/mod.rs:46-47 Seq::from_str: only A, C, G, T */
        for (size_t i = 0; i < len; ++i) {
            uint8_t c = seq_ascii[b + i];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return -(long long)(r + 1);
        }
        /* data/myrseq.rs:119-128 */
        int keep = len != 0;
        if (keep) {
            float sum = 0.0f;
            uint8_t mx = 0;
            for (size_t i = 0; i < len; ++i) {
                uint8_t q = qual_ascii[b + i] >= 33 ? (uint8_t)(qual_ascii[b + i] - 33) : 0; /* io/mod.rs:48 */
                sum = sum + (float)q;
                if (q > mx) mx = q;
            }
            float mean = sum / (float)len;
            keep = (len >= min_length) && (mean >= min_mean_qual) && (mx <= max_qual);
        }
        keep_out[r] = (uint8_t)keep;
        if (keep) {
            for (size_t i = 0; i < len; ++i) {
                uint8_t c = seq_ascii[b + i];
                codes_out[w + i] = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
                qual_out[w + i] = qual_ascii[b + i] >= 33 ? (uint8_t)(qual_ascii[b + i] - 33) : 0;
            }
            w += len;
            kept += 1;
            kept_off_out[kept] = w;
        }
    }
    return (long long)kept;
}

/* ------------------------------------------------------------ k-mer keys */

/* bio-seq `usize::from(&Kmer<Dna,K>)`: A=0,C=1,G=2,T=3, base j of the window at
 * bits 2j..2j+1 (Seq is a BitVec<usize,Lsb0>).  ASSUMED -- bio-seq is not
 * vendored; see header. */
static int g_kmer_msb_first = 0;

/* the alternative packing (base j at bits 2(k-1-j)), for the case the assumption above is wrong:
 * mirrors myrio_ctx_set_kmer_bit_order of the product */
void myo_set_kmer_msb_first(int on) { g_kmer_msb_first = on ? 1 : 0; }

uint64_t myo_kmer_key(const uint8_t *bases, int k) {
    uint64_t key = 0;
    if (g_kmer_msb_first) {
        for (int j = 0; j < k; ++j) key |= (uint64_t)(bases[j] & 3u) << (2 * (k - 1 - j));
        return key;
    }
    for (int j = 0; j < k; ++j) key |= (uint64_t)(bases[j] & 3u) << (2 * j);
    return key;
}

/* ------------------------------------------------------------------ sort */

/* slice::sort_unstable on plain u64 -- only the resulting order matters. */
static void sort_u64(uint64_t *a, size_t n) {
    if (n < 2) return;
    if (n <= 48) {
        for (size_t i = 1; i < n; ++i) {
            uint64_t x = a[i];
            size_t j = i;
            while (j > 0 && a[j - 1] > x) {
                a[j] = a[j - 1];
                --j;
            }
            a[j] = x;
        }
        return;
    }
    uint64_t *tmp = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t all_or = 0, all_and = ~(uint64_t)0;
    for (size_t i = 0; i < n; ++i) {
        all_or |= a[i];
        all_and &= a[i];
    }
    uint64_t varying = all_or ^ all_and;
    uint64_t *src = a, *dst = tmp;
    for (int pass = 0; pass < 8; ++pass) {
        int shift = pass * 8;
        if (((varying >> shift) & 0xFF) == 0) continue;
        size_t hist[256];
        memset(hist, 0, sizeof(hist));
        for (size_t i = 0; i < n; ++i) hist[(src[i] >> shift) & 0xFF]++;
        size_t sum = 0;
        for (int b = 0; b < 256; ++b) {
            size_t c = hist[b];
            hist[b] = sum;
            sum += c;
        }
        for (size_t i = 0; i < n; ++i) dst[hist[(src[i] >> shift) & 0xFF]++] = src[i];
        uint64_t *t = src;
        src = dst;
        dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(uint64_t));
    free(tmp);
}

/* data/sparse.rs:229-272 from_unsorted_singles(singles, 1.0f32) */
size_t myo_sort_reduce_f32(uint64_t *singles, size_t n, uint64_t *keys_out, float *vals_out) {
    if (n == 0) return 0; /* :237-239 */
    sort_u64(singles, n); /* :241 */
    size_t idx = 0;
    uint64_t prev_key = singles[0];
    keys_out[0] = prev_key;
    float val_to_write = 1.0f;
    for (size_t i = 1; i < n; ++i) { /* :253-263 */
        uint64_t key = singles[i];
        if (prev_key != key) {
            vals_out[idx] = val_to_write;
            idx += 1;
            keys_out[idx] = key;
            val_to_write = 1.0f;
            prev_key = key;
        } else {
            val_to_write += 1.0f;
        }
    }
    vals_out[idx] = val_to_write;
    return idx + 1;
}

/* same, T = u16, value = 1; "+=" wraps like a release build */
size_t myo_sort_reduce_u16(uint64_t *singles, size_t n, uint64_t *keys_out, uint16_t *vals_out) {
    if (n == 0) return 0;
    sort_u64(singles, n);
    size_t idx = 0;
    uint64_t prev_key = singles[0];
    keys_out[0] = prev_key;
    uint16_t val_to_write = 1;
    for (size_t i = 1; i < n; ++i) {
        uint64_t key = singles[i];
        if (prev_key != key) {
            vals_out[idx] = val_to_write;
            idx += 1;
            keys_out[idx] = key;
            val_to_write = 1;
            prev_key = key;
        } else {
            val_to_write = (uint16_t)(val_to_write + 1u);
        }
    }
    vals_out[idx] = val_to_write;
    return idx + 1;
}

/* ------------------------------------------------------- K1: read counting */

/* data/myrseq.rs:76-111 */
int myo_count_read(const uint8_t *bases, const uint8_t *qual, size_t n, int k, float cutoff,
                   uint64_t *keys_out, float *vals_out, size_t *d_out, uint64_t *nb_hck_out) {
    if (k < 2 || k > 32) return -1; /* myrio-proc/src/lib.rs:70 panics */
    *d_out = 0;
    *nb_hck_out = 0;
    if (n < (size_t)k) return 0; /* kmers::<K>() yields nothing */
    const float *P = myo_phred_correct_table();
    size_t nb_kmers = n - (size_t)k + 1; /* :94 */

    /* :81-88 conf_score_per_kmer = windows(k).map(product) */
    float *conf = (float *)malloc(nb_kmers * sizeof(float));
    for (size_t i = 0; i < n; ++i)
        if (qual[i] >= 128) { /* table index out of bounds -> panic in the reference */
            free(conf);
            return -2;
        }
    for (size_t i = 0; i < nb_kmers; ++i) {
        float prod = 1.0f;
        for (int j = 0; j < k; ++j) prod = prod * P[qual[i + j]];
        conf[i] = prod;
    }

    /* seq.to_revcomp(): reverse, complement (3 - code) */
    uint8_t *rc = (uint8_t *)malloc(n);
    for (size_t i = 0; i < n; ++i) rc[i] = (uint8_t)(3u - (bases[n - 1 - i] & 3u));

    uint64_t *singles = (uint64_t *)malloc(2 * nb_kmers * sizeof(uint64_t));
    size_t ns = 0;
    uint64_t nb_hck = 0;
    for (size_t idx = 0; idx < nb_kmers; ++idx) { /* :95-105 */
        if (conf[idx] > cutoff) {
            singles[ns++] = myo_kmer_key(bases + idx, k);
            nb_hck += 1;
        }
        if (conf[nb_kmers - 1 - idx] > cutoff) {
            singles[ns++] = myo_kmer_key(rc + idx, k);
            nb_hck += 1;
        }
    }
    *d_out = myo_sort_reduce_f32(singles, ns, keys_out, vals_out); /* :110 */
    *nb_hck_out = nb_hck;
    free(singles);
    free(rc);
    free(conf);
    return 0;
}

int myo_count_reads(const uint8_t *bases, const uint8_t *qual, const uint64_t *base_off,
                    size_t n_reads, int k, float cutoff, int threads, uint64_t *row_off,
                    uint64_t *keys_out, float *vals_out, uint64_t *nb_hck_out) {
    if (k < 2 || k > 32) return -1;
    int nt = threads <= 1 ? 1 : clamp_threads(threads);
    int err = 0;
    if (keys_out == NULL) {
        /* phase 1: sizes */
        uint64_t *d = (uint64_t *)calloc(n_reads + 1, sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
        for (size_t r = 0; r < n_reads; ++r) {
            size_t n = (size_t)(base_off[r + 1] - base_off[r]);
            size_t cap = n >= (size_t)k ? 2 * (n - k + 1) : 0;
            uint64_t *kk = (uint64_t *)malloc((cap + 1) * sizeof(uint64_t));
            float *vv = (float *)malloc((cap + 1) * sizeof(float));
            size_t dd = 0;
            uint64_t hck = 0;
            int e = myo_count_read(bases + base_off[r], qual + base_off[r], n, k, cutoff, kk, vv,
                                   &dd, &hck);
            if (e) {
#pragma omp atomic write
                err = e;
            }
            d[r] = dd;
            if (nb_hck_out) nb_hck_out[r] = hck;
            free(kk);
            free(vv);
        }
        uint64_t s = 0;
        for (size_t r = 0; r < n_reads; ++r) {
            row_off[r] = s;
            s += d[r];
        }
        row_off[n_reads] = s;
        free(d);
        return err;
    }
    /* phase 2: fill (row_off already computed by phase 1) */
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
    for (size_t r = 0; r < n_reads; ++r) {
        size_t n = (size_t)(base_off[r + 1] - base_off[r]);
        size_t cap = n >= (size_t)k ? 2 * (n - k + 1) : 0;
        uint64_t *kk = (uint64_t *)malloc((cap + 1) * sizeof(uint64_t));
        float *vv = (float *)malloc((cap + 1) * sizeof(float));
        size_t dd = 0;
        uint64_t hck = 0;
        int e = myo_count_read(bases + base_off[r], qual + base_off[r], n, k, cutoff, kk, vv, &dd,
                               &hck);
        if (e) {
#pragma omp atomic write
            err = e;
        }
        memcpy(keys_out + row_off[r], kk, dd * sizeof(uint64_t));
        memcpy(vals_out + row_off[r], vv, dd * sizeof(float));
        if (nb_hck_out) nb_hck_out[r] = hck;
        free(kk);
        free(vv);
    }
    return err;
}

/* ------------------------------------------------------- K2: leaf counting */

/* Counter-based stand-in for `SmallRng::from_os_rng` + `choose` (tax/mod.rs:77-97,
 * store.rs:277).  The reference is nondeterministic here; this definition is the
 * shared contract between oracle and CUDA path. splitmix64 finaliser. */
uint32_t myo_resample_pick(uint64_t seed, uint64_t leaf, uint32_t resample, uint32_t pos,
                           uint32_t n_opts) {
    uint64_t x = seed;
    x ^= leaf * 0x9E3779B97F4A7C15ull;
    x ^= ((uint64_t)resample << 32 | (uint64_t)pos) * 0xD1B54A32D192ED03ull;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (uint32_t)(((x >> 32) * (uint64_t)n_opts) >> 32);
}

/* tax/mod.rs:77-97 sample_iupac_nc: options are listed in ascending Dna order
 * (A,C,G,T) = descending bit order of the assumed 4-bit code A=8,C=4,G=2,T=1. */
static uint8_t resolve_iupac(uint8_t code, uint64_t seed, uint64_t leaf, uint32_t resample,
                             uint32_t pos) {
    uint8_t opts[4];
    uint32_t n = 0;
    if (code & 8u) opts[n++] = 0;
    if (code & 4u) opts[n++] = 1;
    if (code & 2u) opts[n++] = 2;
    if (code & 1u) opts[n++] = 3;
    if (n == 1) return opts[0];
    return opts[myo_resample_pick(seed, leaf, resample, pos, n)];
}

/* tax/mod.rs:70-148 */
int myo_count_leaf(const uint8_t *iupac, size_t n_in, int k, uint32_t nb_resamples, uint64_t seed,
                   uint64_t leaf_id, uint64_t *keys_out, uint16_t *vals_out, size_t cap,
                   size_t *d_out, float *rescale_out) {
    if (k < 2 || k > 32) return -1;
    if (cap < 1) return -3;
    /* :98 drop gaps (Iupac::X == 0) */
    uint8_t *bn = (uint8_t *)malloc(n_in + 1);
    size_t n = 0;
    for (size_t i = 0; i < n_in; ++i)
        if ((iupac[i] & 15u) != 0) bn[n++] = iupac[i] & 15u;
    if (n < (size_t)k) { /* :99-102 sentinel */
        keys_out[0] = 0;
        vals_out[0] = 1;
        *d_out = 1;
        *rescale_out = 1.0f;
        free(bn);
        return 0;
    }
    size_t nb = n - (size_t)k + 1;
    size_t nb_kmers_per_seq = 2 * nb; /* :104 */
    size_t total = (size_t)nb_resamples * nb_kmers_per_seq;
    uint64_t *singles = (uint64_t *)malloc((total + 1) * sizeof(uint64_t));
    uint8_t *seq = (uint8_t *)malloc(n);
    uint8_t *rc = (uint8_t *)malloc(n);
    size_t idx = 0;
    for (uint32_t r = 0; r < nb_resamples; ++r) { /* :112-123 */
        for (size_t i = 0; i < n; ++i) seq[i] = resolve_iupac(bn[i], seed, leaf_id, r, (uint32_t)i);
        for (size_t i = 0; i < n; ++i) rc[i] = (uint8_t)(3u - seq[n - 1 - i]);
        for (size_t i = 0; i < nb; ++i) singles[idx++] = myo_kmer_key(seq + i, k);
        for (size_t i = 0; i < nb; ++i) singles[idx++] = myo_kmer_key(rc + i, k);
    }
    uint64_t *tk = (uint64_t *)malloc((total + 1) * sizeof(uint64_t));
    uint16_t *tv = (uint16_t *)malloc((total + 1) * sizeof(uint16_t));
    size_t td = myo_sort_reduce_u16(singles, total, tk, tv);  /* :129 */
    uint16_t cutoff = (uint16_t)((nb_resamples + 5u) >> 3);     /* :130 */
    size_t d = 0;
    int overflow = 0;
    for (size_t i = 0; i < td; ++i) { /* :133-134 */
        if (tv[i] > cutoff) {
            if (d >= cap) {
                overflow = 1;
                break;
            }
            keys_out[d] = tk[i];
            vals_out[d] = tv[i];
            d++;
        }
    }
    free(singles);
    free(seq);
    free(rc);
    free(tk);
    free(tv);
    free(bn);
    if (overflow) return -3;
    if (d == 0) { /* :136-139 */
        keys_out[0] = 0;
        vals_out[0] = 1;
        *d_out = 1;
        *rescale_out = 1.0f;
        return 0;
    }
    uint32_t sum = 0; /* :144-145 sum::<u32>() */
    for (size_t i = 0; i < d; ++i) sum += (uint32_t)vals_out[i];
    *rescale_out = (float)nb_kmers_per_seq / (float)sum;
    *d_out = d;
    return 0;
}

int myo_count_leaves(const uint8_t *iupac, const uint64_t *base_off, size_t n_leaves, int k,
                     uint32_t nb_resamples, uint64_t seed, uint64_t leaf_id_base, int threads,
                     uint64_t *row_off, uint64_t *keys_out, uint16_t *vals_out, float *rescale_out) {
    if (k < 2 || k > 32) return -1;
    int nt = clamp_threads(threads);
    int err = 0;
    int sizing = (keys_out == NULL);
    uint64_t *d = sizing ? (uint64_t *)calloc(n_leaves + 1, sizeof(uint64_t)) : NULL;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (size_t l = 0; l < n_leaves; ++l) {
        size_t n = (size_t)(base_off[l + 1] - base_off[l]);
        size_t cap = n >= (size_t)k ? 2 * (n - k + 1) * (size_t)(nb_resamples ? nb_resamples : 1) : 1;
        if (cap < 1) cap = 1;
        /* distinct keys can never exceed the number of singles */
        uint64_t *kk = (uint64_t *)malloc(cap * sizeof(uint64_t));
        uint16_t *vv = (uint16_t *)malloc(cap * sizeof(uint16_t));
        size_t dd = 0;
        float rs = 1.0f;
        int e = myo_count_leaf(iupac + base_off[l], n, k, nb_resamples, seed, leaf_id_base + l, kk,
                               vv, cap, &dd, &rs);
        if (e) {
#pragma omp atomic write
            err = e;
        }
        if (sizing) {
            d[l] = dd;
        } else {
            memcpy(keys_out + row_off[l], kk, dd * sizeof(uint64_t));
            memcpy(vals_out + row_off[l], vv, dd * sizeof(uint16_t));
        }
        if (rescale_out) rescale_out[l] = rs;
        free(kk);
        free(vv);
    }
    if (sizing) {
        uint64_t s = 0;
        for (size_t l = 0; l < n_leaves; ++l) {
            row_off[l] = s;
            s += d[l];
        }
        row_off[n_leaves] = s;
        free(d);
    }
    return err;
}

/* ------------------------------------------------------- sparse f32 algebra */

/* data/sparse.rs:366-383: norm_l2_squared = values.map(|v| v.powi(2)).sum();
 * magnitude = sqrt; div_assign(magn) on every value */
void myo_normalize_l2(float *vals, size_t n) {
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float sq = vals[i] * vals[i];
        s = s + sq;
    }
    float magn = sqrtf(s);
    for (size_t i = 0; i < n; ++i) vals[i] = vals[i] / magn;
}

/* tax/compute.rs:105-106: svec.apply_into(Float::from).into_normalized_l2() */
void myo_normalize_u16(const uint16_t *in, size_t n, float *out) {
    for (size_t i = 0; i < n; ++i) out[i] = (float)in[i];
    myo_normalize_l2(out, n);
}

void myo_normalize_rows_f32(const uint64_t *row_off, size_t n_rows, float *vals, int threads) {
    int nt = clamp_threads(threads);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (size_t r = 0; r < n_rows; ++r)
        myo_normalize_l2(vals + row_off[r], (size_t)(row_off[r + 1] - row_off[r]));
}

void myo_normalize_rows_u16(const uint64_t *row_off, size_t n_rows, const uint16_t *in, float *out,
                            int threads) {
    int nt = clamp_threads(threads);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (size_t r = 0; r < n_rows; ++r)
        myo_normalize_u16(in + row_off[r], (size_t)(row_off[r + 1] - row_off[r]), out + row_off[r]);
}

/* data/sparse.rs:317-360 */
float myo_dot(const uint64_t *ak, const float *av, size_t an, const uint64_t *bk, const float *bv,
              size_t bn) {
    if (!(an <= bn)) { /* :322-328 the shorter vector drives the outer loop */
        const uint64_t *tk = ak;
        const float *tv = av;
        size_t tn = an;
        ak = bk;
        av = bv;
        an = bn;
        bk = tk;
        bv = tv;
        bn = tn;
    }
    float dot_product = 0.0f;
    size_t i = 0, j = 0;
    while (i < an) { /* :335-357 */
        uint64_t ak_i = ak[i];
        for (;;) {
            if (j >= bn) goto done;
            uint64_t bk_j = bk[j];
            if (bk_j < ak_i) {
                j += 1;
                continue;
            } else if (bk_j == ak_i) {
                float prod = av[i] * bv[j];
                dot_product = dot_product + prod;
                j += 1;
                break;
            } else {
                break;
            }
        }
        i += 1;
    }
done:
    return dot_product;
}

static float norm_l2_squared(const float *v, size_t n) {
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float sq = v[i] * v[i];
        s = s + sq;
    }
    return s;
}

/* SimScore: validate(finite), default 0 (similarity.rs:8-14) */
static float simscore_or_default(float x) { return isfinite(x) ? x : 0.0f; }

/* similarity.rs:172-182 overlap: merge_and_apply(min,max) then two sums */
static float overlap(const uint64_t *ak, const float *av, size_t an, const uint64_t *bk,
                     const float *bv, size_t bn) {
    float sum_min = 0.0f, sum_max = 0.0f;
    size_t i = 0, j = 0;
    while (i < an && j < bn) {
        float x, y;
        if (ak[i] < bk[j]) {
            x = av[i];
            y = 0.0f;
            i++;
        } else if (ak[i] > bk[j]) {
            x = 0.0f;
            y = bv[j];
            j++;
        } else {
            x = av[i];
            y = bv[j];
            i++;
            j++;
        }
        sum_min = sum_min + fminf(x, y);
        sum_max = sum_max + fmaxf(x, y);
    }
    while (i < an) {
        sum_min = sum_min + fminf(av[i], 0.0f);
        sum_max = sum_max + fmaxf(av[i], 0.0f);
        i++;
    }
    while (j < bn) {
        sum_min = sum_min + fminf(0.0f, bv[j]);
        sum_max = sum_max + fmaxf(0.0f, bv[j]);
        j++;
    }
    return simscore_or_default(sum_min / sum_max);
}

/* similarity.rs:40-51 to_simfunc */
float myo_similarity(int sim, int pre_normalized, const uint64_t *ak, const float *av, size_t an,
                     const uint64_t *bk, const float *bv, size_t bn) {
    switch (sim) {
    case MYO_SIM_COSINE:
        if (pre_normalized) {
            /* :87-92 release build: new_unchecked(a.dot(b)) */
            return myo_dot(ak, av, an, bk, bv, bn);
        } else {
            /* :60-65 */
            float d = myo_dot(ak, av, an, bk, bv, bn);
            float na = sqrtf(norm_l2_squared(av, an));
            float nb = sqrtf(norm_l2_squared(bv, bn));
            return simscore_or_default(d / (na * nb));
        }
    case MYO_SIM_JACARD_TANIMOTO:
        if (pre_normalized) { /* :143-149 */
            float d = myo_dot(ak, av, an, bk, bv, bn);
            return simscore_or_default(d / (2.0f - d));
        } else { /* :114-120 */
            float d = myo_dot(ak, av, an, bk, bv, bn);
            float s = norm_l2_squared(av, an) + norm_l2_squared(bv, bn);
            return simscore_or_default(d / (s - d));
        }
    case MYO_SIM_OVERLAP:
        return overlap(ak, av, an, bk, bv, bn);
    default:
        return 0.0f;
    }
}

/* tax/results.rs:82-93: payloads.into_par_iter().map(|sfvec| simfunc(&sfvec,&query)) */
void myo_score_all(int sim, const uint64_t *leaf_row_off, const uint64_t *leaf_keys,
                   const float *leaf_vals, size_t n_leaves, const uint64_t *qk, const float *qv,
                   size_t qn, int threads, float *scores_out) {
    int nt = clamp_threads(threads);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (size_t l = 0; l < n_leaves; ++l) {
        size_t b = (size_t)leaf_row_off[l], e = (size_t)leaf_row_off[l + 1];
        scores_out[l] = myo_similarity(sim, 1, leaf_keys + b, leaf_vals + b, e - b, qk, qv, qn);
    }
}

/* tax/results.rs:310-329: sorted_by_key(-score) (stable) .take(nb_best).
 * The reference's leaf visiting order comes from a HashMap-built tree, so equal
 * scores have no defined order there; the contract is
 * (score desc, payload_id asc). */
typedef struct {
    float s;
    uint32_t id;
} topx_item;

static int topx_better(topx_item a, topx_item b) { /* a ranks before b */
    if (a.s > b.s) return 1;
    if (a.s < b.s) return 0;
    return a.id < b.id;
}

void myo_topx(const float *scores, size_t n, uint32_t id_base, size_t x, float *top_scores,
              uint32_t *top_ids) {
    /* min-heap (worst at root) of the best x so far */
    topx_item *h = (topx_item *)malloc((x + 1) * sizeof(topx_item));
    size_t hn = 0;
    for (size_t i = 0; i < n && x > 0; ++i) {
        topx_item it = {scores[i], (uint32_t)(id_base + i)};
        if (hn < x) {
            size_t c = hn++;
            h[c] = it;
            while (c > 0) {
                size_t p = (c - 1) / 2;
                if (topx_better(h[p], h[c])) {
                    topx_item t = h[p];
                    h[p] = h[c];
                    h[c] = t;
                    c = p;
                } else
                    break;
            }
        } else if (topx_better(it, h[0])) {
            h[0] = it;
            size_t p = 0;
            for (;;) {
                size_t l = 2 * p + 1, r = l + 1, w = p;
                if (l < hn && topx_better(h[w], h[l])) w = l;
                if (r < hn && topx_better(h[w], h[r])) w = r;
                if (w == p) break;
                topx_item t = h[p];
                h[p] = h[w];
                h[w] = t;
                p = w;
            }
        }
    }
    /* pop worst-first into the tail */
    size_t m = hn;
    while (hn > 0) {
        top_scores[hn - 1] = h[0].s;
        top_ids[hn - 1] = h[0].id;
        h[0] = h[--hn];
        size_t p = 0;
        for (;;) {
            size_t l = 2 * p + 1, r = l + 1, w = p;
            if (l < hn && topx_better(h[w], h[l])) w = l;
            if (r < hn && topx_better(h[w], h[r])) w = r;
            if (w == p) break;
            topx_item t = h[p];
            h[p] = h[w];
            h[w] = t;
            p = w;
        }
    }
    for (size_t i = m; i < x; ++i) {
        top_scores[i] = 0.0f;
        top_ids[i] = 0xFFFFFFFFu;
    }
    free(h);
}

/* data/sparse.rs:143-221 merge_and_apply(|x,y| x+y)  (Add, :493-519) */
size_t myo_add(const uint64_t *ak, const float *av, size_t an, const uint64_t *bk, const float *bv,
               size_t bn, uint64_t *ck, float *cv) {
    size_t i = 0, j = 0, w = 0;
    while (i < an && j < bn) {
        if (ak[i] < bk[j]) {
            ck[w] = ak[i];
            cv[w] = av[i] + 0.0f;
            i++;
        } else if (ak[i] > bk[j]) {
            ck[w] = bk[j];
            cv[w] = 0.0f + bv[j];
            j++;
        } else {
            ck[w] = ak[i];
            cv[w] = av[i] + bv[j];
            i++;
            j++;
        }
        w++;
    }
    while (i < an) {
        ck[w] = ak[i];
        cv[w] = av[i] + 0.0f;
        i++;
        w++;
    }
    while (j < bn) {
        ck[w] = bk[j];
        cv[w] = 0.0f + bv[j];
        j++;
        w++;
    }
    return w;
}

/* clustering.rs:44-53: fold(identity, a + b).reduce(..).into_normalized_l2().
 * rayon's split points are not fixed; the sums are integer-valued f32 and exact
 * below 2^24, so a plain left fold gives the same result. */
size_t myo_centroid(const uint64_t *row_off, const uint64_t *keys, const float *vals,
                    const uint64_t *member_ids, size_t m, uint64_t *ck, float *cv) {
    size_t total = 0;
    for (size_t t = 0; t < m; ++t) total += (size_t)(row_off[member_ids[t] + 1] - row_off[member_ids[t]]);
    uint64_t *tk = (uint64_t *)malloc((total + 1) * sizeof(uint64_t));
    float *tv = (float *)malloc((total + 1) * sizeof(float));
    size_t cn = 0;
    for (size_t t = 0; t < m; ++t) {
        size_t b = (size_t)row_off[member_ids[t]], e = (size_t)row_off[member_ids[t] + 1];
        size_t nn = myo_add(ck, cv, cn, keys + b, vals + b, e - b, tk, tv);
        memcpy(ck, tk, nn * sizeof(uint64_t));
        memcpy(cv, tv, nn * sizeof(float));
        cn = nn;
    }
    free(tk);
    free(tv);
    myo_normalize_l2(cv, cn);
    return cn;
}

/* ------------------------------------------------- inverted index structure */

typedef struct {
    uint64_t key;
    uint32_t leaf;
    float val;
} post_t;

static int post_cmp(const void *a, const void *b) {
    const post_t *x = (const post_t *)a, *y = (const post_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->leaf != y->leaf) return x->leaf < y->leaf ? -1 : 1;
    return 0;
}

void myo_invert(const uint64_t *row_off, const uint64_t *keys, const float *vals, size_t n_leaves,
                size_t *n_unique, uint64_t *ukeys, uint64_t *offs, uint32_t *post_leaf,
                float *post_val) {
    size_t P = (size_t)row_off[n_leaves];
    post_t *p = (post_t *)malloc((P + 1) * sizeof(post_t));
    for (size_t l = 0; l < n_leaves; ++l)
        for (size_t i = (size_t)row_off[l]; i < (size_t)row_off[l + 1]; ++i) {
            p[i].key = keys[i];
            p[i].leaf = (uint32_t)l;
            p[i].val = vals[i];
        }
    qsort(p, P, sizeof(post_t), post_cmp);
    size_t u = 0;
    for (size_t i = 0; i < P; ++i) {
        if (i == 0 || p[i].key != p[i - 1].key) {
            if (ukeys) {
                ukeys[u] = p[i].key;
                offs[u] = i;
            }
            u++;
        }
        if (ukeys) {
            post_leaf[i] = p[i].leaf;
            post_val[i] = p[i].val;
        }
    }
    if (ukeys) offs[u] = P;
    *n_unique = u;
    free(p);
}

/* ------------------------------------------------------------- clustering */

typedef struct {
    uint64_t *k;
    float *v;
    size_t n;
} svec_t;

typedef struct {
    svec_t centroid;   /* owned, normalised */
    size_t *members;   /* indices into the kept-read arrays */
    size_t n_members, cap_members;
} cluster_t;

static svec_t svec_clone(const uint64_t *k, const float *v, size_t n) {
    svec_t s;
    s.k = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    s.v = (float *)malloc((n + 1) * sizeof(float));
    memcpy(s.k, k, n * sizeof(uint64_t));
    memcpy(s.v, v, n * sizeof(float));
    s.n = n;
    return s;
}

static void svec_free(svec_t *s) {
    free(s->k);
    free(s->v);
    s->k = NULL;
    s->v = NULL;
    s->n = 0;
}

static void cluster_push(cluster_t *c, size_t idx) {
    if (c->n_members == c->cap_members) {
        c->cap_members = c->cap_members ? 2 * c->cap_members : 64;
        c->members = (size_t *)realloc(c->members, c->cap_members * sizeof(size_t));
    }
    c->members[c->n_members++] = idx;
}

typedef struct {
    /* kept reads (nb_hck > t2), raw and normalised copies (clustering.rs:76-94) */
    size_t n;
    const uint64_t *row_off; /* original CSR */
    const uint64_t *keys;
    const float *raw;
    float *norm;             /* same layout as raw */
    size_t *orig;            /* kept idx -> original read idx */
    int sim;
} cl_ctx;

static inline float cl_sim_centroid_read(const cl_ctx *cx, const svec_t *c, size_t kept) {
    size_t r = cx->orig[kept];
    size_t b = (size_t)cx->row_off[r], e = (size_t)cx->row_off[r + 1];
    return myo_similarity(cx->sim, 1, c->k, c->v, c->n, cx->keys + b, cx->norm + b, e - b);
}

static inline float cl_sim_read_read(const cl_ctx *cx, size_t ka, size_t kb) {
    size_t ra = cx->orig[ka], rb = cx->orig[kb];
    size_t ab = (size_t)cx->row_off[ra], ae = (size_t)cx->row_off[ra + 1];
    size_t bb = (size_t)cx->row_off[rb], be = (size_t)cx->row_off[rb + 1];
    return myo_similarity(cx->sim, 1, cx->keys + ab, cx->norm + ab, ae - ab, cx->keys + bb,
                          cx->norm + bb, be - bb);
}

/* max_by_key(|cluster| simfunc(centroid, read)) -> LAST maximum (clustering.rs:142-147) */
static size_t cl_argmax(const cl_ctx *cx, const cluster_t *cl, size_t nc, size_t kept, float *best_out) {
    size_t best = 0;
    float bs = 0.0f;
    for (size_t c = 0; c < nc; ++c) {
        float s = cl_sim_centroid_read(cx, &cl[c].centroid, kept);
        if (c == 0 || s >= bs) {
            bs = s;
            best = c;
        }
    }
    if (best_out) *best_out = bs;
    return best;
}

/* data/sparse.rs:274-313 from_unsorted_pairs: sort by key (unstable; equal keys are summed, so the order among
 * them only matters for f32 rounding -- this restatement keeps the input order), values of equal keys added.
 * keys_out / vals_out capacity n; returns the number of distinct keys. */
size_t myo_sort_reduce_pairs_f32(const uint64_t *keys, const float *vals, size_t n, uint64_t *keys_out, float *vals_out) {
    if (n == 0) return 0;
    size_t *ord = (size_t *)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) ord[i] = i;
    for (size_t i = 1; i < n; ++i) { /* stable insertion sort */
        size_t v = ord[i], j = i;
        while (j > 0 && keys[ord[j - 1]] > keys[v]) {
            ord[j] = ord[j - 1];
            --j;
        }
        ord[j] = v;
    }
    size_t idx = 0;
    uint64_t prev_key = keys[ord[0]];
    float val_to_write = vals[ord[0]];
    keys_out[0] = prev_key;
    for (size_t i = 1; i < n; ++i) {
        uint64_t key = keys[ord[i]];
        float val = vals[ord[i]];
        if (prev_key != key) {
            vals_out[idx] = val_to_write;
            idx += 1;
            keys_out[idx] = key;
            val_to_write = val;
            prev_key = key;
        } else {
            val_to_write = val_to_write + val;
        }
    }
    vals_out[idx] = val_to_write;
    free(ord);
    return idx + 1;
}

/* myrio-cli/src/app/process/run.rs:243-262: re-weighted local fingerprints -> initial centroid of pass 2.
 * local: CSR of n normalised fingerprints; query: the naive centroid.  out capacity: local nnz. */
size_t myo_reweight_fingerprints(int sim, const uint64_t *row_off, const uint64_t *keys, const float *vals, size_t n,
                                 const uint64_t *qk, const float *qv, size_t qn, float alpha, uint64_t *ck,
                                 float *cv) {
    size_t nnz = (size_t)row_off[n];
    float *scaled = (float *)malloc((nnz + 1) * sizeof(float));
    uint64_t *ids = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    for (size_t i = 0; i < n; ++i) {
        size_t b = (size_t)row_off[i], e = (size_t)row_off[i + 1];
        float score = myo_similarity(sim, 1, keys + b, vals + b, e - b, qk, qv, qn); /* simfunc(local, naive_query) */
        float adjustor = expf(1.0f + alpha * score);
        for (size_t j = b; j < e; ++j) scaled[j] = vals[j] * adjustor; /* local_fingerprint * adjustor */
        ids[i] = i;
    }
    size_t out = myo_centroid(row_off, keys, scaled, ids, n, ck, cv); /* .sum::<SFVec>().into_normalized_l2() */
    free(scaled);
    free(ids);
    return out;
}

/* Full-size checks (tests/test_gpu_configs.py): the O(N^2) silhouette pass of a 100 k-read batch takes
 * the CPU half an hour, so a test may restrict it to a SAMPLE of the reads.  With a mask set
 * (mask[read] != 0 = compute), inner() (:323-336) runs for the masked reads only -- same code, same f32
 * order -- every other read gets NaN, and the trimming step (:244-266), which needs all the scores of
 * a cluster, is skipped.  NULL (the default) restores the reference behaviour. */
static const uint8_t *g_sil_mask = NULL;
void myo_set_silhouette_sample(const uint8_t *mask) { g_sil_mask = mask; }

int myo_cluster(const uint64_t *row_off, const uint64_t *keys, const float *vals,
                const uint64_t *nb_hck, size_t n_reads, const myo_cluster_params *p,
                int32_t *assign_out, uint32_t *n_iters_out, float *sil_out) {
    if (p->n_init == 0 && p->expected_nb_of_clusters == 0) return -10; /* :71-73 panic */
    int nt = clamp_threads(p->threads);
    cl_ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.row_off = row_off;
    cx.keys = keys;
    cx.raw = vals;
    cx.sim = p->similarity;
    cx.orig = (size_t *)malloc((n_reads + 1) * sizeof(size_t));

    /* Step 1 (:78-90): keep reads with nb_hck > t2 */
    for (size_t r = 0; r < n_reads; ++r) {
        if (sil_out) sil_out[r] = NAN;
        if (nb_hck[r] > p->t2) {
            cx.orig[cx.n++] = r;
            assign_out[r] = 0;
        } else {
            assign_out[r] = -1;
        }
    }
    if (cx.n == 0) { /* :107 unwrap on empty iterator panics */
        free(cx.orig);
        return -11;
    }
    /* :93-94 normalised copies */
    size_t nnz = (size_t)row_off[n_reads];
    cx.norm = (float *)malloc((nnz + 1) * sizeof(float));
    memcpy(cx.norm, vals, nnz * sizeof(float));
#pragma omp parallel for schedule(static) num_threads(nt)
    for (size_t i = 0; i < cx.n; ++i) {
        size_t r = cx.orig[i];
        myo_normalize_l2(cx.norm + row_off[r], (size_t)(row_off[r + 1] - row_off[r]));
    }

    /* Step 2 (:99-129) */
    size_t cap_c = p->n_init > p->expected_nb_of_clusters ? p->n_init : p->expected_nb_of_clusters;
    if (cap_c == 0) cap_c = 1;
    cluster_t *cl = (cluster_t *)calloc(cap_c + 1, sizeof(cluster_t));
    size_t nc = 0;
    for (size_t c = 0; c < p->n_init; ++c) {
        size_t b = (size_t)p->init_row_off[c], e = (size_t)p->init_row_off[c + 1];
        cl[nc++].centroid = svec_clone(p->init_keys + b, p->init_vals + b, e - b);
    }
    /* :106-108 max_by_key -> last maximum */
    size_t max_idx = 0;
    uint64_t max_hck = 0;
    for (size_t i = 0; i < cx.n; ++i) {
        uint64_t h = nb_hck[cx.orig[i]];
        if (i == 0 || h >= max_hck) {
            max_hck = h;
            max_idx = i;
        }
    }
    float max_nb_hck = (float)max_hck;
    if (nc == 0) { /* :110-112 from_normalized_centroid: clone + normalise again (:281-286) */
        size_t r = cx.orig[max_idx];
        size_t b = (size_t)row_off[r], e = (size_t)row_off[r + 1];
        cl[nc].centroid = svec_clone(keys + b, cx.norm + b, e - b);
        myo_normalize_l2(cl[nc].centroid.v, cl[nc].centroid.n);
        nc++;
    }
    while (nc < p->expected_nb_of_clusters) { /* :114-129 */
        float *score = (float *)malloc(cx.n * sizeof(float));
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
        for (size_t i = 0; i < cx.n; ++i) {
            float s = 0.0f;
            for (size_t c = 0; c < nc; ++c) s = s + cl_sim_centroid_read(&cx, &cl[c].centroid, i);
            s = s / (float)nc;
            float h = (float)nb_hck[cx.orig[i]];
            float pen = 0.8f + 0.2f * (1.0f - h / max_nb_hck);
            s = s * pen;
            score[i] = s;
        }
        size_t best = 0;
        for (size_t i = 0; i < cx.n; ++i) {
            if (!isfinite(score[i])) { /* :125 SimScore::try_new(score).unwrap() panics */
                free(score);
                goto fail;
            }
            if (score[i] < score[best]) best = i; /* min_by_key -> first minimum */
        }
        free(score);
        size_t r = cx.orig[best];
        size_t b = (size_t)row_off[r], e = (size_t)row_off[r + 1];
        cl[nc].centroid = svec_clone(keys + b, cx.norm + b, e - b);
        myo_normalize_l2(cl[nc].centroid.v, cl[nc].centroid.n);
        nc++;
    }

    /* Step 3 (:132-169) */
    {
        size_t *target = (size_t *)malloc(cx.n * sizeof(size_t));
        float *best_sim = (float *)malloc(cx.n * sizeof(float));
        float previous_mean_wcsss = 1E-6f;
        float improvement_score = 3.40282347e+38f; /* Float::MAX */
        size_t nb_iters = 0;
        while (improvement_score > p->eta_improvement && nb_iters < p->nb_iters_max) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
            for (size_t i = 0; i < cx.n; ++i) target[i] = cl_argmax(&cx, cl, nc, i, &best_sim[i]);
            for (size_t i = 0; i < cx.n; ++i) cluster_push(&cl[target[i]], i);

            /* :159-160; wcsss :307-316 = sum1 over members of simfunc(member, centroid) */
            float sum_w = 0.0f;
            for (size_t c = 0; c < nc; ++c) {
                float w = 0.0f;
                for (size_t t = 0; t < cl[c].n_members; ++t) {
                    size_t i = cl[c].members[t];
                    size_t r = cx.orig[i];
                    size_t b = (size_t)row_off[r], e = (size_t)row_off[r + 1];
                    float s = myo_similarity(cx.sim, 1, keys + b, cx.norm + b, e - b,
                                             cl[c].centroid.k, cl[c].centroid.v, cl[c].centroid.n);
                    if (t == 0)
                        w = s;
                    else
                        w = w + s;
                }
                sum_w = sum_w + w;
            }
            float current_mean_wcsss = sum_w / (float)p->expected_nb_of_clusters;
            improvement_score = (current_mean_wcsss - previous_mean_wcsss) / previous_mean_wcsss;
            previous_mean_wcsss = current_mean_wcsss;

            /* :167 recenter (:288-295) */
            for (size_t c = 0; c < nc; ++c) {
                if (cl[c].n_members > 0) {
                    size_t total = 0;
                    uint64_t *ids = (uint64_t *)malloc(cl[c].n_members * sizeof(uint64_t));
                    for (size_t t = 0; t < cl[c].n_members; ++t) {
                        ids[t] = cx.orig[cl[c].members[t]];
                        total += (size_t)(row_off[ids[t] + 1] - row_off[ids[t]]);
                    }
                    svec_t nc_v;
                    nc_v.k = (uint64_t *)malloc((total + 1) * sizeof(uint64_t));
                    nc_v.v = (float *)malloc((total + 1) * sizeof(float));
                    nc_v.n = myo_centroid(row_off, keys, vals, ids, cl[c].n_members, nc_v.k, nc_v.v);
                    svec_free(&cl[c].centroid);
                    cl[c].centroid = nc_v;
                    free(ids);
                }
                cl[c].n_members = 0;
            }
            nb_iters += 1;
        }
        if (n_iters_out) *n_iters_out = (uint32_t)nb_iters;
        free(best_sim);

        /* Step 5 / first half of step 4 (:200-233): final assignment */
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
        for (size_t i = 0; i < cx.n; ++i) target[i] = cl_argmax(&cx, cl, nc, i, NULL);
        for (size_t i = 0; i < cx.n; ++i) {
            assign_out[cx.orig[i]] = (int32_t)target[i];
            cluster_push(&cl[target[i]], i);
        }
        free(target);
    }

    if (p->has_silhouette && nc != 1) {
        /* Step 4 (:235-267), silhouette :318-365 */
        float ssdcf = p->ssdcf;
        for (size_t c = 0; c < nc; ++c) {
            size_t nm = cl[c].n_members;
            if (nm == 0) continue; /* :241-243 */
            /* :345-351 neighbour = last max over j != c of sim(centroid_c, centroid_j) */
            size_t nb = (size_t)-1;
            float bs = 0.0f;
            for (size_t j = 0; j < nc; ++j) {
                if (j == c) continue;
                float s = myo_similarity(cx.sim, 1, cl[c].centroid.k, cl[c].centroid.v,
                                         cl[c].centroid.n, cl[j].centroid.k, cl[j].centroid.v,
                                         cl[j].centroid.n);
                if (nb == (size_t)-1 || s >= bs) {
                    bs = s;
                    nb = j;
                }
            }
            float *sil = (float *)malloc(nm * sizeof(float));
            const cluster_t *own = &cl[c], *nbr = &cl[nb];
#pragma omp parallel for schedule(dynamic, 8) num_threads(nt)
            for (size_t t = 0; t < nm; ++t) { /* inner() :323-336 */
                size_t i = own->members[t];
                if (g_sil_mask && !g_sil_mask[cx.orig[i]]) {
                    sil[t] = NAN;
                    continue;
                }
                float a = 0.0f;
                for (size_t u = 0; u < own->n_members; ++u) {
                    float d = 1.0f - cl_sim_read_read(&cx, i, own->members[u]);
                    a = a + d;
                }
                a = a / (float)(own->n_members - 1);
                float b = 0.0f;
                for (size_t u = 0; u < nbr->n_members; ++u) {
                    float d = 1.0f - cl_sim_read_read(&cx, i, nbr->members[u]);
                    b = b + d;
                }
                b = b / (float)nbr->n_members;
                sil[t] = (b - a) / fmaxf(a, b);
            }
            if (g_sil_mask) { /* sampled check: scores only, no trimming */
                for (size_t t = 0; t < nm; ++t)
                    if (sil_out) sil_out[cx.orig[own->members[t]]] = sil[t];
                free(sil);
                continue;
            }
            /* :244-249 */
            float n = (float)nm;
            float sum = 0.0f;
            for (size_t t = 0; t < nm; ++t) sum = sum + sil[t];
            float mean = sum / n;
            float sq = 0.0f;
            for (size_t t = 0; t < nm; ++t) {
                float d = sil[t] - mean;
                sq = sq + d * d;
            }
            float std = sqrtf((1.0f / n) * sq);
            float left_cutoff = mean - std * ssdcf;
            for (size_t t = 0; t < nm; ++t) {
                size_t r = cx.orig[own->members[t]];
                if (sil_out) sil_out[r] = sil[t];
                if (sil[t] < left_cutoff) assign_out[r] = -2; /* :257-262 */
            }
            free(sil);
        }
    }

    for (size_t c = 0; c < nc; ++c) {
        svec_free(&cl[c].centroid);
        free(cl[c].members);
    }
    free(cl);
    free(cx.norm);
    free(cx.orig);
    return 0;

fail:
    for (size_t c = 0; c < nc; ++c) {
        svec_free(&cl[c].centroid);
        free(cl[c].members);
    }
    free(cl);
    free(cx.norm);
    free(cx.orig);
    return -12;
}

void myo_pairwise(int sim, const uint64_t *r_off, const uint64_t *r_keys, const float *r_vals,
                  size_t n_rows, const uint64_t *c_off, const uint64_t *c_keys, const float *c_vals,
                  size_t n_cols, int threads, float *out) {
    int nt = clamp_threads(threads);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
    for (size_t i = 0; i < n_rows; ++i) {
        size_t ab = (size_t)r_off[i], ae = (size_t)r_off[i + 1];
        for (size_t j = 0; j < n_cols; ++j) {
            size_t bb = (size_t)c_off[j], be = (size_t)c_off[j + 1];
            out[i * n_cols + j] = myo_similarity(sim, 1, r_keys + ab, r_vals + ab, ae - ab,
                                                 c_keys + bb, c_vals + bb, be - bb);
        }
    }
}
