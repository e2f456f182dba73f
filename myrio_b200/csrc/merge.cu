// merge.cu -- merge-join similarity kernel (see merge.cuh)
#include "merge.cuh"

namespace myrio {

__device__ __forceinline__ float mg_finite_or_zero(float x) { return (fabsf(x) <= 3.402823466e+38f) ? x : 0.0f; }

// similarity.rs:87-92 / :143-149 / :172-182 on pre-normalised inputs
template <int SIM>
__device__ __forceinline__ float merge_similarity(const uint64_t *__restrict__ ak, const float *__restrict__ av,
                                                  uint64_t an, const uint64_t *__restrict__ bk,
                                                  const float *__restrict__ bv, uint64_t bn) {
    uint64_t i = 0, j = 0;
    if (SIM == MYRIO_SIM_OVERLAP) {
        // merge_and_apply(|x, y| (x.min(y), x.max(y))) then the two sequential sums
        float sum_min = 0.0f, sum_max = 0.0f;
        while (i < an && j < bn) {
            const uint64_t a = ak[i], b = bk[j];
            float x = 0.0f, y = 0.0f;
            if (a <= b) x = av[i++];
            if (b <= a) y = bv[j++];
            sum_min = __fadd_rn(sum_min, fminf(x, y));
            sum_max = __fadd_rn(sum_max, fmaxf(x, y));
        }
        for (; i < an; ++i) {
            sum_min = __fadd_rn(sum_min, fminf(av[i], 0.0f));
            sum_max = __fadd_rn(sum_max, fmaxf(av[i], 0.0f));
        }
        for (; j < bn; ++j) {
            sum_min = __fadd_rn(sum_min, fminf(0.0f, bv[j]));
            sum_max = __fadd_rn(sum_max, fmaxf(0.0f, bv[j]));
        }
        return mg_finite_or_zero(__fdiv_rn(sum_min, sum_max));
    }
    float acc = 0.0f;
    while (i < an && j < bn) {
        const uint64_t a = ak[i], b = bk[j];
        if (a < b)
            ++i;
        else if (a > b)
            ++j;
        else {
            acc = __fadd_rn(acc, __fmul_rn(av[i], bv[j]));
            ++i;
            ++j;
        }
    }
    if (SIM == MYRIO_SIM_JACARD_TANIMOTO) return mg_finite_or_zero(__fdiv_rn(acc, __fsub_rn(2.0f, acc)));
    return acc;
}

// consecutive threads take consecutive ROWS against the same column: the column's entries are
// the same addresses across a warp (cache broadcast), the rows stream from L2
template <int SIM>
__global__ void __launch_bounds__(128) merge_pairs_kernel(MergeArgs a) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (uint64_t)a.n_rows * a.n_cols) return;
    const uint32_t ri = (uint32_t)(p % a.n_rows), ci = (uint32_t)(p / a.n_rows);
    const uint32_t r = a.row_ids ? a.row_ids[ri] : ri;
    const uint32_t c = a.col_ids ? a.col_ids[a.c_first + ci] : (a.c_first + ci);
    const uint64_t rb = a.r_off[r], cb = a.c_off[c];
    const uint64_t ce = a.c_end ? a.c_end[c] : a.c_off[c + 1];
    const float s = merge_similarity<SIM>(a.r_keys + rb, a.r_vals + rb, a.r_off[r + 1] - rb, a.c_keys + cb,
                                          a.c_vals + cb, ce - cb);
    if (a.transposed)
        a.out[(size_t)ci * a.ld + ri] = s;
    else
        a.out[(size_t)ri * a.ld + ci] = s;
}

void launch_merge_pairs(myrio_ctx *ctx, const MergeArgs &a, const char *name) {
    const uint64_t np = (uint64_t)a.n_rows * a.n_cols;
    if (np == 0) return;
    const unsigned grid = (unsigned)((np + 127) / 128);
    switch (a.sim) {
    case MYRIO_SIM_COSINE:
        MYRIO_LAUNCH(ctx, name, merge_pairs_kernel<MYRIO_SIM_COSINE>, grid, 128, 0, a);
        break;
    case MYRIO_SIM_JACARD_TANIMOTO:
        MYRIO_LAUNCH(ctx, name, merge_pairs_kernel<MYRIO_SIM_JACARD_TANIMOTO>, grid, 128, 0, a);
        break;
    case MYRIO_SIM_OVERLAP:
        MYRIO_LAUNCH(ctx, name, merge_pairs_kernel<MYRIO_SIM_OVERLAP>, grid, 128, 0, a);
        break;
    default:
        fail(MYRIO_ERR_BAD_ARG, "unknown similarity %d", a.sim);
    }
}

}  // namespace myrio
