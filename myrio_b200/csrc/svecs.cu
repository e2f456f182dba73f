// svecs.cu -- batches of SparseVec<T> in CSR form on the device:
// upload / download / L2 normalisation.
//
// normalize_l2 restated (myrio-core/src/data/sparse.rs:366-388):
//   norm_l2_squared = values.map(|v| v.powi(2)).sum()   -- sequential, key order
//   magnitude = sqrt;  every value /= magnitude
// u16 leaf counts are first converted with Float::from (tax/compute.rs:105-106).
#include "kmer_count.cuh"

namespace myrio {

// one thread per row: the sum must be sequential in key order (f32 is not associative)
template <typename VT>
__global__ void __launch_bounds__(128) row_norm_kernel(const uint64_t *__restrict__ row_off,
                                                       const VT *__restrict__ vals,
                                                       float *__restrict__ magn, uint64_t n_rows) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t b = row_off[row], e = row_off[row + 1];
    float s = 0.0f;
    for (uint64_t i = b; i < e; ++i) {
        const float v = (float)vals[i];
        s = __fadd_rn(s, __fmul_rn(v, v));
    }
    magn[row] = __fsqrt_rn(s);
}

// one warp per row, coalesced
template <typename VT>
__global__ void __launch_bounds__(256) row_scale_kernel(const uint64_t *__restrict__ row_off,
                                                        const VT *__restrict__ vals,
                                                        const float *__restrict__ magn,
                                                        float *__restrict__ out, uint64_t n_rows) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t b = row_off[row], e = row_off[row + 1];
    const float m = magn[row];
    for (uint64_t i = b + lane; i < e; i += 32) out[i] = __fdiv_rn((float)vals[i], m);
}

myrio_svecs *svecs_normalize(myrio_ctx *ctx, const myrio_svecs *in) {
    auto out = std::make_unique<myrio_svecs>();
    out->n_rows = in->n_rows;
    out->nnz = in->nnz;
    out->vtype = MYRIO_VAL_F32;
    out->aux_kind = in->aux_kind;
    out->row_off.alloc(in->n_rows + 1);
    out->keys.alloc(std::max<uint64_t>(in->nnz, 1));
    out->vals_f32.alloc(std::max<uint64_t>(in->nnz, 1));
    MYRIO_CK(cudaMemcpyAsync(out->row_off.p, in->row_off.p, (in->n_rows + 1) * sizeof(uint64_t),
                             cudaMemcpyDeviceToDevice, ctx->stream));
    if (in->nnz)
        MYRIO_CK(cudaMemcpyAsync(out->keys.p, in->keys.p, in->nnz * sizeof(uint64_t),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    if (in->aux_kind == 1) {
        out->aux_u64.alloc(std::max<uint64_t>(in->n_rows, 1));
        MYRIO_CK(cudaMemcpyAsync(out->aux_u64.p, in->aux_u64.p, in->n_rows * sizeof(uint64_t),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    } else if (in->aux_kind == 2) {
        out->aux_f32.alloc(std::max<uint64_t>(in->n_rows, 1));
        MYRIO_CK(cudaMemcpyAsync(out->aux_f32.p, in->aux_f32.p, in->n_rows * sizeof(float),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (in->n_rows && in->nnz) {
        DevBuf<float> magn(in->n_rows);
        const unsigned g1 = (unsigned)((in->n_rows + 127) / 128), g2 = (unsigned)((in->n_rows + 7) / 8);
        if (in->vtype == MYRIO_VAL_U16) {
            MYRIO_LAUNCH(ctx, "row_norm", row_norm_kernel<uint16_t>, g1, 128, 0, in->row_off.p,
                         in->vals_u16.p, magn.p, in->n_rows);
            MYRIO_LAUNCH(ctx, "row_scale", row_scale_kernel<uint16_t>, g2, 256, 0, in->row_off.p,
                         in->vals_u16.p, magn.p, out->vals_f32.p, in->n_rows);
        } else {
            MYRIO_LAUNCH(ctx, "row_norm", row_norm_kernel<float>, g1, 128, 0, in->row_off.p,
                         in->vals_f32.p, magn.p, in->n_rows);
            MYRIO_LAUNCH(ctx, "row_scale", row_scale_kernel<float>, g2, 256, 0, in->row_off.p,
                         in->vals_f32.p, magn.p, out->vals_f32.p, in->n_rows);
        }
        sync(ctx);
    } else {
        sync(ctx);
    }
    return out.release();
}

}  // namespace myrio
