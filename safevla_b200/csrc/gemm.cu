// svla_gemm: validation + dispatch between the tcgen05 tensor-core kernel (gemm_tc.cu, bf16 operands)
// and the fp32-FMA kernel (gemm_simt.cu).
#include "common.cuh"

int svla_gemm_simt(svla_ctx* ctx, const svla_gemm_desc* d, cudaStream_t st);
int svla_gemm_tc(svla_ctx* ctx, const svla_gemm_desc* d, cudaStream_t st);  // returns SVLA_ERR_BAD_SHAPE if it declines
bool svla_gemm_tc_supported(const svla_gemm_desc* d);
bool svla_gemm_tc_fuses_colsum(const svla_gemm_desc* d);

extern "C" int svla_gemm(svla_ctx* ctx, const svla_gemm_desc* d, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && d, "NULL ctx/desc");
  SVLA_CHECK_ARG(d->A && d->B && d->C, "NULL operand");
  SVLA_CHECK_ARG(d->M >= 0 && d->N >= 0 && d->K >= 0, "negative dimension");
  SVLA_CHECK_ARG(!d->accumulate || d->dtypeC == SVLA_F32, "accumulate needs an fp32 C");
  SVLA_CHECK_ARG(d->epilogue != SVLA_EPI_RELU_MASK || d->aux, "RELU_MASK needs aux");
  SVLA_CHECK_ARG(!d->dropout || d->dropout->p == 0.f || d->epilogue == SVLA_EPI_RELU_BITS,
                 "fused dropout exists for the RELU_BITS epilogue only (use svla_dropout_rows elsewhere)");
  SVLA_CHECK_ARG(svla_dropout_ok(d->dropout), "dropout p must be in [0, 1)");
  if (d->epilogue == SVLA_EPI_RELU_BITS || d->epilogue == SVLA_EPI_MASK_BITS) {
    // bit-record epilogues exist on the tensor-core kernels only (64-column block epilogue)
    SVLA_CHECK_ARG(d->aux && (reinterpret_cast<uintptr_t>(d->aux) & 7) == 0 && d->ldaux % 2 == 0 && d->ldaux * 32 >= d->N,
                   "bit-record epilogue: aux = uint32 [M, N / 32], 8-byte aligned rows");
    svla_gemm_desc e = *d;  // the kernels' own aux checks are about the bf16 / fp32 mask operand
    e.aux = nullptr;
    if (d->impl == 1 || d->dtypeC != SVLA_BF16 || d->residual || d->accumulate || d->colsum_a || d->N % 64 != 0 ||
        d->transA || !svla_gemm_tc_supported(&e)) {
      svla_set_error("svla_gemm: bit-record epilogue needs the tcgen05 path (bf16 C, N %% 64 == 0, no residual / "
                     "accumulate): M=%d N=%d K=%d", d->M, d->N, d->K);
      return SVLA_ERR_BAD_SHAPE;
    }
  }
  SVLA_CHECK_ARG(d->lda >= (d->transA ? d->M : d->K), "lda too small");
  SVLA_CHECK_ARG(d->ldb >= (d->transB ? d->K : d->N), "ldb too small");
  SVLA_CHECK_ARG(d->ldc >= d->N, "ldc too small");
  SVLA_CHECK_ARG(!d->colsum_a || d->transA, "colsum_a needs transA (A stored [K, M])");
  if (d->M == 0 || d->N == 0) return SVLA_OK;
  const cudaStream_t st = as_stream(stream);
  // validated above: the tensor-core path takes it.  (Not via svla_gemm_tc_supported(d) below -- that one applies the
  // 16-byte row rule of a bf16 / fp32 mask operand to aux, which a record of N / 32 words need not meet; falling
  // through to the CUDA-core kernel would silently drop the record.)
  if (d->epilogue == SVLA_EPI_RELU_BITS || d->epilogue == SVLA_EPI_MASK_BITS) return svla_gemm_tc(ctx, d, st);
  if (d->colsum_a) {
    // bias gradient: produced by the tensor-core weight-gradient kernel itself when that path runs, else one extra pass
    if (d->impl != 1 && svla_gemm_tc_fuses_colsum(d)) return svla_gemm_tc(ctx, d, st);
    svla_gemm_desc e = *d;
    e.colsum_a = nullptr;
    const int rc = svla_gemm(ctx, &e, stream);
    if (rc) return rc;
    return svla_colsum(ctx, d->A, d->dtypeA, d->K, d->M, d->lda, d->colsum_a, 1, stream);
  }
  if (d->impl == 2) {
    if (!svla_gemm_tc_supported(d)) {
      svla_set_error("svla_gemm: impl=tcgen05 requested for an unsupported shape/dtype (M=%d N=%d K=%d)", d->M, d->N,
                     d->K);
      return SVLA_ERR_BAD_SHAPE;
    }
    return svla_gemm_tc(ctx, d, st);
  }
  if (d->impl == 0 && svla_gemm_tc_supported(d)) return svla_gemm_tc(ctx, d, st);
  return svla_gemm_simt(ctx, d, st);
}

extern "C" int svla_gemm_which(const svla_gemm_desc* d) {
  if (!d) return 0;
  if (d->impl == 1) return 1;
  return svla_gemm_tc_supported(d) ? 2 : 1;
}
