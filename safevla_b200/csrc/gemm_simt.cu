// fp32-FMA GEMM (CUDA cores).  This is the *parity-mode* matmul of the towers (fp32 operands,
// fp32 accumulation: the path the 1e-4 gate of BASELINE config 1 runs on), the skinny/irregular
// matmuls (heads with A+1 columns, K = 384 compressor in fp32) and the fallback for shapes the
// tcgen05 kernel (gemm_tc.cu) does not take.  128x128x16 tiles, 256 threads, 8x8 register tile,
// register-staged double buffering, split-K through the context workspace with a fixed-order
// reduction (deterministic).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

struct GemmArgs {
  svla_gemm_desc d;
  int k_per_split;  // multiple of BK
  int splits;
  float* ws;        // [splits, M, N] when splits > 1
};

__device__ __forceinline__ float ld_any(const void* p, int dt, long long i) {
  return dt == SVLA_F32 ? __ldg(reinterpret_cast<const float*>(p) + i)
                        : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dt, long long i, float v) {
  if (dt == SVLA_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// loads 8 consecutive elements starting at element index i (valid count `nvalid` <= 8, rest zero)
__device__ __forceinline__ void load8(const void* p, int dt, long long i, int nvalid, bool aligned, float* out) {
  if (nvalid == 8 && aligned) {
    if (dt == SVLA_F32) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4));
      out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    } else {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        out[2 * j] = f.x; out[2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = (j < nvalid) ? ld_any(p, dt, i + j) : 0.f;
  }
}

// Tile loader.  `contig_k`: the operand's contiguous dimension is K (A non-transposed / B = nn.Linear
// weight); otherwise the contiguous dimension is the tile's row dimension (M or N).
// smem layout is always S[k][row] (row = m or n), BK x (128 + pad).
template <bool CONTIG_K>
__device__ __forceinline__ void fetch_tile(const void* p, int dt, long long ld, int row0, int nrows, int k0, int kend,
                                           bool aligned, float* regs) {
  const int t = threadIdx.x;
  if (CONTIG_K) {
    const int r = t >> 1, kk = (t & 1) * 8;
    const int row = row0 + r, k = k0 + kk;
    const int nv = (row < nrows) ? max(0, min(8, kend - k)) : 0;
    if (nv > 0) load8(p, dt, (long long)row * ld + k, nv, aligned, regs);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) regs[j] = 0.f;
    }
  } else {
    const int kk = t >> 4, r = (t & 15) * 8;
    const int k = k0 + kk, row = row0 + r;
    const int nv = (k < kend) ? max(0, min(8, nrows - row)) : 0;
    if (nv > 0) load8(p, dt, (long long)k * ld + row, nv, aligned, regs);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) regs[j] = 0.f;
    }
  }
}
template <bool CONTIG_K>
__device__ __forceinline__ void stash_tile(float (*S)[BM + 4], const float* regs) {
  const int t = threadIdx.x;
  if (CONTIG_K) {
    const int r = t >> 1, kk = (t & 1) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) S[kk + j][r] = regs[j];
  } else {
    const int kk = t >> 4, r = (t & 15) * 8;
    *reinterpret_cast<float4*>(&S[kk][r]) = make_float4(regs[0], regs[1], regs[2], regs[3]);
    *reinterpret_cast<float4*>(&S[kk][r + 4]) = make_float4(regs[4], regs[5], regs[6], regs[7]);
  }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const svla_gemm_desc& d = g.d;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_per_split, kend = min(d.K, kbeg + g.k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // A: contiguous in K when !TA ; B: contiguous in K when TB
  const int esA = d.dtypeA == SVLA_F32 ? 4 : 2, esB = d.dtypeB == SVLA_F32 ? 4 : 2;
  const bool alA = ((reinterpret_cast<uintptr_t>(d.A) & 15) == 0) && ((d.lda * esA) % 16 == 0) &&
                   (!TA ? ((kbeg * esA) % 16 == 0) : true);
  const bool alB = ((reinterpret_cast<uintptr_t>(d.B) & 15) == 0) && ((d.ldb * esB) % 16 == 0) &&
                   (TB ? ((kbeg * esB) % 16 == 0) : true);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  const int nk = (kend - kbeg + BK - 1) / BK;
  if (nk > 0) {
    fetch_tile<!TA>(d.A, d.dtypeA, d.lda, m0, d.M, kbeg, kend, alA, ra);
    fetch_tile<TB>(d.B, d.dtypeB, d.ldb, n0, d.N, kbeg, kend, alB, rb);
    stash_tile<!TA>(As[0], ra);
    stash_tile<TB>(Bs[0], rb);
  }
  __syncthreads();
  for (int it = 0; it < nk; ++it) {
    const int cur = it & 1;
    if (it + 1 < nk) {
      const int k0 = kbeg + (it + 1) * BK;
      fetch_tile<!TA>(d.A, d.dtypeA, d.lda, m0, d.M, k0, kend, alA, ra);
      fetch_tile<TB>(d.B, d.dtypeB, d.ldb, n0, d.N, k0, kend, alB, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (it + 1 < nk) {
      stash_tile<!TA>(As[cur ^ 1], ra);
      stash_tile<TB>(Bs[cur ^ 1], rb);
    }
    __syncthreads();
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= d.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= d.N) continue;
      float v = acc[i][j];
      if (g.splits > 1) {
        g.ws[((size_t)blockIdx.z * d.M + m) * d.N + n] = v;
        continue;
      }
      v *= d.alpha;
      if (d.bias) v += __ldg(d.bias + n);
      if (d.epilogue == SVLA_EPI_RELU) v = fmaxf(v, 0.f);
    else if (d.epilogue == SVLA_EPI_GELU) v = gelu_erf(v);
      else if (d.epilogue == SVLA_EPI_GELU) v = gelu_erf(v);
      else if (d.epilogue == SVLA_EPI_RELU_MASK) v = (ld_any(d.aux, d.dtypeAux, (long long)m * d.ldaux + n) > 0.f) ? v : 0.f;
      if (d.residual) v += ld_any(d.residual, d.dtypeR, (long long)m * d.ldr + n);
      const long long ci = (long long)m * d.ldc + n;
      if (d.accumulate) v += ld_any(d.C, d.dtypeC, ci);
      st_any(d.C, d.dtypeC, ci, v);
    }
  }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(GemmArgs g) {
  const svla_gemm_desc& d = g.d;
  const long long total = (long long)d.M * d.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / d.N), n = (int)(i % d.N);
    float v = 0.f;
    for (int s = 0; s < g.splits; ++s) v += g.ws[(size_t)s * total + i];
    v *= d.alpha;
    if (d.bias) v += __ldg(d.bias + n);
    if (d.epilogue == SVLA_EPI_RELU) v = fmaxf(v, 0.f);
    else if (d.epilogue == SVLA_EPI_GELU) v = gelu_erf(v);
    else if (d.epilogue == SVLA_EPI_RELU_MASK) v = (ld_any(d.aux, d.dtypeAux, (long long)m * d.ldaux + n) > 0.f) ? v : 0.f;
    if (d.residual) v += ld_any(d.residual, d.dtypeR, (long long)m * d.ldr + n);
    const long long ci = (long long)m * d.ldc + n;
    if (d.accumulate) v += ld_any(d.C, d.dtypeC, ci);
    st_any(d.C, d.dtypeC, ci, v);
  }
}

}  // namespace

int svla_gemm_simt(svla_ctx* ctx, const svla_gemm_desc* d, cudaStream_t st) {
  SVLA_CHECK_ARG(d->epilogue != SVLA_EPI_RELU_BITS && d->epilogue != SVLA_EPI_MASK_BITS,
                 "the CUDA-core GEMM has no bit-record epilogue");
  GemmArgs g;
  g.d = *d;
  const int tiles = ((d->M + BM - 1) / BM) * ((d->N + BN - 1) / BN);
  int splits = 1;
  const size_t per_split = (size_t)d->M * d->N * sizeof(float);
  if (tiles < ctx->sm_count && d->K >= 4096) {
    splits = std::min({(2 * ctx->sm_count + tiles - 1) / tiles, d->K / (BK * 16), 64});
    splits = (int)std::min<size_t>((size_t)splits, ctx->ws_bytes / std::max<size_t>(per_split, 1));
    splits = std::max(splits, 1);
  }
  int kps = (d->K + splits - 1) / splits;
  kps = (kps + 63) / 64 * 64;  // keeps 16-byte alignment of the K offset for both dtypes
  splits = (d->K + kps - 1) / kps;
  g.k_per_split = kps;
  g.splits = splits;
  g.ws = reinterpret_cast<float*>(ctx->ws);
  dim3 grid((d->N + BN - 1) / BN, (d->M + BM - 1) / BM, splits);
  if (!d->transA && !d->transB) gemm_simt_kernel<false, false><<<grid, NT, 0, st>>>(g);
  else if (!d->transA && d->transB) gemm_simt_kernel<false, true><<<grid, NT, 0, st>>>(g);
  else if (d->transA && !d->transB) gemm_simt_kernel<true, false><<<grid, NT, 0, st>>>(g);
  else gemm_simt_kernel<true, true><<<grid, NT, 0, st>>>(g);
  SVLA_LAUNCH_CHECK();
  if (splits > 1) {
    const long long total = (long long)d->M * d->N;
    const int rb = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8);
    splitk_reduce_kernel<<<rb, 256, 0, st>>>(g);
    SVLA_LAUNCH_CHECK();
  }
  return SVLA_OK;
}
