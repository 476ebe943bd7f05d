// Split-operand staging for the parity-grade tensor-core mode ("bf16x3" / "bf16x6").
//
// An fp32 value x is written as a sum of bf16 parts  x = p0 + p1 (+ p2) + O(2^-17 |x|)  (2^-25 with three parts):
//   p0 = bf16(x),  p1 = bf16(x - p0),  p2 = bf16(x - p0 - p1)          (the subtractions are exact in fp32)
// and an fp32 product A * B becomes a sum of bf16 x bf16 products with fp32 accumulation,
//   two parts,   three products:  p0 q0 + p0 q1 + p1 q0                 (drops p1 q1 ~ 2^-16 relative)
//   three parts, six products:    p0 q0 + p0 q1 + p1 q0 + p1 q1 + p0 q2 + p2 q0
// The products are NOT separate launches: the parts are concatenated along the contraction dimension,
//   A' = [A_a0 | A_a1 | ... ]   B' = [B_b0 | B_b1 | ... ]      (K' = nprod * K)
// so ONE ordinary tcgen05 GEMM launch (gemm_tc.cu / gemm_tc2.cu, bf16 operands, fp32 TMEM accumulators, every fused
// epilogue unchanged) computes the whole sum.  This kernel builds A' (or B') from the fp32 tensor in one pass:
// 4 B read + 2 * nprod B written per element.
//
//   axis = 1 (K-major operand, x [rows, cols = K]):   out[r, j * cols + c] = part_{pat[j]}(x[r, c])
//   axis = 0 (MN-major operand, x [rows = K, cols]):  out[j * rows + r, c] = part_{pat[j]}(x[r, c])
#include <algorithm>

#include "common.cuh"

namespace {

struct SplitArgs {
  const float* x; long long ldx;
  __nv_bfloat16* out; long long ldo;
  long long rows; int cols;
  int axis, nprod;
  int pat[6];
};

__global__ void __launch_bounds__(256) split_concat_kernel(SplitArgs a) {
  const int c4n = a.cols >> 2;
  const long long total = a.rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i - r * c4n) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.x + r * a.ldx + c));
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint2 part[3];
    {
      __nv_bfloat16 h[3][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat16 p0 = __float2bfloat16_rn(f[e]);
        const float r1 = __fsub_rn(f[e], __bfloat162float(p0));
        const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
        const float r2 = __fsub_rn(r1, __bfloat162float(p1));
        h[0][e] = p0; h[1][e] = p1; h[2][e] = __float2bfloat16_rn(r2);
      }
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        __nv_bfloat162 lo2 = __halves2bfloat162(h[p][0], h[p][1]), hi2 = __halves2bfloat162(h[p][2], h[p][3]);
        part[p].x = *reinterpret_cast<uint32_t*>(&lo2);
        part[p].y = *reinterpret_cast<uint32_t*>(&hi2);
      }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (j < a.nprod) {
        const int p = a.pat[j];
        const uint2 u = p == 0 ? part[0] : (p == 1 ? part[1] : part[2]);
        const long long o = a.axis == 1 ? r * a.ldo + (long long)j * a.cols + c : ((long long)j * a.rows + r) * a.ldo + c;
        *reinterpret_cast<uint2*>(a.out + o) = u;
      }
    }
  }
}

}  // namespace

extern "C" int svla_split_concat(svla_ctx* ctx, const float* x, long long ldx, long long rows, int cols, void* out,
                                 long long ldo, int axis, int nprod, const int* pattern, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && out && pattern, "NULL argument");
  SVLA_CHECK_ARG(axis == 0 || axis == 1, "axis must be 0 (stack row blocks) or 1 (concatenate column blocks)");
  SVLA_CHECK_ARG(nprod >= 1 && nprod <= 6, "1..6 products");
  SVLA_CHECK_ARG(rows >= 0 && cols >= 0 && cols % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "cols / ldx / ldo must be multiples of 4");
  SVLA_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, "misaligned buffer");
  SVLA_CHECK_ARG(ldo >= (axis == 1 ? (long long)nprod * cols : (long long)cols), "ldo too small");
  if (rows == 0 || cols == 0) return SVLA_OK;
  SplitArgs a;
  a.x = x; a.ldx = ldx; a.out = reinterpret_cast<__nv_bfloat16*>(out); a.ldo = ldo;
  a.rows = rows; a.cols = cols; a.axis = axis; a.nprod = nprod;
  for (int j = 0; j < 6; ++j) {
    a.pat[j] = j < nprod ? pattern[j] : 0;
    SVLA_CHECK_ARG(a.pat[j] >= 0 && a.pat[j] <= 2, "part index must be 0, 1 or 2");
  }
  const long long total = rows * (cols >> 2);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
  split_concat_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
