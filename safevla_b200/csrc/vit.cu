// Glue kernels of the rollout-side vision preprocessor (reference
// architecture/allenact_preprocessors/dino_preprocessors.py:20-38,224-239): uint8 camera frames -> normalised
// 14 x 14 patches (the operand of the patch-embedding GEMM), token assembly (class token + position embedding),
// and the AdaptiveAvgPool2d((7, 12)) over the final patch tokens.  All HBM-bound, one pass each.
#include <algorithm>

#include "common.cuh"

namespace {

// out[(n*PH + py)*PW + px][c*P*P + dy*P + dx] = (img[n, py*P + dy, x0 + px*P + dx, c] / 255 - mean[c]) / std[c]
// (DataAugmentationPreprocessor.process :231-237 without augmentation, the [:, :, :, 3:-3] crop of :31 and the
// im2col of the ViT's 14 x 14 stride-14 convolution in one pass).  Columns [3*P*P, Kpad) are zero.
template <typename T>
__global__ void __launch_bounds__(256) patchify_u8_kernel(const uint8_t* __restrict__ img, T* __restrict__ out, int N, int H,
                                                          int W, int P, int PH, int PW, int x0, int Kpad, float3 mean,
                                                          float3 inv_std) {
  const int K = 3 * P * P;
  const long long total = (long long)N * PH * PW * Kpad;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % Kpad);
    const long long row = e / Kpad;
    float v = 0.f;
    if (k < K) {
      const int c = k / (P * P), dy = (k / P) % P, dx = k % P;
      const int px = (int)(row % PW), py = (int)((row / PW) % PH);
      const long long n = row / ((long long)PW * PH);
      const uint8_t u = img[((n * H + py * P + dy) * W + x0 + px * P + dx) * 3 + c];
      const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
      const float is = c == 0 ? inv_std.x : (c == 1 ? inv_std.y : inv_std.z);
      v = ((float)u / 255.0f - m) * is;
    }
    out[e] = from_f<T>(v);
  }
}

// x[n, 0, :] = cls + pos[0];  x[n, 1 + p, :] = patches[n*Np + p, :] + pos[1 + p]
template <typename T>
__global__ void __launch_bounds__(256) vit_assemble_kernel(const T* __restrict__ patches, const float* __restrict__ cls,
                                                           const float* __restrict__ pos, T* __restrict__ x, int N, int Np,
                                                           int D) {
  const int D4 = D / 4;
  const long long total = (long long)N * (Np + 1) * D4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(e % D4) * 4;
    const long long tok = e / D4;
    const int t = (int)(tok % (Np + 1));
    const long long n = tok / (Np + 1);
    const float4 pe = *reinterpret_cast<const float4*>(pos + (long long)t * D + d);
    float4 v = t == 0 ? *reinterpret_cast<const float4*>(cls + d) : load4<T>(patches + (n * Np + t - 1) * D + d);
    v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    store4<T>(x + tok * D + d, v);
  }
}

// AdaptiveAvgPool2d((OH, OW)) over the patch tokens: out[n, d, oy, ox] = mean of x[n, 1 + py*PW + px, d] over
// py in [floor(oy*PH/OH), ceil((oy+1)*PH/OH)), px likewise (torch's window rule); out is fp32 NCHW as the
// reference preprocessor returns it (:32-36, .float() at :125).
template <typename T>
__global__ void __launch_bounds__(128) tokens_pool_kernel(const T* __restrict__ x, float* __restrict__ out, int N, int Np,
                                                          int D, int PH, int PW, int OH, int OW) {
  const int cell = blockIdx.x % (OH * OW);
  const long long n = blockIdx.x / (OH * OW);
  const int oy = cell / OW, ox = cell % OW;
  const int y0 = (oy * PH) / OH, y1 = ((oy + 1) * PH + OH - 1) / OH;
  const int x0 = (ox * PW) / OW, x1 = ((ox + 1) * PW + OW - 1) / OW;
  const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int py = y0; py < y1; ++py)
      for (int px = x0; px < x1; ++px) s += to_f<T>(x[(n * (Np + 1) + 1 + py * PW + px) * D + d]);
    out[((n * D + d) * OH + oy) * OW + ox] = s * inv;
  }
}

}  // namespace

extern "C" int svla_patchify_u8(svla_ctx* ctx, const uint8_t* img, int N, int H, int W, int patch, int crop_left,
                                int crop_right, const float* mean3, const float* std3, void* out, int dtype, int Kpad,
                                svla_stream stream) {
  SVLA_CHECK_ARG(ctx && img && out && mean3 && std3, "NULL argument");
  const int PH = H / patch, PW = (W - crop_left - crop_right) / patch;
  SVLA_CHECK_ARG(patch > 0 && PH * patch == H && PW * patch == W - crop_left - crop_right, "image is not a whole number of patches");
  SVLA_CHECK_ARG(Kpad >= 3 * patch * patch, "Kpad smaller than 3 * patch^2");
  if (N <= 0) return SVLA_OK;
  const long long total = (long long)N * PH * PW * Kpad;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
  const float3 mean = make_float3(mean3[0], mean3[1], mean3[2]);
  const float3 inv = make_float3(1.f / std3[0], 1.f / std3[1], 1.f / std3[2]);
  SVLA_DISPATCH_DTYPE(dtype, T, (patchify_u8_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(img, (T*)out, N, H, W, patch, PH,
                                                                                            PW, crop_left, Kpad, mean, inv)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_vit_assemble(svla_ctx* ctx, const void* patches, const float* cls, const float* pos, void* x, int dtype,
                                 int N, int num_patches, int D, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && patches && cls && pos && x, "NULL argument");
  SVLA_CHECK_ARG(D % 4 == 0, "D must be a multiple of 4");
  if (N <= 0) return SVLA_OK;
  const long long total = (long long)N * (num_patches + 1) * (D / 4);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
  SVLA_DISPATCH_DTYPE(dtype, T, (vit_assemble_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)patches, cls, pos,
                                                                                             (T*)x, N, num_patches, D)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}

extern "C" int svla_tokens_pool(svla_ctx* ctx, const void* x, int dtype, float* out, int N, int PH, int PW, int D, int OH,
                                int OW, svla_stream stream) {
  SVLA_CHECK_ARG(ctx && x && out, "NULL argument");
  SVLA_CHECK_ARG(OH >= 1 && OW >= 1 && OH <= PH && OW <= PW, "bad pooled size");
  if (N <= 0) return SVLA_OK;
  SVLA_DISPATCH_DTYPE(dtype, T, (tokens_pool_kernel<T><<<N * OH * OW, 128, 0, as_stream(stream)>>>((const T*)x, out, N, PH * PW,
                                                                                                   D, PH, PW, OH, OW)));
  SVLA_LAUNCH_CHECK();
  return SVLA_OK;
}
