// ViT stem on sm_100a: patch embedding as im2col -> tcgen05 GEMM -> token assembly.
//   im2col   : [B, 3, H, W] fp32 image -> [B*L, 3*P*P] fp16 patch rows (column = c*P*P + py*P + px,
//              the flattening of the Conv2d weight [C, 3, P, P]), 8 pixels per thread
//   GEMM     : patches x W^T + b  (gemm_tn, EPI_BIAS)  -> [B*L, C] fp16   (the fp16 conv output of
//              the autocast reference)
//   assemble : x[b,0] = cls + pos[0];  x[b,1+p] = patch[b,p] + pos[1+p]   -> fp32 residual stream
// Replaces PatchEmbed.proj (Conv2d k=s=16) + cls concat + pos_embed add of the reference
// (models/model_speed_test.py:467-472; timm PatchEmbed), SURVEY.md section 8f rank 2.
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "gemm_tn.cuh"
#include "host_utils.h"
#include "internal.h"
#include "rowwise.cuh"

namespace dyt {

__global__ void __launch_bounds__(256)
im2col_patch_kernel(const float* __restrict__ img, int B, int Cin, int H, int W, int P,
                    __half* __restrict__ out, int ldo) {
  const int w8 = W >> 3;                       // 8-pixel groups per image row
  const long long total = static_cast<long long>(B) * Cin * H * w8;
  const int gw = W / P, gh = H / P;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xg = static_cast<int>(i % w8);
    long long t = i / w8;
    const int y = static_cast<int>(t % H);
    t /= H;
    const int c = static_cast<int>(t % Cin);
    const int b = static_cast<int>(t / Cin);
    const int x = xg << 3;
    const float4* src = reinterpret_cast<const float4*>(
        img + ((static_cast<size_t>(b) * Cin + c) * H + y) * W + x);
    const float4 v0 = src[0], v1 = src[1];
    const int py = y / P, px = x / P;
    const size_t row = (static_cast<size_t>(b) * gh + py) * gw + px;
    const int col = c * P * P + (y - py * P) * P + (x - px * P);
    __half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w);
    __half2 h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(out + row * ldo + col) = u;
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const __half* __restrict__ patch, int ldp, const float* __restrict__ cls,
                       const float* __restrict__ pos, int B, int N, float* __restrict__ x, int ldx) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int rows = B * N;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += gridDim.x * warps_per_block) {
    const int b = r / N, n = r - b * N;
    float4 v[NV];
    if (n == 0) {
      load_row_f32<NV>(cls, lane, v);
    } else {
      const uint2* m = reinterpret_cast<const uint2*>(
          patch + (static_cast<size_t>(b) * (N - 1) + (n - 1)) * ldp);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint2 u = m[i * 32 + lane];
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        v[i] = make_float4(a.x, a.y, c.x, c.y);
      }
    }
    const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<size_t>(n) * (NV * 128));
    float4* o = reinterpret_cast<float4*>(x + static_cast<size_t>(r) * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 q = pp[i * 32 + lane];
      o[i * 32 + lane] = make_float4(v[i].x + q.x, v[i].y + q.y, v[i].z + q.z, v[i].w + q.w);
    }
  }
}

}  // namespace dyt

extern "C" size_t dyt_patch_embed_workspace_bytes(int B, int H, int W, int P, int Cin, int C) {
  if (B <= 0 || P <= 0 || H % P || W % P) return 0;
  const size_t L = static_cast<size_t>(H / P) * (W / P);
  const size_t K = static_cast<size_t>(Cin) * P * P;
  return ((B * L * K * 2 + 255) & ~static_cast<size_t>(255)) + B * L * static_cast<size_t>(C) * 2;
}

extern "C" int dyt_patch_embed_fwd(const float* img, int B, int Cin, int H, int W, int P,
                                   const void* w_f16, const void* bias_f16, const float* cls,
                                   const float* pos, int C, float* x_out, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
  using namespace dyt;
  DYT_CHECK_ARG(img && w_f16 && cls && pos && x_out && workspace, "patch_embed: null buffer");
  DYT_CHECK_ARG(B >= 1 && Cin >= 1 && P >= 8 && P % 8 == 0 && H % P == 0 && W % P == 0 && W % 8 == 0,
                "patch_embed: bad geometry B=%d Cin=%d H=%d W=%d P=%d", B, Cin, H, W, P);
  DYT_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                "patch_embed: image must be 16-byte and workspace 256-byte aligned");
  const size_t need = dyt_patch_embed_workspace_bytes(B, H, W, P, Cin, C);
  DYT_CHECK_ARG(workspace_bytes >= need, "patch_embed: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int L = (H / P) * (W / P);
  const int K = Cin * P * P;
  __half* cols = static_cast<__half*>(workspace);
  __half* patch = reinterpret_cast<__half*>(
      static_cast<char*>(workspace) + ((static_cast<size_t>(B) * L * K * 2 + 255) & ~static_cast<size_t>(255)));
  {
    const long long total = static_cast<long long>(B) * Cin * H * (W / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 32;
    if (blocks > cap) blocks = cap;
    im2col_patch_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(img, B, Cin, H, W, P, cols, K);
    DYT_CUDA(cudaGetLastError());
  }
  int st = gemm_tn(cols, K, static_cast<const __half*>(w_f16), K, B * L, C, K, nullptr, EPI_BIAS,
                   static_cast<const __half*>(bias_f16), patch, C, nullptr, 0, nullptr, 0, 1.0f,
                   stream);
  if (st != 0) return st;
  {
    const int rows = B * (L + 1);
    int grid = (rows + 7) / 8;
    const int cap = sm_count() * 16;
    if (grid > cap) grid = cap;
    switch (C) {
      case 768:
        assemble_tokens_kernel<6><<<grid, 256, 0, stream>>>(patch, C, cls, pos, B, L + 1, x_out, C);
        break;
      case 1024:
        assemble_tokens_kernel<8><<<grid, 256, 0, stream>>>(patch, C, cls, pos, B, L + 1, x_out, C);
        break;
      case 384:
        assemble_tokens_kernel<3><<<grid, 256, 0, stream>>>(patch, C, cls, pos, B, L + 1, x_out, C);
        break;
      case 128:
        assemble_tokens_kernel<1><<<grid, 256, 0, stream>>>(patch, C, cls, pos, B, L + 1, x_out, C);
        break;
      default:
        return fail(DYT_EUNSUPPORTED, "patch_embed: embed dim %d not instantiated", C);
    }
    DYT_CUDA(cudaGetLastError());
  }
  return DYT_OK;
}
