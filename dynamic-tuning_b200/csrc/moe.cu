// MoE-adapter (BASELINE configs[3]; DyT paper, arXiv 2403.11808: "inspired by the mixture-of-experts
// mechanism, we introduce an enhanced adapter").  NOT part of the reference repository (SURVEY.md
// section 0.6): there is no reference code to replace and no reference parity -- the arithmetic is
// pinned only to this repository's own restatement, oracle/dyt_oracle.py `moe_adapter`.
//
//   alpha[b, :] = softmax(router(mean_tokens(x1[b])))                     per image, E experts
//   W_down_mix = sum_i alpha_i W_down^i,  W_up_mix = sum_i alpha_i W_up^i  (biases likewise)
//   adapt = scale * (relu(x1 W_down_mix^T + b_down_mix) W_up_mix^T + b_up_mix)
//
// Both projections are linear in their weights, so the mixture is applied to the expert OUTPUTS
// instead of building per-image weights:  x1 W_down_mix^T = sum_i alpha_i (x1 W_down^i^T)  -- one
// ordinary GEMM over the concatenated experts [E * K, C] -- and  d W_up_mix^T = [alpha_1 d | ... |
// alpha_E d | alpha] [W_up^1 | ... | W_up^E | b_up^1 .. b_up^E]^T  -- one ordinary GEMM with
// K' = E * K + E (the mixed bias rides along as E extra columns).  Two small kernels sit in between:
//   moe_mean_kernel     mean over the tokens of an image (fp32)
//   moe_route_kernel    router logits with fp16 operands / fp32 accumulation / one fp16 rounding,
//                       softmax in fp32
//   moe_combine_kernel  d = relu(f16(sum_i alpha_i h_i + sum_i alpha_i b_down^i)); writes the
//                       expanded operand [alpha_i * d | alpha] of the up GEMM (fp16)
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "internal.h"
#include "ptx.cuh"

namespace dyt {

constexpr int MOE_MAX_E = 8;

// mean over the tokens of an image, fp32: CTA = (image, slab of 256 columns), thread = (row group
// tid / 64, four columns): every row group reads 1 KB contiguous per row, four rows in flight per
// thread; the four row groups are summed through shared memory.  (One CTA per image walking its
// columns thread by thread took 192 us per layer at 128 x 197 x 1024.)
__global__ void __launch_bounds__(256)
moe_mean_kernel(const float* __restrict__ x1, int ldx, int N, int C, float* __restrict__ mean) {
  __shared__ float4 part[4][64];
  const int b = blockIdx.x;
  const int rg = threadIdx.x >> 6, c4 = threadIdx.x & 63;
  const int col = blockIdx.y * 256 + c4 * 4;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
  if (col < C) {
    const float* xb = x1 + static_cast<size_t>(b) * N * ldx + col;
    int n = rg;
    for (; n + 12 < N; n += 16) {
      const float4 a = *reinterpret_cast<const float4*>(xb + static_cast<size_t>(n) * ldx);
      const float4 bb = *reinterpret_cast<const float4*>(xb + static_cast<size_t>(n + 4) * ldx);
      const float4 c = *reinterpret_cast<const float4*>(xb + static_cast<size_t>(n + 8) * ldx);
      const float4 d = *reinterpret_cast<const float4*>(xb + static_cast<size_t>(n + 12) * ldx);
      s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
      s1.x += bb.x; s1.y += bb.y; s1.z += bb.z; s1.w += bb.w;
      s2.x += c.x; s2.y += c.y; s2.z += c.z; s2.w += c.w;
      s3.x += d.x; s3.y += d.y; s3.z += d.z; s3.w += d.w;
    }
    for (; n < N; n += 4) {
      const float4 a = *reinterpret_cast<const float4*>(xb + static_cast<size_t>(n) * ldx);
      s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
    }
  }
  part[rg][c4] = make_float4((s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y),
                             (s0.z + s1.z) + (s2.z + s3.z), (s0.w + s1.w) + (s2.w + s3.w));
  __syncthreads();
  if (rg == 0 && col < C) {
    const float inv_n = 1.0f / static_cast<float>(N);
    const float4 p0 = part[0][c4], p1 = part[1][c4], p2 = part[2][c4], p3 = part[3][c4];
    *reinterpret_cast<float4*>(mean + static_cast<size_t>(b) * C + col) =
        make_float4(((p0.x + p1.x) + (p2.x + p3.x)) * inv_n, ((p0.y + p1.y) + (p2.y + p3.y)) * inv_n,
                    ((p0.z + p1.z) + (p2.z + p3.z)) * inv_n, ((p0.w + p1.w) + (p2.w + p3.w)) * inv_n);
  }
}

// router logits (fp16 operands, fp32 accumulation, one fp16 rounding: the autocast Linear) and the
// softmax (fp32) of one image: a warp per expert
__global__ void __launch_bounds__(256)
moe_route_kernel(const float* __restrict__ mean, int C, const float* __restrict__ rw,
                 const float* __restrict__ rb, int E, float* __restrict__ alpha) {
  __shared__ float logit_s[MOE_MAX_E];
  const int b = blockIdx.x;
  const float* mb = mean + static_cast<size_t>(b) * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = warp; e < E; e += blockDim.x >> 5) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32)
      acc = fmaf(round_f16(mb[c]), round_f16(rw[static_cast<size_t>(e) * C + c]), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logit_s[e] = round_f16(acc + round_f16(rb != nullptr ? rb[e] : 0.f));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = logit_s[0];
    for (int e = 1; e < E; ++e) m = fmaxf(m, logit_s[e]);
    float ex[MOE_MAX_E], sum = 0.f;
    for (int e = 0; e < E; ++e) {
      ex[e] = expf(logit_s[e] - m);
      sum += ex[e];
    }
    for (int e = 0; e < E; ++e) alpha[static_cast<size_t>(b) * E + e] = ex[e] / sum;
  }
}

// one thread per (token, pair of bottleneck columns)
__global__ void __launch_bounds__(256)
moe_combine_kernel(const __half* __restrict__ hid, int ldh, const float* __restrict__ alpha,
                   const __half* __restrict__ bd, int T, int N, int E, int K, __half* __restrict__ aup,
                   int lda, int kup) {
  const int half_k = K >> 1;
  const size_t total = static_cast<size_t>(T) * half_k;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / half_k);
    const int c = static_cast<int>(i - static_cast<size_t>(t) * half_k) * 2;
    const float* al = alpha + static_cast<size_t>(t / N) * E;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    for (int e = 0; e < E; ++e) {
      const float w = al[e];
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(hid + static_cast<size_t>(t) * ldh + e * K + c));
      a0 = fmaf(w, h.x, a0);
      a1 = fmaf(w, h.y, a1);
      if (bd != nullptr) {
        const float2 bb = __half22float2(*reinterpret_cast<const __half2*>(bd + e * K + c));
        b0 = fmaf(w, bb.x, b0);
        b1 = fmaf(w, bb.y, b1);
      }
    }
    const float d0 = fmaxf(round_f16(a0 + b0), 0.f), d1 = fmaxf(round_f16(a1 + b1), 0.f);
    __half* row = aup + static_cast<size_t>(t) * lda;
    for (int e = 0; e < E; ++e)
      *reinterpret_cast<__half2*>(row + e * K + c) = __floats2half2_rn(al[e] * d0, al[e] * d1);
    if (c == 0) {   // the mixture weights themselves (bias columns of the up GEMM) and the padding
      for (int e = 0; e < E; ++e) row[E * K + e] = __float2half_rn(al[e]);
      for (int j = E * K + E; j < kup; ++j) row[j] = __float2half_rn(0.f);
    }
  }
}

size_t moe_workspace_bytes(int B, int N, int E, int K) {
  const size_t T = static_cast<size_t>(B) * N;
  const size_t kup = (static_cast<size_t>(E) * K + E + 7) & ~static_cast<size_t>(7);
  auto a256 = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  return a256(static_cast<size_t>(B) * E * 4) + a256(T * E * K * 2) + a256(T * kup * 2) +
         a256(static_cast<size_t>(B) * 1024 * 4);   // + the token means [B, C <= 1024]
}

// adapt[T, C] = scale * MoE-adapter(x1): x1 fp32 [B*N, C], x1h its fp16 copy; weights:
// down_cat [E*K, C] fp16, down_b [E, K] fp16, up_cat [C, kup] fp16 = [W_up^1 | .. | W_up^E | b_up^1 .. b_up^E | 0],
// router_w [E, C] fp32, router_b [E] fp32.
int moe_adapter_fwd(const float* x1, int ldx, const __half* x1h, int ldxh, int B, int N, int C, int E,
                    int K, const float* router_w, const float* router_b, const __half* down_cat,
                    const __half* down_b, const __half* up_cat, float scale, __half* adapt, int ld_adapt,
                    void* ws, size_t ws_bytes, cudaStream_t stream) {
  DYT_CHECK_ARG(x1 && x1h && router_w && down_cat && up_cat && adapt && ws, "moe_adapter: null buffer");
  DYT_CHECK_ARG(E >= 1 && E <= MOE_MAX_E && K >= 8 && K % 8 == 0 && C % 8 == 0 && C <= 1024 && ldx % 4 == 0,
                "moe_adapter: 1..8 experts, bottleneck multiple of 8, C <= 1024 (E=%d K=%d C=%d)", E, K, C);
  DYT_CHECK_ARG(ws_bytes >= moe_workspace_bytes(B, N, E, K), "moe_adapter: workspace too small");
  const int T = B * N;
  const int kup = (E * K + E + 7) & ~7;
  auto a256 = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  char* base = static_cast<char*>(ws);
  float* alpha = reinterpret_cast<float*>(base);
  __half* hid = reinterpret_cast<__half*>(base + a256(static_cast<size_t>(B) * E * 4));
  __half* aup = reinterpret_cast<__half*>(base + a256(static_cast<size_t>(B) * E * 4) +
                                          a256(static_cast<size_t>(T) * E * K * 2));
  float* mean = reinterpret_cast<float*>(base + a256(static_cast<size_t>(B) * E * 4) +
                                         a256(static_cast<size_t>(T) * E * K * 2) +
                                         a256(static_cast<size_t>(T) * kup * 2));
  moe_mean_kernel<<<dim3(B, (C + 255) / 256), 256, 0, stream>>>(x1, ldx, N, C, mean);
  DYT_CUDA(cudaGetLastError());
  moe_route_kernel<<<B, 256, 0, stream>>>(mean, C, router_w, router_b, E, alpha);
  DYT_CUDA(cudaGetLastError());
  // every expert's down projection in one GEMM (no bias, no activation: both are applied after the mixture)
  int st = gemm_tn(x1h, ldxh, down_cat, C, T, E * K, C, nullptr, 0 /* EPI_BIAS */, nullptr, hid, E * K,
                   nullptr, 0, nullptr, 0, 1.0f, stream);
  if (st != DYT_OK) return st;
  {
    const size_t total = static_cast<size_t>(T) * (K / 2);
    int grid = static_cast<int>((total + 255) / 256);
    const int cap = sm_count() * 16;
    if (grid > cap) grid = cap;
    moe_combine_kernel<<<grid, 256, 0, stream>>>(hid, E * K, alpha, down_b, T, N, E, K, aup, kup, kup);
    DYT_CUDA(cudaGetLastError());
  }
  // mixed up projection (+ mixed bias through the alpha columns), f16(f16(acc) * scale)
  return gemm_tn(aup, kup, up_cat, kup, T, C, kup, nullptr, 0 /* EPI_BIAS */, nullptr, adapt, ld_adapt,
                 nullptr, 0, nullptr, 0, scale, stream);
}

}  // namespace dyt

extern "C" size_t dyt_moe_workspace_bytes(int B, int N, int E, int K) {
  if (B < 1 || N < 1 || E < 1 || E > dyt::MOE_MAX_E || K < 8) return 0;
  return dyt::moe_workspace_bytes(B, N, E, K);
}

extern "C" int dyt_moe_adapter_fwd(const float* x1, int ldx, const void* x1_f16, int ldxh, int B, int N,
                                   int C, int E, int K, const float* router_w, const float* router_b,
                                   const void* down_cat_f16, const void* down_b_f16,
                                   const void* up_cat_f16, float scale, void* adapt_f16, int ld_adapt,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  return dyt::moe_adapter_fwd(x1, ldx, static_cast<const __half*>(x1_f16), ldxh, B, N, C, E, K, router_w,
                              router_b, static_cast<const __half*>(down_cat_f16),
                              static_cast<const __half*>(down_b_f16),
                              static_cast<const __half*>(up_cat_f16), scale,
                              static_cast<__half*>(adapt_f16), ld_adapt, workspace, workspace_bytes,
                              static_cast<cudaStream_t>(stream));
}
