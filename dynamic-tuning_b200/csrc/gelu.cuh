// GELU (exact-erf form) and its derivative for fp16 data, evaluated with one polynomial + MUFU.EX2
// (timm Mlp act_layer = nn.GELU, reference models/vision_transformer_IN21K.py:261).
#pragma once
#include "ptx.cuh"

namespace dyt {

// Gaussian upper tail Q(a) = erfc(a / sqrt 2) / 2 = 2^P(a) for a = min(|x|, 5.75): P is a degree-7
// polynomial fitted to log2 Q on [0, 5.75] (relative error of Q <= 3.3e-6; beyond 5.75 the tail is
// below half the smallest fp16 subnormal once multiplied by |x|).
__device__ __forceinline__ float gauss_tail_q(float a) {
  constexpr float kC[8] = {
      -9.999953348e-01f,
      -1.151250054e+00f,
      -4.584681700e-01f,
      -5.395489181e-02f,
      8.504621978e-03f,
      -9.294938285e-04f,
      6.144065649e-05f,
      -1.825521560e-06f};
  float pl = kC[7];
#pragma unroll
  for (int i = 6; i >= 0; --i) pl = fmaf(pl, a, kC[i]);
  return ex2_approx(pl);
}

// gelu(x) = max(x, 0) - |x| * Q(|x|).  Over all 63488 finite fp16 inputs the fp16-rounded result
// equals that of an fp32 erf evaluation except for 162 inputs that land on the neighbouring fp16
// value (the same order as fp32 erf itself against float64).  9 FMA-pipe instructions + 1 MUFU per
// element: the fc1 epilogue is bound by instruction issue, so this is what sets that GEMM's speed.
// NaN propagates (max.NaN), +-inf give +inf / -0.
__device__ __forceinline__ float gelu_f16(float x) {
  const float a = fminf(fabsf(x), 5.75f);
  const float q = gauss_tail_q(a);
  float relu;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(relu) : "f"(x));
  return fmaf(-a, q, relu);
}

// Two GELUs at once on the packed fp32 FMA (fma.rn.f32x2 -> FFMA2): the polynomial of both elements
// takes 7 issue slots instead of 14.  Same arithmetic as gelu_f16 element by element (the
// polynomial is evaluated in b = -a with the odd coefficients negated, which is exact: only signs
// change), so the results are bit-identical.
__device__ __forceinline__ float2 gelu_f16_x2(float x0, float x1) {
  // b = -min(|x|, 5.75)
  const float b0 = fmaxf(-fabsf(x0), -5.75f), b1 = fmaxf(-fabsf(x1), -5.75f);
  const unsigned long long b2 = f32x2_pack(b0, b1);
  unsigned long long pl = f32x2_pack(1.825521560e-06f, 1.825521560e-06f);            // -c7
  pl = f32x2_fma(pl, b2, f32x2_pack(6.144065649e-05f, 6.144065649e-05f));            //  c6
  pl = f32x2_fma(pl, b2, f32x2_pack(9.294938285e-04f, 9.294938285e-04f));            // -c5
  pl = f32x2_fma(pl, b2, f32x2_pack(8.504621978e-03f, 8.504621978e-03f));            //  c4
  pl = f32x2_fma(pl, b2, f32x2_pack(5.395489181e-02f, 5.395489181e-02f));            // -c3
  pl = f32x2_fma(pl, b2, f32x2_pack(-4.584681700e-01f, -4.584681700e-01f));          //  c2
  pl = f32x2_fma(pl, b2, f32x2_pack(1.151250054e+00f, 1.151250054e+00f));            // -c1
  pl = f32x2_fma(pl, b2, f32x2_pack(-9.999953348e-01f, -9.999953348e-01f));          //  c0
  float p0, p1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(pl));
  const unsigned long long q2 = f32x2_pack(ex2_approx(p0), ex2_approx(p1));
  float r0, r1;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r0) : "f"(x0));
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r1) : "f"(x1));
  const unsigned long long g = f32x2_fma(b2, q2, f32x2_pack(r0, r1));   // relu - a q
  float2 out;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(out.x), "=f"(out.y) : "l"(g));
  return out;
}

// gelu'(x) = Phi(x) + x * phi(x):  Phi(x) = 1 - Q(|x|) (x >= 0) or Q(|x|);  phi(x) = 2^(-x^2 log2(e)/2
// - log2 sqrt(2 pi)).  Absolute error <= 4e-6: far below the fp16 rounding of the gradient.
__device__ __forceinline__ float gelu_grad_f16(float x) {
  const float ax = fabsf(x);
  const float q = gauss_tail_q(fminf(ax, 5.75f));
  const float cdf = x >= 0.f ? 1.f - q : q;
  const float pdf = ex2_approx(fmaf(-0.72134752044f * x, x, -1.32574806473f));
  return fmaf(x, pdf, cdf);
}

}  // namespace dyt
