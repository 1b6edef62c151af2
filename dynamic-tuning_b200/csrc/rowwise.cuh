// Warp-per-row helpers for the HBM-bound row kernels (LayerNorm, dispatcher, scatter-merge).
// A row of C fp32 values (C = 128 * NV) lives in registers: lane l holds float4 chunks
// i = 0..NV-1 covering columns (i*32 + l)*4 .. +3, so every warp-wide access is one contiguous
// 512-byte (fp32) or 256-byte (fp16) segment: 128-bit coalesced loads/stores.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dyt {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ row, int lane,
                                             float4 (&v)[NV]) {
  const float4* p = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = p[i * 32 + lane];
}

// LayerNorm of a register-resident row; biased variance, two-pass (mean, then squared deviations)
template <int NV>
__device__ __forceinline__ void row_layernorm(float4 (&v)[NV], const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps, int lane) {
  constexpr float inv_c = 1.0f / static_cast<float>(NV * 128);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * inv_c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = g4[i * 32 + lane];
    const float4 b = b4[i * 32 + lane];
    v[i].x = (v[i].x - mean) * rstd * g.x + b.x;
    v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
    v[i].z = (v[i].z - mean) * rstd * g.z + b.z;
    v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
  }
}

template <int NV>
__device__ __forceinline__ void store_row_f16(__half* __restrict__ row, int lane,
                                              const float4 (&v)[NV]) {
  uint2* p = reinterpret_cast<uint2*>(row);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    __half2 lo = __floats2half2_rn(v[i].x, v[i].y);
    __half2 hi = __floats2half2_rn(v[i].z, v[i].w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    p[i * 32 + lane] = u;
  }
}

}  // namespace dyt
