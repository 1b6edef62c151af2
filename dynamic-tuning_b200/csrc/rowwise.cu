// HBM-bound row kernels: LayerNorm -> fp16 (optionally gathering rows by index) and the vectorised
// scatter-merge that rebuilds the [B, N, C] residual stream from the adapter output and the packed
// MLP output.  One warp per row, 128-bit coalesced accesses (see rowwise.cuh).
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "ptx.cuh"
#include "rowwise.cuh"

namespace dyt {

// ---------------------------------------------------------------------------------------------
// LayerNorm (fp32 in, fp16 out): reference nn.LayerNorm(eps=1e-6) norm1 / norm2
// (models/vision_transformer_IN21K.py:110, :123, :262); autocast keeps the statistics in fp32 and
// the consumer Linear casts the result to fp16, which is the single rounding done here.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_f16_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ row_idx,
                     const int* __restrict__ n_rows_dev, int n_rows, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, __half* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  int rows = n_rows;
  if (n_rows_dev != nullptr) {
    const int nd = *n_rows_dev;
    rows = nd < rows ? nd : rows;
  }
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += gridDim.x * warps_per_block) {
    const int src = row_idx != nullptr ? row_idx[r] : r;
    float4 v[NV];
    load_row_f32<NV>(x + static_cast<size_t>(src) * ldx, lane, v);
    row_layernorm<NV>(v, gamma, beta, eps, lane);
    store_row_f16<NV>(out + static_cast<size_t>(r) * ldo, lane, v);
  }
}

int layernorm_f16(const float* x, int ldx, const int* row_idx, const int* n_rows_dev, int n_rows,
                  int C, const float* gamma, const float* beta, float eps, __half* out, int ldo,
                  cudaStream_t stream) {
  DYT_CHECK_ARG(x && gamma && beta && out, "layernorm: null buffer");
  DYT_CHECK_ARG(n_rows >= 0 && ldx >= C && ldo >= C && ldx % 4 == 0 && ldo % 4 == 0,
                "layernorm: bad sizes");
  if (n_rows == 0) return DYT_OK;
  int grid = (n_rows + 7) / 8;
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  cudaError_t err = cudaSuccess;
  switch (C) {
    case 768:
      err = launch_pdl(layernorm_f16_kernel<6>, dim3(grid), dim3(256), 0, stream, x, ldx, row_idx,
                       n_rows_dev, n_rows, gamma, beta, eps, out, ldo);
      break;
    case 1024:
      err = launch_pdl(layernorm_f16_kernel<8>, dim3(grid), dim3(256), 0, stream, x, ldx, row_idx,
                       n_rows_dev, n_rows, gamma, beta, eps, out, ldo);
      break;
    case 384:
      err = launch_pdl(layernorm_f16_kernel<3>, dim3(grid), dim3(256), 0, stream, x, ldx, row_idx,
                       n_rows_dev, n_rows, gamma, beta, eps, out, ldo);
      break;
    case 128:
      err = launch_pdl(layernorm_f16_kernel<1>, dim3(grid), dim3(256), 0, stream, x, ldx, row_idx,
                       n_rows_dev, n_rows, gamma, beta, eps, out, ldo);
      break;
    default:
      return fail(DYT_EUNSUPPORTED, "layernorm: embed dim %d not instantiated (128/384/768/1024)", C);
  }
  return cuda_status(err, "layernorm_f16_kernel launch");
}

// ---------------------------------------------------------------------------------------------
// Video pooling head input: k_in = f16(LN_k(LN(x))), v_in = f16(LN_v(LN(x))) for every token
// (reference video_models/video_vision_transformer_IN21K.py:474 `self.norm`, :44-45 `norm_k` /
// `norm_v` of AttentiveBlock; all three LayerNorms run in fp32 under autocast, the consumer Linear
// casts to fp16).  One pass over the fp32 stream instead of three.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
double_layernorm_f16_kernel(const float* __restrict__ x, int ldx, int n_rows,
                            const float* __restrict__ g0, const float* __restrict__ b0,
                            const float* __restrict__ gk, const float* __restrict__ bk,
                            const float* __restrict__ gv, const float* __restrict__ bv, float eps,
                            __half* __restrict__ out_k, __half* __restrict__ out_v, int ldo) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rows;
       r += gridDim.x * warps_per_block) {
    float4 v[NV], k[NV];
    load_row_f32<NV>(x + static_cast<size_t>(r) * ldx, lane, v);
    row_layernorm<NV>(v, g0, b0, eps, lane);
#pragma unroll
    for (int i = 0; i < NV; ++i) k[i] = v[i];
    row_layernorm<NV>(k, gk, bk, eps, lane);
    store_row_f16<NV>(out_k + static_cast<size_t>(r) * ldo, lane, k);
    row_layernorm<NV>(v, gv, bv, eps, lane);
    store_row_f16<NV>(out_v + static_cast<size_t>(r) * ldo, lane, v);
  }
}

int double_layernorm_f16(const float* x, int ldx, int n_rows, int C, const float* g0,
                         const float* b0, const float* gk, const float* bk, const float* gv,
                         const float* bv, float eps, __half* out_k, __half* out_v, int ldo,
                         cudaStream_t stream) {
  DYT_CHECK_ARG(x && g0 && b0 && gk && bk && gv && bv && out_k && out_v, "double_layernorm: null buffer");
  DYT_CHECK_ARG(n_rows >= 0 && ldx >= C && ldo >= C && ldx % 4 == 0 && ldo % 4 == 0,
                "double_layernorm: bad sizes");
  if (n_rows == 0) return DYT_OK;
  int grid = (n_rows + 7) / 8;
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
#define DYT_LAUNCH_DLN(NV)                                                                          \
  double_layernorm_f16_kernel<NV><<<grid, 256, 0, stream>>>(x, ldx, n_rows, g0, b0, gk, bk, gv, bv, \
                                                            eps, out_k, out_v, ldo)
  switch (C) {
    case 768: DYT_LAUNCH_DLN(6); break;
    case 1024: DYT_LAUNCH_DLN(8); break;
    case 384: DYT_LAUNCH_DLN(3); break;
    case 128: DYT_LAUNCH_DLN(1); break;
    default:
      return fail(DYT_EUNSUPPORTED, "double_layernorm: embed dim %d not instantiated", C);
  }
#undef DYT_LAUNCH_DLN
  return cuda_status(cudaGetLastError(), "double_layernorm_f16_kernel launch");
}

// ---------------------------------------------------------------------------------------------
// scatter-merge: out[t] = adapt[t] + (x1[t] + (kept(t) ? mlp_packed[pos[t]] : 0))
// Replaces zeros() + index_put + two adds (reference models/model_speed_test.py:302-308).
// Optionally also emits LayerNorm(out) in fp16 with the NEXT block's norm1 (or the final norm), so
// the following kernel does not have to re-read the fp32 stream.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
scatter_merge_kernel(const float* __restrict__ x1, int ldx, const __half* __restrict__ adapt, int lda,
                     const __half* __restrict__ mlp_packed, int ldm, const int* __restrict__ token_pos,
                     int n_rows, float* __restrict__ out, int ldo, const float* __restrict__ nln_w,
                     const float* __restrict__ nln_b, float eps, __half* __restrict__ nln_out,
                     int ldn) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rows;
       r += gridDim.x * warps_per_block) {
    float4 v[NV];
    load_row_f32<NV>(x1 + static_cast<size_t>(r) * ldx, lane, v);
    const int pos = token_pos[r];
    if (pos >= 0) {
      const uint2* m = reinterpret_cast<const uint2*>(mlp_packed + static_cast<size_t>(pos) * ldm);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint2 u = m[i * 32 + lane];
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        v[i].x += a.x; v[i].y += a.y; v[i].z += b.x; v[i].w += b.y;
      }
    }
    {
      const uint2* ad = reinterpret_cast<const uint2*>(adapt + static_cast<size_t>(r) * lda);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint2 u = ad[i * 32 + lane];
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        v[i].x += a.x; v[i].y += a.y; v[i].z += b.x; v[i].w += b.y;
      }
    }
    float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(r) * ldo);
#pragma unroll
    for (int i = 0; i < NV; ++i) o[i * 32 + lane] = v[i];
    if (nln_out != nullptr) {
      row_layernorm<NV>(v, nln_w, nln_b, eps, lane);
      store_row_f16<NV>(nln_out + static_cast<size_t>(r) * ldn, lane, v);
    }
  }
}

int scatter_merge(const float* x1, int ldx, const __half* adapt, int lda, const __half* mlp_packed,
                  int ldm, const int* token_pos, int n_rows, int C, float* out, int ldo,
                  const float* nln_w, const float* nln_b, float eps, __half* nln_out, int ldn,
                  cudaStream_t stream) {
  DYT_CHECK_ARG(x1 && adapt && mlp_packed && token_pos && out, "scatter_merge: null buffer");
  DYT_CHECK_ARG(ldx % 4 == 0 && lda % 4 == 0 && ldm % 4 == 0 && ldo % 4 == 0, "scatter_merge: strides");
  DYT_CHECK_ARG(nln_out == nullptr || (nln_w && nln_b && ldn % 4 == 0), "scatter_merge: next-LN args");
  if (n_rows == 0) return DYT_OK;
  int grid = (n_rows + 7) / 8;
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  cudaError_t err = cudaSuccess;
#define DYT_LAUNCH_MERGE(NV)                                                                       \
  err = launch_pdl(scatter_merge_kernel<NV>, dim3(grid), dim3(256), 0, stream, x1, ldx, adapt, lda, \
                   mlp_packed, ldm, token_pos, n_rows, out, ldo, nln_w, nln_b, eps, nln_out, ldn)
  switch (C) {
    case 768: DYT_LAUNCH_MERGE(6); break;
    case 1024: DYT_LAUNCH_MERGE(8); break;
    case 384: DYT_LAUNCH_MERGE(3); break;
    case 128: DYT_LAUNCH_MERGE(1); break;
    default:
      return fail(DYT_EUNSUPPORTED, "scatter_merge: embed dim %d not instantiated", C);
  }
#undef DYT_LAUNCH_MERGE
  return cuda_status(err, "scatter_merge_kernel launch");
}

}  // namespace dyt

extern "C" int dyt_layernorm_f16(const float* x, int ldx, const int* row_idx, const int* n_rows_dev,
                                 int n_rows, int C, const float* gamma, const float* beta, float eps,
                                 void* out_f16, int ldo, void* stream) {
  return dyt::layernorm_f16(x, ldx, row_idx, n_rows_dev, n_rows, C, gamma, beta, eps,
                            static_cast<__half*>(out_f16), ldo, static_cast<cudaStream_t>(stream));
}

extern "C" int dyt_pool_layernorm_f16(const float* x, int ldx, int n_rows, int C, const float* g0,
                                      const float* b0, const float* gk, const float* bk,
                                      const float* gv, const float* bv, float eps, void* out_k_f16,
                                      void* out_v_f16, int ldo, void* stream) {
  return dyt::double_layernorm_f16(x, ldx, n_rows, C, g0, b0, gk, bk, gv, bv, eps,
                                   static_cast<__half*>(out_k_f16), static_cast<__half*>(out_v_f16),
                                   ldo, static_cast<cudaStream_t>(stream));
}

extern "C" int dyt_scatter_merge_fwd(const float* x1, int ldx, const void* adapt_f16, int lda,
                                     const void* mlp_packed_f16, int ldm, const int* token_pos,
                                     int n_rows, int C, float* out, int ldo, const float* next_ln_w,
                                     const float* next_ln_b, float eps, void* next_ln_out_f16,
                                     int ldn, void* stream) {
  return dyt::scatter_merge(x1, ldx, static_cast<const __half*>(adapt_f16), lda,
                            static_cast<const __half*>(mlp_packed_f16), ldm, token_pos, n_rows, C,
                            out, ldo, next_ln_w, next_ln_b, eps,
                            static_cast<__half*>(next_ln_out_f16), ldn,
                            static_cast<cudaStream_t>(stream));
}
