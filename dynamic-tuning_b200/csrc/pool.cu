// Single-query cross attention of the video pooling head (sm_100a, HBM-bound): one learned query
// per clip attends over all t*N tokens of the clip, per head.
//   scores_j = f16(<q_h, k_j,h>)   (q already scaled by head_dim^-0.5 and rounded to fp16)
//   p        = softmax(scores) in fp32, rounded to fp16 for the value product
//   out_h    = f16(sum_j p_j v_j,h)
// which are the rounding points of CrossAttention.forward under fp16 autocast (reference
// video_models/video_vision_transformer_IN21K.py:92-110: fp16 `q @ k^T`, fp32 softmax, fp16
// `attn @ v`).  One CTA per (clip, head); K and V are read exactly once, 128-bit coalesced.
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "rowwise.cuh"

namespace dyt {

constexpr int QA_THREADS = 256;

__global__ void __launch_bounds__(QA_THREADS)
query_attn_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k,
                  const __half* __restrict__ v, int ld_kv, int n_keys, int H,
                  __half* __restrict__ out, int ldo) {
  extern __shared__ float s_scores[];  // [n_keys]
  __shared__ float s_red[QA_THREADS / 32];
  __shared__ float s_out[QA_THREADS / 32][64];
  const int b = blockIdx.x / H;
  const int h = blockIdx.x - b * H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __half* kb = k + static_cast<size_t>(b) * n_keys * ld_kv + h * 64;
  const __half* vb = v + static_cast<size_t>(b) * n_keys * ld_kv + h * 64;

  // query of this head in registers (fp32)
  float qf[64];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(q + static_cast<size_t>(b) * ldq + h * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = q4[i];
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hp[j]);
        qf[i * 8 + j * 2] = f.x;
        qf[i * 8 + j * 2 + 1] = f.y;
      }
    }
  }
  // ---- scores ----
  float mx = -INFINITY;
  for (int j = tid; j < n_keys; j += QA_THREADS) {
    const uint4* k4 = reinterpret_cast<const uint4*>(kb + static_cast<size_t>(j) * ld_kv);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = k4[i];
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float2 f = __half22float2(hp[jj]);
        acc = fmaf(qf[i * 8 + jj * 2], f.x, acc);
        acc = fmaf(qf[i * 8 + jj * 2 + 1], f.y, acc);
      }
    }
    const float s = __half2float(__float2half_rn(acc));
    s_scores[j] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < QA_THREADS / 32; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  // ---- softmax (fp32) ----
  float sum = 0.f;
  for (int j = tid; j < n_keys; j += QA_THREADS) {
    const float e = __expf(s_scores[j] - mx);
    s_scores[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < QA_THREADS / 32; ++w) sum += s_red[w];
  const float inv = 1.0f / sum;
  // ---- out = P V : warp w takes keys j = w (mod 8), lane = two head dims ----
  float2 acc = make_float2(0.f, 0.f);
  for (int j = warp; j < n_keys; j += QA_THREADS / 32) {
    const float pj = __half2float(__float2half_rn(s_scores[j] * inv));
    const __half2 hv = *reinterpret_cast<const __half2*>(vb + static_cast<size_t>(j) * ld_kv + lane * 2);
    const float2 f = __half22float2(hv);
    acc.x = fmaf(pj, f.x, acc.x);
    acc.y = fmaf(pj, f.y, acc.y);
  }
  s_out[warp][lane * 2] = acc.x;
  s_out[warp][lane * 2 + 1] = acc.y;
  __syncthreads();
  if (tid < 64) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < QA_THREADS / 32; ++w) o += s_out[w][tid];
    out[static_cast<size_t>(b) * ldo + h * 64 + tid] = __float2half_rn(o);
  }
}

}  // namespace dyt

extern "C" int dyt_query_attn_fwd(const void* q_f16, int ldq, const void* k_f16, const void* v_f16,
                                  int ld_kv, int num_clips, int n_keys, int num_heads, int head_dim,
                                  void* out_f16, int ldo, void* stream) {
  using namespace dyt;
  DYT_CHECK_ARG(q_f16 && k_f16 && v_f16 && out_f16, "query_attn: null buffer");
  DYT_CHECK_ARG(head_dim == 64, "query_attn: only head_dim 64 is implemented (got %d)", head_dim);
  DYT_CHECK_ARG(num_clips >= 0 && n_keys >= 1 && num_heads >= 1, "query_attn: bad sizes");
  DYT_CHECK_ARG(ld_kv >= num_heads * 64 && ld_kv % 8 == 0 && (ldq == 0 || ldq % 8 == 0) &&
                    ldo >= num_heads * 64,
                "query_attn: bad leading dimensions");
  DYT_CHECK_ARG(((reinterpret_cast<uintptr_t>(q_f16) | reinterpret_cast<uintptr_t>(k_f16) |
                  reinterpret_cast<uintptr_t>(v_f16)) & 15) == 0,
                "query_attn: operands must be 16-byte aligned");
  if (n_keys > 12000)
    return fail(DYT_EUNSUPPORTED, "query_attn: more than 12000 keys per clip not implemented (%d)", n_keys);
  if (num_clips == 0) return DYT_OK;
  const size_t smem = static_cast<size_t>(n_keys) * sizeof(float);
  query_attn_kernel<<<num_clips * num_heads, QA_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(q_f16), ldq, static_cast<const __half*>(k_f16),
      static_cast<const __half*>(v_f16), ld_kv, n_keys, num_heads, static_cast<__half*>(out_f16), ldo);
  return cuda_status(cudaGetLastError(), "query_attn_kernel launch");
}
