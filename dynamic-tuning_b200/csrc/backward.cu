// Backward kernels of the DyT block for parameter-efficient fine-tuning (SURVEY.md section 8f
// rank 1): the backbone is frozen, so the backward is data gradients through every frozen
// Linear / LayerNorm / GELU (dgrad GEMMs reuse gemm_tn with the transposed weight), weight
// gradients only for the adapter, the selector and the head, and the straight-through gradient of
// the hard gate (reference models/dynamic_adapter.py:46-51).
//
//   layernorm_bwd      g_x = resid + dLN(g_y)            (+ row_scale[r] * axpy[:], + fp16 copy)
//   merge_bwd          g16 = f16(g_out); gm16 = mask * g16; g_logit = <g16, mlp_x> * dgate/dlogit
//   rowscale_colsum    d_w[:] += sum_t s[t] * X[t,:], d_b += sum_t s[t]     (selector wgrad)
//   gelu fwd / bwd, relu-dropout bwd, dropout fwd       (elementwise, 128-bit accesses)
//   wgrad              dW[Nout,Kin] += alpha * G[T,Nout]^T X[T,Kin], db += alpha * colsum(G)
//                      (HMMA through nvcuda::wmma on zero-padded shared-memory tiles)
#include <mma.h>
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "gelu.cuh"
#include "host_utils.h"
#include "rowwise.cuh"

namespace dyt {

// ---------------------------------------------------------------------------------------------
// LayerNorm backward w.r.t. the input (gamma / beta are frozen: norm1 / norm2 / norm are backbone
// parameters, main_image.py:242-256).  y = xhat * gamma + beta, xhat = (x - mean) * rstd:
//   gh = g_y * gamma;  g_x = rstd * (gh - mean(gh) - xhat * mean(gh * xhat)).
// Statistics are recomputed from x (fp32) instead of being saved by the forward.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __half* __restrict__ gy, int ldg, const float* __restrict__ x, int ldx,
                     const int* __restrict__ row_idx, const int* __restrict__ n_rows_dev, int n_rows,
                     const float* __restrict__ gamma, float eps, const float* resid, int ldr,
                     const float* __restrict__ row_scale, const float* __restrict__ axpy, float* out,
                     int ldo, __half* __restrict__ out_h, int ldoh) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  constexpr float inv_c = 1.0f / static_cast<float>(NV * 128);
  if (n_rows_dev != nullptr) {
    const int nd = *n_rows_dev;
    n_rows = nd < n_rows ? nd : n_rows;
  }
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rows;
       r += gridDim.x * warps_per_block) {
    const size_t src = row_idx != nullptr ? static_cast<size_t>(row_idx[r]) : static_cast<size_t>(r);
    float4 v[NV];
    load_row_f32<NV>(x + src * ldx, lane, v);
    // every load of the row is issued before the first reduction (the kernel is HBM-bound and its
    // four warp reductions per row are a serial chain: three load phases in a row ran at 0.6 of the
    // copy bandwidth)
    uint2 gu[NV];
    float4 rr4[NV];
    {
      const uint2* gp = reinterpret_cast<const uint2*>(gy + static_cast<size_t>(r) * ldg);
#pragma unroll
      for (int i = 0; i < NV; ++i) gu[i] = gp[i * 32 + lane];
      if (resid != nullptr) {
        const float4* r4 = reinterpret_cast<const float4*>(resid + src * ldr);
#pragma unroll
        for (int i = 0; i < NV; ++i) rr4[i] = r4[i * 32 + lane];
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    float4 g[NV];
    const float4* gm4 = reinterpret_cast<const float4*>(gamma);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint2 u = gu[i];
      const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      const float4 gm = gm4[i * 32 + lane];
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;      // xhat
      g[i] = make_float4(lo.x * gm.x, lo.y * gm.y, hi.x * gm.z, hi.y * gm.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
    const float rs = row_scale != nullptr ? row_scale[r] : 0.f;
    const float4* a4 = reinterpret_cast<const float4*>(axpy);
    float4* o4 = reinterpret_cast<float4*>(out + src * ldo);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rstd * (g[i].x - s1 - v[i].x * s2);
      o.y = rstd * (g[i].y - s1 - v[i].y * s2);
      o.z = rstd * (g[i].z - s1 - v[i].z * s2);
      o.w = rstd * (g[i].w - s1 - v[i].w * s2);
      if (resid != nullptr) {
        const float4 rr = rr4[i];
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
      if (row_scale != nullptr) {
        const float4 a = a4[i * 32 + lane];
        o.x += rs * a.x; o.y += rs * a.y; o.z += rs * a.z; o.w += rs * a.w;
      }
      o4[i * 32 + lane] = o;
      g[i] = o;
    }
    if (out_h != nullptr) store_row_f16<NV>(out_h + src * ldoh, lane, g);
  }
}

template <int NV>
static void launch_ln_bwd(int grid, cudaStream_t st, const __half* gy, int ldg, const float* x,
                          int ldx, const int* row_idx, const int* n_rows_dev, int n_rows,
                          const float* gamma, float eps, const float* resid, int ldr,
                          const float* row_scale, const float* axpy, float* out, int ldo,
                          __half* out_h, int ldoh) {
  layernorm_bwd_kernel<NV><<<grid, 256, 0, st>>>(gy, ldg, x, ldx, row_idx, n_rows_dev, n_rows, gamma,
                                                 eps, resid, ldr, row_scale, axpy, out, ldo, out_h,
                                                 ldoh);
}

// ---------------------------------------------------------------------------------------------
// Backward of  out = (x1 + mask * mlp_x) + adapt  (reference vision_transformer_IN21K.py:161-163)
// and of the straight-through gate  ret = y_hard - y_soft.detach() + y_soft
// (models/dynamic_adapter.py:46-51): d ret / d logit = y(1-y)/tau (train) or y(1-y) (eval form).
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
merge_bwd_kernel(const float* __restrict__ g_out, int ldg, const __half* __restrict__ mlp_x, int ldm,
                 const float* __restrict__ mask, const float* __restrict__ logits,
                 const float* __restrict__ n1, const float* __restrict__ n2, float tau,
                 const float* __restrict__ g_sel, const float* __restrict__ g_logits_ext, int T,
                 int N, __half* __restrict__ g16, int ld16, __half* __restrict__ gm16, int ldgm,
                 float* __restrict__ g_logit, const int* __restrict__ token_pos,
                 const float* __restrict__ sel_w, float* __restrict__ gx_init, int ldgx) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < T;
       t += gridDim.x * warps_per_block) {
    float4 v[NV];
    load_row_f32<NV>(g_out + static_cast<size_t>(t) * ldg, lane, v);
    float4 raw[NV];
    if (gx_init != nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) raw[i] = v[i];
    }
    // the gradient reaching the fp16 tensors (mlp_x, adapt) is g_out rounded to fp16
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x = __half2float(__float2half_rn(v[i].x));
      v[i].y = __half2float(__float2half_rn(v[i].y));
      v[i].z = __half2float(__float2half_rn(v[i].z));
      v[i].w = __half2float(__float2half_rn(v[i].w));
    }
    store_row_f16<NV>(g16 + static_cast<size_t>(t) * ld16, lane, v);
    if (gm16 == nullptr) continue;  // complete_model: no mask on the path
    const float m = mask[t];
    const uint2* mp = reinterpret_cast<const uint2*>(mlp_x + static_cast<size_t>(t) * ldm);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint2 u = mp[i * 32 + lane];
      const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      dot += (v[i].x * lo.x + v[i].y * lo.y) + (v[i].z * hi.x + v[i].w * hi.y);
      v[i].x *= m; v[i].y *= m; v[i].z *= m; v[i].w *= m;
    }
    dot = warp_sum(dot);
    if (token_pos == nullptr) {
      store_row_f16<NV>(gm16 + static_cast<size_t>(t) * ldgm, lane, v);
    } else {  // packed: kept rows only, at their position in the compacted order
      const int pos = token_pos[t];
      if (pos >= 0) store_row_f16<NV>(gm16 + static_cast<size_t>(pos) * ldgm, lane, v);
    }
    float gl_row = 0.f;
    if (lane == 0) {
      const int n = t % N;
      float gl = 0.f;
      if (n != 0) {  // the cls slot is a constant one
        const size_t li = static_cast<size_t>(t / N) * (N - 1) + (n - 1);
        const float l = logits[li];
        float z = l, dz = 1.f;
        if (n1 != nullptr) {
          z = ((l + n1[li]) - n2[li]) / tau;
          dz = 1.f / tau;
        }
        const float y = 1.f / (1.f + __expf(-z));
        const float up = dot + (g_sel != nullptr ? g_sel[t] : 0.f);
        gl = up * y * (1.f - y) * dz;
        if (g_logits_ext != nullptr) gl += g_logits_ext[li];
      }
      g_logit[t] = gl;
      gl_row = gl;
    }
    if (gx_init != nullptr) {
      // start of the x1 gradient: residual path + the selector's data gradient g_logit * w
      gl_row = __shfl_sync(0xffffffffu, gl_row, 0);
      const float4* w4 = reinterpret_cast<const float4*>(sel_w);
      float4* o4 = reinterpret_cast<float4*>(gx_init + static_cast<size_t>(t) * ldgx);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 w = w4[i * 32 + lane];
        o4[i * 32 + lane] = make_float4(raw[i].x + gl_row * w.x, raw[i].y + gl_row * w.y,
                                        raw[i].z + gl_row * w.z, raw[i].w + gl_row * w.w);
      }
    }
  }
}

// out[r, :] = a[r, :] * gelu'(pre[row_idx[r], :]) for the first *n_rows_dev packed rows
__global__ void __launch_bounds__(256)
gelu_bwd_rows_kernel(const __half* __restrict__ a, int lda, const __half* __restrict__ pre, int ldp,
                     const int* __restrict__ row_idx, const int* __restrict__ n_rows_dev, int n_rows,
                     int H, __half* __restrict__ out, int ldo) {
  if (n_rows_dev != nullptr) {
    const int nd = *n_rows_dev;
    n_rows = nd < n_rows ? nd : n_rows;
  }
  const int groups = H >> 3;
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const uint4* pa = reinterpret_cast<const uint4*>(a + static_cast<size_t>(r) * lda);
    const uint4* pp = reinterpret_cast<const uint4*>(pre + static_cast<size_t>(row_idx[r]) * ldp);
    uint4* po = reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * ldo);
    for (int gi = threadIdx.x; gi < groups; gi += blockDim.x) {
      const uint4 ua = pa[gi], up = pp[gi];
      const __half2* ha = reinterpret_cast<const __half2*>(&ua);
      const __half2* hp = reinterpret_cast<const __half2*>(&up);
      uint4 uo;
      __half2* ho = reinterpret_cast<__half2*>(&uo);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = __half22float2(ha[j]);
        const float2 fp = __half22float2(hp[j]);
        ho[j] = __floats2half2_rn(fa.x * gelu_grad_f16(fp.x), fa.y * gelu_grad_f16(fp.y));
      }
      po[gi] = uo;
    }
  }
}

// d_w[c] += sum_t s[t] * X[t, c];  d_b += sum_t s[t]   (selector mlp_head weight / bias gradient)
__global__ void __launch_bounds__(256)
rowscale_colsum_kernel(const float* __restrict__ s, const __half* __restrict__ X, int ldx, int T,
                       int C, int rows_per_cta, float* __restrict__ out_w, float* __restrict__ out_b) {
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(T, t0 + rows_per_cta);
  const int pairs = C >> 1;
  float2 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
  float sb = 0.f;
  for (int t = t0; t < t1; ++t) {
    const float st = s[t];
    if (st == 0.f) continue;
    sb += st;
    const __half2* row = reinterpret_cast<const __half2*>(X + static_cast<size_t>(t) * ldx);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int p = threadIdx.x + k * 256;
      if (p < pairs) {
        const float2 xv = __half22float2(row[p]);
        acc[k].x += st * xv.x;
        acc[k].y += st * xv.y;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = threadIdx.x + k * 256;
    if (p < pairs && (acc[k].x != 0.f || acc[k].y != 0.f)) {
      atomicAdd(out_w + 2 * p, acc[k].x);
      atomicAdd(out_w + 2 * p + 1, acc[k].y);
    }
  }
  if (threadIdx.x == 0 && out_b != nullptr && sb != 0.f) atomicAdd(out_b, sb);
}

// Vectorised form (C % 8 == 0, 16-byte aligned rows): a thread owns 8 columns, C / 8 threads cover a
// row, the CTA's 384 threads form 384 / (C / 8) row groups that stride over the CTA's rows four
// rows at a time (four 16-byte loads in flight per thread); groups are summed through shared
// memory, one atomicAdd per column and CTA.  (The scalar kernel above walks its rows one after the
// other with 4-byte loads: 30 us for 12.6k x 768, latency-bound; this one is HBM-bound.)
constexpr int RC_THREADS = 384;
__global__ void __launch_bounds__(RC_THREADS)
rowscale_colsum_vec_kernel(const float* __restrict__ s, const __half* __restrict__ X, int ldx, int T,
                           int C, int rows_per_cta, float* __restrict__ out_w,
                           float* __restrict__ out_b) {
  extern __shared__ float rc_smem[];   // [groups][C] + [groups]
  const int tpr = C >> 3;              // threads per row
  const int groups = RC_THREADS / tpr;
  const int grp = threadIdx.x / tpr;
  const int c8 = (threadIdx.x - grp * tpr) * 8;
  const int t0 = blockIdx.x * rows_per_cta;
  const int t1 = min(T, t0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  float sb = 0.f;
  if (grp < groups) {
    for (int t = t0 + grp; t < t1; t += 4 * groups) {
      float st[4];
      uint4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tt = t + u * groups;
        st[u] = tt < t1 ? s[tt] : 0.f;
        xv[u] = make_uint4(0u, 0u, 0u, 0u);
        if (tt < t1 && st[u] != 0.f)
          xv[u] = *reinterpret_cast<const uint4*>(X + static_cast<size_t>(tt) * ldx + c8);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sb += st[u];
        const __half2* h = reinterpret_cast<const __half2*>(&xv[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(h[k]);
          acc[2 * k] = fmaf(st[u], f.x, acc[2 * k]);
          acc[2 * k + 1] = fmaf(st[u], f.y, acc[2 * k + 1]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) rc_smem[grp * C + c8 + k] = acc[k];
    if (c8 == 0) rc_smem[groups * C + grp] = sb;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += RC_THREADS) {
    float v = 0.f;
    for (int g = 0; g < groups; ++g) v += rc_smem[g * C + c];
    if (v != 0.f) atomicAdd(out_w + c, v);
  }
  if (threadIdx.x == 0 && out_b != nullptr) {
    float v = 0.f;
    for (int g = 0; g < groups; ++g) v += rc_smem[groups * C + g];
    if (v != 0.f) atomicAdd(out_b, v);
  }
}

// ---------------------------------------------------------------------------------------------
// Elementwise fp16 kernels (8 halves per thread per step)
// ---------------------------------------------------------------------------------------------
enum { EW_GELU_FWD = 0, EW_GELU_BWD = 1, EW_RELU_DROP_BWD = 2, EW_MUL = 3 };

template <int OP>
__global__ void __launch_bounds__(256)
eltwise_f16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                   const uint4* __restrict__ c, uint4* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 ua = a[i];
    uint4 ub = make_uint4(0, 0, 0, 0), uc = make_uint4(0, 0, 0, 0);
    if (OP != EW_GELU_FWD) ub = b[i];
    if (OP == EW_RELU_DROP_BWD && c != nullptr) uc = c[i];
    const __half2* pa = reinterpret_cast<const __half2*>(&ua);
    const __half2* pb = reinterpret_cast<const __half2*>(&ub);
    const __half2* pc = reinterpret_cast<const __half2*>(&uc);
    uint4 uo;
    __half2* po = reinterpret_cast<__half2*>(&uo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __half22float2(pa[j]);
      const float2 fb = __half22float2(pb[j]);
      float2 r;
      if (OP == EW_GELU_FWD) {            // a = pre-activation
        r.x = gelu_f16(fa.x);
        r.y = gelu_f16(fa.y);
      } else if (OP == EW_GELU_BWD) {     // a = g_h, b = pre-activation
        r.x = fa.x * gelu_grad_f16(fb.x);
        r.y = fa.y * gelu_grad_f16(fb.y);
      } else if (OP == EW_RELU_DROP_BWD) {  // a = g, b = relu output (after dropout), c = multiplier
        float2 fc = make_float2(1.f, 1.f);
        if (c != nullptr) fc = __half22float2(pc[j]);
        r.x = fb.x > 0.f ? fa.x * fc.x : 0.f;
        r.y = fb.y > 0.f ? fa.y * fc.y : 0.f;
      } else {                             // product
        r.x = fa.x * fb.x;
        r.y = fa.y * fb.y;
      }
      po[j] = __floats2half2_rn(r.x, r.y);
    }
    out[i] = uo;
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient  dW[Nout, Kin] += alpha * sum_t G[t, Nout]^T X[t, Kin]  (contraction over tokens)
// for the trainable Linears: adapter down / up (models/dynamic_adapter.py:127-130) and the head.
// CTA = 4 warps, 64 x 64 output tile, a chunk of tokens; tiles of 32 tokens are staged zero-padded
// in shared memory and multiplied with HMMA (wmma 16x16x16, fp32 accumulate); the partial tile is
// added to dW with 16-byte fp32 reductions (red.global.add.v4.f32; scalar atomics for unaligned dW
// rows) -- dW zero-initialised / carried by the caller.
// ---------------------------------------------------------------------------------------------
constexpr int WG_TILE = 64, WG_TOK = 32, WG_LD = 72;

// up to 8 halves starting at p (only the first `remaining` exist), missing ones read as zero
__device__ __forceinline__ uint4 load8_guarded(const __half* p, int remaining, bool vec) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (remaining >= 8 && vec) {
    v = *reinterpret_cast<const uint4*>(p);
  } else if (remaining > 0) {
    __half* h = reinterpret_cast<__half*>(&v);
    for (int i = 0; i < 8; ++i)
      if (i < remaining) h[i] = p[i];
  }
  return v;
}

__global__ void __launch_bounds__(128)
wgrad_kernel(const __half* __restrict__ G, int ldg, const __half* __restrict__ X, int ldx, int T,
             int Nout, int Kin, int tok_per_cta, float alpha, float* __restrict__ dW, int ldw,
             float* __restrict__ db) {
  using namespace nvcuda;
  __shared__ __align__(32) __half Gs[WG_TOK * WG_LD];
  __shared__ __align__(32) __half Xs[WG_TOK * WG_LD];
  __shared__ __align__(32) float Cs[4][16 * 20];
  const int n0 = blockIdx.y * WG_TILE;
  const int k0 = blockIdx.x * WG_TILE;
  const int t_begin = blockIdx.z * tok_per_cta;
  const int t_end = min(T, t_begin + tok_per_cta);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wn = (warp >> 1) * 32, wk = (warp & 1) * 32;  // warp's 32 x 32 sub-tile
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) wmma::fill_fragment(acc[i][j], 0.f);
  float bsum = 0.f;  // thread c < 64 sums column n0 + c of G (only the k-tile 0 CTAs report it)
  const bool g_vec = (ldg % 8 == 0) && ((reinterpret_cast<uintptr_t>(G) & 15) == 0);
  const bool x_vec = (ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  // each thread stages two 16-byte groups of G and of X per 32-token tile; the next tile's global
  // loads are issued before the current tile's MMAs (register double buffering)
  const int sr0 = threadIdx.x >> 3, scg = (threadIdx.x & 7) * 8;  // rows sr0 and sr0 + 16
  uint4 pg[2], px[2];
  auto fetch = [&](int t0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = t0 + sr0 + h * 16;
      pg[h] = make_uint4(0, 0, 0, 0);
      px[h] = pg[h];
      if (t < t_end) {
        pg[h] = load8_guarded(G + static_cast<size_t>(t) * ldg + n0 + scg, Nout - (n0 + scg), g_vec);
        px[h] = load8_guarded(X + static_cast<size_t>(t) * ldx + k0 + scg, Kin - (k0 + scg), x_vec);
      }
    }
  };
  fetch(t_begin);
  for (int t0 = t_begin; t0 < t_end; t0 += WG_TOK) {
    // stage G[t0..t0+32, n0..n0+64] and X[t0..t0+32, k0..k0+64] (zero outside the matrices)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<uint4*>(Gs + (sr0 + h * 16) * WG_LD + scg) = pg[h];
      *reinterpret_cast<uint4*>(Xs + (sr0 + h * 16) * WG_LD + scg) = px[h];
    }
    __syncthreads();
    if (t0 + WG_TOK < t_end) fetch(t0 + WG_TOK);
    if (db != nullptr && blockIdx.x == 0 && threadIdx.x < WG_TILE) {
#pragma unroll 8
      for (int r = 0; r < WG_TOK; ++r) bsum += __half2float(Gs[r * WG_LD + threadIdx.x]);
    }
#pragma unroll
    for (int kk = 0; kk < WG_TOK; kk += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::col_major> a[2];
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> b[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) wmma::load_matrix_sync(a[i], Gs + kk * WG_LD + wn + i * 16, WG_LD);
#pragma unroll
      for (int j = 0; j < 2; ++j) wmma::load_matrix_sync(b[j], Xs + kk * WG_LD + wk + j * 16, WG_LD);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) wmma::mma_sync(acc[i][j], a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* stg = Cs[warp];
  const int rr = lane >> 1, cb = (lane & 1) * 8;
  const bool vec_red = (ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      wmma::store_matrix_sync(stg, acc[i][j], 20, wmma::mem_row_major);
      __syncwarp();
      const int n = n0 + wn + i * 16 + rr;
      if (n < Nout) {
#pragma unroll
        for (int c4 = 0; c4 < 8; c4 += 4) {
          const int k = k0 + wk + j * 16 + cb + c4;
          const float v0 = stg[rr * 20 + cb + c4] * alpha, v1 = stg[rr * 20 + cb + c4 + 1] * alpha;
          const float v2 = stg[rr * 20 + cb + c4 + 2] * alpha, v3 = stg[rr * 20 + cb + c4 + 3] * alpha;
          float* dst = dW + static_cast<size_t>(n) * ldw + k;
          if (vec_red && k + 3 < Kin) {
            // one 16-byte reduction instead of four scalar atomics: the splits of a tile all add
            // to the same addresses, and the kernel is bound by that traffic
            if (v0 != 0.f || v1 != 0.f || v2 != 0.f || v3 != 0.f)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v0), "f"(v1),
                           "f"(v2), "f"(v3)
                           : "memory");
          } else {
            if (k < Kin && v0 != 0.f) atomicAdd(dst, v0);
            if (k + 1 < Kin && v1 != 0.f) atomicAdd(dst + 1, v1);
            if (k + 2 < Kin && v2 != 0.f) atomicAdd(dst + 2, v2);
            if (k + 3 < Kin && v3 != 0.f) atomicAdd(dst + 3, v3);
          }
        }
      }
      __syncwarp();
    }
  if (db != nullptr && blockIdx.x == 0 && threadIdx.x < WG_TILE && n0 + threadIdx.x < Nout &&
      bsum != 0.f)
    atomicAdd(db + n0 + threadIdx.x, bsum * alpha);
}

static int row_grid(int n_rows) {
  int grid = (n_rows + 7) / 8;
  const int cap = sm_count() * 16;
  return grid > cap ? cap : (grid < 1 ? 1 : grid);
}

}  // namespace dyt

using namespace dyt;

extern "C" int dyt_layernorm_bwd(const void* gy_f16, int ldg, const float* x, int ldx,
                                 const int* row_idx, const int* n_rows_dev, int n_rows, int C,
                                 const float* gamma,
                                 float eps, const float* resid, int ldr, const float* row_scale,
                                 const float* axpy, float* out, int ldo, void* out_f16, int ldoh,
                                 void* stream) {
  DYT_CHECK_ARG(gy_f16 && x && gamma && out, "layernorm_bwd: null buffer");
  DYT_CHECK_ARG(n_rows >= 0 && ldg >= C && ldx >= C && ldo >= C && ldg % 4 == 0 && ldx % 4 == 0 &&
                    ldo % 4 == 0,
                "layernorm_bwd: bad sizes");
  DYT_CHECK_ARG(resid == nullptr || (ldr >= C && ldr % 4 == 0), "layernorm_bwd: bad residual stride");
  DYT_CHECK_ARG((row_scale == nullptr) == (axpy == nullptr), "layernorm_bwd: row_scale and axpy go together");
  DYT_CHECK_ARG(out_f16 == nullptr || (ldoh >= C && ldoh % 4 == 0), "layernorm_bwd: bad fp16 stride");
  if (n_rows == 0) return DYT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* gy = static_cast<const __half*>(gy_f16);
  __half* oh = static_cast<__half*>(out_f16);
  const int grid = row_grid(n_rows);
#define DYT_LNB(NV)                                                                              \
  launch_ln_bwd<NV>(grid, st, gy, ldg, x, ldx, row_idx, n_rows_dev, n_rows, gamma, eps, resid, ldr, \
                    row_scale, axpy, out, ldo, oh, ldoh)
  switch (C) {
    case 768: DYT_LNB(6); break;
    case 1024: DYT_LNB(8); break;
    case 384: DYT_LNB(3); break;
    case 128: DYT_LNB(1); break;
    default:
      return fail(DYT_EUNSUPPORTED, "layernorm_bwd: embed dim %d not instantiated (128/384/768/1024)", C);
  }
#undef DYT_LNB
  return cuda_status(cudaGetLastError(), "layernorm_bwd_kernel launch");
}

extern "C" int dyt_merge_bwd(const float* g_out, int ldg, const void* mlp_f16, int ldm,
                             const float* mask, const float* logits, const float* noise1,
                             const float* noise2, float tau, const float* g_token_select,
                             const float* g_token_logits, int B, int N, int C, void* g_f16, int ld16,
                             void* g_masked_f16, int ldgm, float* g_logit, const int* token_pos,
                             const float* sel_w, float* gx_init, int ldgx, void* stream) {
  DYT_CHECK_ARG(g_out && g_f16, "merge_bwd: null buffer");
  DYT_CHECK_ARG(B >= 0 && N >= 1 && ldg >= C && ld16 >= C && ldg % 4 == 0 && ld16 % 4 == 0,
                "merge_bwd: bad sizes");
  if (g_masked_f16 != nullptr) {
    DYT_CHECK_ARG(mlp_f16 && mask && logits && g_logit && ldm >= C && ldm % 4 == 0 && ldgm >= C &&
                      ldgm % 4 == 0,
                  "merge_bwd: the masked path needs mlp / mask / logits / g_logit");
    DYT_CHECK_ARG((noise1 == nullptr) == (noise2 == nullptr), "merge_bwd: noise1 and noise2 go together");
    DYT_CHECK_ARG(noise1 == nullptr || tau > 0.f, "merge_bwd: tau must be positive");
  } else {
    DYT_CHECK_ARG(token_pos == nullptr && gx_init == nullptr,
                  "merge_bwd: token_pos / gx_init belong to the masked form");
  }
  DYT_CHECK_ARG(gx_init == nullptr || (sel_w != nullptr && ldgx >= C && ldgx % 4 == 0),
                "merge_bwd: gx_init needs the selector weight and a valid stride");
  const int T = B * N;
  if (T == 0) return DYT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = row_grid(T);
#define DYT_MB(NV)                                                                                 \
  merge_bwd_kernel<NV><<<grid, 256, 0, st>>>(                                                      \
      g_out, ldg, static_cast<const __half*>(mlp_f16), ldm, mask, logits, noise1, noise2, tau,     \
      g_token_select, g_token_logits, T, N, static_cast<__half*>(g_f16), ld16,                     \
      static_cast<__half*>(g_masked_f16), ldgm, g_logit, token_pos, sel_w, gx_init, ldgx)
  switch (C) {
    case 768: DYT_MB(6); break;
    case 1024: DYT_MB(8); break;
    case 384: DYT_MB(3); break;
    case 128: DYT_MB(1); break;
    default:
      return fail(DYT_EUNSUPPORTED, "merge_bwd: embed dim %d not instantiated (128/384/768/1024)", C);
  }
#undef DYT_MB
  return cuda_status(cudaGetLastError(), "merge_bwd_kernel launch");
}

extern "C" int dyt_rowscale_colsum(const float* s, const void* x_f16, int ldx, int T, int C,
                                   float* out_w, float* out_b, void* stream) {
  DYT_CHECK_ARG(s && x_f16 && out_w, "rowscale_colsum: null buffer");
  DYT_CHECK_ARG(T >= 0 && C > 0 && C % 2 == 0 && C <= 2048 && ldx >= C && ldx % 2 == 0,
                "rowscale_colsum: bad sizes (C even, <= 2048)");
  if (T == 0) return DYT_OK;
  if (C % 8 == 0 && C >= 64 && C <= 1024 && ldx % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(x_f16) & 15) == 0) {
    const int groups = dyt::RC_THREADS / (C / 8);
    int rows_per_cta = (T + sm_count() * 2 - 1) / (sm_count() * 2);
    if (rows_per_cta < 4 * groups) rows_per_cta = 4 * groups;
    const int grid = (T + rows_per_cta - 1) / rows_per_cta;
    const size_t smem = (static_cast<size_t>(groups) * C + groups) * sizeof(float);
    rowscale_colsum_vec_kernel<<<grid, dyt::RC_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        s, static_cast<const __half*>(x_f16), ldx, T, C, rows_per_cta, out_w, out_b);
    return cuda_status(cudaGetLastError(), "rowscale_colsum_vec_kernel launch");
  }
  int rows_per_cta = (T + sm_count() * 2 - 1) / (sm_count() * 2);
  if (rows_per_cta < 8) rows_per_cta = 8;
  const int grid = (T + rows_per_cta - 1) / rows_per_cta;
  rowscale_colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      s, static_cast<const __half*>(x_f16), ldx, T, C, rows_per_cta, out_w, out_b);
  return cuda_status(cudaGetLastError(), "rowscale_colsum_kernel launch");
}

extern "C" int dyt_eltwise_f16(int op, const void* a, const void* b, const void* c, void* out,
                               size_t n, void* stream) {
  DYT_CHECK_ARG(a && out, "eltwise: null buffer");
  DYT_CHECK_ARG(n % 8 == 0, "eltwise: element count must be a multiple of 8");
  DYT_CHECK_ARG(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                  reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "eltwise: buffers must be 16-byte aligned");
  DYT_CHECK_ARG(op == DYT_EW_GELU_FWD || b != nullptr, "eltwise: second operand missing");
  if (n == 0) return DYT_OK;
  const size_t n8 = n / 8;
  size_t blocks = (n8 + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint4* pa = static_cast<const uint4*>(a);
  const uint4* pb = static_cast<const uint4*>(b);
  const uint4* pc = static_cast<const uint4*>(c);
  uint4* po = static_cast<uint4*>(out);
  const int g = static_cast<int>(blocks);
  switch (op) {
    case DYT_EW_GELU_FWD: eltwise_f16_kernel<EW_GELU_FWD><<<g, 256, 0, st>>>(pa, pb, pc, po, n8); break;
    case DYT_EW_GELU_BWD: eltwise_f16_kernel<EW_GELU_BWD><<<g, 256, 0, st>>>(pa, pb, pc, po, n8); break;
    case DYT_EW_RELU_DROP_BWD:
      eltwise_f16_kernel<EW_RELU_DROP_BWD><<<g, 256, 0, st>>>(pa, pb, pc, po, n8);
      break;
    case DYT_EW_MUL: eltwise_f16_kernel<EW_MUL><<<g, 256, 0, st>>>(pa, pb, pc, po, n8); break;
    default: return fail(DYT_EINVAL, "eltwise: unknown op %d", op);
  }
  return cuda_status(cudaGetLastError(), "eltwise_f16_kernel launch");
}

extern "C" int dyt_gelu_bwd_rows(const void* g_f16, int ldg, const void* pre_f16, int ldp,
                                 const int* row_idx, const int* n_rows_dev, int n_rows, int H,
                                 void* out_f16, int ldo, void* stream) {
  DYT_CHECK_ARG(g_f16 && pre_f16 && row_idx && out_f16, "gelu_bwd_rows: null buffer");
  DYT_CHECK_ARG(n_rows >= 0 && H > 0 && H % 8 == 0 && ldg >= H && ldp >= H && ldo >= H &&
                    ldg % 8 == 0 && ldp % 8 == 0 && ldo % 8 == 0,
                "gelu_bwd_rows: H and the strides must be multiples of 8");
  DYT_CHECK_ARG(((reinterpret_cast<uintptr_t>(g_f16) | reinterpret_cast<uintptr_t>(pre_f16) |
                  reinterpret_cast<uintptr_t>(out_f16)) & 15) == 0,
                "gelu_bwd_rows: buffers must be 16-byte aligned");
  if (n_rows == 0) return DYT_OK;
  int grid = n_rows;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  gelu_bwd_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(g_f16), ldg, static_cast<const __half*>(pre_f16), ldp, row_idx,
      n_rows_dev, n_rows, H, static_cast<__half*>(out_f16), ldo);
  return cuda_status(cudaGetLastError(), "gelu_bwd_rows_kernel launch");
}

extern "C" int dyt_wgrad_f16(const void* g_f16, int ldg, const void* x_f16, int ldx, int T, int Nout,
                             int Kin, float alpha, float* dW, int ldw, float* db, void* stream) {
  DYT_CHECK_ARG(g_f16 && x_f16 && dW, "wgrad: null buffer");
  DYT_CHECK_ARG(T >= 0 && Nout > 0 && Kin > 0 && ldg >= Nout && ldx >= Kin && ldw >= Kin,
                "wgrad: bad sizes");
  if (T == 0) return DYT_OK;
  const int tk = (Kin + WG_TILE - 1) / WG_TILE, tn = (Nout + WG_TILE - 1) / WG_TILE;
  // token splits: about six CTAs per SM.  More splits shorten the serial tile loop of a CTA but add
  // reductions to the same dW addresses; measured per launch at 12.6k tokens (adapter down / up of
  // ViT-B, 16-byte reductions): 12.0 / 11.4 / 11.2 / 13.8 us for 4 / 5 / 6 / 8 CTAs per SM (scalar
  // atomics: 17.1 us at 4, 23.5 us at 8).
  int splits = (sm_count() * 6 + tk * tn - 1) / (tk * tn);
  int tok = (T + splits - 1) / splits;
  tok = (tok + WG_TOK - 1) / WG_TOK * WG_TOK;
  if (tok < 4 * WG_TOK) tok = 4 * WG_TOK;
  splits = (T + tok - 1) / tok;
  DYT_CHECK_ARG(tn <= 65535 && splits <= 65535, "wgrad: grid too large");
  dim3 grid(tk, tn, splits);
  wgrad_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(g_f16), ldg, static_cast<const __half*>(x_f16), ldx, T, Nout, Kin,
      tok, alpha, dW, ldw, db);
  return cuda_status(cudaGetLastError(), "wgrad_kernel launch");
}
