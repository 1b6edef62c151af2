// Fused token dispatcher (sm_100a, HBM-bound): one kernel, one CTA per image, no host sync and no
// grid-wide barrier.
//   gate    : logit[b,n] = <x1[b,n,:], w> + bias for the N-1 patch tokens -- inside a block the dot
//             products arrive as column partial sums from the proj GEMM epilogue, stand-alone the
//             CTA's warps compute them (128-bit coalesced row loads, fp16-rounded operands / fp32
//             accumulate / one rounding of the logit when the activation dtype is fp16);
//             keep = sigmoid(logit) > threshold evaluated in the logit dtype, implemented as the
//             equivalent monotone test `logit >= min_kept` (the host derives min_kept from torch's
//             own sigmoid, see dyt_b200/gate.py); train mode adds the two caller-drawn Gumbel terms
//             and the temperature first.  cls token always kept.
//   prefix  : the image's packed base = kept tokens of all previous images, by decoupled look-back
//             over per-image status words (aggregate / inclusive prefix, tagged with a launch
//             epoch kept in the workspace, images taken in ticket order so predecessors always run)
//   compact : ballot/popc ranks -> packed_idx (ascending flat index == nonzero() order), token_pos
//             (inverse map), cu_seqlens, n_kept
//   pack    : for every kept token LayerNorm2(x1 row) -> fp16 row of the packed buffer (the A
//             operand of the MLP fc1 GEMM), the CTA's warps striding over the image's kept rows
//
// Replaces TokenSelect.forward + _gumbel_sigmoid (reference models/dynamic_adapter.py:25-77,
// models/model_speed_test.py:27-60), the nonzero()/gather glue (models/model_speed_test.py:297-301)
// and norm2 on the gathered rows (:303).
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "ptx.cuh"
#include "rowwise.cuh"

namespace dyt {

struct DispatchParams {
  const float* x1;      // [B*N, ldx] fp32 residual stream after attention
  int ldx;
  const float* sel_w;   // [C] selector weight (fp32 master copy)
  const float* sel_b;   // [1] selector bias
  int logit_fp16;       // 1: emulate fp16 autocast rounding points; 0: pure fp32
  float min_kept;       // keep iff gate input >= min_kept
  const float* noise1;  // optional [B, N-1] Gumbel draws (train mode); nullptr in eval
  const float* noise2;
  float tau;
  int B, N;
  const float* ln_w;    // norm2 weight / bias
  const float* ln_b;
  float eps;
  const float* forced_mask;  // optional [B, N]: imposed keep mask (BASELINE config 1), cls forced
  float* mask;          // [B, N]   out: 1.0 kept / 0.0 dropped, cls = 1 (after forcing)
  float* gate_out;      // optional [B, N]: the selector's own decision, ignoring forced_mask
  float* logits;        // [B, N-1] out
  int* packed_idx;      // [B*N]    out (first n_kept entries valid)
  int* token_pos;       // [B*N]    out: packed position or -1
  int* cu_seqlens;      // [B+1]    out
  int* n_kept;          // [1]      out
  __half* packed;       // [B*N, ldp] out: LayerNorm2 of the kept rows, fp16
  int ldp;
  const float* partials;  // optional [B*N, n_partials]: the score dot products, already computed in
  int n_partials;         // column slices by the producer of x1 (proj GEMM epilogue); summed in order
  unsigned int* ctl;            // workspace: [0] ticket, [1] finished CTAs, [2] launch epoch
  unsigned long long* status;   // workspace: [B] per-image (epoch tag << 32 | flag << 30 | count)
};

constexpr int DISPATCH_MAX_N = 2048;

__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr unsigned long long ST_AGG = 1ull << 30;     // count of this image only
constexpr unsigned long long ST_PREFIX = 2ull << 30;  // inclusive count up to this image
constexpr unsigned int ST_VALUE_MASK = (1u << 30) - 1u;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Workspace: ctl[0] image ticket, ctl[1] finished CTAs, ctl[2] launch epoch, status[B].  It must
// be zero before the first launch; the last CTA of a launch resets ticket / finished and bumps the
// epoch, so stale status words of earlier launches never match (graph-replay safe: no host state).
template <int NV>
__device__ __forceinline__ void dispatch_body(const DispatchParams& p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int N = p.N;

  __shared__ int s_b, s_cnt, s_base;
  __shared__ unsigned int s_tag;
  __shared__ short s_list[DISPATCH_MAX_N];       // kept token indices of the image, ascending
  __shared__ unsigned char s_keep[DISPATCH_MAX_N];
  __shared__ int s_wsum[8];

  pdl_launch_dependents();
  pdl_wait();  // x1 / the score partials of the preceding GEMM
  if (threadIdx.x == 0) {
    s_b = static_cast<int>(atomicAdd(&p.ctl[0], 1u));   // ticket order: predecessors are running
    s_tag = *reinterpret_cast<volatile unsigned int*>(&p.ctl[2]) + 1u;
  }
  __syncthreads();
  const int b = s_b;
  const unsigned long long tag = static_cast<unsigned long long>(s_tag) << 32;
  float bias = p.sel_b[0];
  if (p.logit_fp16) bias = r16(bias);
  const float* x_img = p.x1 + static_cast<size_t>(b) * N * p.ldx;

  // ------------------------------- gate -------------------------------
  auto decide = [&](int n, float acc) {   // one thread per token: logit -> keep flag + outputs
    const size_t t = static_cast<size_t>(b) * N + n;
    float logit = acc + bias;
    if (p.logit_fp16) logit = r16(logit);
    float g = logit;
    const size_t li = static_cast<size_t>(b) * (N - 1) + (n - 1);
    if (p.noise1 != nullptr) {
      // (logits + g1 - g2) / tau, every op rounded in the logit dtype (dynamic_adapter.py:41)
      if (p.logit_fp16) {
        g = r16(g + r16(p.noise1[li]));
        g = r16(g - r16(p.noise2[li]));
        g = r16(g / r16(p.tau));
      } else {
        g = ((g + p.noise1[li]) - p.noise2[li]) / p.tau;
      }
    }
    bool keep = g >= p.min_kept;  // NaN -> dropped, +inf -> kept
    if (p.gate_out != nullptr) p.gate_out[t] = keep ? 1.0f : 0.0f;
    if (p.forced_mask != nullptr) keep = p.forced_mask[t] != 0.0f;
    p.logits[li] = logit;
    p.mask[t] = keep ? 1.0f : 0.0f;
    s_keep[n] = keep ? 1 : 0;
  };
  if (threadIdx.x == 0) {
    p.mask[static_cast<size_t>(b) * N] = 1.0f;
    if (p.gate_out != nullptr) p.gate_out[static_cast<size_t>(b) * N] = 1.0f;
    s_keep[0] = 1;
  }
  if (p.partials != nullptr) {
    // the dot products come from the proj GEMM epilogue as column partial sums
    for (int n = 1 + threadIdx.x; n < N; n += 256) {
      const float* pp = p.partials + (static_cast<size_t>(b) * N + n) * p.n_partials;
      float acc = 0.f;
      for (int i = 0; i < p.n_partials; ++i) acc += pp[i];
      decide(n, acc);
    }
  } else {
    float4 w[NV];
    load_row_f32<NV>(p.sel_w, lane, w);
    if (p.logit_fp16) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        w[i].x = r16(w[i].x); w[i].y = r16(w[i].y); w[i].z = r16(w[i].z); w[i].w = r16(w[i].w);
      }
    }
    for (int n = 1 + warp; n < N; n += 8) {
      float4 v[NV];
      load_row_f32<NV>(x_img + static_cast<size_t>(n) * p.ldx, lane, v);
      float acc = 0.f;
      if (p.logit_fp16) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          acc = fmaf(r16(v[i].x), w[i].x, acc);
          acc = fmaf(r16(v[i].y), w[i].y, acc);
          acc = fmaf(r16(v[i].z), w[i].z, acc);
          acc = fmaf(r16(v[i].w), w[i].w, acc);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          acc = fmaf(v[i].x, w[i].x, acc);
          acc = fmaf(v[i].y, w[i].y, acc);
          acc = fmaf(v[i].z, w[i].z, acc);
          acc = fmaf(v[i].w, w[i].w, acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) decide(n, acc);
    }
  }
  __syncthreads();

  // ------------------------------- compact (inside the image) -------------------------------
  if (warp == 0) {
    int cnt = 0;
    for (int n0 = 0; n0 < N; n0 += 32) {
      const int n = n0 + lane;
      const bool k = n < N && s_keep[n] != 0;
      const unsigned ballot = __ballot_sync(0xffffffffu, k);
      if (k) s_list[cnt + __popc(ballot & ((1u << lane) - 1u))] = static_cast<short>(n);
      cnt += __popc(ballot);
    }
    if (lane == 0) s_cnt = cnt;
  }
  __syncthreads();   // the image's list of kept tokens is complete
  // While warp 0 walks the look-back (it waits for the images before this one), the other warps load
  // and normalise their first two kept rows: only the DESTINATION of a packed row needs the base.
  const bool do_pack = p.packed != nullptr;
  float4 v0[NV], v1[NV];
  bool preloaded = false;
  if (warp != 0 && do_pack) {
    const int c0 = s_cnt;
    if (warp < c0) {
      load_row_f32<NV>(x_img + static_cast<size_t>(s_list[warp]) * p.ldx, lane, v0);
      if (warp + 8 < c0) load_row_f32<NV>(x_img + static_cast<size_t>(s_list[warp + 8]) * p.ldx, lane, v1);
      row_layernorm<NV>(v0, p.ln_w, p.ln_b, p.eps, lane);
      if (warp + 8 < c0) row_layernorm<NV>(v1, p.ln_w, p.ln_b, p.eps, lane);
      preloaded = true;
    }
  }
  if (warp == 0) {
    const int cnt = s_cnt;
    // ------------------------------- prefix over the images: decoupled look-back ----------------
    int excl = 0;
    if (b > 0) {
      if (lane == 0) st_status(&p.status[b], tag | ST_AGG | static_cast<unsigned int>(cnt));
      int j = b - 1;
      long long t0 = 0;
      while (j >= 0) {
        const int idx = j - lane;
        unsigned long long s = 0;
        bool ready;
        unsigned int polls = 0;
        do {  // all (valid) lanes of the window have published something in this launch
          if (idx >= 0) s = ld_status(&p.status[idx]);
          ready = idx < 0 || (s >> 32) == (tag >> 32);
          if ((++polls & 1023u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000ll) __trap();
          }
        } while (!__all_sync(0xffffffffu, ready));
        const bool is_prefix = idx >= 0 && (s & ST_PREFIX) != 0;
        const unsigned pb = __ballot_sync(0xffffffffu, is_prefix);
        const int first = pb ? __ffs(pb) - 1 : 32;   // nearest predecessor holding a prefix
        const int val = (idx >= 0 && lane <= first) ? static_cast<int>(s & ST_VALUE_MASK) : 0;
        excl += warp_sum_int(val);
        if (pb) break;
        j -= 32;
      }
    }
    if (lane == 0) {
      st_status(&p.status[b], tag | ST_PREFIX | static_cast<unsigned int>(excl + cnt));
      s_base = excl;
      p.cu_seqlens[b] = excl;
      if (b == p.B - 1) {
        p.cu_seqlens[p.B] = excl + cnt;
        p.n_kept[0] = excl + cnt;
      }
    }
  }
  __syncthreads();
  const int cnt = s_cnt, base = s_base;

  // ------------------------------- index maps -------------------------------
  for (int n = threadIdx.x; n < N; n += 256) p.token_pos[static_cast<size_t>(b) * N + n] = -1;
  __syncthreads();
  for (int r = threadIdx.x; r < cnt; r += 256) {
    const int t = b * N + s_list[r];
    p.packed_idx[base + r] = t;
    p.token_pos[t] = base + r;
  }

  // ------------------------------- pack -------------------------------
  // LayerNorm2 of the kept rows into the packed fp16 buffer, two rows in flight per warp
  if (do_pack) {
    for (int r = warp; r < cnt; r += 16) {
      const int r1 = r + 8;
      if (!preloaded) {
        load_row_f32<NV>(x_img + static_cast<size_t>(s_list[r]) * p.ldx, lane, v0);
        if (r1 < cnt) load_row_f32<NV>(x_img + static_cast<size_t>(s_list[r1]) * p.ldx, lane, v1);
        row_layernorm<NV>(v0, p.ln_w, p.ln_b, p.eps, lane);
        if (r1 < cnt) row_layernorm<NV>(v1, p.ln_w, p.ln_b, p.eps, lane);
      }
      preloaded = false;
      store_row_f16<NV>(p.packed + static_cast<size_t>(base + r) * p.ldp, lane, v0);
      if (r1 < cnt) store_row_f16<NV>(p.packed + static_cast<size_t>(base + r1) * p.ldp, lane, v1);
    }
  }

  // last CTA out: next launch starts with ticket 0 and a new epoch
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(&p.ctl[1], 1u);
    if (prev == gridDim.x - 1) {
      p.ctl[0] = 0u;
      p.ctl[1] = 0u;
      p.ctl[2] = s_tag;   // epoch + 1
      __threadfence();
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
dispatch_kernel(const DispatchParams p) {
  dispatch_body<NV>(p);
}
template <int NV>
static int launch_dispatch(const DispatchParams& p, cudaStream_t stream) {
  // one CTA per image, taken in ticket order
  return cuda_status(launch_pdl(dispatch_kernel<NV>, dim3(p.B), dim3(256), 0, stream, p),
                     "dispatch_kernel launch");
}

}  // namespace dyt

namespace dyt {
// ---------------------------------------------------------------------------------------------
// TokenSelect.forward alone (score + gate, no compaction): one warp per token over the whole grid.
// Same arithmetic as the dispatcher's gate above; used by the train-mode forward, where the dense
// masked block needs only the mask and the logits (models/dynamic_adapter.py:70-77) and a batch of
// 64 images would leave most SMs idle under the one-CTA-per-image dispatcher.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
token_select_kernel(const float* __restrict__ x1, int ldx, const float* __restrict__ sel_w,
                    const float* __restrict__ sel_b, int logit_fp16, float min_kept,
                    const float* __restrict__ noise1, const float* __restrict__ noise2, float tau,
                    int T, int N, float* __restrict__ mask, float* __restrict__ logits,
                    int* __restrict__ row_of) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  float bias = sel_b[0];
  if (logit_fp16) bias = r16(bias);
  float4 w[NV];
  load_row_f32<NV>(sel_w, lane, w);
  if (logit_fp16) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      w[i].x = r16(w[i].x); w[i].y = r16(w[i].y); w[i].z = r16(w[i].z); w[i].w = r16(w[i].w);
    }
  }
  for (int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < T;
       t += gridDim.x * warps_per_block) {
    const int n = t % N;
    if (n == 0) {  // cls: always kept, no logit
      if (lane == 0) {
        mask[t] = 1.0f;
        if (row_of != nullptr) row_of[t] = t;
      }
      continue;
    }
    float4 v[NV];
    load_row_f32<NV>(x1 + static_cast<size_t>(t) * ldx, lane, v);
    float acc = 0.f;
    if (logit_fp16) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        acc = fmaf(r16(v[i].x), w[i].x, acc);
        acc = fmaf(r16(v[i].y), w[i].y, acc);
        acc = fmaf(r16(v[i].z), w[i].z, acc);
        acc = fmaf(r16(v[i].w), w[i].w, acc);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        acc = fmaf(v[i].x, w[i].x, acc);
        acc = fmaf(v[i].y, w[i].y, acc);
        acc = fmaf(v[i].z, w[i].z, acc);
        acc = fmaf(v[i].w, w[i].w, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float logit = acc + bias;
      if (logit_fp16) logit = r16(logit);
      float g = logit;
      const size_t li = static_cast<size_t>(t / N) * (N - 1) + (n - 1);
      if (noise1 != nullptr) {
        if (logit_fp16) {
          g = r16(g + r16(noise1[li]));
          g = r16(g - r16(noise2[li]));
          g = r16(g / r16(tau));
        } else {
          g = ((g + noise1[li]) - noise2[li]) / tau;
        }
      }
      logits[li] = logit;
      const bool keep = g >= min_kept;
      mask[t] = keep ? 1.0f : 0.0f;
      if (row_of != nullptr) row_of[t] = keep ? t : -1;  // dense "token_pos" for dyt_scatter_merge_fwd
    }
  }
}

}  // namespace dyt

extern "C" int dyt_token_select_fwd(const float* x1, int ldx, const float* sel_w, const float* sel_b,
                                    int logit_fp16, float min_kept, const float* noise1,
                                    const float* noise2, float tau, int B, int N, int C, float* mask,
                                    float* logits, int* row_of, void* stream) {
  using namespace dyt;
  DYT_CHECK_ARG(x1 && sel_w && sel_b && mask && logits, "token_select: null buffer");
  DYT_CHECK_ARG(B >= 0 && N >= 2 && ldx >= C && ldx % 4 == 0, "token_select: bad sizes");
  DYT_CHECK_ARG((noise1 == nullptr) == (noise2 == nullptr), "token_select: need both noise tensors");
  DYT_CHECK_ARG(noise1 == nullptr || tau > 0.f, "token_select: tau must be positive");
  const int T = B * N;
  if (T == 0) return DYT_OK;
  int grid = (T + 7) / 8;
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define DYT_TS(NV)                                                                                 \
  token_select_kernel<NV><<<grid, 256, 0, st>>>(x1, ldx, sel_w, sel_b, logit_fp16, min_kept, noise1, \
                                                noise2, tau, T, N, mask, logits, row_of)
  switch (C) {
    case 768: DYT_TS(6); break;
    case 1024: DYT_TS(8); break;
    case 384: DYT_TS(3); break;
    case 128: DYT_TS(1); break;
    default:
      return fail(DYT_EUNSUPPORTED, "token_select: embed dim %d not instantiated (128/384/768/1024)", C);
  }
#undef DYT_TS
  return cuda_status(cudaGetLastError(), "token_select_kernel launch");
}

extern "C" size_t dyt_dispatch_workspace_bytes(int B) {
  return static_cast<size_t>(B) * sizeof(unsigned long long) + 64;
}

namespace dyt {
int dispatch_fwd(const float* x1, int ldx, const float* sel_w, const float* sel_b, int logit_fp16,
                 float min_kept, const float* noise1, const float* noise2, float tau, int B, int N,
                 int C, const float* ln_w, const float* ln_b, float eps, const float* forced_mask,
                 float* mask, float* gate_out, float* logits, int* packed_idx, int* token_pos,
                 int* cu_seqlens, int* n_kept, void* packed_f16, int ldp, void* workspace,
                 void* stream, const float* partials, int n_partials);
}

extern "C" int dyt_dispatch_fwd(const float* x1, int ldx, const float* sel_w, const float* sel_b,
                                int logit_fp16, float min_kept, const float* noise1,
                                const float* noise2, float tau, int B, int N, int C,
                                const float* ln_w, const float* ln_b, float eps,
                                const float* forced_mask, float* mask, float* gate_out,
                                float* logits, int* packed_idx, int* token_pos, int* cu_seqlens, int* n_kept,
                                void* packed_f16, int ldp, void* workspace, void* stream) {
  return dyt::dispatch_fwd(x1, ldx, sel_w, sel_b, logit_fp16, min_kept, noise1, noise2, tau, B, N, C,
                           ln_w, ln_b, eps, forced_mask, mask, gate_out, logits, packed_idx, token_pos,
                           cu_seqlens, n_kept, packed_f16, ldp, workspace, stream, nullptr, 0);
}

// Internal entry: `partials` ([B*N, n_partials] fp32) replaces the score pass over x1 by the column
// partial sums the proj GEMM epilogue has already produced (block.cu).
int dyt::dispatch_fwd(const float* x1, int ldx, const float* sel_w, const float* sel_b,
                      int logit_fp16, float min_kept, const float* noise1, const float* noise2,
                      float tau, int B, int N, int C, const float* ln_w, const float* ln_b,
                      float eps, const float* forced_mask, float* mask, float* gate_out,
                      float* logits, int* packed_idx, int* token_pos, int* cu_seqlens, int* n_kept,
                      void* packed_f16, int ldp, void* workspace, void* stream,
                      const float* partials, int n_partials) {
  using namespace dyt;
  DYT_CHECK_ARG(x1 && sel_w && sel_b && mask && logits && packed_idx && token_pos && cu_seqlens &&
                    n_kept && workspace,
                "dispatch: null buffer");
  DYT_CHECK_ARG(B >= 1 && N >= 1 && N <= DISPATCH_MAX_N, "dispatch: bad B=%d N=%d", B, N);
  DYT_CHECK_ARG(ldx >= C && ldx % 4 == 0, "dispatch: bad ldx");
  DYT_CHECK_ARG((noise1 == nullptr) == (noise2 == nullptr), "dispatch: need both noise tensors");
  DYT_CHECK_ARG(packed_f16 == nullptr || (ln_w && ln_b && ldp >= C && ldp % 4 == 0),
                "dispatch: packed output needs norm2 parameters");
  DYT_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "dispatch: workspace alignment");
  DispatchParams p;
  p.x1 = x1; p.ldx = ldx; p.sel_w = sel_w; p.sel_b = sel_b;
  p.logit_fp16 = logit_fp16; p.min_kept = min_kept;
  p.noise1 = noise1; p.noise2 = noise2; p.tau = tau;
  p.B = B; p.N = N;
  p.ln_w = ln_w; p.ln_b = ln_b; p.eps = eps;
  p.forced_mask = forced_mask;
  p.mask = mask; p.gate_out = gate_out; p.logits = logits; p.packed_idx = packed_idx; p.token_pos = token_pos;
  p.cu_seqlens = cu_seqlens; p.n_kept = n_kept;
  p.packed = static_cast<__half*>(packed_f16); p.ldp = ldp;
  p.partials = partials; p.n_partials = n_partials;
  p.ctl = static_cast<unsigned int*>(workspace);
  p.status = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + 64);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 768: return launch_dispatch<6>(p, s);
    case 1024: return launch_dispatch<8>(p, s);
    case 384: return launch_dispatch<3>(p, s);
    case 128: return launch_dispatch<1>(p, s);
    default:
      return fail(DYT_EUNSUPPORTED, "dispatch: embed dim %d not instantiated (128/384/768/1024)", C);
  }
}
