// Fused token dispatcher (sm_100a, HBM-bound): one persistent kernel, no host sync.
//   phase 1  score  : logit[b,n] = <x1[b,n,:], w> + bias for the N-1 patch tokens (warp per token,
//                     128-bit coalesced row loads, fp16-rounded operands / fp32 accumulate / one
//                     rounding of the logit when the activation dtype is fp16)
//            gate   : keep = sigmoid(logit) > threshold evaluated in the logit dtype, implemented as
//                     the equivalent monotone test `logit >= min_kept` (the host derives min_kept
//                     from torch's own sigmoid, see dyt_b200/gate.py); train mode adds the two
//                     caller-drawn Gumbel terms and the temperature first.  cls token always kept.
//   grid barrier    : all CTAs are co-resident (grid <= occupancy * #SMs)
//   phase 2  compact: per-image exclusive base from the per-image counts, ballot/popc warp scan of
//                     the keep flags -> packed_idx (ascending flat index == nonzero() order),
//                     token_pos (inverse map), cu_seqlens, n_kept
//            pack   : for every kept token LayerNorm2(x1 row) -> fp16 row of the packed buffer
//                     (the A operand of the MLP fc1 GEMM)
//
// Replaces TokenSelect.forward + _gumbel_sigmoid (reference models/dynamic_adapter.py:25-77,
// models/model_speed_test.py:27-60), the nonzero()/gather glue (models/model_speed_test.py:297-301)
// and norm2 on the gathered rows (:303).
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
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
  int* counts;          // [B] workspace (see dispatch_kernel)
  unsigned int* sync;   // [2] workspace; zero before the first launch, left zeroed by the kernel
};

constexpr int DISPATCH_MAX_N = 2048;
constexpr int DISPATCH_CHUNK = 32;   // tokens per phase-2 (compaction) work item
constexpr int DISPATCH_SMEM_B = 1024;  // largest batch whose per-image prefix lives in smem

__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Work distribution: phase 1 strides warps over all B*N rows (consecutive warps read consecutive
// rows); phase 2 strides warps over (image, 16-row chunk) items, each warp deriving the packed
// base of its image from the per-image counts and its rank inside the image from the mask.
// Workspace: sync[0] barrier arrivals, sync[1] finished CTAs, counts[B] kept tokens per image.
// All of it must be zero at launch; the last CTA to finish zeroes it again, so a workspace that was
// zero-filled once stays valid for every later launch, whatever its B.
template <int NV>
__global__ void __launch_bounds__(256)
dispatch_kernel(const DispatchParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int N = p.N;
  const int gw = blockIdx.x * 8 + warp;
  const int nw = gridDim.x * 8;
  const int T = p.B * N;

  __shared__ int s_last;
  int* cur = p.counts;

  // selector weight in registers, laid out like a row
  float4 w[NV];
  load_row_f32<NV>(p.sel_w, lane, w);
  float bias = p.sel_b[0];
  if (p.logit_fp16) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      w[i].x = r16(w[i].x); w[i].y = r16(w[i].y); w[i].z = r16(w[i].z); w[i].w = r16(w[i].w);
    }
    bias = r16(bias);
  }

  // ------------------------------- phase 1: score + gate -------------------------------
  if (p.partials != nullptr) {
    // the dot products come from the proj GEMM epilogue: one thread per token, no pass over x1
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
      const int b = t / N;
      const int n = t - b * N;
      if (n == 0) {
        p.mask[t] = 1.0f;
        if (p.gate_out != nullptr) p.gate_out[t] = 1.0f;
        atomicAdd(cur + b, 1);
        continue;
      }
      float acc = 0.f;
      for (int i = 0; i < p.n_partials; ++i) acc += p.partials[static_cast<size_t>(t) * p.n_partials + i];
      float logit = acc + bias;
      if (p.logit_fp16) logit = r16(logit);
      float g = logit;
      const size_t li = static_cast<size_t>(b) * (N - 1) + (n - 1);
      if (p.noise1 != nullptr) {
        if (p.logit_fp16) {
          g = r16(g + r16(p.noise1[li]));
          g = r16(g - r16(p.noise2[li]));
          g = r16(g / r16(p.tau));
        } else {
          g = ((g + p.noise1[li]) - p.noise2[li]) / p.tau;
        }
      }
      bool keep = g >= p.min_kept;
      if (p.gate_out != nullptr) p.gate_out[t] = keep ? 1.0f : 0.0f;
      if (p.forced_mask != nullptr) keep = p.forced_mask[t] != 0.0f;
      p.logits[li] = logit;
      p.mask[t] = keep ? 1.0f : 0.0f;
      if (keep) atomicAdd(cur + b, 1);
    }
  } else
  for (int t = gw; t < T; t += nw) {
    const int b = t / N;
    const int n = t - b * N;
    if (n == 0) {
      if (lane == 0) {
        p.mask[t] = 1.0f;
        if (p.gate_out != nullptr) p.gate_out[t] = 1.0f;
        atomicAdd(cur + b, 1);
      }
      continue;
    }
    float4 v[NV];
    load_row_f32<NV>(p.x1 + static_cast<size_t>(t) * p.ldx, lane, v);
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
    if (lane == 0 && p.gate_out != nullptr) p.gate_out[t] = keep ? 1.0f : 0.0f;
    if (p.forced_mask != nullptr) keep = p.forced_mask[t] != 0.0f;
    if (lane == 0) {
      p.logits[li] = logit;
      p.mask[t] = keep ? 1.0f : 0.0f;
      if (keep) atomicAdd(cur + b, 1);
    }
  }

  // ------------------------------- grid barrier -------------------------------
  // (monotonic arrival counter: the k-th barrier waits for k * gridDim.x arrivals)
  auto grid_barrier = [&](unsigned int k) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&p.sync[0], 1u);
      const long long t0 = clock64();
      while (atomicAdd(&p.sync[0], 0u) < k * gridDim.x) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000ll) {
          printf("dyt: dispatcher grid barrier timeout (block %d)\n", (int)blockIdx.x);
          __trap();
        }
      }
      __threadfence();
    }
    __syncthreads();
  };
  grid_barrier(1u);

  // ------------------------------- phase 2: compact + pack -------------------------------
  // exclusive prefix of the per-image counts, once per CTA (falls back to a per-item sum for
  // batches beyond the smem table)
  __shared__ int s_base[DISPATCH_SMEM_B + 1];
  const bool smem_base = p.B <= DISPATCH_SMEM_B;
  if (smem_base) {
    __shared__ int s_warp[8];
    int carry = 0;
    for (int b0 = 0; b0 < p.B; b0 += 256) {
      const int i = b0 + threadIdx.x;
      const int c = i < p.B ? __ldcg(cur + i) : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      int woff = 0;
      for (int w = 0; w < warp; ++w) woff += s_warp[w];
      if (i < p.B) s_base[i] = carry + woff + incl - c;
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      carry += tot;
      __syncthreads();
    }
    if (threadIdx.x == 0) s_base[p.B] = carry;
    __syncthreads();
  }
  const int cpi = (N + DISPATCH_CHUNK - 1) / DISPATCH_CHUNK;  // chunks per image
  const int items = p.B * cpi;
  for (int item = gw; item < items; item += nw) {
    const int b = item / cpi;
    const int n0 = (item - b * cpi) * DISPATCH_CHUNK;
    // packed base of the image = kept tokens of all previous images
    int base;
    if (smem_base) {
      base = s_base[b];
    } else {
      int part = 0;
      for (int i = lane; i < b; i += 32) part += __ldcg(cur + i);
      base = warp_sum_int(part);
    }
    // kept tokens of this image before the chunk: every lane counts its share of the mask row
    const float* mrow = p.mask + static_cast<size_t>(b) * N;
    int mine = 0;
    for (int n = lane; n < n0; n += 32) mine += (__ldcg(mrow + n) != 0.0f) ? 1 : 0;
    const int before = warp_sum_int(mine);
    const int n = n0 + lane;
    const bool valid = lane < DISPATCH_CHUNK && n < N;
    const bool keep = valid && (__ldcg(mrow + n) != 0.0f);
    unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int rank = __popc(ballot & ((1u << lane) - 1u));
    const int first = base + before;
    if (valid) {
      const size_t t = static_cast<size_t>(b) * N + n;
      if (keep) {
        p.packed_idx[first + rank] = static_cast<int>(t);
        p.token_pos[t] = first + rank;
      } else {
        p.token_pos[t] = -1;
      }
    }
    if (n0 == 0 && lane == 0) {
      p.cu_seqlens[b] = base;
      if (b == p.B - 1) {
        const int total = base + __ldcg(cur + b);
        p.cu_seqlens[p.B] = total;
        p.n_kept[0] = total;
      }
    }
  }

  // ------------------------------- phase 3: pack -------------------------------
  // LayerNorm2 of the kept rows into the packed fp16 buffer: warps stride over packed positions
  // (perfectly balanced, whatever the per-image keep counts), two rows in flight per warp.
  if (p.packed != nullptr) {
    grid_barrier(2u);
    int total;
    if (smem_base) {
      total = s_base[p.B];
    } else {
      total = __ldcg(p.n_kept);
    }
    for (int j = gw; j < total; j += 2 * nw) {
      const int j1 = j + nw;
      const int t0 = __ldcg(p.packed_idx + j);
      const int t1 = j1 < total ? __ldcg(p.packed_idx + j1) : -1;
      float4 v0[NV], v1[NV];
      load_row_f32<NV>(p.x1 + static_cast<size_t>(t0) * p.ldx, lane, v0);
      if (t1 >= 0) load_row_f32<NV>(p.x1 + static_cast<size_t>(t1) * p.ldx, lane, v1);
      row_layernorm<NV>(v0, p.ln_w, p.ln_b, p.eps, lane);
      store_row_f16<NV>(p.packed + static_cast<size_t>(j) * p.ldp, lane, v0);
      if (t1 >= 0) {
        row_layernorm<NV>(v1, p.ln_w, p.ln_b, p.eps, lane);
        store_row_f16<NV>(p.packed + static_cast<size_t>(j1) * p.ldp, lane, v1);
      }
    }
  }

  // last CTA out: leave the workspace zeroed for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(&p.sync[1], 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < p.B; i += blockDim.x) cur[i] = 0;
    if (threadIdx.x == 0) {
      p.sync[0] = 0u;
      p.sync[1] = 0u;
    }
    __threadfence();
  }
}

template <int NV>
static int launch_dispatch(const DispatchParams& p, cudaStream_t stream) {
  static int max_blocks = 0;
  if (max_blocks == 0) {
    int per_sm = 0;
    DYT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dispatch_kernel<NV>, 256, 0));
    if (per_sm < 1) return fail(DYT_EDRIVER, "dispatch kernel does not fit on an SM");
    if (per_sm > 4) per_sm = 4;
    max_blocks = per_sm * sm_count();  // all CTAs co-resident: required by the grid barrier
  }
  const int want = (p.B * p.N + 7) / 8;
  const int grid = want < max_blocks ? want : max_blocks;
  dispatch_kernel<NV><<<grid, 256, 0, stream>>>(p);
  return cuda_status(cudaGetLastError(), "dispatch_kernel launch");
}

}  // namespace dyt

extern "C" size_t dyt_dispatch_workspace_bytes(int B) {
  return static_cast<size_t>(B) * sizeof(int) + 64;
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
  p.sync = static_cast<unsigned int*>(workspace);
  p.counts = reinterpret_cast<int*>(static_cast<char*>(workspace) + 64);
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
