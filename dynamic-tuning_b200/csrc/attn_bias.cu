// Long-sequence attention with an additive per-head bias (first, correctness-first version):
//   out = softmax(f16(q * scale) k^T + bias[h]) v      per (image, head), any sequence length
// for the segmentation backbone's eager attention path with a relative-position bias
// (reference dense_tasks/Segmentation/backbone/segmentation_vision_transformer_IN21K.py:181-203:
// 1025 tokens at 512 x 512, bias [num_heads, N, N] gathered from relative_position_bias_table).
// Rounding points of that path under fp16 autocast: q * scale and the scores q k^T are fp16
// tensors, the bias add promotes to fp32, softmax runs in fp32, the probabilities are cast to fp16
// for the PV product (fp32 accumulate, fp16 output).
//
// Flash-style: one CTA (4 warps) per (image, head, block of 64 queries); keys / values stream
// through two shared-memory buffers in blocks of 64 (cp.async, the next block in flight); S and PV on HMMA through nvcuda::wmma; the online softmax
// works on the accumulator registers (m16n16k16 fp32 layout, verified at run time like
// attn_bwd.cu).  The tcgen05 kernel (attn_varlen.cu) holds all keys of a sequence in one TMEM tile
// and stops at 256 keys; this kernel has no such limit and is the base for SURVEY section 8f rank 5.
#include <mma.h>
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"

namespace dyt {

constexpr int FB_Q = 64, FB_K = 64, FB_LD = 72, FB_WARPS = 4;
constexpr int FB_TILE = FB_K * FB_LD;                       // halves per 64-row tile
// Q, 2 x K, 2 x V (double-buffered with cp.async), per-warp P tiles, per-warp fp32 staging
constexpr size_t FB_SMEM = (5 * FB_TILE + FB_WARPS * 16 * FB_LD) * 2 + FB_WARPS * 16 * 20 * 4;

// 16-byte asynchronous global -> shared copy; src_bytes = 0 writes zeros (rows past the sequence)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

using namespace nvcuda;
typedef wmma::fragment<wmma::accumulator, 16, 16, 16, float> FbC;
typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> FbA;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> FbB;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> FbBt;

__global__ void __launch_bounds__(FB_WARPS * 32)
attn_bias_fwd_kernel(const __half* __restrict__ qkv, int ld_qkv, const float* __restrict__ bias,
                     int N, int H, int C, float scale, __half* __restrict__ out, int ldo) {
  extern __shared__ __align__(128) unsigned char fb_smem[];
  __half* Qs = reinterpret_cast<__half*>(fb_smem);
  __half* Kbuf = Qs + FB_TILE;              // [2][FB_TILE]
  __half* Vbuf = Kbuf + 2 * FB_TILE;        // [2][FB_TILE]
  __half* Pall = Vbuf + 2 * FB_TILE;        // [FB_WARPS][16 * FB_LD]
  float* stg_base = reinterpret_cast<float*>(Pall + FB_WARPS * 16 * FB_LD);

  const int qblocks = (N + FB_Q - 1) / FB_Q;
  const int qb = blockIdx.x % qblocks;
  const int h = (blockIdx.x / qblocks) % H;
  const int b = blockIdx.x / (qblocks * H);
  const int q0 = qb * FB_Q;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  float* stg = stg_base + warp * 16 * 20;
  __half* Pw = Pall + warp * 16 * FB_LD;
  const size_t row_base = static_cast<size_t>(b) * N;

  // accumulator layout self-check (see attn_bwd.cu)
  if (warp == 0) {
    const int rr = lane >> 1, cb = (lane & 1) * 8;
#pragma unroll
    for (int c = 0; c < 8; ++c) stg[rr * 20 + cb + c] = static_cast<float>(rr * 16 + cb + c);
    __syncwarp();
    FbC chk;
    wmma::load_matrix_sync(chk, stg, 20, wmma::mem_row_major);
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      ok = ok && chk.x[i] == static_cast<float>((g + 8 * ((i >> 1) & 1)) * 16 + 2 * t4 + (i & 1) + 8 * (i >> 2));
    if (!ok) __trap();
    __syncwarp();
  }

  // Q block, scaled: q * scale is an fp16 tensor in the reference (:190)
  for (int e = tid; e < FB_Q * 8; e += FB_WARPS * 32) {
    const int r = e >> 3, c = e & 7;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (q0 + r < N) {
      q = *reinterpret_cast<const uint4*>(qkv + (row_base + q0 + r) * ld_qkv + h * 64 + c * 8);
      __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hq[j]);
        hq[j] = __floats2half2_rn(f.x * scale, f.y * scale);
      }
    }
    *reinterpret_cast<uint4*>(Qs + r * FB_LD + c * 8) = q;
  }
  __syncthreads();
  FbA aq[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) wmma::load_matrix_sync(aq[kk], Qs + warp * 16 * FB_LD + kk * 16, FB_LD);

  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;  // the two query rows of this lane
  const float* bias0 = bias != nullptr ? bias + (static_cast<size_t>(h) * N + min(row0, N - 1)) * N : nullptr;
  const float* bias1 = bias != nullptr ? bias + (static_cast<size_t>(h) * N + min(row1, N - 1)) * N : nullptr;
  float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;
  FbC acc[4];
#pragma unroll
  for (int dn = 0; dn < 4; ++dn) wmma::fill_fragment(acc[dn], 0.f);
  __half* pw0 = Pw + g * FB_LD + 2 * t4;
  __half* pw1 = pw0 + 8 * FB_LD;

  // K / V blocks stream through two shared-memory buffers: block i+1 is in flight (cp.async) while
  // block i is multiplied
  auto load_block = [&](int k0, int buf) {
    for (int e = tid; e < FB_K * 8; e += FB_WARPS * 32) {
      const int r = e >> 3, c = e & 7;
      const bool ok = k0 + r < N;
      const __half* row = qkv + (row_base + (ok ? k0 + r : 0)) * ld_qkv + h * 64 + c * 8;
      cp_async16(Kbuf + buf * FB_TILE + r * FB_LD + c * 8, row + C, ok ? 16 : 0);
      cp_async16(Vbuf + buf * FB_TILE + r * FB_LD + c * 8, row + 2 * C, ok ? 16 : 0);
    }
    cp_async_commit();
  };
  load_block(0, 0);
  int buf = 0;
  for (int k0 = 0; k0 < N; k0 += FB_K, buf ^= 1) {
    __syncthreads();  // every warp is done with the block that used the other buffer
    if (k0 + FB_K < N) {
      load_block(k0 + FB_K, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();  // this block has landed for every thread
    const __half* Ks = Kbuf + buf * FB_TILE;
    const __half* Vs = Vbuf + buf * FB_TILE;

    // scores of this warp's 16 rows against the 64 keys: fp16-rounded q k^T, then + bias in fp32
    float sv[4][8];
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      FbC s;
      wmma::fill_fragment(s, 0.f);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        FbBt bk;
        wmma::load_matrix_sync(bk, Ks + jn * 16 * FB_LD + kk * 16, FB_LD);
        wmma::mma_sync(s, aq[kk], bk, s);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = k0 + jn * 16 + 8 * (i >> 2) + 2 * t4 + (i & 1);
        float v = __half2float(__float2half_rn(s.x[i]));
        if (bias != nullptr && j < N) v += ((i >> 1) & 1) ? bias1[j] : bias0[j];
        v = j < N ? v : -1e30f;
        sv[jn][i] = v;
        if ((i >> 1) & 1) mx1 = fmaxf(mx1, v); else mx0 = fmaxf(mx0, v);
      }
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, off));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, off));
    }
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = __expf(m0 - mn0), a1 = __expf(m1 - mn1);
    m0 = mn0;
    m1 = mn1;
    l0 *= a0;
    l1 *= a1;
#pragma unroll
    for (int dn = 0; dn < 4; ++dn)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[dn].x[i] *= ((i >> 1) & 1) ? a1 : a0;
    // P = exp(s - m) -> fp16 into this warp's tile (the operand of the PV product)
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      float pv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool r1 = (i >> 1) & 1;
        pv[i] = __expf(sv[jn][i] - (r1 ? m1 : m0));
        if (r1) l1 += pv[i]; else l0 += pv[i];   // fp32 softmax denominator (:199)
      }
      *reinterpret_cast<__half2*>(pw0 + jn * 16) = __floats2half2_rn(pv[0], pv[1]);
      *reinterpret_cast<__half2*>(pw0 + jn * 16 + 8) = __floats2half2_rn(pv[4], pv[5]);
      *reinterpret_cast<__half2*>(pw1 + jn * 16) = __floats2half2_rn(pv[2], pv[3]);
      *reinterpret_cast<__half2*>(pw1 + jn * 16 + 8) = __floats2half2_rn(pv[6], pv[7]);
    }
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      FbA ap;
      wmma::load_matrix_sync(ap, Pw + kk * 16, FB_LD);
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        FbB bv;
        wmma::load_matrix_sync(bv, Vs + kk * 16 * FB_LD + dn * 16, FB_LD);
        wmma::mma_sync(acc[dn], ap, bv, acc[dn]);
      }
    }
    __syncwarp();  // Ps is rewritten by the next block
  }

#pragma unroll
  for (int off = 1; off <= 2; off <<= 1) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, off);
    l1 += __shfl_xor_sync(0xffffffffu, l1, off);
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int rr = lane >> 1, cb = (lane & 1) * 8;
  __half* orow = out + (row_base + q0 + warp * 16) * ldo + h * 64;
#pragma unroll
  for (int dn = 0; dn < 4; ++dn) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[dn].x[i] *= ((i >> 1) & 1) ? i1 : i0;
    wmma::store_matrix_sync(stg, acc[dn], 20, wmma::mem_row_major);
    __syncwarp();
    if (q0 + warp * 16 + rr < N) {
      uint4 u;
      __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        hh[c] = __floats2half2_rn(stg[rr * 20 + cb + 2 * c], stg[rr * 20 + cb + 2 * c + 1]);
      *reinterpret_cast<uint4*>(orow + static_cast<size_t>(rr) * ldo + dn * 16 + cb) = u;
    }
    __syncwarp();
  }
}

}  // namespace dyt

extern "C" int dyt_attn_bias_fwd(const void* qkv, int ld_qkv, const float* bias, int num_seqs,
                                 int seq_len, int num_heads, int head_dim, void* out, int ldo,
                                 void* stream) {
  using namespace dyt;
  DYT_CHECK_ARG(qkv && out, "attn_bias: null buffer");
  DYT_CHECK_ARG(head_dim == 64, "attn_bias: head_dim must be 64 (got %d)", head_dim);
  DYT_CHECK_ARG(num_seqs >= 0 && seq_len >= 1 && num_heads > 0, "attn_bias: bad sizes");
  const int C = num_heads * head_dim;
  DYT_CHECK_ARG(ld_qkv >= 3 * C && ldo >= C && ld_qkv % 8 == 0 && ldo % 8 == 0,
                "attn_bias: strides must cover the row and be multiples of 8");
  DYT_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "attn_bias: buffers must be 16-byte aligned");
  if (num_seqs == 0) return DYT_OK;
  const long blocks = static_cast<long>(num_seqs) * num_heads * ((seq_len + FB_Q - 1) / FB_Q);
  DYT_CHECK_ARG(blocks < (1l << 31), "attn_bias: grid too large");
  static SmemAttrCache smem_cache;
  {
    const int st = ensure_dyn_smem(attn_bias_fwd_kernel, static_cast<int>(FB_SMEM), smem_cache);
    if (st != DYT_OK) return st;
  }
  attn_bias_fwd_kernel<<<static_cast<unsigned>(blocks), FB_WARPS * 32, FB_SMEM, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(qkv), ld_qkv, bias, seq_len, num_heads, C, 1.0f / 8.0f,
      static_cast<__half*>(out), ldo);
  return cuda_status(cudaGetLastError(), "attn_bias_fwd_kernel launch");
}
