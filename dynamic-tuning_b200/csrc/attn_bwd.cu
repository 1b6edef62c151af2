// Attention backward (first version): dQ, dK, dV of softmax(Q K^T * scale) V per (sequence, head),
// recomputing the probabilities from Q and K instead of saving them (reference
// models/vision_transformer_IN21K.py:61-65, the backward of F.scaled_dot_product_attention).
//
// One CTA (16 warps) per (sequence, head); Q, K, V, dO of the head and the whole P / dS matrix stay
// in shared memory (fp16), every contraction runs on HMMA through nvcuda::wmma with fp32
// accumulation.  The softmax / dS row arithmetic works directly on the accumulator registers (the
// m16n16k16 fp32 accumulator layout of two m16n8 HMMAs: lane (g = lane/4, t = lane%4) holds rows g and
// g+8, columns 2t, 2t+1, 8+2t, 9+2t; verified at run time against wmma::load_matrix_sync, trap on
// mismatch); only the [16 x 64] output tiles pass through a per-warp fp32 staging tile.
// Phases (separated by CTA barriers):
//   1  per 16-query block: S = Q K^T twice (row max / sum, then P = exp(S - m) / l  -> smem)
//   2  per 16-key block:   dV = P^T dO
//   3  per 16-query block: D = rowsum(dO o O); dP = dO V^T; dS = P o (dP - D) * scale -> smem
//                          (in place of P); dQ = dS K
//   4  per 16-key block:   dK = dS^T Q
// Sequences up to 208 tokens (13 blocks of 16), head_dim 64.  This is the correctness-first
// kernel of the backward row; the tcgen05 version is the next step.
#include <mma.h>
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"

namespace dyt {

constexpr int AB_MAXN = 208;   // padded sequence length held in shared memory
constexpr int AB_LDQ = 72;     // row stride (halves) of the Q / K / V / dO tiles
constexpr int AB_LDP = 216;    // row stride (halves) of the P / dS matrix
constexpr int AB_WARPS = 16;
constexpr int AB_STG = 16 * 20;  // floats per warp staging tile
constexpr size_t AB_QBYTES = static_cast<size_t>(AB_MAXN) * AB_LDQ * 2;
constexpr size_t AB_PBYTES = static_cast<size_t>(AB_MAXN) * AB_LDP * 2;
constexpr size_t AB_SMEM = 4 * AB_QBYTES + AB_PBYTES + AB_WARPS * AB_STG * 4;

using namespace nvcuda;
typedef wmma::fragment<wmma::accumulator, 16, 16, 16, float> FragC;
typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> FragA;
typedef wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::col_major> FragAt;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> FragB;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> FragBt;

// [16 x 64] accumulators (4 fragments) -> fp16 rows of the gradient tensor, rows < n only
__device__ __forceinline__ void store_rows64(FragC (&acc)[4], float* stg, int lane, int row0, int n,
                                             __half* dst, int ld) {
  const int rr = lane >> 1, cb = (lane & 1) * 8;
#pragma unroll
  for (int dn = 0; dn < 4; ++dn) {
    wmma::store_matrix_sync(stg, acc[dn], 20, wmma::mem_row_major);
    __syncwarp();
    if (row0 + rr < n) {
      uint4 u;
      __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        h[c] = __floats2half2_rn(stg[rr * 20 + cb + 2 * c], stg[rr * 20 + cb + 2 * c + 1]);
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(row0 + rr) * ld + dn * 16 + cb) = u;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(AB_WARPS * 32, 1)
attn_bwd_kernel(const __half* __restrict__ qkv, int ld_qkv, const __half* __restrict__ o, int ldo,
                const __half* __restrict__ d_o, int ld_do, const int* __restrict__ cu_seqlens,
                int uniform_len, int H, int C, float scale, __half* __restrict__ dqkv, int ld_dqkv) {
  extern __shared__ __align__(128) unsigned char smem[];
  __half* Qs = reinterpret_cast<__half*>(smem);
  __half* Ks = reinterpret_cast<__half*>(smem + AB_QBYTES);
  __half* Vs = reinterpret_cast<__half*>(smem + 2 * AB_QBYTES);
  __half* Gs = reinterpret_cast<__half*>(smem + 3 * AB_QBYTES);   // dO
  __half* Ps = reinterpret_cast<__half*>(smem + 4 * AB_QBYTES);
  float* stg_all = reinterpret_cast<float*>(smem + 4 * AB_QBYTES + AB_PBYTES);

  const int seq = blockIdx.x / H, h = blockIdx.x % H;
  int start, n;
  if (cu_seqlens != nullptr) {
    start = cu_seqlens[seq];
    n = cu_seqlens[seq + 1] - start;
  } else {
    start = seq * uniform_len;
    n = uniform_len;
  }
  if (n <= 0) return;
  if (n > AB_MAXN) n = AB_MAXN;  // rejected on the host; never index past the tiles
  const int nb = (n + 15) >> 4;
  const int np = nb * 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* stg = stg_all + warp * AB_STG;
  const int rr = lane >> 1, cb = (lane & 1) * 8;

  // ---- stage Q, K, V, dO of this head (rows >= n are zero) ----
  for (int e = tid; e < np * 8; e += AB_WARPS * 32) {
    const int r = e >> 3, c = e & 7;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, g = q;
    if (r < n) {
      const __half* row = qkv + static_cast<size_t>(start + r) * ld_qkv + h * 64 + c * 8;
      q = *reinterpret_cast<const uint4*>(row);
      k = *reinterpret_cast<const uint4*>(row + C);
      v = *reinterpret_cast<const uint4*>(row + 2 * C);
      g = *reinterpret_cast<const uint4*>(d_o + static_cast<size_t>(start + r) * ld_do + h * 64 + c * 8);
    }
    *reinterpret_cast<uint4*>(Qs + r * AB_LDQ + c * 8) = q;
    *reinterpret_cast<uint4*>(Ks + r * AB_LDQ + c * 8) = k;
    *reinterpret_cast<uint4*>(Vs + r * AB_LDQ + c * 8) = v;
    *reinterpret_cast<uint4*>(Gs + r * AB_LDQ + c * 8) = g;
  }
  __syncthreads();

  // ---- accumulator layout self-check (warp 0): element i of lane (g, t) is (g + 8*((i>>1)&1),
  //      2t + (i&1) + 8*(i>>2)) ----
  const int g = lane >> 2, t4 = lane & 3;
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) stg[rr * 20 + cb + c] = static_cast<float>(rr * 16 + cb + c);
    __syncwarp();
    FragC chk;
    wmma::load_matrix_sync(chk, stg, 20, wmma::mem_row_major);
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      ok = ok && chk.x[i] == static_cast<float>((g + 8 * ((i >> 1) & 1)) * 16 + 2 * t4 + (i & 1) + 8 * (i >> 2));
    if (!ok) __trap();
    __syncwarp();
  }

  // ---- phase 1: P ----
  for (int ib = warp; ib < nb; ib += AB_WARPS) {
    const int i0 = ib * 16;
    FragA aq[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) wmma::load_matrix_sync(aq[kk], Qs + i0 * AB_LDQ + kk * 16, AB_LDQ);
    float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f, inv0 = 0.f, inv1 = 0.f;
    const bool live0 = i0 + g < n, live1 = i0 + g + 8 < n;
    __half* prow0 = Ps + (i0 + g) * AB_LDP + 2 * t4;
    __half* prow1 = prow0 + 8 * AB_LDP;
    for (int pass = 0; pass < 2; ++pass) {
      for (int jb = 0; jb < nb; ++jb) {
        FragC acc;
        wmma::fill_fragment(acc, 0.f);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          FragBt bk;
          wmma::load_matrix_sync(bk, Ks + jb * 16 * AB_LDQ + kk * 16, AB_LDQ);
          wmma::mma_sync(acc, aq[kk], bk, acc);
        }
        float sv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = jb * 16 + 8 * (i >> 2) + 2 * t4 + (i & 1);
          sv[i] = j < n ? acc.x[i] * scale : -1e30f;
        }
        if (pass == 0) {
          const float mx0 = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[4], sv[5]));
          const float mx1 = fmaxf(fmaxf(sv[2], sv[3]), fmaxf(sv[6], sv[7]));
          const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
          l0 = l0 * __expf(m0 - mn0) + (__expf(sv[0] - mn0) + __expf(sv[1] - mn0)) +
               (__expf(sv[4] - mn0) + __expf(sv[5] - mn0));
          l1 = l1 * __expf(m1 - mn1) + (__expf(sv[2] - mn1) + __expf(sv[3] - mn1)) +
               (__expf(sv[6] - mn1) + __expf(sv[7] - mn1));
          m0 = mn0;
          m1 = mn1;
        } else {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool r1 = (i >> 1) & 1;
            pv[i] = (r1 ? live1 : live0) ? __expf(sv[i] - (r1 ? m1 : m0)) * (r1 ? inv1 : inv0) : 0.f;
          }
          *reinterpret_cast<__half2*>(prow0 + jb * 16) = __floats2half2_rn(pv[0], pv[1]);
          *reinterpret_cast<__half2*>(prow0 + jb * 16 + 8) = __floats2half2_rn(pv[4], pv[5]);
          *reinterpret_cast<__half2*>(prow1 + jb * 16) = __floats2half2_rn(pv[2], pv[3]);
          *reinterpret_cast<__half2*>(prow1 + jb * 16 + 8) = __floats2half2_rn(pv[6], pv[7]);
        }
      }
      if (pass == 0) {  // merge the four lanes that share rows g and g + 8
#pragma unroll
        for (int off = 1; off <= 2; off <<= 1) {
          const float mo0 = __shfl_xor_sync(0xffffffffu, m0, off);
          const float lo0 = __shfl_xor_sync(0xffffffffu, l0, off);
          const float mo1 = __shfl_xor_sync(0xffffffffu, m1, off);
          const float lo1 = __shfl_xor_sync(0xffffffffu, l1, off);
          const float mt0 = fmaxf(m0, mo0), mt1 = fmaxf(m1, mo1);
          l0 = l0 * __expf(m0 - mt0) + lo0 * __expf(mo0 - mt0);
          l1 = l1 * __expf(m1 - mt1) + lo1 * __expf(mo1 - mt1);
          m0 = mt0;
          m1 = mt1;
        }
        inv0 = 1.f / l0;
        inv1 = 1.f / l1;
      }
    }
  }
  __syncthreads();

  __half* dq_out = dqkv + static_cast<size_t>(start) * ld_dqkv + h * 64;
  __half* dk_out = dq_out + C;
  __half* dv_out = dq_out + 2 * C;

  // ---- phase 2: dV_j = sum_i P_ij^T dO_i ----
  for (int jb = warp; jb < nb; jb += AB_WARPS) {
    FragC acc[4];
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) wmma::fill_fragment(acc[dn], 0.f);
    for (int ib = 0; ib < nb; ++ib) {
      FragAt a;
      wmma::load_matrix_sync(a, Ps + ib * 16 * AB_LDP + jb * 16, AB_LDP);
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        FragB b;
        wmma::load_matrix_sync(b, Gs + ib * 16 * AB_LDQ + dn * 16, AB_LDQ);
        wmma::mma_sync(acc[dn], a, b, acc[dn]);
      }
    }
    store_rows64(acc, stg, lane, jb * 16, n, dv_out, ld_dqkv);
  }
  __syncthreads();

  // ---- phase 3: dS (in place of P) and dQ ----
  for (int ib = warp; ib < nb; ib += AB_WARPS) {
    const int i0 = ib * 16;
    // D = sum_d dO[row, d] * O[row, d] for rows g and g + 8; lane t covers 16 of the 64 columns
    float dsum[2] = {0.f, 0.f};
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int row = i0 + g + 8 * hrow;
      if (row < n) {
        const __half* orow = o + static_cast<size_t>(start + row) * ldo + h * 64 + t4 * 16;
        const __half* grow = Gs + row * AB_LDQ + t4 * 16;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 uo = *reinterpret_cast<const uint4*>(orow + c * 8);
          const uint4 ug = *reinterpret_cast<const uint4*>(grow + c * 8);
          const __half2* po = reinterpret_cast<const __half2*>(&uo);
          const __half2* pg = reinterpret_cast<const __half2*>(&ug);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fo = __half22float2(po[e]);
            const float2 fg = __half22float2(pg[e]);
            dsum[hrow] += fo.x * fg.x + fo.y * fg.y;
          }
        }
      }
      dsum[hrow] += __shfl_xor_sync(0xffffffffu, dsum[hrow], 1);
      dsum[hrow] += __shfl_xor_sync(0xffffffffu, dsum[hrow], 2);
    }
    FragA ag[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) wmma::load_matrix_sync(ag[kk], Gs + i0 * AB_LDQ + kk * 16, AB_LDQ);
    __half* prow0 = Ps + (i0 + g) * AB_LDP + 2 * t4;
    __half* prow1 = prow0 + 8 * AB_LDP;
    for (int jb = 0; jb < nb; ++jb) {
      FragC acc;
      wmma::fill_fragment(acc, 0.f);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        FragBt bv;
        wmma::load_matrix_sync(bv, Vs + jb * 16 * AB_LDQ + kk * 16, AB_LDQ);
        wmma::mma_sync(acc, ag[kk], bv, acc);
      }
      __half2* q00 = reinterpret_cast<__half2*>(prow0 + jb * 16);
      __half2* q01 = reinterpret_cast<__half2*>(prow0 + jb * 16 + 8);
      __half2* q10 = reinterpret_cast<__half2*>(prow1 + jb * 16);
      __half2* q11 = reinterpret_cast<__half2*>(prow1 + jb * 16 + 8);
      const float2 p00 = __half22float2(*q00), p01 = __half22float2(*q01);
      const float2 p10 = __half22float2(*q10), p11 = __half22float2(*q11);
      *q00 = __floats2half2_rn(p00.x * (acc.x[0] - dsum[0]) * scale, p00.y * (acc.x[1] - dsum[0]) * scale);
      *q01 = __floats2half2_rn(p01.x * (acc.x[4] - dsum[0]) * scale, p01.y * (acc.x[5] - dsum[0]) * scale);
      *q10 = __floats2half2_rn(p10.x * (acc.x[2] - dsum[1]) * scale, p10.y * (acc.x[3] - dsum[1]) * scale);
      *q11 = __floats2half2_rn(p11.x * (acc.x[6] - dsum[1]) * scale, p11.y * (acc.x[7] - dsum[1]) * scale);
    }
    __syncwarp();  // this warp's dS rows are complete before its fragment loads read them
    FragC acc[4];
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) wmma::fill_fragment(acc[dn], 0.f);
    for (int jb = 0; jb < nb; ++jb) {
      FragA a;
      wmma::load_matrix_sync(a, Ps + i0 * AB_LDP + jb * 16, AB_LDP);
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        FragB b;
        wmma::load_matrix_sync(b, Ks + jb * 16 * AB_LDQ + dn * 16, AB_LDQ);
        wmma::mma_sync(acc[dn], a, b, acc[dn]);
      }
    }
    store_rows64(acc, stg, lane, i0, n, dq_out, ld_dqkv);
  }
  __syncthreads();

  // ---- phase 4: dK_j = sum_i dS_ij^T Q_i ----
  for (int jb = warp; jb < nb; jb += AB_WARPS) {
    FragC acc[4];
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) wmma::fill_fragment(acc[dn], 0.f);
    for (int ib = 0; ib < nb; ++ib) {
      FragAt a;
      wmma::load_matrix_sync(a, Ps + ib * 16 * AB_LDP + jb * 16, AB_LDP);
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        FragB b;
        wmma::load_matrix_sync(b, Qs + ib * 16 * AB_LDQ + dn * 16, AB_LDQ);
        wmma::mma_sync(acc[dn], a, b, acc[dn]);
      }
    }
    store_rows64(acc, stg, lane, jb * 16, n, dk_out, ld_dqkv);
  }
}

}  // namespace dyt

using namespace dyt;

extern "C" int dyt_attn_varlen_bwd(const void* qkv, int ld_qkv, const void* out, int ldo,
                                   const void* d_out, int ld_do, const int* cu_seqlens,
                                   int num_seqs, int uniform_len, int max_seqlen, int num_heads,
                                   int head_dim, void* d_qkv, int ld_dqkv, void* stream) {
  DYT_CHECK_ARG(qkv && out && d_out && d_qkv, "attn_bwd: null buffer");
  DYT_CHECK_ARG(head_dim == 64, "attn_bwd: head_dim must be 64 (got %d)", head_dim);
  DYT_CHECK_ARG(num_seqs >= 0 && num_heads > 0, "attn_bwd: bad sizes");
  DYT_CHECK_ARG(cu_seqlens != nullptr || uniform_len > 0, "attn_bwd: need cu_seqlens or uniform_len");
  const int mx = cu_seqlens != nullptr ? max_seqlen : uniform_len;
  if (mx > AB_MAXN)
    return fail(DYT_EUNSUPPORTED, "attn_bwd: sequences up to %d tokens (got %d)", AB_MAXN, mx);
  const int C = num_heads * head_dim;
  DYT_CHECK_ARG(ld_qkv >= 3 * C && ld_dqkv >= 3 * C && ldo >= C && ld_do >= C &&
                    ld_qkv % 8 == 0 && ld_dqkv % 8 == 0 && ldo % 8 == 0 && ld_do % 8 == 0,
                "attn_bwd: strides must cover the row and be multiples of 8");
  DYT_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) |
                  reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_qkv)) & 15) == 0,
                "attn_bwd: buffers must be 16-byte aligned");
  if (num_seqs == 0) return DYT_OK;
  static SmemAttrCache smem_cache;
  {
    const int st = ensure_dyn_smem(attn_bwd_kernel, static_cast<int>(AB_SMEM), smem_cache);
    if (st != DYT_OK) return st;
  }
  const float scale = 0.125f;  // head_dim^-0.5
  attn_bwd_kernel<<<num_seqs * num_heads, AB_WARPS * 32, AB_SMEM, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(qkv), ld_qkv, static_cast<const __half*>(out), ldo,
      static_cast<const __half*>(d_out), ld_do, cu_seqlens, uniform_len, num_heads, C, scale,
      static_cast<__half*>(d_qkv), ld_dqkv);
  return cuda_status(cudaGetLastError(), "attn_bwd_kernel launch");
}
