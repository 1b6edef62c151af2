// Variable-length multi-head attention over a packed [T, 3, H, 64] fp16 qkv buffer (sm_100a).
//   S = Q K^T  : tcgen05.mma kind::f16, A = Q tile (TMA, 128B-swizzled K-major), B = K tile
//                (TMA, K-major), fp32 accumulator in TMEM columns [0, nk)
//   softmax    : one thread per query row (TMEM lane): two passes over the TMEM row (max; exp2 +
//                fp32 row sum), P written back to TMEM as packed fp16 aliasing S columns [0, nk/2)
//   O = P V    : tcgen05.mma with A = P from TMEM, B = V tile (TMA, 128B-swizzled MN-major),
//                accumulator in TMEM columns [128, 192); normalised by the fp32 row sum, rounded
//                once to fp16 and stored as [T, H*64].
// One CTA per (sequence, head, 128-query tile); sequences of up to 256 keys (one KV tile), so no
// online-softmax rescaling is needed.  256 TMEM columns and ~70 KB smem per CTA -> 2 CTAs / SM, so
// one CTA's softmax overlaps the other's TMA/MMA.
//
// Replaces F.scaled_dot_product_attention in Attention.forward of the reference
// (models/model_speed_test.py:145-166; models/vision_transformer_IN21K.py:54-75): non-causal,
// scale = head_dim^-0.5, dropout 0, q_norm/k_norm = Identity.
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "ptx.cuh"

namespace dyt {

struct AttnParams {
  const int* cu_seqlens;  // [B+1] device int32, or nullptr -> every sequence has uniform_len tokens
  int uniform_len;
  int nk_box;  // rows of the K/V TMA box: round_up(max_seqlen, 16) <= 256
  int C;       // H * 64
  __half* out;
  int ldo;
  float scale_log2e;  // head_dim^-0.5 * log2(e)
};

constexpr int ATT_BM = 128;
constexpr int ATT_D = 64;
constexpr int ATT_TMEM_COLS = 256;
constexpr int ATT_O_COL = 128;

__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,
                const __grid_constant__ CUtensorMap tmap_kv, const AttnParams p) {
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  int seq_start, seq_len;
  if (p.cu_seqlens != nullptr) {
    seq_start = p.cu_seqlens[b];
    seq_len = p.cu_seqlens[b + 1] - seq_start;
  } else {
    seq_start = b * p.uniform_len;
    seq_len = p.uniform_len;
  }
  if (seq_len > p.nk_box) seq_len = p.nk_box;
  if (t * ATT_BM >= seq_len) return;  // uniform for the whole CTA, before any barrier / TMEM use
  const int nk = (seq_len + 15) & ~15;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_BM * 128;
  uint8_t* sV = sK + p.nk_box * 128;  // nk_box is a multiple of 16 -> 2 KB granularity keeps 1 KB alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.nk_box * 128);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 5 && lane == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp_idx == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_kv);
    }
    tmem_alloc(tmem_ptr_smem, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 4) {
    if (lane == 0) {
      const uint32_t kv_bytes = static_cast<uint32_t>(p.nk_box) * 128u;
      mbar_arrive_expect_tx(bar_qk, ATT_BM * 128 + kv_bytes);
      tma_load_2d(sQ, &tmap_q, bar_qk, h * ATT_D, seq_start + t * ATT_BM);
      tma_load_2d(sK, &tmap_kv, bar_qk, p.C + h * ATT_D, seq_start);
      mbar_arrive_expect_tx(bar_v, kv_bytes);
      tma_load_2d(sV, &tmap_kv, bar_v, 2 * p.C + h * ATT_D, seq_start);
    }
  } else if (warp_idx == 5) {
    if (lane == 0) {
      // ---- S = Q K^T ----
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_f16(ATT_BM, nk, 0, 0);
      const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
      const uint64_t k_desc = umma_desc_sw128(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k)
        umma_ss_f16(tmem_base, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
      umma_commit(bar_s);
      // ---- O = P V ----
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = umma_idesc_f16(ATT_BM, ATT_D, 0, 1);  // B = V is MN-major
      const uint64_t v_desc = umma_desc_sw128(smem_u32(sV));
      for (int kk = 0; kk < nk / 16; ++kk) {
        // 16 keys per MMA: 8 TMEM columns of packed fp16 P; 16 V rows = 2048 B = +128 (16 B units)
        umma_ts_f16(tmem_base + ATT_O_COL, tmem_base + kk * 8, v_desc + kk * 128, idesc_o,
                    kk != 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
  } else {
    // ===================== softmax + output (warps 0..3, TMEM lane quarter = warp_idx) ==========
    const int q = warp_idx;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nchunks = (nk + 31) >> 5;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float v = (c * 32 + j < seq_len) ? __uint_as_float(r[j]) : -INFINITY;
        mx = fmaxf(mx, v);
      }
    }
    const float sl2 = p.scale_log2e;
    const float mb = mx * sl2;
    float sum = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c * 32 + 2 * j;
        float e0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), sl2, -mb));
        float e1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -mb));
        e0 = (col < seq_len) ? e0 : 0.f;
        e1 = (col + 1 < seq_len) ? e1 : 0.f;
        sum += e0 + e1;
        pk[j] = pack_half2(e0, e1);
      }
      tmem_st16(taddr + c * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld32(taddr + ATT_O_COL, o0);
    tmem_ld32(taddr + ATT_O_COL + 32, o1);
    tmem_ld_wait();
    const int qrow = t * ATT_BM + row;
    if (qrow < seq_len) {
      const float inv = 1.0f / sum;
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(seq_start + qrow) * p.ldo +
                                            h * ATT_D);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        v.x = pack_half2(__uint_as_float(o0[8 * j + 0]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
        v.y = pack_half2(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
        v.z = pack_half2(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
        v.w = pack_half2(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
        dst[j] = v;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        v.x = pack_half2(__uint_as_float(o1[8 * j + 0]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
        v.y = pack_half2(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
        v.z = pack_half2(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
        v.w = pack_half2(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
        dst[4 + j] = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

int attn_varlen_fwd(const __half* qkv, int ld_qkv, const int* cu_seqlens, int num_seqs,
                    int uniform_len, int max_seqlen, int total_tokens, int num_heads, int head_dim,
                    __half* out, int ldo, cudaStream_t stream) {
  DYT_CHECK_ARG(qkv != nullptr && out != nullptr, "attn: null buffer");
  DYT_CHECK_ARG(head_dim == 64, "attn: only head_dim 64 is implemented (got %d)", head_dim);
  DYT_CHECK_ARG(num_seqs >= 0 && num_heads > 0 && total_tokens >= 0, "attn: bad sizes");
  DYT_CHECK_ARG(max_seqlen >= 1, "attn: max_seqlen must be >= 1");
  if (max_seqlen > 256)
    return fail(DYT_EUNSUPPORTED, "attn: sequences longer than 256 tokens are not implemented (%d)",
                max_seqlen);
  DYT_CHECK_ARG(cu_seqlens != nullptr || uniform_len == max_seqlen,
                "attn: uniform_len must equal max_seqlen when cu_seqlens is null");
  const int C = num_heads * head_dim;
  DYT_CHECK_ARG(ld_qkv >= 3 * C && ldo >= C && ldo % 8 == 0, "attn: bad leading dimensions");
  if (num_seqs == 0 || total_tokens == 0) return DYT_OK;

  const int nk_box = (max_seqlen + 15) & ~15;
  CUtensorMap tq, tkv;
  int s = make_tmap_f16_sw128(&tq, qkv, static_cast<uint64_t>(total_tokens),
                              static_cast<uint64_t>(3 * C), static_cast<uint64_t>(ld_qkv), ATT_BM);
  if (s != DYT_OK) return s;
  s = make_tmap_f16_sw128(&tkv, qkv, static_cast<uint64_t>(total_tokens),
                          static_cast<uint64_t>(3 * C), static_cast<uint64_t>(ld_qkv),
                          static_cast<uint32_t>(nk_box));
  if (s != DYT_OK) return s;

  AttnParams p;
  p.cu_seqlens = cu_seqlens;
  p.uniform_len = uniform_len;
  p.nk_box = nk_box;
  p.C = C;
  p.out = out;
  p.ldo = ldo;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  const int smem_bytes = 1024 + ATT_BM * 128 + 2 * nk_box * 128 + 64;
  static int configured_smem = 0;
  if (smem_bytes > configured_smem) {
    DYT_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  smem_bytes));
    configured_smem = smem_bytes;
  }
  dim3 grid((max_seqlen + ATT_BM - 1) / ATT_BM, num_heads, num_seqs);
  attn_fwd_kernel<<<grid, 192, smem_bytes, stream>>>(tq, tkv, p);
  return cuda_status(cudaGetLastError(), "attn_fwd_kernel launch");
}

}  // namespace dyt

extern "C" int dyt_attn_varlen_fwd(const void* qkv, int ld_qkv, const int* cu_seqlens, int num_seqs,
                                   int uniform_len, int max_seqlen, int total_tokens, int num_heads,
                                   int head_dim, void* out, int ldo, void* stream) {
  return dyt::attn_varlen_fwd(static_cast<const __half*>(qkv), ld_qkv, cu_seqlens, num_seqs,
                              uniform_len, max_seqlen, total_tokens, num_heads, head_dim,
                              static_cast<__half*>(out), ldo, static_cast<cudaStream_t>(stream));
}
