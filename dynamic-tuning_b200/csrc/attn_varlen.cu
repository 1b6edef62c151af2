// Variable-length multi-head attention over a packed [T, 3, H, 64] fp16 qkv buffer (sm_100a).
//
// Persistent kernel, one CTA per SM, work unit = (sequence, head); sequences of up to 256 keys
// (one KV tile, so no online-softmax rescaling), i.e. up to two 128-query tiles A and B per unit.
//   warp 0       TMA producer: Q (two 128-row boxes), K, V of the next unit into a 2-stage smem ring
//   warp 1       tcgen05.mma issuer (one thread), fixed order with blocking barrier waits:
//                  S = Q K^T : A = Q tile, B = K tile (both 128B-swizzled K-major smem), fp32
//                              accumulator in TMEM
//                  O = P V   : A = P from TMEM (packed fp16 aliasing the first half of S),
//                              B = V tile (128B-swizzled MN-major smem); issued in two parts, the
//                              first 128 keys as soon as the softmax has produced them
//                S_A(u0) S_B(u0) | PV_A(u) S_A(u+1) PV_B(u) S_B(u+1) | ...
//   warp 2       TMEM allocator (512 columns)
//   warps 4-7    softmax warpgroup of tile A, warps 8-11 of tile B: one thread per query row (TMEM
//                lane); pass 1 row max (three TMEM loads in flight), pass 2 exp2 + fp32 row sum, P
//                written back to TMEM as fp16, 1/rowsum handed to the output warps through smem.
//                The softmax warps never wait for the PV product: they go straight to the next S.
//   warps 12-15  output warpgroup: O read back from TMEM, normalised, rounded once to fp16,
//                transposed through smem and stored as full 128-byte rows of [T, H*64].
// Measured on B200 (scripts/ubench/tmem_mufu.cu): MUFU.EX2 8 clk per warp instruction per SM
// sub-partition, tcgen05.ld x32 ~180 clk latency, a TMEM-sourced 128x64x16 MMA ~85 clk: the kernel
// is bound by the ex2 pipe (~3300 clk per unit per sub-partition at N = 197) and by the PV MMAs.
// TMEM columns: S_A at 0, S_B at nk; O_A / O_B live outside the S regions when they fit
// (2*nk + 128 <= 512), otherwise inside their own S region at +128 (free once P is complete), in
// which case the next S of that tile has to wait for the O read-out.
//
// Replaces F.scaled_dot_product_attention in Attention.forward of the reference
// (models/model_speed_test.py:145-166; models/vision_transformer_IN21K.py:54-75): non-causal,
// scale = head_dim^-0.5, dropout 0, q_norm/k_norm = Identity.
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "internal.h"
#include "ptx.cuh"

namespace dyt {

struct AttnParams {
  const int* cu_seqlens;  // [B+1] device int32, or nullptr -> every sequence has uniform_len tokens
  int uniform_len;
  int nk_box;     // rows of the K/V TMA box: round_up(max_seqlen, 16) <= 256
  int C;          // H * 64
  int H;
  int num_units;  // num_seqs * H
  __half* out;
  int ldo;
  float scale_log2e;  // head_dim^-0.5 * log2(e)
  int s_col[2];       // TMEM column of S for tile A / B
  int o_col[2];       // TMEM column of O for tile A / B
  int o_alias[2];     // 1: O lives inside the tile's own S region
};

constexpr int ATT_BM = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 512;
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_Q_BYTES = 2 * ATT_BM * 128;  // both query tiles of a unit
constexpr int ATT_OSTAGE_BYTES = 4 * 32 * 128;  // per-output-warp transpose slabs
constexpr int ATT_INV_BYTES = 2 * 2 * ATT_BM * 4;  // 1/rowsum: [unit parity][tile][row]
constexpr int ATT_SPLIT_KEYS = 128;  // keys covered by the early first part of the PV product

__device__ __forceinline__ void unit_span(const AttnParams& p, int unit, int& b, int& h,
                                          int& seq_start, int& seq_len) {
  b = unit / p.H;
  h = unit - b * p.H;
  if (p.cu_seqlens != nullptr) {
    seq_start = __ldg(p.cu_seqlens + b);
    seq_len = __ldg(p.cu_seqlens + b + 1) - seq_start;
  } else {
    seq_start = b * p.uniform_len;
    seq_len = p.uniform_len;
  }
  if (seq_len > p.nk_box) seq_len = p.nk_box;
  if (seq_len < 0) seq_len = 0;
}

// ---- softmax building blocks.  A row (one thread) walks its keys in full 32-key chunks (no
// masking at all) and then in at most two 16-key tail pieces whose out-of-sequence keys are set to
// -inf first, so there is one copy of the arithmetic and the loops stay inside the instruction cache.
__device__ __forceinline__ void max32(const uint32_t (&r)[32], float (&mx)[4]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(r[j]));
}
__device__ __forceinline__ void mask16(uint32_t (&r)[16], int k0, int seq_len) {
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (k0 + j >= seq_len) r[j] = 0xff800000u;  // -inf: ignored by max, exp2 -> +0
}
// exponentials of one chunk (back-to-back MUFU), row-sum partials (four chains), fp16 P -> TMEM
__device__ __forceinline__ void exp32(uint32_t (&r)[32], float sl2, float mb, float (&sum)[4],
                                      uint32_t p_addr) {
#pragma unroll
  for (int j = 0; j < 32; ++j)
    r[j] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(r[j]), sl2, -mb)));
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float e0 = __uint_as_float(r[2 * j]), e1 = __uint_as_float(r[2 * j + 1]);
    sum[j & 3] += e0 + e1;
    pk[j] = pack_half2(e0, e1);
  }
  tmem_st16(p_addr, pk);
}
__device__ __forceinline__ void exp16(uint32_t (&r)[16], float sl2, float mb, float (&sum)[4],
                                      uint32_t p_addr) {
#pragma unroll
  for (int j = 0; j < 16; ++j)
    r[j] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(r[j]), sl2, -mb)));
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float e0 = __uint_as_float(r[2 * j]), e1 = __uint_as_float(r[2 * j + 1]);
    sum[j & 3] += e0 + e1;
    pk[j] = pack_half2(e0, e1);
  }
  tmem_st8(p_addr, pk);
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,
                const __grid_constant__ CUtensorMap tmap_kv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t kv_bytes = static_cast<uint32_t>(p.nk_box) * 128u;
  const uint32_t stage_bytes = ATT_Q_BYTES + 2 * kv_bytes;  // multiple of 2 KB
  uint8_t* o_stage = smem + 2 * stage_bytes;                // 4 warps x 32 rows x 128 B
  const uint32_t inv_addr = smem_u32(o_stage + ATT_OSTAGE_BYTES);  // float [2][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_stage + ATT_OSTAGE_BYTES + ATT_INV_BYTES);
  uint64_t* full_qk = bars + 0;     // [2] TMA -> MMA
  uint64_t* full_v = bars + 2;      // [2] TMA -> MMA
  uint64_t* qk_empty = bars + 4;    // [2] MMA -> TMA: both tiles' S MMAs have read Q / K
  uint64_t* s_full = bars + 6;      // [2: tile] MMA -> softmax
  uint64_t* p_half = bars + 8;      // [2: tile] softmax -> MMA (first ATT_SPLIT_KEYS keys of P)
  uint64_t* p_full = bars + 10;     // [2: tile] softmax -> MMA (all of P)
  uint64_t* o_full = bars + 12;     // [2: tile] MMA -> output warps
  uint64_t* o_free = bars + 14;     // [2: tile] output warps -> MMA (O read out)
  uint64_t* v_empty = bars + 16;    // [2] MMA -> TMA: both tiles' PV MMAs have read V
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  // kernel parameters are not indexed dynamically (would force a local-memory copy)
  auto s_col = [&](int tile) { return static_cast<uint32_t>(tile ? p.s_col[1] : p.s_col[0]); };
  auto o_col = [&](int tile) { return static_cast<uint32_t>(tile ? p.o_col[1] : p.o_col[0]); };
  auto o_alias = [&](int tile) { return (tile ? p.o_alias[1] : p.o_alias[0]) != 0; };

  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full_qk[i], 1);
      mbar_init(&full_v[i], 1);
      mbar_init(&qk_empty[i], 2);  // one arrival per tile stream
      mbar_init(&v_empty[i], 2);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_half[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 4);
    }
    fence_mbar_init();
  }
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // qkv (and cu_seqlens) of the previous kernel are read from here on

  // register budget per warpgroup (512 threads start at 128 each): control 56, softmax 168,
  // output 96 -> 128 * (56 + 2 * 168 + 96) = 62464 <= 65536

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    reg_dealloc<56>();
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        int b, h, seq_start, seq_len;
        unit_span(p, unit, b, h, seq_start, seq_len);
        if (seq_len == 0) continue;
        const int s = it & 1;
        mbar_wait(&qk_empty[s], ((it >> 1) & 1) ^ 1);
        uint8_t* sQ = smem + s * stage_bytes;
        uint8_t* sK = sQ + ATT_Q_BYTES;
        uint8_t* sV = sK + kv_bytes;
        const bool has_b = seq_len > ATT_BM;
        mbar_arrive_expect_tx(&full_qk[s], (has_b ? 2u : 1u) * ATT_BM * 128u + kv_bytes);
        tma_load_2d(sQ, &tmap_q, &full_qk[s], h * ATT_D, seq_start);
        tma_load_2d(sK, &tmap_kv, &full_qk[s], p.C + h * ATT_D, seq_start);
        if (has_b) tma_load_2d(sQ + ATT_BM * 128, &tmap_q, &full_qk[s], h * ATT_D, seq_start + ATT_BM);
        mbar_wait(&v_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_v[s], kv_bytes);
        tma_load_2d(sV, &tmap_kv, &full_v[s], 2 * p.C + h * ATT_D, seq_start);
        ++it;
      }
    }
  } else if (warp_idx == 1 || warp_idx == 3) {
    // ===================== MMA issuers: warp 1 drives tile A, warp 3 tile B =====================
    // Each tile is an independent stream S(u) -> [softmax] -> PV(u) -> S(u') ... with its own TMEM
    // regions; tcgen05.commit tracks the issuing thread's MMAs only, so the two streams never wait
    // for each other and the two softmax warpgroups drift out of phase: one is in its exponentials
    // (MUFU) while the other waits for its PV / next S.  Q/K and V of a smem stage are released
    // separately, each once both streams are done with it (two arrivals per barrier).
    reg_dealloc<56>();
    {
      // The whole warp runs this loop with warp-uniform values (descriptors and addresses live in
      // uniform registers); one elected lane issues the MMAs.  Issuing from inside an
      // `if (lane == 0)` region instead costs ~100 clk of scalar code per MMA -- more than a
      // 128x64x16 MMA itself takes.
      const int tile = warp_idx == 3 ? 1 : 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc_o = umma_idesc_f16(ATT_BM, ATT_D, 0, 1);  // B = V is MN-major
      const uint32_t d_tmem = tmem_u + o_col(tile);
      const uint32_t a_tmem = tmem_u + s_col(tile);
      const bool alias = o_alias(tile);
      const uint32_t smem_base_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      uint32_t k = 0;  // tile-units completed so far (-> barrier parities)
      int it = 0;      // index among the CTA's non-empty units (-> smem stage, ring parity)
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        int b, h, seq_start, seq_len;
        unit_span(p, unit, b, h, seq_start, seq_len);
        seq_len = __shfl_sync(0xffffffffu, seq_len, 0);
        if (seq_len == 0) continue;
        const int s = it & 1;
        const uint32_t ring_par = (it >> 1) & 1;
        ++it;
        mbar_wait(&full_qk[s], ring_par);
        if (seq_len <= tile * ATT_BM) {  // no such tile in this unit: only the stage accounting
          if (elect_one()) {
            mbar_arrive(&qk_empty[s]);
            mbar_arrive(&v_empty[s]);
          }
          __syncwarp();
          continue;
        }
        const int nk = (seq_len + 15) & ~15;
        const int steps = nk / 16;  // PV: 16 keys per MMA
        // first part of PV as soon as the first ATT_SPLIT_KEYS keys of P exist -- not when O lives
        // inside this tile's S region, which is still being read as S until P is complete
        const int early = (!alias && (seq_len >> 5) >= ATT_SPLIT_KEYS / 32) ? ATT_SPLIT_KEYS / 16 : 0;
        const uint32_t stage_addr = smem_base_u + s * stage_bytes;
        // ---- S = Q K^T ----
        if (alias) mbar_wait(&o_free[tile], (k & 1) ^ 1);  // O(k-1) sits inside this S region
        tc_fence_after();
        {
          const uint32_t idesc_s = umma_idesc_f16(ATT_BM, nk, 0, 0);
          const uint64_t q_desc = umma_desc_sw128(stage_addr + tile * ATT_BM * 128);
          const uint64_t k_desc = umma_desc_sw128(stage_addr + ATT_Q_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k16 = 0; k16 < ATT_D / 16; ++k16)
              umma_ss_f16(a_tmem, q_desc + 2 * k16, k_desc + 2 * k16, idesc_s, k16 != 0 ? 1u : 0u);
            umma_commit(&s_full[tile]);
            umma_commit(&qk_empty[s]);  // Q / K of this stage are free once both tiles' S retire
          }
          __syncwarp();
        }
        // ---- O = P V ----  (16 keys per MMA: 8 TMEM columns of packed fp16 P; 16 V rows = +128)
        const uint64_t v_desc = umma_desc_sw128(stage_addr + ATT_Q_BYTES + kv_bytes);
        mbar_wait(&p_half[tile], k & 1);
        mbar_wait(&full_v[s], ring_par);
        if (!alias) mbar_wait(&o_free[tile], (k & 1) ^ 1);  // O(k-1) read out
        tc_fence_after();
        if (elect_one()) {
          for (int kk = 0; kk < early; ++kk)
            umma_ts_f16(d_tmem, a_tmem + kk * 8, v_desc + kk * 128, idesc_o, kk != 0 ? 1u : 0u);
        }
        __syncwarp();
        mbar_wait(&p_full[tile], k & 1);
        tc_fence_after();
        if (elect_one()) {
          for (int kk = early; kk < steps; ++kk)
            umma_ts_f16(d_tmem, a_tmem + kk * 8, v_desc + kk * 128, idesc_o, kk != 0 ? 1u : 0u);
          umma_commit(&o_full[tile]);
          umma_commit(&v_empty[s]);  // V of this stage is free once both tiles' PV retire
        }
        __syncwarp();
        ++k;
      }
    }
  } else if (warp_idx == 2) {
    reg_dealloc<56>();
  } else if (warp_idx < 12) {
    // ===================== softmax =====================
    reg_alloc<168>();
    const int tile = (warp_idx - 4) >> 2;
    const int q = warp_idx & 3;  // TMEM lane quarter this warp may access
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + s_col(tile);
    const float sl2 = p.scale_log2e;
    uint32_t cnt = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int b, h, seq_start, seq_len;
      unit_span(p, unit, b, h, seq_start, seq_len);
      if (seq_len <= tile * ATT_BM) continue;  // this tile does not exist for the unit
      const bool active = tile * ATT_BM + q * 32 < seq_len;  // warp-uniform
      const int nfull = seq_len >> 5;                         // full 32-key chunks
      const int ntail = (((seq_len + 15) & ~15) - nfull * 32) >> 4;  // 16-key tail pieces (0..2)
      const bool split = nfull >= ATT_SPLIT_KEYS / 32;        // the issuer derives the same flag

      mbar_wait(&s_full[tile], cnt & 1);
      tc_fence_after();
      bool half_sent = false;
      if (active) {
        float mx;
        {
          // ---- pass 1: row max; three chunk loads in flight (a tcgen05.ld takes ~180 clk) ----
          uint32_t ra[32], rb[32], rc[32];
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          const int n1 = nfull;
          // (loads past the row's keys stay inside the 512 allocated columns and are ignored)
          tmem_ld32(s_addr, ra);
          tmem_ld32(s_addr + 32, rb);
          tmem_ld32(s_addr + 64, rc);
          tmem_ld_wait();
#pragma unroll 1
          for (int c = 0; c < n1; c += 3) {
            max32(ra, mx4);
            if (c + 3 < n1) tmem_ld32(s_addr + (c + 3) * 32, ra);
            if (c + 1 < n1) {
              max32(rb, mx4);
              if (c + 4 < n1) tmem_ld32(s_addr + (c + 4) * 32, rb);
            }
            if (c + 2 < n1) {
              max32(rc, mx4);
              if (c + 5 < n1) tmem_ld32(s_addr + (c + 5) * 32, rc);
            }
            tmem_ld_wait();
          }
#pragma unroll 1
          for (int t = 0; t < ntail; ++t) {
            uint32_t rt[16];
            const int k0 = nfull * 32 + t * 16;
            tmem_ld16(s_addr + k0, rt);
            tmem_ld_wait();
            mask16(rt, k0, seq_len);
#pragma unroll
            for (int j = 0; j < 16; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(rt[j]));
          }
          mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        }
        // ---- pass 2: exponentials, row sum, P -> TMEM (aliasing the consumed part of S) ----
        uint32_t ra[32], rb[32];
        const float mb = mx * sl2;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        tmem_ld32(s_addr, ra);
        tmem_ld_wait();
#pragma unroll 1
        for (int c = 0; c < nfull; c += 2) {
          if (c + 1 < nfull) tmem_ld32(s_addr + (c + 1) * 32, rb);
          exp32(ra, sl2, mb, sum4, s_addr + c * 16);
          tmem_ld_wait();
          if (c + 1 < nfull) {
            if (c + 2 < nfull) tmem_ld32(s_addr + (c + 2) * 32, ra);
            exp32(rb, sl2, mb, sum4, s_addr + (c + 1) * 16);
            tmem_ld_wait();
          }
          if (split && c + 2 == ATT_SPLIT_KEYS / 32) {
            // the first ATT_SPLIT_KEYS keys of P are in TMEM: let their PV MMAs start
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_half[tile]);
            half_sent = true;
          }
        }
#pragma unroll 1
        for (int t = 0; t < ntail; ++t) {
          uint32_t rt[16];
          const int k0 = nfull * 32 + t * 16;
          tmem_ld16(s_addr + k0, rt);
          tmem_ld_wait();
          mask16(rt, k0, seq_len);
          exp16(rt, sl2, mb, sum4, s_addr + (k0 >> 1));
        }
        tmem_st_wait();
        sts32(inv_addr + (((cnt & 1) * 2 + tile) * ATT_BM + q * 32 + lane) * 4,
              __float_as_uint(1.0f / ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]))));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!half_sent) mbar_arrive(&p_half[tile]);
        mbar_arrive(&p_full[tile]);
      }
      ++cnt;
    }
  } else {
    // ===================== output =====================
    reg_dealloc<96>();
    const int q = warp_idx & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t stg = smem_u32(o_stage) + q * (32 * 128);  // this warp's 32 x 128 B transpose slab
    uint32_t cnt_a = 0, cnt_b = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int b, h, seq_start, seq_len;
      unit_span(p, unit, b, h, seq_start, seq_len);
#pragma unroll 1
      for (int tile = 0; tile < 2; ++tile) {
        if (seq_len <= tile * ATT_BM) continue;
        const uint32_t cnt = tile ? cnt_b : cnt_a;
        const bool active = tile * ATT_BM + q * 32 < seq_len;  // warp-uniform
        mbar_wait(&o_full[tile], cnt & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        if (active) {
          const uint32_t o_addr = tmem_base + lane_off + o_col(tile);
          tmem_ld32(o_addr, o0);
          tmem_ld32(o_addr + 32, o1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[tile]);
        if (active) {
          // normalise, round once to fp16, transpose through the warp's smem slab (16-byte chunks
          // XOR-swizzled by row: conflict-free both ways) so that every global store instruction
          // writes four full 128-byte rows instead of 32 scattered 16-byte pieces.
          const float inv = __uint_as_float(
              lds32(inv_addr + (((cnt & 1) * 2 + tile) * ATT_BM + q * 32 + lane) * 4));
          const uint32_t my = stg + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v;
            v.x = pack_half2(__uint_as_float(o0[8 * j + 0]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
            v.y = pack_half2(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
            v.z = pack_half2(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
            v.w = pack_half2(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
            sts128(my + ((j ^ (lane & 7)) << 4), v);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v;
            v.x = pack_half2(__uint_as_float(o1[8 * j + 0]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
            v.y = pack_half2(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
            v.z = pack_half2(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
            v.w = pack_half2(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
            sts128(my + (((4 + j) ^ (lane & 7)) << 4), v);
          }
          __syncwarp();
          const int ch = lane & 7;    // 16-byte chunk of the 128-byte row
          const int rs = lane >> 3;   // row within a group of four
          const int row0 = tile * ATT_BM + q * 32;
          __half* gbase = p.out + static_cast<size_t>(seq_start) * p.ldo + h * ATT_D + ch * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + rs;
            const uint4 v = lds128(stg + r * 128 + ((ch ^ (r & 7)) << 4));
            if (row0 + r < seq_len)
              *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(row0 + r) * p.ldo) = v;
          }
          __syncwarp();  // the slab is rewritten by the next tile
        }
        if (tile) ++cnt_b; else ++cnt_a;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
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

  if (attn_split_option().load(std::memory_order_relaxed) != 0 &&
      attn_split_supported(cu_seqlens, uniform_len, head_dim))
    return attn_split_fwd(qkv, ld_qkv, num_seqs, uniform_len, total_tokens, num_heads, out, ldo, stream);

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
  p.H = num_heads;
  p.num_units = num_seqs * num_heads;
  p.out = out;
  p.ldo = ldo;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  // TMEM plan (512 columns).  The softmax reads S in 32-column chunks, so an S region is read up
  // to round_up(nk, 32) columns: reads past nk only touch other live regions, never write them.
  if (2 * nk_box + 2 * ATT_D <= ATT_TMEM_COLS) {
    p.s_col[0] = 0; p.s_col[1] = nk_box;
    p.o_col[0] = 2 * nk_box; p.o_col[1] = 2 * nk_box + ATT_D;
    p.o_alias[0] = 0; p.o_alias[1] = 0;
  } else if (2 * nk_box + ATT_D <= ATT_TMEM_COLS) {  // 192 < nk <= 224, e.g. 197 tokens
    p.s_col[0] = 0; p.s_col[1] = nk_box;
    p.o_col[0] = 2 * nk_box; p.o_col[1] = nk_box + 128;
    p.o_alias[0] = 0; p.o_alias[1] = 1;
  } else {
    p.s_col[0] = 0; p.s_col[1] = 256;
    p.o_col[0] = 128; p.o_col[1] = 256 + 128;
    p.o_alias[0] = 1; p.o_alias[1] = 1;
  }
  const int smem_bytes = 1024 + 2 * (ATT_Q_BYTES + 2 * nk_box * 128) + ATT_OSTAGE_BYTES + ATT_INV_BYTES + 256;
  static SmemAttrCache smem_cache;  // per device; always the kernel's maximum (227 KB)
  {
    const int st = ensure_dyn_smem(attn_fwd_kernel, 232448, smem_cache);
    if (st != DYT_OK) return st;
  }
  const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
  return cuda_status(launch_pdl(attn_fwd_kernel, dim3(grid), dim3(ATT_THREADS), smem_bytes, stream,
                                tq, tkv, p),
                     "attn_fwd_kernel launch");
}

}  // namespace dyt

extern "C" int dyt_attn_varlen_fwd(const void* qkv, int ld_qkv, const int* cu_seqlens, int num_seqs,
                                   int uniform_len, int max_seqlen, int total_tokens, int num_heads,
                                   int head_dim, void* out, int ldo, void* stream) {
  return dyt::attn_varlen_fwd(static_cast<const __half*>(qkv), ld_qkv, cu_seqlens, num_seqs,
                              uniform_len, max_seqlen, total_tokens, num_heads, head_dim,
                              static_cast<__half*>(out), ldo, static_cast<cudaStream_t>(stream));
}
