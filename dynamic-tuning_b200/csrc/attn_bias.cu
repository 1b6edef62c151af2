// Long-sequence attention with an optional additive per-head bias on tcgen05 (sm_100a):
//   out = softmax(f16(q * scale) k^T + bias[h]) v      per (image, head), any sequence length
// for the segmentation backbone's eager attention path with a relative-position bias
// (reference dense_tasks/Segmentation/backbone/segmentation_vision_transformer_IN21K.py:181-203:
// 1025 tokens at 512 x 512, bias [num_heads, N, N] gathered from relative_position_bias_table).
// Rounding points of that path under fp16 autocast: q * scale and the scores q k^T are fp16
// tensors (scale = 2^-3 commutes with the fp32 accumulation, so the score is f16(acc * scale)), the
// bias add promotes to fp32, softmax runs in fp32, the probabilities are cast to fp16 for the PV
// product (fp32 accumulate, fp16 output).
//
// Flash-style over key tiles of 128, persistent CTAs, unit = (image, head, 128-query tile):
//   warp 0       TMA producer: the unit's Q tile, then K_j / V_j tiles through a 3-stage ring
//   warp 1       tcgen05.mma issuer: S_j = Q K_j^T into one of two TMEM score buffers (S_{j+2} is
//                issued right behind PV_j, so the next scores are ready before the softmax asks),
//                O_part += P_j[:, part] V_j[part] with P read from TMEM (fp16, written over the
//                consumed scores)
//   warp 2       TMEM allocator
//   warps 4-11   softmax: thread = (query row, half of the tile's keys): 64 scores in registers.
//                The two halves of a row are INDEPENDENT flash-decoding streams: own reference max
//                and sum, own accumulator in TMEM (2 x 64 columns), own PV MMAs, own barriers -- no
//                exchange between warps inside the key loop and no CTA-wide barrier anywhere, so
//                the two softmax warps of a sub-partition drift apart and one's exponentials cover
//                the other's TMEM reads.  The reference max of a stream only moves when a tile's
//                max exceeds it by more than 2^8; then the thread rescales its accumulator in TMEM
//                first (skipped warp-wide otherwise: the common case after the first key tile; with
//                a plain running max some row of a warp grows in almost every tile).  The bias
//                values of the next tile are requested one tile ahead.  At the end of the unit the
//                halves are combined exactly, out = (O_0 e^{m_0-m} + O_1 e^{m_1-m}) /
//                (l_0 e^{m_0-m} + l_1 e^{m_1-m}), rounded once to fp16 and stored from registers
//                (64 contiguous bytes per thread).
// Measured on B200 at 16 x 12 x 1025: 136 us without bias (round-1 kernel: HMMA through
// nvcuda::wmma + cp.async, 343 us), 180 us with an fp32 bias whose rows are padded to 16 bytes,
// 289 us with the dense [H, N, N] bias of an odd N (470 us); the steps that led here and the variants
// that were slower (16 softmax warps with a per-tile max exchange: 208 us; four key parts: 154 us)
// are in profiles/r2_attention_long.md.  Also serves sequences beyond the 256 tokens of
// dyt_attn_varlen_fwd.
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "ptx.cuh"

namespace dyt {

struct LongAttnParams {
  int N;          // tokens per sequence (uniform)
  int C;          // H * 64
  int H;
  int q_tiles;    // ceil(N / 128)
  int kv_tiles;   // ceil(N / 128)
  int num_units;  // num_seqs * H * q_tiles
  const float* bias;  // [H, N, ld_bias] fp32 or nullptr
  int ld_bias;        // floats between consecutive rows of the bias (>= N)
  __half* out;
  int ldo;
};

#ifndef DYT_LA_NP
#define DYT_LA_NP 2
#endif
constexpr int LA_NP = DYT_LA_NP;               // key parts per row = softmax warps per lane quarter
constexpr int LA_PK = 128 / LA_NP;             // keys of a tile per thread
constexpr int LA_NH = LA_PK / 32;              // ... in register blocks of 32
constexpr int LA_CW = 64 / LA_NP;              // output columns per thread at the unit end
static_assert(LA_NP == 2 || LA_NP == 4, "two or four key parts (TMEM: 2 x 128 score + LA_NP x 64 accumulator columns)");
// setmaxnreg moves registers inside the launch allocation only: 4 control warps at 56 and
// 4 LA_NP softmax warps must fit threads x compiled registers (640 x 96 / 384 x 168)
constexpr int LA_SOFTMAX_REGS = LA_NP == 4 ? 104 : 208;
constexpr int LA_THREADS = 128 + LA_NP * 128;
constexpr int LA_STAGES = 3;                   // K / V ring
constexpr int LA_TILE = 128 * 128;             // bytes of a [128 x 64] fp16 tile
constexpr int LA_TMEM_COLS = 512;
constexpr int LA_S0 = 0, LA_O = 256;           // two score buffers of 128 columns, one accumulator of 64 per part
constexpr int LA_XCH = 2 * 2 * LA_NP * 128 * 4;  // two sets of (row max, row sum) per part and row (unit end)
constexpr int LA_SMEM = 1024 + LA_TILE * (1 + 2 * LA_STAGES) + LA_XCH + 256;
constexpr float LA_LOG2E = 1.4426950408889634f;
constexpr float LA_RESCALE_LOG2 = 8.0f;        // rescale threshold, log2 units

__device__ __forceinline__ void la_named_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }

// The 32 bias values of (my row, my keys of the tile at column `col`): 16-byte loads when the
// address allows (always, with a bias pitch that is a multiple of 4 floats), scalar ones otherwise.
__device__ __forceinline__ void la_load_bias(float4 (&bv)[8], const float* __restrict__ brow, int col,
                                             int nvalid) {
  const float* bp = brow + col;
  if (brow == nullptr) {   // a row past the sequence inside a live warp
#pragma unroll
    for (int i = 0; i < 8; ++i) bv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (nvalid >= 32 && (reinterpret_cast<uintptr_t>(bp) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(bp) + i);
  } else {   // ragged tail; rows of an odd N are only 4-byte aligned
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      bv[i].x = 4 * i + 0 < nvalid ? __ldg(bp + 4 * i + 0) : 0.f;
      bv[i].y = 4 * i + 1 < nvalid ? __ldg(bp + 4 * i + 1) : 0.f;
      bv[i].z = 4 * i + 2 < nvalid ? __ldg(bp + 4 * i + 2) : 0.f;
      bv[i].w = 4 * i + 3 < nvalid ? __ldg(bp + 4 * i + 3) : 0.f;
    }
  }
}

// fp32 accumulators of q k^T -> scores as the reference's fp16 autocast produces them, in place, and
// their maximum.  With a bias: score = f16(acc) / 8 + bias (fp32).  Without: the function leaves
// f16(acc) and the caller carries the factor 1/8 in its constants.  f16(acc / 8) == f16(acc) / 8 for
// every accumulator whose score is a normal fp16 number (a power-of-two scale commutes with the
// rounding); below 6.1e-5 the two differ by < 2^-25.
template <bool BIAS, bool MASK>
__device__ __forceinline__ float la_scores(uint32_t (&r)[32], const float4 (&bv)[8], int nvalid) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = __half22float2(__floats2half2_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])));
    float s0 = x.x, s1 = x.y;
    if (BIAS) {
      const float4 b4 = bv[i >> 1];
      s0 = fmaf(s0, 0.125f, (i & 1) ? b4.z : b4.x);
      s1 = fmaf(s1, 0.125f, (i & 1) ? b4.w : b4.y);
    }
    if (MASK) {   // keys past the sequence
      if (2 * i >= nvalid) s0 = -INFINITY;
      if (2 * i + 1 >= nvalid) s1 = -INFINITY;
    }
    r[2 * i] = __float_as_uint(s0);
    r[2 * i + 1] = __float_as_uint(s1);
    mx = fmaxf(mx, fmaxf(s0, s1));
  }
  return mx;
}

// p = exp(score - m), packed to fp16 pairs and written to TMEM columns [t_p, t_p + 16) eight at a
// time (keeps the packed values out of the register peak); returns the fp32 sum of the part
template <bool BIAS>
__device__ __forceinline__ float la_exps(const uint32_t (&r)[32], uint32_t t_p, float m) {
  const float c = BIAS ? LA_LOG2E : 0.125f * LA_LOG2E;
  const float mb = m * LA_LOG2E;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float e0 = ex2_approx(fmaf(__uint_as_float(r[16 * hf + 2 * i]), c, -mb));
      const float e1 = ex2_approx(fmaf(__uint_as_float(r[16 * hf + 2 * i + 1]), c, -mb));
      s0 += e0;
      s1 += e1;
      pk[i] = pack_half2(e0, e1);
    }
    tmem_st8(t_p + 8 * hf, pk);
  }
  return s0 + s1;
}

__global__ void __launch_bounds__(LA_THREADS, 1)
attn_long_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const LongAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sKV = sQ + LA_TILE;                        // stage s: K at +2s tiles, V at +2s+1
  const uint32_t xch = sKV + 2 * LA_STAGES * LA_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LA_TILE * (1 + 2 * LA_STAGES) + LA_XCH);
  uint64_t* q_full = bars + 0;
  uint64_t* q_free = bars + 1;     // MMA commit: every S of the unit has read Q
  uint64_t* kv_full = bars + 2;    // [3]
  uint64_t* kv_free = bars + 5;    // [3] MMA commit after PV_j
  uint64_t* s_full = bars + 8;     // [2]
  uint64_t* o_read = bars + 10;    // softmax -> MMA: the accumulators of the unit read out (count 4 LA_NP)
  uint64_t* o_full = bars + 11;    // MMA commit after the unit's last PV
  uint64_t* pv_done = bars + 12;   // [parts] MMA commit after the part's PV_j
  uint64_t* p_ready = bars + 12 + LA_NP;   // [2 buffers][parts] softmax -> MMA (count 4: the part's warps)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12 + 3 * LA_NP);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_free, 1);
    for (int i = 0; i < LA_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < LA_NP; ++i) mbar_init(&pv_done[i], 1);
    for (int i = 0; i < 2 * LA_NP; ++i) mbar_init(&p_ready[i], 4);
    mbar_init(o_read, 4 * LA_NP);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp_idx == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, LA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int T = p.kv_tiles;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    reg_dealloc<56>();
    if (lane == 0) {
      uint32_t uc = 0, g = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
        const int qt = unit % p.q_tiles;
        const int bh = unit / p.q_tiles;
        const int h = bh % p.H, b = bh / p.H;
        const int row0 = b * p.N;
        mbar_wait(q_free, (uc & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, LA_TILE);
        tma_load_2d(smem, &tmap_qkv, q_full, h * 64, row0 + qt * 128);
        for (int j = 0; j < T; ++j, ++g) {
          const uint32_t st = g % LA_STAGES;
          mbar_wait(&kv_free[st], ((g / LA_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], 2 * LA_TILE);
          uint8_t* dst = smem + LA_TILE * (1 + 2 * st);
          tma_load_2d(dst, &tmap_qkv, &kv_full[st], p.C + h * 64, row0 + j * 128);
          tma_load_2d(dst + LA_TILE, &tmap_qkv, &kv_full[st], 2 * p.C + h * 64, row0 + j * 128);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    reg_dealloc<56>();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
    const uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 1);  // B = V is MN-major
    uint32_t uc = 0, g = 0;   // g = key tiles processed so far by this CTA (ring / buffer parities)
    auto issue_s = [&](uint32_t gt) {   // S of global key tile gt into score buffer gt & 1
      const uint32_t st = gt % LA_STAGES;
      mbar_wait(&kv_full[st], (gt / LA_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a = umma_desc_sw128(sQ);
        const uint64_t b = umma_desc_sw128(sKV + 2 * st * LA_TILE);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16)
          umma_ss_f16(tmem_u + LA_S0 + (gt & 1) * 128, a + 2 * k16, b + 2 * k16, idesc_s,
                      k16 != 0 ? 1u : 0u);
        umma_commit(&s_full[gt & 1]);
      }
      __syncwarp();
    };
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
      mbar_wait(q_full, uc & 1);
      issue_s(g);
      if (T > 1) issue_s(g + 1);
      if (T <= 2) {
        if (elect_one()) umma_commit(q_free);
        __syncwarp();
      }
      for (int j = 0; j < T; ++j, ++g) {
        // ---- O_pt (+)= P_j[:, part keys] V_j[part keys]  (each part: own P, own accumulator) ----
        const uint32_t st = g % LA_STAGES;
        const uint64_t v = umma_desc_sw128(sKV + (2 * st + 1) * LA_TILE);
#pragma unroll
        for (int pt = 0; pt < LA_NP; ++pt) {
          mbar_wait(&p_ready[(g & 1) * LA_NP + pt], (g >> 1) & 1);
          if (j == 0 && pt == 0) mbar_wait(o_read, (uc & 1) ^ 1);   // O of the previous unit read out
          tc_fence_after();
          if (elect_one()) {
            // P of the part: 32 columns of packed fp16 at the start of its own score columns
            const uint32_t a_tmem = tmem_u + LA_S0 + (g & 1) * 128 + pt * LA_PK;
#pragma unroll
            for (int kk = 0; kk < LA_PK / 16; ++kk)   // 16 keys per MMA: 8 P columns, 16 V rows = +2048 B
              umma_ts_f16(tmem_u + LA_O + pt * 64, a_tmem + kk * 8, v + (pt * (LA_PK / 16) + kk) * 128,
                          idesc_o, (j | kk) != 0 ? 1u : 0u);
            umma_commit(&pv_done[pt]);
            if (pt == LA_NP - 1) {
              umma_commit(&kv_free[st]);
              if (j == T - 1) umma_commit(o_full);
            }
          }
          __syncwarp();
        }
        // ---- S_{j+2} into the score buffer PV_j has just been issued on (executes after it) ----
        if (j + 2 < T) {
          issue_s(g + 2);
          if (j + 3 == T) {   // last S of the unit: Q is free once it retires
            if (elect_one()) umma_commit(q_free);
            __syncwarp();
          }
        }
      }
    }
  } else if (warp_idx == 2 || warp_idx == 3) {
    reg_dealloc<56>();
  } else {
    // ===================== softmax / correction / output =====================
    reg_alloc<LA_SOFTMAX_REGS>();
    const int q = warp_idx & 3;             // TMEM lane quarter
    const int pt = (warp_idx - 4) >> 2;     // key part: keys [LA_PK pt, LA_PK pt + LA_PK) of every tile
    const int rit = q * 32 + lane;          // row inside the query tile = TMEM lane
    const bool has_bias = p.bias != nullptr;
    const float m_unit = has_bias ? 1.0f : 0.125f;   // la_scores leaves raw (unscaled) scores without bias
    const uint32_t row_slot = xch + rit * 4;         // [m | l][part][row]
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_acc = t_row + LA_O + pt * 64;   // my part's accumulator
    uint64_t* const my_pv_done = &pv_done[pt];
    uint32_t uc = 0, g = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
      const int qt = unit % p.q_tiles;
      const int bh = unit / p.q_tiles;
      const int h = bh % p.H, b = bh / p.H;
      const int row = qt * 128 + rit;       // token of the sequence
      // warp-uniform: a warp whose 32 rows all lie past the sequence (last query tile) only keeps
      // the barrier protocol going; rows past N inside a live warp compute on whatever the TMA
      // delivered (the next image's tokens or zero fill) and are not stored
      const bool rows_live = qt * 128 + q * 32 < p.N;
      const float* brow = has_bias && row < p.N
                              ? p.bias + (static_cast<size_t>(h) * p.N + row) * p.ld_bias + pt * LA_PK
                              : nullptr;
      float m_run = -INFINITY, l_run = 0.f;   // of my part of the keys
      // bias of (my row, my keys) of the coming tile: requested one tile ahead, right after the
      // previous tile's values have been consumed, so the L2 latency sits behind the exponentials
      float4 bv[LA_NH][8];
      if (has_bias && rows_live && pt * LA_PK < p.N) {
#pragma unroll
        for (int hf = 0; hf < LA_NH; ++hf) la_load_bias(bv[hf], brow, 32 * hf, p.N - pt * LA_PK - 32 * hf);
      }
      for (int j = 0; j < T; ++j, ++g) {
        const int key0 = j * 128 + pt * LA_PK;       // first key of my part
        const bool live = rows_live && key0 < p.N;   // warp-uniform
        const int nvalid = p.N - key0;
        const uint32_t t_s = t_row + LA_S0 + (g & 1) * 128 + pt * LA_PK;   // my score columns
        if (live) {
          uint32_t r[LA_NH][32];
          mbar_wait(&s_full[g & 1], (g >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int hf = 0; hf < LA_NH; ++hf) tmem_ld32(t_s + 32 * hf, r[hf]);
          tmem_ld_wait();
          float mx = -INFINITY;
          if (has_bias) {
            if (nvalid >= LA_PK) {
#pragma unroll
              for (int hf = 0; hf < LA_NH; ++hf) mx = fmaxf(mx, la_scores<true, false>(r[hf], bv[hf], 32));
            } else {
#pragma unroll
              for (int hf = 0; hf < LA_NH; ++hf) mx = fmaxf(mx, la_scores<true, true>(r[hf], bv[hf], nvalid - 32 * hf));
            }
            if (nvalid > 128) {   // my keys of the next tile exist
#pragma unroll
              for (int hf = 0; hf < LA_NH; ++hf)
                la_load_bias(bv[hf], brow, (j + 1) * 128 + 32 * hf, nvalid - 128 - 32 * hf);
            }
          } else {
            if (nvalid >= LA_PK) {
#pragma unroll
              for (int hf = 0; hf < LA_NH; ++hf) mx = fmaxf(mx, la_scores<false, false>(r[hf], bv[hf], 32));
            } else {
#pragma unroll
              for (int hf = 0; hf < LA_NH; ++hf) mx = fmaxf(mx, la_scores<false, true>(r[hf], bv[hf], nvalid - 32 * hf));
            }
          }
          // The reference max of the part only moves when the tile's max exceeds it by more than
          // 2^8 (any reference gives the same softmax; p <= 256 is exact enough in fp16 and far from
          // its range): with a plain running max some row of a warp grows in almost every tile and
          // every tile pays the accumulator rescale and its wait for the previous PV.
          const float mx_s = mx * m_unit;
          const float m_new = (mx_s - m_run) * LA_LOG2E > LA_RESCALE_LOG2 ? mx_s : m_run;
          const float alpha = ex2_approx((m_run - m_new) * LA_LOG2E);   // 0 on the first tile (m_run = -inf)
          // P over the first half of my own (consumed) score columns
          float sum = 0.f;
          if (has_bias) {
#pragma unroll
            for (int hf = 0; hf < LA_NH; ++hf) sum += la_exps<true>(r[hf], t_s + 16 * hf, m_new);
          } else {
#pragma unroll
            for (int hf = 0; hf < LA_NH; ++hf) sum += la_exps<false>(r[hf], t_s + 16 * hf, m_new);
          }
          l_run = l_run * alpha + sum;
          // rescale my accumulator when a reference max moved (not on the first tile: PV_0 overwrites
          // it).  PV_{j-2} is complete (S_j was issued behind it), so the parity wait cannot alias.
          if (j > 0 && __any_sync(0xffffffffu, m_new > m_run)) {
            mbar_wait(my_pv_done, (g - 1) & 1);   // PV_{j-1} of my part has accumulated
            tc_fence_after();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t o[32];
              tmem_ld32(t_acc + 32 * hf, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(t_acc + 32 * hf, o);
            }
          }
          m_run = m_new;
          tmem_st_wait();
        } else {
          mbar_wait(&s_full[g & 1], (g >> 1) & 1);   // S_j has been written: P may go over it
          tc_fence_after();
          if (rows_live) {   // my keys lie past the sequence: P = 0 (V there is another image's)
            uint32_t z[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = 0u;
#pragma unroll
            for (int i = 0; i < LA_PK / 16; ++i) tmem_st8(t_s + 8 * i, z);
            tmem_st_wait();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[(g & 1) * LA_NP + pt]);
      }

      // ---- unit end: combine the parts, O / l -> fp16 -> staging tile -> full rows ----
      mbar_wait(o_full, uc & 1);   // the unit's last PV (own barrier: pv_done's parity could alias here)
      tc_fence_after();
      const uint32_t uslot = row_slot + (uc & 1) * (LA_XCH / 2);   // slot sets alternate between units
      sts32(uslot + pt * 512, __float_as_uint(m_run));
      sts32(uslot + (LA_NP + pt) * 512, __float_as_uint(l_run));
      la_named_sync(1 + q, LA_NP * 32);
      if (rows_live) {
        float m = -INFINITY;
#pragma unroll
        for (int o = 0; o < LA_NP; ++o) m = fmaxf(m, __uint_as_float(lds32(uslot + o * 512)));
        float f[LA_NP], den = 0.f;   // e^{m_part - m}: 0 for a part without keys
#pragma unroll
        for (int o = 0; o < LA_NP; ++o) {
          f[o] = ex2_approx((__uint_as_float(lds32(uslot + o * 512)) - m) * LA_LOG2E);
          den = fmaf(__uint_as_float(lds32(uslot + (LA_NP + o) * 512)), f[o], den);
        }
        const float inv = 1.0f / den;
        // my LA_CW output columns of every accumulator
        float y[LA_CW];
#pragma unroll
        for (int i = 0; i < LA_CW; ++i) y[i] = 0.f;
#pragma unroll
        for (int o = 0; o < LA_NP; ++o) {
          uint32_t a[LA_CW];
          tmem_ld_cols(t_row + LA_O + o * 64 + pt * LA_CW, a);
          tmem_ld_wait();
          const float fo = f[o] * inv;
#pragma unroll
          for (int i = 0; i < LA_CW; ++i) y[i] = fmaf(__uint_as_float(a[i]), fo, y[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_read);
        // my LA_CW columns of the row straight from registers: 2 x LA_CW contiguous bytes per
        // thread in 16-byte stores (whole sectors).  No staging tile and no CTA-wide barrier here:
        // the softmax warps stay out of phase across units.
        if (row < p.N) {
          __half* dst = p.out + (static_cast<size_t>(b) * p.N + row) * p.ldo + h * 64 + pt * LA_CW;
#pragma unroll
          for (int c = 0; c < LA_CW / 8; ++c) {
            uint4 v;
            v.x = pack_half2(y[8 * c + 0], y[8 * c + 1]);
            v.y = pack_half2(y[8 * c + 2], y[8 * c + 3]);
            v.z = pack_half2(y[8 * c + 4], y[8 * c + 5]);
            v.w = pack_half2(y[8 * c + 6], y[8 * c + 7]);
            *reinterpret_cast<uint4*>(dst + 8 * c) = v;
          }
        }
      } else {
        if (lane == 0) mbar_arrive(o_read);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, LA_TMEM_COLS);
  }
}

}  // namespace dyt

extern "C" int dyt_attn_bias_fwd(const void* qkv, int ld_qkv, const float* bias, int ld_bias,
                                 int num_seqs, int seq_len, int num_heads, int head_dim, void* out,
                                 int ldo, void* stream) {
  using namespace dyt;
  DYT_CHECK_ARG(qkv != nullptr && out != nullptr, "attn_bias: null buffer");
  DYT_CHECK_ARG(head_dim == 64, "attn_bias: head_dim must be 64 (got %d)", head_dim);
  DYT_CHECK_ARG(num_seqs >= 0 && seq_len >= 1 && num_heads >= 1, "attn_bias: bad sizes");
  const int C = num_heads * head_dim;
  DYT_CHECK_ARG(ld_qkv >= 3 * C && ldo >= C && ld_qkv % 8 == 0 && ldo % 8 == 0,
                "attn_bias: strides must cover the row and be multiples of 8");
  DYT_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "attn_bias: buffers must be 16-byte aligned");
  DYT_CHECK_ARG(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 3) == 0,
                "attn_bias: bias must be 4-byte aligned");
  if (ld_bias == 0) ld_bias = seq_len;
  DYT_CHECK_ARG(bias == nullptr || ld_bias >= seq_len, "attn_bias: ld_bias (%d) < seq_len (%d)", ld_bias,
                seq_len);
  if (num_seqs == 0) return DYT_OK;
  const long total = static_cast<long>(num_seqs) * seq_len;
  DYT_CHECK_ARG(total < (1l << 31), "attn_bias: too many tokens");
  CUtensorMap tq;
  int s = make_tmap_f16_sw128(&tq, qkv, static_cast<uint64_t>(total), static_cast<uint64_t>(3 * C),
                              static_cast<uint64_t>(ld_qkv), 128);
  if (s != DYT_OK) return s;
  LongAttnParams p;
  p.N = seq_len;
  p.C = C;
  p.H = num_heads;
  p.q_tiles = (seq_len + 127) / 128;
  p.kv_tiles = p.q_tiles;
  const long units = static_cast<long>(num_seqs) * num_heads * p.q_tiles;
  DYT_CHECK_ARG(units < (1l << 31), "attn_bias: grid too large");
  p.num_units = static_cast<int>(units);
  p.bias = bias;
  p.ld_bias = ld_bias;
  p.out = static_cast<__half*>(out);
  p.ldo = ldo;
  static SmemAttrCache smem_cache;
  s = ensure_dyn_smem(attn_long_kernel, LA_SMEM, smem_cache);
  if (s != DYT_OK) return s;
  const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
  attn_long_kernel<<<grid, LA_THREADS, LA_SMEM, static_cast<cudaStream_t>(stream)>>>(tq, p);
  return cuda_status(cudaGetLastError(), "attn_long_kernel launch");
}
