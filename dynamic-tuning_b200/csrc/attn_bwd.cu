// Attention backward on tcgen05 (sm_100a): dQ, dK, dV of softmax(Q K^T * scale) V per (sequence,
// head), recomputing the probabilities from Q and K instead of saving them (reference
// models/vision_transformer_IN21K.py:61-65: the backward of F.scaled_dot_product_attention as
// engine_finetune.py:47-76 runs it).  Sequences of up to 256 tokens = at most two 128-query tiles.
//
// Persistent kernel, one CTA per SM, unit = (sequence, head).  Q, K, V and dO of the unit are
// TMA-staged once (128-byte-swizzled [tokens, 64] tiles); all five contractions run on tcgen05.mma
// with fp32 accumulators in TMEM, operands addressed by shared-memory descriptors -- the same
// tile is read K-major or MN-major as the product needs it, nothing is transposed in memory:
//   per query tile t (128 rows):
//     S    = Q_t K^T          A = Q_t   (K-major)   B = K     (K-major)   -> TMEM X[0, nk)
//     P    = softmax(S scale)                        math warps: TMEM -> registers -> fp16 smem tile
//     dV  += P^T dO_t         A = P     (MN-major)  B = dO_t  (MN-major)  -> TMEM DV (key tiles 0, 1)
//     dP   = dO_t V^T         A = dO_t  (K-major)   B = V     (K-major)   -> TMEM X[0, nk)  (S is consumed)
//     dS   = P o (dP - D) scale,  D = rowsum(dO o O)  math warps, written over P in shared memory
//     dQ_t = dS K             A = dS    (K-major)   B = K     (MN-major)  -> TMEM X[0, 64)  (dP is consumed)
//     dK  += dS^T Q_t         A = dS    (MN-major)  B = Q_t   (MN-major)  -> TMEM DK
//   TMEM: X 256 columns (S -> dP -> dQ in turn), DV 2 x 64, DK 2 x 64 = 512 columns.
//   warp 0 TMA producer, warp 1 MMA issuer (warp-uniform, one elected lane), warp 2 TMEM allocator,
//   warps 4-19 math: thread = (query row, quarter of the keys), row max / sum / D exchanged between
//   the four parts of a row through shared memory (named barrier per TMEM lane quarter); four warps
//   per SM sub-partition keep MUFU / TMEM-load latencies covered.
// The P / dS tile is [128 rows][256 keys] fp16 in four 64-key blocks of 16 KB (row pitch 128 B,
// 16-byte chunks XOR-swizzled by row & 7 = the SWIZZLE_128B pattern TMA writes), so it is a legal
// K-major A operand (dQ) and, with a leading-dimension byte offset of 16 KB between the 64-key
// blocks, a legal MN-major A operand (dV, dK).
// Replaces the round-1 kernel that ran the contractions on HMMA through nvcuda::wmma (191 us at
// 64 x 12 x 197; sequences <= 208).
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "ptx.cuh"

namespace dyt {

#ifdef DYT_AB_BUILD
__device__ long long* g_bwd_trace = nullptr;   // development builds: clock64 timeline of CTA 0
#define BW_TRACE(slot, tile, ev)                                                          \
  do {                                                                                    \
    if (bw_tr != nullptr && (tile) < 16) bw_tr[((slot) * 16 + (tile)) * 8 + (ev)] = clock64(); \
  } while (0)
#define BW_TRACE_INIT long long* const bw_tr = blockIdx.x == 0 ? g_bwd_trace : nullptr
#else
#define BW_TRACE(slot, tile, ev) do { } while (0)
#define BW_TRACE_INIT do { } while (0)
#endif

struct BwdParams {
  const int* cu_seqlens;  // [B+1] device int32 or nullptr -> uniform_len tokens per sequence
  int uniform_len;
  int nk_box;  // rows of the TMA boxes: round_up(max_seqlen, 16) <= 256
  int kq_bufs; // 2: the next unit's [K Q0] is prefetched (fits up to 208-token boxes); 1: no prefetch
  int C;       // H * 64
  int H;
  int num_units;
  const __half* o;
  int ldo;
  __half* dqkv;
  int ld_dqkv;
  float scale;        // head_dim^-0.5
  float scale_log2e;  // scale * log2(e)
};

constexpr int BW_NP = 4;                   // key parts per query row = math warps per TMEM lane quarter
constexpr int BW_THREADS = 128 + BW_NP * 128;
constexpr int BW_TMEM_COLS = 512;
constexpr int BW_X = 0, BW_DV = 256, BW_DK = 384;
constexpr int BW_BLK = 16384;              // one 64-key block of the P / dS tile: 128 rows x 128 B
constexpr int BW_PTILE = 4 * BW_BLK;
constexpr int BW_XCHG = 3 * BW_NP * 128 * 4;  // row max, row sum, D partial: per key part and row
constexpr int BW_MAX_SMEM = 232448;

// MN-major operand whose MN extent spans several 64-element blocks `lbo_bytes` apart
__device__ __forceinline__ uint64_t umma_desc_sw128_lbo(uint32_t smem_addr_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void bw_unit_span(const BwdParams& p, int unit, int& h, int& start, int& n) {
  const int b = unit / p.H;
  h = unit - b * p.H;
  if (p.cu_seqlens != nullptr) {
    start = __ldg(p.cu_seqlens + b);
    n = __ldg(p.cu_seqlens + b + 1) - start;
  } else {
    start = b * p.uniform_len;
    n = p.uniform_len;
  }
  if (n > p.nk_box) n = p.nk_box;
  if (n < 0) n = 0;
}
__device__ __forceinline__ void bw_named_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// TMEM -> registers of chunk C (32 columns, or the 16-column tail) of a half row of n columns
template <int HC, int C>
__device__ __forceinline__ void bw_load_chunk(uint32_t taddr, int n, uint32_t (&r)[32]) {
  if constexpr (C * 32 + 32 <= HC) {
    if (C * 32 + 32 <= n) {
      tmem_ld32(taddr + C * 32, r);
      return;
    }
  }
  if constexpr (C * 32 + 16 <= HC) {
    if (C * 32 + 16 <= n) tmem_ld16(taddr + C * 32, reinterpret_cast<uint32_t(&)[16]>(r));
  }
}

// my 16 fp32 columns (key part pt) of row rit -> fp16 into a [128 rows][128 B] staging tile, 16-byte
// chunks XOR-swizzled by the row (conflict-free both ways)
__device__ __forceinline__ void bw_stage16(uint32_t tile, int rit, int pt, const uint32_t (&a)[16]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint4 v;
    v.x = pack_half2(__uint_as_float(a[8 * c + 0]), __uint_as_float(a[8 * c + 1]));
    v.y = pack_half2(__uint_as_float(a[8 * c + 2]), __uint_as_float(a[8 * c + 3]));
    v.z = pack_half2(__uint_as_float(a[8 * c + 4]), __uint_as_float(a[8 * c + 5]));
    v.w = pack_half2(__uint_as_float(a[8 * c + 6]), __uint_as_float(a[8 * c + 7]));
    sts128(tile + rit * 128 + (((pt * 2 + c) ^ (rit & 7)) << 4), v);
  }
}
// staged tile -> rows [row0, row0 + 128) of dst (row stride ld halves), rows < nrows only; math warp
// mw (0..15) writes eight rows, every store instruction four full 128-byte rows
__device__ __forceinline__ void bw_flush(uint32_t tile, int mw, int lane, __half* dst, int ld, int row0,
                                         int nrows) {
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int rr = mw * 8 + it * 4 + (lane >> 3);
    const int ch = lane & 7;
    const uint4 v = lds128(tile + rr * 128 + ((ch ^ (rr & 7)) << 4));
    if (row0 + rr < nrows)
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(row0 + rr) * ld + ch * 8) = v;
  }
}

template <int HC>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_kv,   // qkv, box = nk_box rows (K, V)
                const __grid_constant__ CUtensorMap tmap_q0,   // qkv, box = min(128, nk_box) rows
                const __grid_constant__ CUtensorMap tmap_q1,   // qkv, box = nk_box - 128 rows (or unused)
                const __grid_constant__ CUtensorMap tmap_g0,   // d_out, like tmap_q0
                const __grid_constant__ CUtensorMap tmap_g1,   // d_out, like tmap_q1
                const BwdParams p) {
  constexpr int NCH = (HC + 31) / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // shared memory: [K Q0] x 2 (the next unit's K and first query tile are prefetched while the
  // current unit computes) | V | dO0 | Q1 | dO1 | P / dS tile | exchange slots | barriers
  const uint32_t tile_bytes = static_cast<uint32_t>(p.nk_box) * 128u;  // K (or V) of a unit
  const uint32_t t0_bytes = (p.nk_box < 128 ? static_cast<uint32_t>(p.nk_box) : 128u) * 128u;
  const uint32_t t1_bytes = tile_bytes - t0_bytes;        // second query tile (rows 128..)
  const uint32_t kq_bytes = tile_bytes + t0_bytes;
  const uint32_t sKQ = smem_u32(smem);                    // + (unit parity) * kq_bytes: K, then Q0
  const uint32_t sV = sKQ + p.kq_bufs * kq_bytes, sG0 = sV + tile_bytes, sQ1 = sG0 + t0_bytes, sG1 = sQ1 + t1_bytes;
  const uint32_t sP = sG1 + t1_bytes;                     // P / dS tile, 1024-byte aligned
  const uint32_t xch = sP + BW_PTILE;                     // float [3 quantities][BW_NP][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.kq_bufs * kq_bytes + tile_bytes + t0_bytes +
                                               2 * t1_bytes + BW_PTILE + BW_XCHG);
  uint64_t* kq_full = bars + 0;     // [2: bars + 0, bars + 12] TMA -> MMA: K and the first query tile of Q
  uint64_t* vg_full = bars + 10;    // TMA -> MMA: V and the first tile of dO
  uint64_t* t1_full = bars + 11;    // TMA -> MMA: second query tile of Q and dO
  uint64_t* opnd_free = bars + 1;   // MMA (commit) -> TMA
  uint64_t* s_full = bars + 2;      // MMA -> math
  uint64_t* p_ready = bars + 3;     // math -> MMA (P in shared memory)
  uint64_t* dp_full = bars + 4;     // MMA -> math (dV accumulated, dP in TMEM)
  uint64_t* ds_ready = bars + 5;    // math -> MMA (dS in shared memory)
  uint64_t* dq_full = bars + 6;     // MMA -> math (dQ in TMEM; after the last tile also dV, dK)
  uint64_t* x_free = bars + 7;      // math -> MMA (dQ read out)
  uint64_t* acc_free = bars + 8;    // math -> MMA (dV, dK read out)
  uint64_t* kq_full1 = bars + 12;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  BW_TRACE_INIT;

  if (warp_idx == 1 && lane == 0) {
    mbar_init(kq_full, 1);
    mbar_init(kq_full1, 1);
    mbar_init(vg_full, 1);
    mbar_init(t1_full, 1);
    mbar_init(opnd_free, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4 * BW_NP);
    mbar_init(dp_full, 1);
    mbar_init(ds_ready, 4 * BW_NP);
    mbar_init(dq_full, 1);
    mbar_init(x_free, 4 * BW_NP);
    mbar_init(acc_free, 4 * BW_NP);
    fence_mbar_init();
  }
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_q0);
    tma_prefetch_desc(&tmap_g0);
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, BW_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    reg_dealloc<56>();
    if (lane == 0) {
      // arrival groups in the order the MMA warp needs them: (K, Q rows 0-127) for S of the first
      // tile -- loaded one unit AHEAD into the other [K Q0] buffer --, (V, dO rows 0-127) for
      // dV / dP, (Q, dO rows 128-) for the second tile
      const uint32_t nb = static_cast<uint32_t>(p.kq_bufs);
      auto load_kq = [&](int h, int start, uint32_t u) {
        uint64_t* bar = (u % nb) ? kq_full1 : kq_full;
        uint8_t* dst = smem + (u % nb) * kq_bytes;
        mbar_arrive_expect_tx(bar, kq_bytes);
        tma_load_2d(dst, &tmap_kv, bar, p.C + h * 64, start);
        tma_load_2d(dst + tile_bytes, &tmap_q0, bar, h * 64, start);
      };
      auto next_unit = [&](int unit, int& h, int& start, int& n) {   // next non-empty unit or -1
        for (unit += gridDim.x; unit < p.num_units; unit += gridDim.x) {
          bw_unit_span(p, unit, h, start, n);
          if (n > 0) return unit;
        }
        return -1;
      };
      uint32_t uc = 0;
      int h = 0, start = 0, n = 0;
      int unit = next_unit(static_cast<int>(blockIdx.x) - static_cast<int>(gridDim.x), h, start, n);
      if (unit >= 0 && nb == 2) load_kq(h, start, 0);
      uint8_t* const pV = smem + nb * kq_bytes;
      while (unit >= 0) {
        // the single-buffered operands of this unit: free once every MMA of the previous unit retired
        mbar_wait(opnd_free, (uc & 1) ^ 1);
        if (nb == 1) load_kq(h, start, uc);
        mbar_arrive_expect_tx(vg_full, tile_bytes + t0_bytes);
        tma_load_2d(pV, &tmap_kv, vg_full, 2 * p.C + h * 64, start);
        tma_load_2d(pV + tile_bytes, &tmap_g0, vg_full, h * 64, start);
        if (n > 128) {
          mbar_arrive_expect_tx(t1_full, 2 * t1_bytes);
          tma_load_2d(pV + tile_bytes + t0_bytes, &tmap_q1, t1_full, h * 64, start + 128);
          tma_load_2d(pV + tile_bytes + t0_bytes + t1_bytes, &tmap_g1, t1_full, h * 64, start + 128);
        }
        // K and Q0 of the NEXT unit into the other buffer (its previous user, unit uc - 1, is done)
        unit = next_unit(unit, h, start, n);
        ++uc;
        if (unit >= 0 && nb == 2) load_kq(h, start, uc);
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    reg_dealloc<56>();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_kk = 0;  // placeholder to keep the descriptors below in one place
    (void)idesc_kk;
    const uint32_t idesc_acc = umma_idesc_f16(128, 64, 1, 1);  // dV / dK: A and B MN-major
    const uint32_t idesc_dq = umma_idesc_f16(128, 64, 0, 1);   // dQ: A K-major, B MN-major
    uint32_t uc = 0, u2 = 0, tc = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int h, start, n;
      bw_unit_span(p, unit, h, start, n);
      n = __shfl_sync(0xffffffffu, n, 0);
      if (n == 0) continue;
      const int nk = (n + 15) & ~15;
      const int ntiles = n > 128 ? 2 : 1;
      const int nmt = nk > 128 ? 2 : 1;  // 128-key tiles of dV / dK
      const uint32_t idesc_s = umma_idesc_f16(128, nk, 0, 0);
      const uint32_t kb = uc % static_cast<uint32_t>(p.kq_bufs);
      const uint32_t sK = sKQ + kb * kq_bytes;             // this unit's [K Q0] buffer
      for (int t = 0; t < ntiles; ++t, ++tc) {
        const uint32_t sQt = t == 0 ? sK + tile_bytes : sQ1;   // query tile t of Q / dO
        const uint32_t sGt = t == 0 ? sG0 : sG1;
        // ---- S = Q_t K^T ----
        if (t == 0) {
          mbar_wait(kb ? kq_full1 : kq_full, (uc / static_cast<uint32_t>(p.kq_bufs)) & 1);
        } else {
          mbar_wait(t1_full, u2 & 1);
          ++u2;
        }
        mbar_wait(x_free, (tc & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) BW_TRACE(0, tc, 0);
        if (elect_one()) {
          const uint64_t a = umma_desc_sw128(sQt);
          const uint64_t b = umma_desc_sw128(sK);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16)
            umma_ss_f16(tmem_u + BW_X, a + 2 * k16, b + 2 * k16, idesc_s, k16 != 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
        // ---- dV += P^T dO_t ; dP = dO_t V^T ----
        mbar_wait(p_ready, tc & 1);
        if (t == 0) {
          mbar_wait(vg_full, uc & 1);
          mbar_wait(acc_free, (uc & 1) ^ 1);  // dV / dK of the previous unit read out
        }
        tc_fence_after();
        if (lane == 0) BW_TRACE(0, tc, 1);
        if (elect_one()) {
          for (int mt = 0; mt < nmt; ++mt) {
            const uint64_t a = umma_desc_sw128_lbo(sP + mt * 2 * BW_BLK, BW_BLK);
            const uint64_t b = umma_desc_sw128(sGt);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)  // 16 query rows per step: +2048 B in both operands
              umma_ss_f16(tmem_u + BW_DV + mt * 64, a + ks * 128, b + ks * 128, idesc_acc,
                          (t | ks) != 0 ? 1u : 0u);
          }
          const uint64_t a = umma_desc_sw128(sGt);
          const uint64_t b = umma_desc_sw128(sV);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16)
            umma_ss_f16(tmem_u + BW_X, a + 2 * k16, b + 2 * k16, idesc_s, k16 != 0 ? 1u : 0u);
          umma_commit(dp_full);
        }
        __syncwarp();
        // ---- dQ_t = dS K ; dK += dS^T Q_t ----
        mbar_wait(ds_ready, tc & 1);
        tc_fence_after();
        if (lane == 0) BW_TRACE(0, tc, 2);
        if (elect_one()) {
          const uint64_t bk = umma_desc_sw128(sK);
          const int steps = nk >> 4;
          for (int kk = 0; kk < steps; ++kk) {  // 16 keys per step
            const uint64_t a = umma_desc_sw128(sP + (kk >> 2) * BW_BLK + (kk & 3) * 32);
            umma_ss_f16(tmem_u + BW_X, a, bk + kk * 128, idesc_dq, kk != 0 ? 1u : 0u);
          }
          for (int mt = 0; mt < nmt; ++mt) {
            const uint64_t a = umma_desc_sw128_lbo(sP + mt * 2 * BW_BLK, BW_BLK);
            const uint64_t b = umma_desc_sw128(sQt);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_ss_f16(tmem_u + BW_DK + mt * 64, a + ks * 128, b + ks * 128, idesc_acc,
                          (t | ks) != 0 ? 1u : 0u);
          }
          umma_commit(dq_full);
          if (t == ntiles - 1) umma_commit(opnd_free);  // every MMA of the unit has read its operands
        }
        __syncwarp();
      }
      ++uc;
    }
  } else if (warp_idx == 2 || warp_idx == 3) {
    reg_dealloc<56>();
  } else {
    // ===================== math: softmax, dS, read-out =====================
    // setmaxnreg moves registers inside the CTA's own launch allocation (640 threads x 96):
    // 16 x 32 x 104 + 4 x 32 x 56 = 60416 <= 61440
    reg_alloc<104>();
    const int q = warp_idx & 3;             // TMEM lane quarter
    const int pt = (warp_idx - 4) >> 2;     // key part (and output column part)
    const int rit = q * 32 + lane;          // row inside the query tile = TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_x = tmem_base + lane_off + BW_X;
    uint32_t r[NCH][32];
    uint32_t uc = 0, tc = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int h, start, n;
      bw_unit_span(p, unit, h, start, n);
      if (n == 0) continue;
      const int nk = (n + 15) & ~15;
      const int ntiles = n > 128 ? 2 : 1;
      // the nk / 16 key groups are dealt to the BW_NP parts as evenly as possible (197 tokens:
      // 64 + 48 + 48 + 48 keys); my keys: [k0, k0 + nh)
      const int grp = nk >> 4, gbase = grp / BW_NP, grem = grp % BW_NP;
      const int k0 = 16 * (pt * gbase + (pt < grem ? pt : grem));
      const int nh = 16 * (gbase + (pt < grem ? 1 : 0));
      int nvalid = n - k0;
      nvalid = nvalid < 0 ? 0 : (nvalid > nh ? nh : nvalid);
      for (int t = 0; t < ntiles; ++t, ++tc) {
        const int row = t * 128 + rit;       // token of the sequence
        const bool row_ok = row < n;
        // ---- S half row -> P ----
        if (warp_idx == 4 && lane == 0) BW_TRACE(1, tc, 0);
        mbar_wait(s_full, tc & 1);
        tc_fence_after();
        if (warp_idx == 4 && lane == 0) BW_TRACE(1, tc, 1);
        bw_load_chunk<HC, 0>(t_x + k0, nh, r[0]);
        if constexpr (NCH > 1) bw_load_chunk<HC, 1>(t_x + k0, nh, r[1]);
        if constexpr (NCH > 2) bw_load_chunk<HC, 2>(t_x + k0, nh, r[2]);
        if constexpr (NCH > 3) bw_load_chunk<HC, 3>(t_x + k0, nh, r[3]);
        tmem_ld_wait();
        if (warp_idx == 4 && lane == 0) BW_TRACE(2, tc, 0);
        // keys outside the sequence / this half -> -inf (only the last chunks can be affected:
        // warp-uniform tests keep the common chunks free of per-element predicates)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c * 32 + 32 > nvalid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < HC && c * 32 + j >= nvalid) r[c][j] = 0xff800000u;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < HC) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(r[c][j]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // three exchange slots per part and row (max, sum, D): a slot is rewritten only after a
        // barrier of the row's warps has passed since its last read
        sts32(xch + ((0 * BW_NP + pt) * 128 + rit) * 4, __float_as_uint(mx));
        bw_named_sync(1 + q, BW_NP * 32);
        if (warp_idx == 4 && lane == 0) BW_TRACE(2, tc, 1);
        float m = mx;
#pragma unroll
        for (int o = 1; o < BW_NP; ++o)
          m = fmaxf(m, __uint_as_float(lds32(xch + ((0 * BW_NP + ((pt + o) % BW_NP)) * 128 + rit) * 4)));
        const float mb = m * p.scale_log2e;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < HC) {
              const float e = ex2_approx(fmaf(__uint_as_float(r[c][j]), p.scale_log2e, -mb));
              r[c][j] = __float_as_uint(e);
              sum4[j & 3] += e;
            }
        const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
        sts32(xch + ((1 * BW_NP + pt) * 128 + rit) * 4, __float_as_uint(sum));
        bw_named_sync(1 + q, BW_NP * 32);
        if (warp_idx == 4 && lane == 0) BW_TRACE(2, tc, 2);
        float tot = 0.f;    // same summation order in every part: identical P normalisation
#pragma unroll
        for (int o = 0; o < BW_NP; ++o)
          tot += __uint_as_float(lds32(xch + ((1 * BW_NP + o) * 128 + rit) * 4));
        // rows outside the sequence contribute nothing to dV / dK: P = 0
        const float inv = row_ok ? 1.0f / tot : 0.f;
        // P (fp16) -> shared memory tile, 8 keys = one 16-byte chunk at a time
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8)
            if (c * 32 + g8 * 8 < HC) {
              const int kk = k0 + c * 32 + g8 * 8;
              if (c * 32 + g8 * 8 < nh) {
                uint4 v;
                v.x = pack_half2(__uint_as_float(r[c][g8 * 8 + 0]) * inv, __uint_as_float(r[c][g8 * 8 + 1]) * inv);
                v.y = pack_half2(__uint_as_float(r[c][g8 * 8 + 2]) * inv, __uint_as_float(r[c][g8 * 8 + 3]) * inv);
                v.z = pack_half2(__uint_as_float(r[c][g8 * 8 + 4]) * inv, __uint_as_float(r[c][g8 * 8 + 5]) * inv);
                v.w = pack_half2(__uint_as_float(r[c][g8 * 8 + 6]) * inv, __uint_as_float(r[c][g8 * 8 + 7]) * inv);
                sts128(sP + (kk >> 6) * BW_BLK + rit * 128 + ((((kk & 63) >> 3) ^ (rit & 7)) << 4), v);
              }
            }
        if (warp_idx == 4 && lane == 0) BW_TRACE(2, tc, 3);
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the MMA's async-proxy reads
        if (warp_idx == 4 && lane == 0) BW_TRACE(2, tc, 4);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        if (warp_idx == 4 && lane == 0) BW_TRACE(1, tc, 2);
        // ---- dP half row -> dS, written over P ----
        mbar_wait(dp_full, tc & 1);
        tc_fence_after();
        if (warp_idx == 4 && lane == 0) BW_TRACE(1, tc, 3);
        bw_load_chunk<HC, 0>(t_x + k0, nh, r[0]);
        if constexpr (NCH > 1) bw_load_chunk<HC, 1>(t_x + k0, nh, r[1]);
        if constexpr (NCH > 2) bw_load_chunk<HC, 2>(t_x + k0, nh, r[2]);
        if constexpr (NCH > 3) bw_load_chunk<HC, 3>(t_x + k0, nh, r[3]);
        tmem_ld_wait();
        if (warp_idx == 4 && lane == 0) BW_TRACE(3, tc, 0);
        // D = rowsum(dO o O) = sum_j P_ij dP_ij (O = P V, dP = dO V^T): from what is on chip already,
        // no read of O; the two key halves add their partial sums
        float dpart4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8)
            if (c * 32 + g8 * 8 < HC) {
              const int kk = k0 + c * 32 + g8 * 8;
              if (c * 32 + g8 * 8 < nh) {
                const uint4 pv = lds128(sP + (kk >> 6) * BW_BLK + rit * 128 + ((((kk & 63) >> 3) ^ (rit & 7)) << 4));
                const __half2* ph = reinterpret_cast<const __half2*>(&pv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 pf = __half22float2(ph[e]);
                  dpart4[e] = fmaf(pf.x, __uint_as_float(r[c][g8 * 8 + 2 * e]),
                                   fmaf(pf.y, __uint_as_float(r[c][g8 * 8 + 2 * e + 1]), dpart4[e]));
                }
              }
            }
        const float dpart = (dpart4[0] + dpart4[1]) + (dpart4[2] + dpart4[3]);
        sts32(xch + ((2 * BW_NP + pt) * 128 + rit) * 4, __float_as_uint(dpart));
        bw_named_sync(1 + q, BW_NP * 32);
        if (warp_idx == 4 && lane == 0) BW_TRACE(3, tc, 1);
        float dsum = 0.f;
#pragma unroll
        for (int o = 0; o < BW_NP; ++o)
          dsum += __uint_as_float(lds32(xch + ((2 * BW_NP + o) * 128 + rit) * 4));
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8)
            if (c * 32 + g8 * 8 < HC) {
              const int kk = k0 + c * 32 + g8 * 8;
              if (c * 32 + g8 * 8 < nh) {
                const uint32_t addr = sP + (kk >> 6) * BW_BLK + rit * 128 + ((((kk & 63) >> 3) ^ (rit & 7)) << 4);
                const uint4 pv = lds128(addr);
                const __half2* ph = reinterpret_cast<const __half2*>(&pv);
                uint32_t out[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 pf = __half22float2(ph[e]);
                  const float d0 = (__uint_as_float(r[c][g8 * 8 + 2 * e]) - dsum) * p.scale;
                  const float d1 = (__uint_as_float(r[c][g8 * 8 + 2 * e + 1]) - dsum) * p.scale;
                  out[e] = pack_half2(pf.x * d0, pf.y * d1);   // P = 0 outside the sequence -> dS = 0
                }
                sts128(addr, make_uint4(out[0], out[1], out[2], out[3]));
              }
            }
        if (warp_idx == 4 && lane == 0) BW_TRACE(3, tc, 2);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_ready);
        if (warp_idx == 4 && lane == 0) BW_TRACE(1, tc, 4);
        // ---- dQ_t: my 16 of the 64 columns -> fp16 -> smem staging tile (the P / dS tile is free:
        //      dq_full covers the dK MMAs too) -> full 128-byte rows to global.  A thread-per-row
        //      store (32 lanes x 16 B, 4.6 KB apart) costs one L2 transaction per lane.
        mbar_wait(dq_full, tc & 1);
        tc_fence_after();
        if (warp_idx == 4 && lane == 0) BW_TRACE(1, tc, 5);
        __half* const unit_base = p.dqkv + static_cast<size_t>(start) * p.ld_dqkv + h * 64;
        {
          uint32_t a[16];
          tmem_ld16(t_x + pt * 16, a);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(x_free);
          bw_stage16(sP, rit, pt, a);
        }
        bw_named_sync(5, BW_NP * 128);
        bw_flush(sP, warp_idx - 4, lane, unit_base, p.ld_dqkv, t * 128, n);
        bw_named_sync(5, BW_NP * 128);   // the staging tile is rewritten (dK / dV, or P of the next tile)
        if (t == ntiles - 1) {
          // dK and dV of the unit (TMEM lane = key): blocks 0, 1 <- dK key tiles, 2, 3 <- dV key tiles
          const int nmt = nk > 128 ? 2 : 1;
          for (int mt = 0; mt < nmt; ++mt) {
            uint32_t a[16];
            tmem_ld16(tmem_base + lane_off + BW_DK + mt * 64 + pt * 16, a);
            tmem_ld_wait();
            bw_stage16(sP + mt * BW_BLK, rit, pt, a);
            tmem_ld16(tmem_base + lane_off + BW_DV + mt * 64 + pt * 16, a);
            tmem_ld_wait();
            bw_stage16(sP + (2 + mt) * BW_BLK, rit, pt, a);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_free);
          bw_named_sync(5, BW_NP * 128);
          for (int mt = 0; mt < nmt; ++mt) {
            bw_flush(sP + mt * BW_BLK, warp_idx - 4, lane, unit_base + p.C, p.ld_dqkv, mt * 128, n);
            bw_flush(sP + (2 + mt) * BW_BLK, warp_idx - 4, lane, unit_base + 2 * p.C, p.ld_dqkv, mt * 128, n);
          }
          bw_named_sync(5, BW_NP * 128);
        }
      }
      ++uc;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BW_TMEM_COLS);
  }
}

template <int HC>
static int launch_bwd(const CUtensorMap (&tm)[5], const BwdParams& p, int smem_bytes,
                      cudaStream_t stream) {
  static SmemAttrCache smem_cache;
  const int st = ensure_dyn_smem(attn_bwd_kernel<HC>, BW_MAX_SMEM, smem_cache);
  if (st != DYT_OK) return st;
  const int grid = p.num_units < sm_count() ? p.num_units : sm_count();
  attn_bwd_kernel<HC><<<grid, BW_THREADS, smem_bytes, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], p);
  return cuda_status(cudaGetLastError(), "attn_bwd_kernel launch");
}

}  // namespace dyt

using namespace dyt;

#ifdef DYT_AB_BUILD
extern "C" int dyt_debug_bwd_trace(void* dev_buf) {
  long long* ptr = static_cast<long long*>(dev_buf);
  return static_cast<int>(cudaMemcpyToSymbol(dyt::g_bwd_trace, &ptr, sizeof(ptr)));
}
#endif

extern "C" int dyt_attn_varlen_bwd(const void* qkv, int ld_qkv, const void* out, int ldo,
                                   const void* d_out, int ld_do, const int* cu_seqlens,
                                   int num_seqs, int uniform_len, int max_seqlen, int total_tokens,
                                   int num_heads, int head_dim, void* d_qkv, int ld_dqkv,
                                   void* stream) {
  DYT_CHECK_ARG(qkv && out && d_out && d_qkv, "attn_bwd: null buffer");
  DYT_CHECK_ARG(head_dim == 64, "attn_bwd: head_dim must be 64 (got %d)", head_dim);
  DYT_CHECK_ARG(num_seqs >= 0 && num_heads > 0 && total_tokens >= 0, "attn_bwd: bad sizes");
  DYT_CHECK_ARG(cu_seqlens != nullptr || uniform_len > 0, "attn_bwd: need cu_seqlens or uniform_len");
  const int mx = cu_seqlens != nullptr ? max_seqlen : uniform_len;
  DYT_CHECK_ARG(mx >= 1, "attn_bwd: max_seqlen must be >= 1");
  if (mx > 256)
    return fail(DYT_EUNSUPPORTED, "attn_bwd: sequences up to 256 tokens (got %d)", mx);
  const int C = num_heads * head_dim;
  DYT_CHECK_ARG(ld_qkv >= 3 * C && ld_dqkv >= 3 * C && ldo >= C && ld_do >= C &&
                    ld_qkv % 8 == 0 && ld_dqkv % 8 == 0 && ldo % 8 == 0 && ld_do % 8 == 0,
                "attn_bwd: strides must cover the row and be multiples of 8");
  DYT_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) |
                  reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_qkv)) & 15) == 0,
                "attn_bwd: buffers must be 16-byte aligned");
  if (num_seqs == 0 || total_tokens == 0) return DYT_OK;

  const int nk_box = (mx + 15) & ~15;
  // K / V: one box of nk_box rows; Q and dO: rows [0, 128) and [128, nk_box) as separate boxes so
  // that the first query tile can start before the second one has arrived
  const uint32_t rows0 = nk_box < 128 ? nk_box : 128;
  const uint32_t rows1 = nk_box > 128 ? nk_box - 128 : 16;   // (unused when the launch has one tile)
  const uint32_t boxes[5] = {static_cast<uint32_t>(nk_box), rows0, rows1, rows0, rows1};
  CUtensorMap tm[5];
  for (int i = 0; i < 5; ++i) {
    const bool is_do = i >= 3;
    const int s = make_tmap_f16_sw128(&tm[i], is_do ? d_out : qkv, static_cast<uint64_t>(total_tokens),
                                      static_cast<uint64_t>(is_do ? C : 3 * C),
                                      static_cast<uint64_t>(is_do ? ld_do : ld_qkv), boxes[i]);
    if (s != DYT_OK) return s;
  }
  BwdParams p;
  p.cu_seqlens = cu_seqlens;
  p.uniform_len = uniform_len;
  p.nk_box = nk_box;
  p.C = C;
  p.H = num_heads;
  p.num_units = num_seqs * num_heads;
  p.o = static_cast<const __half*>(out);
  p.ldo = ldo;
  p.dqkv = static_cast<__half*>(d_qkv);
  p.ld_dqkv = ld_dqkv;
  p.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  p.scale_log2e = p.scale * 1.4426950408889634f;
  const int t0b = (nk_box < 128 ? nk_box : 128) * 128, tb = nk_box * 128;
  // [K Q0] x kq_bufs + V + dO0 + Q1 + dO1 + P tile + exchange slots + barriers
  auto smem_need = [&](int bufs) {
    return 1024 + bufs * (tb + t0b) + tb + t0b + 2 * (tb - t0b) + BW_PTILE + BW_XCHG + 128;
  };
  p.kq_bufs = smem_need(2) <= BW_MAX_SMEM ? 2 : 1;
  const int smem_bytes = smem_need(p.kq_bufs);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // widest key part = score registers per math thread
  const int hc = 16 * (((nk_box >> 4) + BW_NP - 1) / BW_NP);
  if (hc <= 16) return launch_bwd<16>(tm, p, smem_bytes, st);
  if (hc <= 32) return launch_bwd<32>(tm, p, smem_bytes, st);
  if (hc <= 48) return launch_bwd<48>(tm, p, smem_bytes, st);
  return launch_bwd<64>(tm, p, smem_bytes, st);
}
