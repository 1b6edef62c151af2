// Persistent warp-specialised tcgen05 GEMM for sm_100a:   OUT = epilogue(A[M,K] * B[N,K]^T)
//   A : fp16 row-major activations (K contiguous)        -> TMA, 128B-swizzled K-major smem tiles
//   B : fp16 row-major nn.Linear weight [out,in]=[N,K]    -> TMA, 128B-swizzled K-major smem tiles
//   D : fp32 accumulators in TMEM, two stages of BN columns so the epilogue of tile i overlaps the
//       MMA main loop of tile i+1.
// CTA pairs (clusters of two, tcgen05 cta_group::2) work on one 256 x BN tile: each CTA stages its
// own 128 rows of A and one half of the B tile, the leader CTA issues M = 256 MMAs that read both
// CTAs' shared memory and accumulate each CTA's 128 rows in its own TMEM.  Per 128xBNx64 MMA block
// an SM then receives A + B/2 instead of A + B: with BN = 256 that is 32 KB per 512 tensor-clocks,
// which the ~70 B/clk an SM can take in sustains; a lone CTA (48 KB) is capped at ~2/3 of the
// tensor peak, and multicasting B inside a cluster of two does not help (same bytes per SM).
// Warp roles (384 threads): warp0 = TMA producer, warp1 = MMA issuer (one elected thread),
// warp2 = TMEM allocator, warps4.. = epilogue: two or four warps per TMEM lane quarter, each
// owning a slice of the tile's columns.  The epilogue arithmetic (bias / GELU / ReLU / scale, one rounding to
// fp16 per Linear output) runs in the TMEM register layout (thread = row, 32 columns); only the
// fp16 result goes through a bank-conflict-free smem transpose (64-byte rows, 16-byte chunks XORed
// with (row>>1)&3) to reach coalesced 128-bit global accesses.  The main loop's operand reads
// already use about half of the shared-memory pipe, so the epilogue must stay light on it: the
// earlier fp32 transpose cost 3x the wavefronts and made the kernel shared-memory bound.
//
// Replaces the cuBLASLt calls behind nn.Linear in the reference block
// (reference models/model_speed_test.py:147 qkv, :164 proj, :106-111 adapter, timm Mlp fc1/fc2).
#pragma once
#include "gelu.cuh"
#include "ptx.cuh"

namespace dyt {

enum GemmEpilogue : int {
  EPI_BIAS = 0,        // out_h = f16(acc + bias); if scale != 1: out_h = f16(out_h * scale)
  EPI_BIAS_GELU = 1,   // out_h = f16(gelu_erf(f16(acc + bias)))
  EPI_BIAS_RELU = 2,   // out_h = f16(relu(f16(acc + bias)))
  EPI_BIAS_RESID = 3,  // v = f16(acc + bias); if scale != 1: v = f16(v * scale);
                       // out_f = resid + v   (fp32 residual stream); optional out_h = f16(out_f)
  EPI_BIAS_GELU_KEEP = 4,  // like EPI_BIAS_GELU, and aux = f16(acc + bias): the pre-activation the
                           // backward needs (train-mode fc1)
  EPI_DGELU = 5,       // out_h = f16(f16(acc + bias) * gelu'(aux)): fc2 data gradient times the
                       // GELU derivative at the saved pre-activation
};

struct GemmParams {
  int M;             // row capacity of A / OUT
  int N;             // output features (multiple of 8)
  int K;             // reduction length (multiple of 8)
  const int* m_dev;  // optional: device-resident effective row count (<= M); nullptr -> M
  const __half* bias;  // [N] fp16 or nullptr
  __half* out_h;
  float* out_f;
  const float* resid;
  int ldo_h;   // row stride (elements) of out_h
  int ldo_f;   // row stride of out_f
  int ld_res;  // row stride of resid
  float scale;
  int vec8;    // 1: out_h rows are 16-byte aligned and N % 8 == 0 -> 128-bit fp16 stores
  // EPI_BIAS_RESID only, optional: per-row dot products <out row, dot_w> accumulated over each
  // epilogue warp's column slice (the token selector's score Linear fused into the proj GEMM,
  // reference models/dynamic_adapter.py:70-72).  dot_out[row * dot_ld + slice], slice =
  // n_tile * PARTS + part; dot_f16 = 1 rounds the row values to fp16 first (autocast policy).
  const float* dot_w;  // [N] fp32 or nullptr
  float* dot_out;
  int dot_ld;
  int dot_f16;
  __half* aux;   // EPI_BIAS_GELU_KEEP: written; EPI_DGELU: read ([M, N] fp16, row stride ld_aux)
  int ld_aux;
  int tail_split;  // 1 (BN = 256 only): tiles of the last, partial round are cut into 2 or 4 column
                   // sub-tiles when that lets every cluster take one (wave quantisation)
  int half_grid;   // launcher only: 1 = at most half of the clusters (a GEMM that shares the machine)
  int reverse_m;   // 1: row pairs are taken from the last to the first (the rows the producing kernel
                   // wrote last are still in the L2 when this kernel starts; the results are the same)
};

// One work item of the persistent tile loop: a full BM x BN tile pair, or -- in the last round,
// when the leftover tiles number at most half / a quarter of the clusters -- a column sub-tile of
// width BN / 2 or BN / 4 (own TMA box for B, narrower MMA, fewer epilogue chunks).  With e.g. 297
// tile pairs on 74 clusters the fifth round held ONE tile and cost a whole round; cut in four it
// costs the A-bound time of a 64-wide tile.
struct GemmItems {
  int n_tiles, full, split, total, last_pair;   // last_pair >= 0: row pairs in descending order
  __device__ __forceinline__ void init(int total_tiles, int n_tiles_, int clusters, int allow,
                                       int reverse = 0) {
    n_tiles = n_tiles_;
    last_pair = reverse ? total_tiles / n_tiles_ - 1 : -1;
    full = (total_tiles / clusters) * clusters;
    const int rem = total_tiles - full;
    split = 1;
    if (allow && rem > 0) {
      if (rem * 4 <= clusters) split = 4;
      else if (rem * 2 <= clusters) split = 2;
    }
    total = full + rem * split;
  }
  // item -> row pair index, first column, width
  __device__ __forceinline__ void decode(int item, int bn_full, int& m_pair, int& n0, int& bn) const {
    int tile = item, sub = 0;
    bn = bn_full;
    if (item >= full && split > 1) {
      tile = full + (item - full) / split;
      sub = (item - full) % split;
      bn = bn_full / split;
    }
    m_pair = tile / n_tiles;
    n0 = (tile % n_tiles) * bn_full + sub * bn;
    if (last_pair >= 0) m_pair = last_pair - m_pair;
  }
};

// EW = epilogue warps: 8 (two per TMEM lane quarter) or 16 (four per quarter, for epilogue-bound
// shapes: GELU, K <= 128 -- the arithmetic epilogue is latency-bound and wants more warps in flight)
// RS (residual staging, EPI_BIAS_RESID on full 256-wide tiles only): every epilogue lane owns a
// 128-byte slot in shared memory that cp.async fills with its piece of the residual one chunk
// ahead (one pipeline stage is given up for the 32 KB); without it the residual loads of a chunk
// are issued at the top of that chunk and their latency is only partly covered.
template <int BN, int EW = 8, bool RS = false>
struct GemmCfg {
  static_assert(BN == 64 || BN == 128 || BN == 192 || BN == 256, "tile widths: 64 / 128 / 192 / 256");
  static_assert(EW == 8 || (EW == 16 && (BN == 128 || BN == 256)),
                "epilogue warps: 8, or 16 for BN = 128 / 256 (column slices of 32-column chunks)");
  static constexpr int BM = 128;
  static constexpr int BK = 64;  // 64 halves = 128 B = one swizzle row
  static_assert(!RS || (BN == 256 && EW == 8), "residual staging: the 256-wide tile with 8 epilogue warps");
  static constexpr int STAGES =
      (BN == 256) ? ((EW == 16 || RS) ? 5 : 6) : (BN == 192 ? 7 : (BN == 128 && EW == 16 ? 7 : 8));
  static constexpr int A_BYTES = BM * BK * 2;         // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * BK * 2;   // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = EW;
  static constexpr int PARTS = EW / 4;          // column slices of the tile, one per warp of a quarter
  static constexpr int PART_COLS = BN / PARTS;  // columns owned by one epilogue warp
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
  static constexpr int SLAB_BYTES = 32 * 64;   // 32 rows x 32 fp16 (one 32-column chunk), per warp
  static constexpr int BIAS_BYTES = 2 * PART_COLS * 4;  // fp32 bias of the warp's columns + the fused
                                                        // row-dot's weights of the same columns
  static constexpr int RSTAGE_BYTES = RS ? 32 * 128 : 0;   // per warp: 32 rows x 32 fp32 of the residual
  static constexpr int EPI_BYTES = EPI_WARPS * (SLAB_BYTES + BIAS_BYTES + RSTAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);  // power of two
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
};

template <int BN, int EPI, int EW, bool RS = false>
__global__ void __launch_bounds__(GemmCfg<BN, EW, RS>::THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a,
               const __grid_constant__ CUtensorMap tmap_b,    // box = BN / 2 rows of W per CTA
               const __grid_constant__ CUtensorMap tmap_b2,   // box = BN / 4 rows (half-width sub-tiles)
               const __grid_constant__ CUtensorMap tmap_b4,   // box = BN / 8 rows (quarter-width)
               const GemmParams p) {
  using Cfg = GemmCfg<BN, EW, RS>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // used in the leader CTA: its producer's expect_tx arrival
      mbar_init(&empty_bar[s], 1);  // the leader's commit, multicast to both CTAs
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);                       // the leader's commit, multicast
      mbar_init(&tmem_empty_bar[a], 2 * Cfg::EPI_WARPS);     // leader CTA: epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_cg2(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // everything above overlapped the previous kernel's tail; its results (A, the residual, the
  // device-side row count) are read from here on
  pdl_wait();
  int m_eff = p.M;
  if (p.m_dev != nullptr) {
    int md = *p.m_dev;
    m_eff = md < p.M ? md : p.M;
  }
  const int m_tiles = (m_eff + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  // work item = a pair of vertically adjacent tiles, one per CTA of the cluster (the odd last tile
  // row pairs with an all-masked dummy: its TMA boxes are out of bounds and read as zeros)
  const int total_tiles = ((m_tiles + 1) / 2) * n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  GemmItems items;
  items.init(total_tiles, n_tiles, num_clusters, BN == 256 ? p.tail_split : 0, p.reverse_m);

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < items.total; item += num_clusters) {
        int m_pair, n0, bn;
        items.decode(item, BN, m_pair, n0, bn);
        const int m0 = (m_pair * 2 + cta_rank) * BM;
        const CUtensorMap* tb = bn == BN ? &tmap_b : (bn * 2 == BN ? &tmap_b2 : &tmap_b4);
        const uint32_t tx_bytes = 2u * (Cfg::A_BYTES + (bn / 2) * BK * 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          // all four boxes of the pair (2 x A, 2 x B half) complete on the leader's barrier
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_cg2(sa, &tmap_a, lead_bar, kb * BK, m0);
          tma_load_2d_cg2(sb, tb, lead_bar, kb * BK, n0 + cta_rank * (bn / 2));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the pipeline with warp-uniform state (descriptors in uniform registers)
    // and one elected lane issues; issuing from inside an `if (lane == 0)` region makes the
    // compiler wrap every tcgen05.mma in a lane-serialising loop (~100 clk of scalar code each).
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const int items_u = __shfl_sync(0xffffffffu, items.total, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; cta_rank == 0 && item < items_u; item += num_clusters) {
      int m_pair, n0, bn;
      items.decode(item, BN, m_pair, n0, bn);
      const uint32_t idesc = umma_idesc_f16(2 * BM, bn, 0, 0);  // M = 256 over the CTA pair
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u + stage * Cfg::STAGE_BYTES;
        const uint64_t a_desc = umma_desc_sw128(sa);
        const uint64_t b_desc = umma_desc_sw128(sa + Cfg::A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 halves = 32 B along K inside the swizzle row: +2 in 16-byte units
            umma_ss_f16_cg2(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_cg2_mc(&empty_bar[stage], 0x3);  // to both producers, once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit_cg2_mc(&tmem_full_bar[acc], 0x3);  // both CTAs' epilogues
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    constexpr int HALF = Cfg::PART_COLS;  // columns owned by this warp
    constexpr int NCH = HALF / 32;        // 32-column chunks per warp per tile
    const int e = warp_idx - 4;
    const int q = e & 3;              // == warp_idx % 4: TMEM lane quarter this warp may access
    const int half = e >> 2;          // which column slice of the tile
    const uint32_t slab = smem_u32(epi_smem) + e * (Cfg::SLAB_BYTES + Cfg::BIAS_BYTES + Cfg::RSTAGE_BYTES);
    const uint32_t bias_s = slab + Cfg::SLAB_BYTES;
    const uint32_t my_row = slab + lane * 64;
    const int swz_w = (lane >> 1) & 3;  // chunk swizzle of the row this lane writes
    int acc = 0;
    uint32_t acc_phase = 0;
    // RS: this lane's staging slots (row it*4 + lane/8, 16 bytes) and the asynchronous fetch of the
    // residual piece of (tile rows m0.., 32 columns from column cfirst)
    const uint32_t rstage = slab + Cfg::SLAB_BYTES + Cfg::BIAS_BYTES + (lane >> 3) * 128 + (lane & 7) * 16;
    auto rs_issue = [&](int m0_, int cfirst) {
      if constexpr (RS) {
        const int col = cfirst + (lane & 7) * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          int grow = m0_ + q * 32 + it * 4 + (lane >> 3);
          const bool in = grow < m_eff && col < p.N;
          grow = in ? grow : 0;
          const float* src = p.resid + static_cast<size_t>(grow) * p.ld_res + (in ? col : 0);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(rstage + it * 512), "l"(src),
                       "r"(in ? 16u : 0u)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    };
    if constexpr (RS) {
      if (cluster_id < items.total) {
        int mp, nt, b0;
        items.decode(cluster_id, BN, mp, nt, b0);
        rs_issue((mp * 2 + cta_rank) * BM, nt + half * (b0 >> 5) / Cfg::PARTS * 32);
      }
    }
    for (int item = cluster_id; item < items.total; item += num_clusters) {
      int m_pair, n_tile0, bn;
      items.decode(item, BN, m_pair, n_tile0, bn);
      const int m0 = (m_pair * 2 + cta_rank) * BM;
      // this warp's 32-column chunks of the item: NCH of a full tile, fewer (or none) of a sub-tile
      const int chunks_total = bn >> 5;
      int cpp = chunks_total / Cfg::PARTS;
      if (cpp == 0) cpp = 1;
      const int c_first = half * cpp;
      const int my_n = c_first < chunks_total ? cpp : 0;
      const int n0 = n_tile0 + c_first * 32;

      // fp32 bias of this warp's columns -> smem (read back as warp-wide broadcasts); in flight
      // while the main loop of this tile finishes
      for (int i = lane; i < my_n * 8; i += 32) {
        const int col = n0 + i * 4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr && col < p.N) {
          const uint2 bb = *reinterpret_cast<const uint2*>(p.bias + col);
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&bb.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&bb.y));
          bv = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        sts128(bias_s + i * 16, make_uint4(__float_as_uint(bv.x), __float_as_uint(bv.y),
                                           __float_as_uint(bv.z), __float_as_uint(bv.w)));
      }
      if constexpr (EPI == EPI_BIAS_RESID) {
        // the fused row-dot's weights of this warp's columns (rounded as the selector Linear sees
        // them) next to the bias: read per chunk straight from global memory they cost an exposed
        // L2 round trip per chunk (13 % of the proj GEMM's stall samples)
        if (p.dot_w != nullptr) {
          for (int i = lane; i < my_n * 8; i += 32) {
            const int col = n0 + i * 4;
            float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col < p.N) {
              dv = *reinterpret_cast<const float4*>(p.dot_w + col);
              if (p.dot_f16)
                dv = make_float4(round_f16(dv.x), round_f16(dv.y), round_f16(dv.z), round_f16(dv.w));
            }
            sts128(bias_s + HALF * 4 + i * 16, make_uint4(__float_as_uint(dv.x), __float_as_uint(dv.y),
                                                          __float_as_uint(dv.z), __float_as_uint(dv.w)));
          }
        }
      }
      __syncwarp();

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc * BN + c_first * 32);
      float dot_acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // RESID row-dot partials
      if (my_n == 0) {  // sub-tile narrower than the warps' column split: only the hand-back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0));
      }
#pragma unroll 1
      for (int c = 0; c < my_n; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        // residual rows of this chunk (coalesced layout, see below): in flight during the math
        float4 res[8];
        if constexpr (EPI == EPI_BIAS_RESID && !RS) {
          const int col = n0 + c * 32 + (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int grow = m0 + q * 32 + it * 4 + (lane >> 3);
            res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < m_eff && col < p.N)
              res[it] = *reinterpret_cast<const float4*>(p.resid +
                                                         static_cast<size_t>(grow) * p.ld_res + col);
          }
        }
        // bias of the chunk's 32 columns: warp-wide broadcast loads, in flight with the TMEM load
        uint4 bq[8];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) bq[j4] = lds128(bias_s + (c * 32 + j4 * 4) * 4);
        // EPI_DGELU: the saved pre-activations of this thread's row (TMEM layout: 32 columns = 64 B)
        uint4 ax[4];
        if constexpr (EPI == EPI_DGELU) {
          const int grow = m0 + q * 32 + lane;
          const int col = n0 + c * 32;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ax[k] = make_uint4(0, 0, 0, 0);
            if (grow < m_eff && col + k * 8 < p.N)
              ax[k] = *reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(grow) * p.ld_aux +
                                                      col + k * 8);
          }
        }
        tmem_ld_wait();
        if (c == my_n - 1) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty_bar[acc]), 0));
        }
        // ---- arithmetic in the TMEM layout: this thread = row (q*32 + lane), 32 columns ----
        // Per pair of columns: one F2FP rounds both Linear outputs to fp16 (the single rounding of
        // the autocast Linear); activations that need the rounded value unpack it again.
        uint32_t pk[16];
        uint32_t keep[EPI == EPI_BIAS_GELU_KEEP ? 16 : 1];
        const bool plain = (EPI == EPI_BIAS || EPI == EPI_BIAS_RESID) && p.scale == 1.0f;
#pragma unroll
        for (int j2 = 0; j2 < 16; ++j2) {
          const uint4 bv = bq[j2 >> 1];
          const float b0 = __uint_as_float((j2 & 1) ? bv.z : bv.x);
          const float b1 = __uint_as_float((j2 & 1) ? bv.w : bv.y);
          const __half2 h = __floats2half2_rn(__uint_as_float(r[2 * j2]) + b0,
                                              __uint_as_float(r[2 * j2 + 1]) + b1);
          __half2 o = h;
          if constexpr (EPI == EPI_BIAS_RELU) {
            o = __hmax2(h, __float2half2_rn(0.f));
          } else if constexpr (EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_GELU_KEEP) {
            const float2 gl = gelu_f16_x2(__low2float(h), __high2float(h));
            o = __floats2half2_rn(gl.x, gl.y);
            if constexpr (EPI == EPI_BIAS_GELU_KEEP) keep[j2] = *reinterpret_cast<const uint32_t*>(&h);
          } else if constexpr (EPI == EPI_DGELU) {
            const uint4 av = ax[j2 >> 2];
            const uint32_t aw = (j2 & 3) == 0 ? av.x : ((j2 & 3) == 1 ? av.y : ((j2 & 3) == 2 ? av.z : av.w));
            const float2 pre = __half22float2(*reinterpret_cast<const __half2*>(&aw));
            o = __floats2half2_rn(__low2float(h) * gelu_grad_f16(pre.x),
                                  __high2float(h) * gelu_grad_f16(pre.y));
          } else {
            if (!plain)  // f16(f16(v) * scale), the multiply in fp32 like torch's fp16-tensor * float
              o = __floats2half2_rn(__low2float(h) * p.scale, __high2float(h) * p.scale);
          }
          pk[j2] = *reinterpret_cast<const uint32_t*>(&o);
        }
        // ---- fp16 transpose through the slab ----
        auto slab_write = [&](const uint32_t* v) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            sts128(my_row + ((k ^ swz_w) << 4),
                   make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
          __syncwarp();
        };
        // lane -> (row it*8 + lane/4, columns (lane%4)*8 .. +7): 64-byte fp16 row segments
        auto slab_to_global = [&](__half* base, int ld) {
          const int col = n0 + c * 32 + (lane & 3) * 8;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + (lane >> 2);
            const int grow = m0 + q * 32 + rr;
            const uint4 hv = lds128(slab + rr * 64 + (((lane & 3) ^ ((rr >> 1) & 3)) << 4));
            if (grow < m_eff && col < p.N) {
              __half* dst = base + static_cast<size_t>(grow) * ld + col;
              if (p.vec8) {
                *reinterpret_cast<uint4*>(dst) = hv;
              } else {  // row pitch / column count only 8-byte aligned
                *reinterpret_cast<uint2*>(dst) = make_uint2(hv.x, hv.y);
                if (col + 4 < p.N) *reinterpret_cast<uint2*>(dst + 4) = make_uint2(hv.z, hv.w);
              }
            }
          }
        };
        slab_write(pk);
        if constexpr (RS) {
          // the staged residual of this chunk -> registers; the slots are refilled at once with the
          // next chunk's piece (of this tile, or the first chunk of this warp's next tile)
          asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const uint4 rv = lds128(rstage + it * 512);
            res[it] = make_float4(__uint_as_float(rv.x), __uint_as_float(rv.y), __uint_as_float(rv.z),
                                  __uint_as_float(rv.w));
          }
          if (c + 1 < my_n) {
            rs_issue(m0, n0 + (c + 1) * 32);
          } else if (item + num_clusters < items.total) {
            int mp, nt, b1;
            items.decode(item + num_clusters, BN, mp, nt, b1);
            rs_issue((mp * 2 + cta_rank) * BM, nt + c_first * 32);
          }
        }
        if constexpr (EPI == EPI_BIAS_RESID) {
          // lane -> (row it*4 + lane/8, columns (lane%8)*4 .. +3): 128-byte fp32 row segments
          const int col = n0 + c * 32 + (lane & 7) * 4;
          const int p8 = lane & 7;
          float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.dot_w != nullptr) {
            const uint4 dq = lds128(bias_s + HALF * 4 + (c * 32 + p8 * 4) * 4);
            dw = make_float4(__uint_as_float(dq.x), __uint_as_float(dq.y), __uint_as_float(dq.z),
                             __uint_as_float(dq.w));
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3);
            const int grow = m0 + q * 32 + rr;
            uint2 hv;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                         : "=r"(hv.x), "=r"(hv.y)
                         : "r"(slab + rr * 64 + (((p8 >> 1) ^ ((rr >> 1) & 3)) << 4) + (p8 & 1) * 8)
                         : "memory");
            if (grow < m_eff && col < p.N) {
              const float2 v01 = __half22float2(*reinterpret_cast<const __half2*>(&hv.x));
              const float2 v23 = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
              const float4 o = make_float4(res[it].x + v01.x, res[it].y + v01.y, res[it].z + v23.x,
                                           res[it].w + v23.y);
              *reinterpret_cast<float4*>(p.out_f + static_cast<size_t>(grow) * p.ldo_f + col) = o;
              const uint2 oh = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
              if (p.out_h != nullptr)
                *reinterpret_cast<uint2*>(p.out_h + static_cast<size_t>(grow) * p.ldo_h + col) = oh;
              if (p.dot_w != nullptr) {
                float4 s = o;
                if (p.dot_f16) {
                  const float2 s01 = __half22float2(*reinterpret_cast<const __half2*>(&oh.x));
                  const float2 s23 = __half22float2(*reinterpret_cast<const __half2*>(&oh.y));
                  s = make_float4(s01.x, s01.y, s23.x, s23.y);
                }
                dot_acc[it] = fmaf(s.x, dw.x, fmaf(s.y, dw.y, fmaf(s.z, dw.z, fmaf(s.w, dw.w, dot_acc[it]))));
              }
            }
          }
        } else {
          slab_to_global(p.out_h, p.ldo_h);
          if constexpr (EPI == EPI_BIAS_GELU_KEEP) {
            __syncwarp();  // every lane has read the activations before the slab is reused
            slab_write(keep);
            slab_to_global(p.aux, p.ld_aux);
          }
        }
        __syncwarp();  // the slab is overwritten by the next chunk
      }
      if constexpr (EPI == EPI_BIAS_RESID) {
        if (p.dot_w != nullptr) {
          // the eight lanes that share a row add up their column partials, lane%8 == 0 stores
          const int slice = (n_tile0 / BN) * Cfg::PARTS + half;  // (full tiles only: no split with dot_w)
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            float v = dot_acc[it];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            const int grow = m0 + q * 32 + it * 4 + (lane >> 3);
            if ((lane & 7) == 0 && grow < m_eff)
              p.dot_out[static_cast<size_t>(grow) * p.dot_ld + slice] = v;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA exits while its peer can still multicast into it
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dyt
