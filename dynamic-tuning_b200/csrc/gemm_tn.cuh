// Persistent warp-specialised tcgen05 GEMM for sm_100a:   OUT = epilogue(A[M,K] * B[N,K]^T)
//   A : fp16 row-major activations (K contiguous)        -> TMA, 128B-swizzled K-major smem tiles
//   B : fp16 row-major nn.Linear weight [out,in]=[N,K]    -> TMA, 128B-swizzled K-major smem tiles
//   D : fp32 accumulators in TMEM, two stages of BN columns so the epilogue of tile i overlaps the
//       MMA main loop of tile i+1.
// Warp roles (384 threads): warp0 = TMA producer, warp1 = MMA issuer (one elected thread),
// warp2 = TMEM allocator, warps4-11 = epilogue: two warps per TMEM lane quarter, each owning half
// of the tile's columns (TMEM -> regs -> swizzled smem transpose -> coalesced 128-bit global
// accesses, with bias / GELU / ReLU / scale / fp32-residual fused).  Bias and residual loads are
// issued before the accumulator wait so their latency overlaps the main loop.
//
// Replaces the cuBLASLt calls behind nn.Linear in the reference block
// (reference models/model_speed_test.py:147 qkv, :164 proj, :106-111 adapter, timm Mlp fc1/fc2).
#pragma once
#include "ptx.cuh"

namespace dyt {

enum GemmEpilogue : int {
  EPI_BIAS = 0,        // out_h = f16(acc + bias); if scale != 1: out_h = f16(out_h * scale)
  EPI_BIAS_GELU = 1,   // out_h = f16(gelu_erf(f16(acc + bias)))
  EPI_BIAS_RELU = 2,   // out_h = f16(relu(f16(acc + bias)))
  EPI_BIAS_RESID = 3,  // v = f16(acc + bias); if scale != 1: v = f16(v * scale);
                       // out_f = resid + v   (fp32 residual stream); optional out_h = f16(out_f)
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
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;  // 64 halves = 128 B = one swizzle row
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
  static constexpr int SLAB_BYTES = 32 * 128;  // 32 rows x 32 fp32, one per epilogue warp
  static constexpr int EPI_BYTES = EPI_WARPS * SLAB_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // 128 / 256 / 512: all powers of two
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
};

// exact-erf GELU, branch-free: erf via Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7, far below
// the fp16 rounding applied to the result), 2 MUFU (rcp, ex2) + ~12 FMA per element.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p, e, 1.0f);          // erf(|x|/sqrt2)
  const float erf_s = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_s);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GemmCfg<BN>::THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a,
               const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
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

  int m_eff = p.M;
  if (p.m_dev != nullptr) {
    int md = *p.m_dev;
    m_eff = md < p.M ? md : p.M;
  }
  const int m_tiles = (m_eff + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], Cfg::EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
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
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN, 0, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const int tiles_u = __shfl_sync(0xffffffffu, total_tiles, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < tiles_u; tile += gridDim.x) {
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
            umma_ss_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);  // accumulator ready for the epilogue
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    constexpr int HALF = BN / 2;      // columns owned by this warp
    constexpr int NCH = HALF / 32;    // 32-column chunks per warp per tile
    const int e = warp_idx - 4;
    const int q = e & 3;              // == warp_idx % 4: TMEM lane quarter this warp may access
    const int half = e >> 2;
    const uint32_t slab = smem_u32(epi_smem) + e * Cfg::SLAB_BYTES;
    const int cl = lane & 7;     // 16-byte chunk (4 fp32 columns) within the 32-column slab row
    const int rsub = lane >> 3;  // row within a group of 4
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN + half * HALF;
      const int row_base = m0 + q * 32 + rsub;

      // bias for every chunk of this tile: issued before the accumulator is ready
      float bias_r[NCH][4];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = n0 + c * 32 + cl * 4;
        bias_r[c][0] = bias_r[c][1] = bias_r[c][2] = bias_r[c][3] = 0.f;
        if (p.bias != nullptr && col < p.N) {
          const uint2 bb = *reinterpret_cast<const uint2*>(p.bias + col);
          const __half2 h01 = *reinterpret_cast<const __half2*>(&bb.x);
          const __half2 h23 = *reinterpret_cast<const __half2*>(&bb.y);
          bias_r[c][0] = __low2float(h01);
          bias_r[c][1] = __high2float(h01);
          bias_r[c][2] = __low2float(h23);
          bias_r[c][3] = __high2float(h23);
        }
      }
      // residual of the first chunk: in flight while the main loop of this tile finishes
      float4 res[8];
      if constexpr (EPI == EPI_BIAS_RESID) {
        const int col = n0 + cl * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = row_base + it * 4;
          res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow < m_eff && col < p.N)
            res[it] = *reinterpret_cast<const float4*>(p.resid +
                                                       static_cast<size_t>(grow) * p.ld_res + col);
        }
      }

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc * BN + half * HALF);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        if (c == NCH - 1) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        {
          const uint32_t rowp = slab + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(rowp + ((j ^ (lane & 7)) << 4),
                   make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
        }
        __syncwarp();
        const int col = n0 + c * 32 + cl * 4;
        const bool col_ok = col < p.N;
        float4 a[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rsub;
          const uint4 u = lds128(slab + rr * 128 + ((cl ^ (rr & 7)) << 4));
          a[it] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z),
                              __uint_as_float(u.w));
        }
        __syncwarp();  // slab may be overwritten by the next chunk
        // prefetch the next chunk's residual while this chunk is processed
        float4 res_next[8];
        if constexpr (EPI == EPI_BIAS_RESID) {
          if (c + 1 < NCH) {
            const int ncol = col + 32;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int grow = row_base + it * 4;
              res_next[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (grow < m_eff && ncol < p.N)
                res_next[it] = *reinterpret_cast<const float4*>(
                    p.resid + static_cast<size_t>(grow) * p.ld_res + ncol);
            }
          }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = row_base + it * 4;
          if (grow < m_eff && col_ok) {
            float v0 = round_f16(a[it].x + bias_r[c][0]), v1 = round_f16(a[it].y + bias_r[c][1]);
            float v2 = round_f16(a[it].z + bias_r[c][2]), v3 = round_f16(a[it].w + bias_r[c][3]);
            if constexpr (EPI == EPI_BIAS) {
              if (p.scale != 1.0f) {
                v0 = round_f16(v0 * p.scale); v1 = round_f16(v1 * p.scale);
                v2 = round_f16(v2 * p.scale); v3 = round_f16(v3 * p.scale);
              }
            }
            if constexpr (EPI == EPI_BIAS_GELU) {
              v0 = gelu_erf(v0); v1 = gelu_erf(v1); v2 = gelu_erf(v2); v3 = gelu_erf(v3);
            } else if constexpr (EPI == EPI_BIAS_RELU) {
              v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
            }
            if constexpr (EPI == EPI_BIAS_RESID) {
              if (p.scale != 1.0f) {
                v0 = round_f16(v0 * p.scale); v1 = round_f16(v1 * p.scale);
                v2 = round_f16(v2 * p.scale); v3 = round_f16(v3 * p.scale);
              }
              const float4 o = make_float4(res[it].x + v0, res[it].y + v1, res[it].z + v2,
                                           res[it].w + v3);
              *reinterpret_cast<float4*>(p.out_f + static_cast<size_t>(grow) * p.ldo_f + col) = o;
              if (p.out_h != nullptr) {
                const uint2 oh = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
                *reinterpret_cast<uint2*>(p.out_h + static_cast<size_t>(grow) * p.ldo_h + col) = oh;
              }
            } else {
              const uint2 oh = make_uint2(pack_half2(v0, v1), pack_half2(v2, v3));
              *reinterpret_cast<uint2*>(p.out_h + static_cast<size_t>(grow) * p.ldo_h + col) = oh;
            }
          }
        }
        if constexpr (EPI == EPI_BIAS_RESID) {
          if (c + 1 < NCH) {
#pragma unroll
            for (int it = 0; it < 8; ++it) res[it] = res_next[it];
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dyt
