// Adapter up-projection fused into the scatter-merge (and the next block's LayerNorm):
//   adapt = f16(f16(down * Wup^T + b_up) * scale)          (never written to HBM)
//   out   = adapt + (x1 + (kept(t) ? mlp_packed[pos[t]] : 0))
//   ln    = f16(LayerNorm(out))                            optional
// Replaces the up Linear of the adapter (reference models/model_speed_test.py:106-111, called at
// :291), torch.zeros + index_put + the two adds (:302-308) and the next block's norm1.  Before this
// kernel the [T, C] fp16 adapter output made a round trip through HBM (155 MB per layer at B = 256)
// and the up GEMM was a launch of its own: measured 79 us per layer of the step (ablation).
//
// Persistent CTA per SM, tile = 128 rows x all C columns:
//   warp 0    TMA: the 128 x 64 `down` tile of a row tile, then Wup in chunks of 128 output columns
//             (16 KB each, ring of three; Wup is re-streamed from the L2 for every tile so that the
//             shared memory can hold the epilogue's staging buffers instead)
//   warp 1    tcgen05.mma issuer: per tile C/128 accumulator chunks of 128 columns (K = 64: four
//             MMAs each) into a ring of four TMEM buffers
//   warp 2    TMEM allocator
//   warps 4.. sixteen epilogue warps, warp = (TMEM lane quarter q, 32-column part of every chunk).
//             Pass 1 per chunk: accumulator (thread = row) -> bias, fp16 rounding, scale -> fp16
//             transpose through a per-warp slab -> coalesced layout (8 lanes per row, 128-byte
//             segments): x1 + mlp + adapt, fp32 store of `out`, running mean / M2 of the row
//             (Chan's pairwise update: as accurate as the two-pass variance of rowwise.cuh).
//             Row statistics are merged over the 8 lanes and the 4 part-warps (shared memory,
//             named barrier per quarter).  Pass 2: every lane re-reads the `out` values it wrote
//             itself (L2 hits, same-thread read-after-write) and stores f16((v - mean) rstd g + b).
// The kernel is HBM-bound and its loads are 128-byte pieces of 32 rows per instruction: what
// matters is bytes in flight.  Every lane therefore owns a private 192-byte staging slot in shared
// memory that cp.async fills one chunk ahead (x1 / mlp in pass 1, `out` in pass 2): the loads of
// chunk j+1 are in flight while chunk j is transposed, added and stored, without holding registers.
// (Register-only versions: 178 us with the loads of a chunk serialised by a scoreboard alias,
// 112 us with eight row loads in flight per warp; L2 prefetch of whole rows made it slower.)
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"
#include "internal.h"
#include "ptx.cuh"

namespace dyt {

constexpr int MU_BM = 128;
constexpr int MU_CH = 128;                     // accumulator chunk (columns)
constexpr int MU_EW = 16;                      // epilogue warps
constexpr int MU_THREADS = 128 + MU_EW * 32;   // 640
constexpr int MU_ABYTES = MU_BM * 128;         // one `down` tile: 128 rows x 64 halves
constexpr int MU_WBYTES = MU_CH * 128;         // one Wup chunk: 128 output columns x 64 halves
constexpr int MU_WSTAGES = 3;
constexpr int MU_SLAB = 32 * 64;               // 32 rows x 32 fp16 per epilogue warp
constexpr int MU_STAGE = 32 * 128 + 32 * 64;   // per warp: 32 rows x (32 fp32 + 32 fp16)
constexpr int MU_MAXC = 1024;
constexpr int MU_EPI_REGS = 104, MU_AUX_REGS = 56;
constexpr int MU_TBUF = 3;                      // TMEM accumulator buffers of the up projection (3 x 128 columns)
constexpr int MU_DCOL = MU_TBUF * MU_CH;        // TMEM column of the down projection's 128 x 64 accumulator
constexpr int MU_KBUF = 3;                      // fp16 x1 chunks (128 x 64) in flight to the down MMAs
constexpr int MU_WDBYTES = 64 * 128;            // one Wdown chunk: 64 outputs x 64 halves of K

struct MergeUpParams {
  int T, C;
  const __half* bias;   // [C] fp16 or nullptr
  float scale;
  const float* x1;
  int ldx;
  const __half* mlp;    // packed MLP output
  int ldm;
  const int* token_pos; // [T]: row of mlp, or -1
  float* out;
  int ldo;
  const float* ln_w;    // optional next LayerNorm
  const float* ln_b;
  float eps;
  __half* ln_out;
  int ldn;
  // fused adapter down-projection (tmap_wd valid): down = relu(f16(f16(x1) Wd^T + bd)) is computed
  // from the tile's x1 rows in a pre-pass instead of being read from HBM
  int fuse_down;
  const __half* down_bias;   // [Kd] fp16 or nullptr
  int Kd;                    // adapter bottleneck (<= 64)
};

static inline int mu_smem_bytes(int C) {
  return MU_WSTAGES * MU_WBYTES + MU_ABYTES + MU_EW * (MU_SLAB + MU_STAGE) + 3 * C * 4 + 64 * 4 +
         2 * 4 * MU_BM * 8 + 512 + 1024;
}

__device__ __forceinline__ void mu_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// 16-byte asynchronous copy global -> shared, L2 only
__device__ __forceinline__ void mu_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// 8-byte copy; src_bytes = 0 writes zeros instead
__device__ __forceinline__ void mu_cp8z(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void mu_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mu_cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mu_cp_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__global__ void __launch_bounds__(MU_THREADS, 1)
merge_up_kernel(const __grid_constant__ CUtensorMap tmap_a,   // down [T, K], box 128 x 64
                const __grid_constant__ CUtensorMap tmap_w,   // Wup  [C, K], box 128 x 64
                const __grid_constant__ CUtensorMap tmap_wd,  // Wdown [K, C], box 64 x 64 (fuse_down)
                const MergeUpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int C = p.C;
  const int nch = C / MU_CH;
  uint8_t* w_smem = smem;                                   // ring of Wup chunks
  uint8_t* a_smem = w_smem + MU_WSTAGES * MU_WBYTES;        // one `down` tile
  uint8_t* slabs = a_smem + MU_ABYTES;
  uint8_t* stages = slabs + MU_EW * MU_SLAB;
  float* bias_s = reinterpret_cast<float*>(stages + MU_EW * MU_STAGE);   // [C]
  float* gamma_s = bias_s + C;
  float* beta_s = gamma_s + C;
  float* dbias_s = beta_s + C;                                          // [64]
  float2* stats_s = reinterpret_cast<float2*>(dbias_s + 64);           // [2][4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stats_s + 2 * 4 * MU_BM);
  uint64_t* a_full = bars;            // [1]
  uint64_t* a_empty = bars + 1;       // [1]
  uint64_t* w_full = bars + 2;        // [3]
  uint64_t* w_empty = bars + 5;       // [3]
  uint64_t* t_full = bars + 8;        // [3]
  uint64_t* t_empty = bars + 12;      // [3]
  uint64_t* ak_ready = bars + 16;     // [3] epilogue warps -> MMA: fp16 x1 chunk written
  uint64_t* ak_free = bars + 19;      // [3] MMA -> epilogue warps
  uint64_t* down_done = bars + 22;    // [1] MMA -> epilogue warps: the down accumulator is complete
  uint64_t* down_ready = bars + 23;   // [1] epilogue warps -> MMA: the `down` tile is in shared memory
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 24);
  uint8_t* ak_smem = stages;          // the x1 chunks live in the staging area (idle during the pre-pass)

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    if (p.fuse_down) tma_prefetch_desc(&tmap_wd);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < MU_WSTAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < MU_TBUF; ++b) {
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], MU_EW);
    }
    for (int b = 0; b < MU_KBUF; ++b) {
      mbar_init(&ak_ready[b], MU_EW);
      mbar_init(&ak_free[b], 1);
    }
    mbar_init(down_done, 1);
    mbar_init(down_ready, MU_EW);
    fence_mbar_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  if (warp_idx >= 4) {
    // parameters (not produced by the preceding kernels) -> shared memory
    for (int i = threadIdx.x - 128; i < C; i += MU_EW * 32) {
      bias_s[i] = p.bias != nullptr ? __half2float(p.bias[i]) : 0.f;
      gamma_s[i] = p.ln_w != nullptr ? p.ln_w[i] : 1.f;
      beta_s[i] = p.ln_b != nullptr ? p.ln_b[i] : 0.f;
    }
    if (threadIdx.x - 128 < 64) {
      const int i = threadIdx.x - 128;
      dbias_s[i] = (p.fuse_down && p.down_bias != nullptr && i < p.Kd) ? __half2float(p.down_bias[i]) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int num_tiles = (p.T + MU_BM - 1) / MU_BM;

  if (warp_idx < 4) {
    reg_dealloc<MU_AUX_REGS>();
    if (warp_idx == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        pdl_wait();   // `down` comes from the preceding kernels
        int ws = 0;
        uint32_t wph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
          if (!p.fuse_down) {
            mbar_wait(a_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u);
            mbar_arrive_expect_tx(a_full, MU_ABYTES);
            tma_load_2d(a_smem, &tmap_a, a_full, 0, tile * MU_BM);
          } else {
            // the weight ring first carries Wdown in K chunks of 64 columns (8 KB each)
            for (int kc = 0; kc < C / 64; ++kc) {
              mbar_wait(&w_empty[ws], wph ^ 1u);
              mbar_arrive_expect_tx(&w_full[ws], MU_WDBYTES);
              tma_load_2d(w_smem + ws * MU_WBYTES, &tmap_wd, &w_full[ws], kc * 64, 0);
              if (++ws == MU_WSTAGES) {
                ws = 0;
                wph ^= 1u;
              }
            }
          }
          for (int j = 0; j < nch; ++j) {
            mbar_wait(&w_empty[ws], wph ^ 1u);
            mbar_arrive_expect_tx(&w_full[ws], MU_WBYTES);
            tma_load_2d(w_smem + ws * MU_WBYTES, &tmap_w, &w_full[ws], 0, j * MU_CH);
            if (++ws == MU_WSTAGES) {
              ws = 0;
              wph ^= 1u;
            }
          }
        }
      }
    } else if (warp_idx == 1) {
      // ===================== MMA issuer (warp-uniform, one elected lane issues) =====================
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t w_u = __shfl_sync(0xffffffffu, smem_u32(w_smem), 0);
      const uint32_t a_u = __shfl_sync(0xffffffffu, smem_u32(a_smem), 0);
      constexpr uint32_t idesc = umma_idesc_f16(MU_BM, MU_CH, 0, 0);
      const uint64_t a_desc = umma_desc_sw128(a_u);
      int it = 0;
      int ws = 0;
      uint32_t wph = 0;
      uint32_t buf = 0, bph = 0;   // TMEM accumulator ring of the up projection
      uint32_t ab = 0, aph = 0;    // x1 chunk ring of the down projection
      constexpr uint32_t idesc_d = umma_idesc_f16(MU_BM, 64, 0, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        if (p.fuse_down) {
          // down[128, 64] = f16(x1)[128, C] Wdown[64, C]^T, K in chunks of 64 written by the epilogue warps
          const uint32_t ak_u = __shfl_sync(0xffffffffu, smem_u32(ak_smem), 0);
          const int nkc = C / 64;
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&w_full[ws], wph);
            mbar_wait(&ak_ready[ab], aph);
            tc_fence_after();
            const uint64_t xa_desc = umma_desc_sw128(ak_u + ab * MU_ABYTES);
            const uint64_t wd_desc = umma_desc_sw128(w_u + ws * MU_WBYTES);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss_f16(tmem_u + MU_DCOL, xa_desc + 2 * k, wd_desc + 2 * k, idesc_d, (kc | k) != 0 ? 1u : 0u);
              umma_commit(&w_empty[ws]);
              umma_commit(&ak_free[ab]);
              if (kc == nkc - 1) umma_commit(down_done);
            }
            __syncwarp();
            if (++ws == MU_WSTAGES) {
              ws = 0;
              wph ^= 1u;
            }
            if (++ab == MU_KBUF) {
              ab = 0;
              aph ^= 1u;
            }
          }
          mbar_wait(down_ready, static_cast<uint32_t>(it) & 1u);   // the `down` tile is in a_smem
        } else {
          mbar_wait(a_full, static_cast<uint32_t>(it) & 1u);
        }
        tc_fence_after();
        for (int j = 0; j < nch; ++j) {
          mbar_wait(&w_full[ws], wph);
          mbar_wait(&t_empty[buf], bph ^ 1u);
          tc_fence_after();
          const uint64_t b_desc = umma_desc_sw128(w_u + ws * MU_WBYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss_f16(tmem_u + buf * MU_CH, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0 ? 1u : 0u);
            umma_commit(&t_full[buf]);
            umma_commit(&w_empty[ws]);
            if (j == nch - 1 && !p.fuse_down) umma_commit(a_empty);
          }
          __syncwarp();
          if (++ws == MU_WSTAGES) {
            ws = 0;
            wph ^= 1u;
          }
          if (++buf == MU_TBUF) {
            buf = 0;
            bph ^= 1u;
          }
        }
      }
    }
  } else {
    reg_alloc<MU_EPI_REGS>();
    // ===================== epilogue =====================
    const int e = warp_idx - 4;
    const int q = e & 3;       // == warp_idx % 4: the TMEM lane quarter this warp may access
    const int part = e >> 2;   // 32-column part of every 128-column chunk
    const uint32_t slab = smem_u32(slabs) + e * MU_SLAB;
    const uint32_t my_row = slab + lane * 64;
    const int swz_w = (lane >> 1) & 3;
    const int p8 = lane & 7;
    const int r8 = lane >> 3;
    // this lane's private staging slots: row it*4 + r8, 16 B of x1 / out, 8 B of mlp
    const uint32_t stage_x = smem_u32(stages) + e * MU_STAGE + r8 * 128 + p8 * 16;
    const uint32_t stage_m = smem_u32(stages) + e * MU_STAGE + 32 * 128 + r8 * 64 + p8 * 8;
    // slab read address of (row it*4 + r8, this lane's 4 columns): the chunk swizzle depends on
    // it & 1 only
    const uint32_t slab_rd0 = slab + r8 * 64 + (((p8 >> 1) ^ ((r8 >> 1) & 3)) << 4) + (p8 & 1) * 8;
    const uint32_t slab_rd1 = slab + r8 * 64 + (((p8 >> 1) ^ (((r8 >> 1) + 2) & 3)) << 4) + (p8 & 1) * 8;
    const bool plain = p.scale == 1.0f;
    const bool do_ln = p.ln_out != nullptr;
    const uint32_t bias_u = smem_u32(bias_s), gamma_u = smem_u32(gamma_s), beta_u = smem_u32(beta_s);
    const int col_lane = part * 32 + p8 * 4;   // + j * 128
    const int last_row = p.T - 1;
    pdl_wait();

    // token_pos of a quarter's 32 rows: ONE coalesced load per tile, handed out by shuffles.
    // (Loaded straight into an array, the registers stay tied to a load scoreboard slot that later
    // loads reuse: every read of pos[it] then waited for all loads issued before it, which
    // serialised the eight row loads of a chunk -- 10k clocks per chunk.)
    auto load_pl = [&](int tile) -> int {
      const int r = tile * MU_BM + q * 32 + lane;
      return (tile < num_tiles && r < p.T) ? p.token_pos[r] : -1;
    };
    // Rows past the end are clamped to the last row for every load (their results are computed and
    // dropped), so the per-row code has no bounds branches; only the stores look at `live`.
    // pass-1 inputs of chunk j -> staging (asynchronous); dropped rows get zeros for the mlp part
    auto issue_p1 = [&](int tile, int pl, int j) {
      if (tile < num_tiles) {
        const int rbase = tile * MU_BM + q * 32 + r8;
        const int col = j * MU_CH + col_lane;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = min(rbase + it * 4, last_row);
          const int ps = __shfl_sync(0xffffffffu, pl, it * 4 + r8);
          mu_cp16(stage_x + it * 512, p.x1 + (static_cast<uint32_t>(grow) * p.ldx + col));
          mu_cp8z(stage_m + it * 256, p.mlp + (static_cast<uint32_t>(ps < 0 ? 0 : ps) * p.ldm + col),
                  ps < 0 ? 0u : 8u);
        }
      }
      mu_cp_commit();
    };
    // pass-2 input: the `out` values this lane stored in pass 1.  Pass 2 needs neither the mlp
    // slots nor the transpose slab, which together make a second buffer: two chunks in flight.
    const uint32_t stage_b0 = smem_u32(stages) + e * MU_STAGE + 32 * 128 + lane * 16;   // it 0..3
    const uint32_t stage_b1 = slab + lane * 16;                                         // it 4..7
    auto p2_slot = [&](int j, int it) -> uint32_t {
      return (j & 1) ? ((it < 4 ? stage_b0 : stage_b1) + (it & 3) * 512) : stage_x + it * 512;
    };
    auto issue_p2 = [&](int tile, int j) {
      if (j < nch) {
        const int rbase = tile * MU_BM + q * 32 + r8;
        const int col = j * MU_CH + col_lane;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = min(rbase + it * 4, last_row);
          mu_cp16(p2_slot(j, it), p.out + (static_cast<uint32_t>(grow) * p.ldo + col));
        }
      }
      mu_cp_commit();   // (possibly empty: every pass-2 step commits exactly one group)
    };

    const bool fuse_down = p.fuse_down != 0;
    // ---- pre-pass (fuse_down): this warp turns rows e*8 .. e*8+7 of the tile's x1 into the fp16
    // A operand of the down projection, one 64-column K chunk at a time: lane -> (row 2i + lane/16,
    // 4 columns), 16 lanes cover the 256 bytes of a row; the chunk is written in the canonical
    // K-major 128B-swizzled layout (16-byte unit index XOR row % 8) the MMA descriptors expect.
    const int pr = lane >> 4, pc4 = (lane & 15) * 4;
    uint32_t ka = 0, kph = 0;   // x1 chunk ring position / phase (runs over the whole kernel)
    auto load_x1_chunk = [&](int tile, int kc, float4 (&v)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int grow = min(tile * MU_BM + e * 8 + 2 * i + pr, last_row);
        v[i] = *reinterpret_cast<const float4*>(p.x1 + (static_cast<uint32_t>(grow) * p.ldx + kc * 64 + pc4));
      }
    };
    auto store_x1_chunk = [&](const float4 (&v)[4]) {
      const uint32_t base = smem_u32(ak_smem) + ka * MU_ABYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = e * 8 + 2 * i + pr;   // row of the tile; r / 8 == e, r % 8 == 2i + pr
        const uint32_t addr = base + e * 1024 + (2 * i + pr) * 128 +
                              ((((pc4 >> 3) ^ (2 * i + pr)) & 7) << 4) + (pc4 & 7) * 2;
        (void)r;
        sts64(addr, pack_half2(v[i].x, v[i].y), pack_half2(v[i].z, v[i].w));
      }
    };

    int pl = load_pl(blockIdx.x);
    if (!fuse_down) issue_p1(blockIdx.x, pl, 0);
    uint32_t buf = 0, bph = 0;   // TMEM accumulator ring of the up projection
    int it_tile = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it_tile) {
      const int rbase = tile * MU_BM + q * 32 + r8;
      const int next_tile = tile + gridDim.x;
      const int pl_next = load_pl(next_tile);   // in flight for the whole tile
      if (fuse_down) {
        // every warp is done with the previous tile's staged data: the staging area becomes the
        // x1 chunk ring of the pre-pass
        mu_cp_wait();
        mu_bar_sync(5, MU_EW * 32);
        const int nkc = C / 64;
        // four chunks of this warp's x1 rows in flight (registers), MU_KBUF converted chunks in
        // flight to the MMAs (shared memory)
        float4 va[4], vb[4], vc[4], vd[4];
        load_x1_chunk(tile, 0, va);
        load_x1_chunk(tile, 1, vb);
        load_x1_chunk(tile, 2, vc);
        load_x1_chunk(tile, 3, vd);
#pragma unroll 1
        for (int kc = 0; kc < nkc; kc += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (kc + u < nkc) {
              mbar_wait(&ak_free[ka], kph ^ 1u);
              float4 (&v)[4] = u == 0 ? va : (u == 1 ? vb : (u == 2 ? vc : vd));
              store_x1_chunk(v);
              fence_proxy_async_smem();   // generic-proxy stores -> visible to the MMA's async-proxy reads
              __syncwarp();
              if (lane == 0) mbar_arrive(&ak_ready[ka]);
              if (kc + u + 4 < nkc) load_x1_chunk(tile, kc + u + 4, v);
              if (++ka == MU_KBUF) {
                ka = 0;
                kph ^= 1u;
              }
            }
          }
        }
        // the down accumulator is complete (and the MMAs have read the last chunk: the staging area
        // is free again): first chunk of pass 1 on its way, then bias + ReLU -> the `down` tile
        mbar_wait(down_done, static_cast<uint32_t>(it_tile) & 1u);
        tc_fence_after();
        issue_p1(tile, pl, 0);
        {
          uint32_t r[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + MU_DCOL + part * 16, r);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j2 = 0; j2 < 8; ++j2) {
            const float b0 = dbias_s[part * 16 + 2 * j2], b1 = dbias_s[part * 16 + 2 * j2 + 1];
            const __half2 h = __floats2half2_rn(__uint_as_float(r[2 * j2]) + b0,
                                                __uint_as_float(r[2 * j2 + 1]) + b1);
            const __half2 o = __hmax2(h, __float2half2_rn(0.f));
            pk[j2] = *reinterpret_cast<const uint32_t*>(&o);
          }
          // row q*32 + lane of the 128 x 64 fp16 tile, 16-byte units 2*part and 2*part + 1
          const int row = q * 32 + lane;
          const uint32_t rowaddr = smem_u32(a_smem) + (row >> 3) * 1024 + (row & 7) * 128;
          sts128(rowaddr + ((((2 * part) ^ row) & 7) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          sts128(rowaddr + ((((2 * part + 1) ^ row) & 7) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(down_ready);
        }
      }
      unsigned live = 0;                         // bit it: row it*4 + r8 of this quarter exists
#pragma unroll
      for (int it = 0; it < 8; ++it) live |= (rbase + it * 4 < p.T ? 1u : 0u) << it;
      float mean[8], m2[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) mean[it] = m2[it] = 0.f;

#pragma unroll 1
      for (int j = 0; j < nch; ++j) {
        const int c0 = j * MU_CH + part * 32;   // first column of this warp's 32
        {
          uint32_t pk[16];
          {
            uint32_t r[32];
            mbar_wait(&t_full[buf], bph);
            tc_fence_after();
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * MU_CH + part * 32, r);
            uint4 bq[8];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) bq[j4] = lds128(bias_u + (c0 + j4 * 4) * 4);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
            if (++buf == MU_TBUF) {
              buf = 0;
              bph ^= 1u;
            }
            // thread = row: one rounding to fp16 per Linear output, then f16(v * scale)
#pragma unroll
            for (int j2 = 0; j2 < 16; ++j2) {
              const uint4 bv = bq[j2 >> 1];
              const float b0 = __uint_as_float((j2 & 1) ? bv.z : bv.x);
              const float b1 = __uint_as_float((j2 & 1) ? bv.w : bv.y);
              __half2 h = __floats2half2_rn(__uint_as_float(r[2 * j2]) + b0,
                                            __uint_as_float(r[2 * j2 + 1]) + b1);
              if (!plain) h = __floats2half2_rn(__low2float(h) * p.scale, __high2float(h) * p.scale);
              pk[j2] = *reinterpret_cast<const uint32_t*>(&h);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            sts128(my_row + ((k ^ swz_w) << 4),
                   make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]));
        }
        __syncwarp();
        // staged inputs of this chunk -> registers; the slots are refilled at once
        mu_cp_wait();
        uint4 xv[8];
        uint2 mv[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          xv[it] = lds128(stage_x + it * 512);
          mv[it] = lds64(stage_m + it * 256);
        }
        if (j + 1 < nch) issue_p1(tile, pl, j + 1);
        else if (!do_ln && !fuse_down) issue_p1(next_tile, pl_next, 0);
        // coalesced layout: lane -> (row it*4 + lane/8, columns (lane%8)*4 .. +3)
        const int col = c0 + p8 * 4;
        const float w1 = 1.0f / static_cast<float>(j + 1);
        const float w2 = 4.0f * static_cast<float>(j) * w1;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const uint2 hv = lds64(((it & 1) ? slab_rd1 : slab_rd0) + it * 256);
          // (x1 + mlp) + adapt on packed fp32 pairs; a dropped row adds +0 for the mlp term
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&mv[it].x));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&mv[it].y));
          const float2 a01 = __half22float2(*reinterpret_cast<const __half2*>(&hv.x));
          const float2 a23 = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
          const unsigned long long v01 =
              f32x2_add(f32x2_add(f32x2_pack(__uint_as_float(xv[it].x), __uint_as_float(xv[it].y)),
                                  f32x2_pack(a.x, a.y)), f32x2_pack(a01.x, a01.y));
          const unsigned long long v23 =
              f32x2_add(f32x2_add(f32x2_pack(__uint_as_float(xv[it].z), __uint_as_float(xv[it].w)),
                                  f32x2_pack(b.x, b.y)), f32x2_pack(a23.x, a23.y));
          if ((live >> it) & 1u) {
            const float2 lo = f32x2_unpack(v01), hi = f32x2_unpack(v23);
            *reinterpret_cast<float4*>(p.out + (static_cast<uint32_t>(rbase + it * 4) * p.ldo + col)) =
                make_float4(lo.x, lo.y, hi.x, hi.y);
          }
          // running mean / sum of squared deviations of this lane's columns of the row
          const float2 s2 = f32x2_unpack(f32x2_add(v01, v23));
          const float m4 = 0.25f * (s2.x + s2.y);
          const unsigned long long nm = f32x2_pack(-m4, -m4);
          const unsigned long long d01 = f32x2_add(v01, nm), d23 = f32x2_add(v23, nm);
          const float2 q2 = f32x2_unpack(f32x2_fma(d01, d01, f32x2_mul(d23, d23)));
          const float q4 = q2.x + q2.y;
          const float dl = m4 - mean[it];
          mean[it] = fmaf(dl, w1, mean[it]);
          m2[it] += fmaf(dl * dl, w2, q4);
        }
        __syncwarp();   // the slab is overwritten by the next chunk
      }
      if (!do_ln) {
        pl = pl_next;
        continue;
      }
      issue_p2(tile, 0);   // in flight during the exchange of the row statistics
      issue_p2(tile, 1);

      // ---- row statistics: 8 lanes of a row, then the 4 part-warps of the quarter ----
      float cnt = static_cast<float>(4 * nch);
#pragma unroll
      for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const float mo = __shfl_xor_sync(0xffffffffu, mean[it], o);
          const float qo = __shfl_xor_sync(0xffffffffu, m2[it], o);
          const float d = mean[it] - mo;
          m2[it] = (m2[it] + qo) + d * d * (0.5f * cnt);
          mean[it] = 0.5f * (mean[it] + mo);
        }
        cnt *= 2.f;
      }
      float2* st = stats_s + (it_tile & 1) * 4 * MU_BM;
      if (p8 == 0) {
#pragma unroll
        for (int it = 0; it < 8; ++it)
          st[part * MU_BM + q * 32 + it * 4 + r8] = make_float2(mean[it], m2[it]);
      }
      mu_bar_sync(1 + q, 128);
      float rstd[8];
      const float inv_c = 1.0f / static_cast<float>(C);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = q * 32 + it * 4 + r8;
        const float2 s0 = st[0 * MU_BM + rr], s1 = st[1 * MU_BM + rr];
        const float2 s2 = st[2 * MU_BM + rr], s3 = st[3 * MU_BM + rr];
        const float d01 = s0.x - s1.x, d23 = s2.x - s3.x;
        const float m01 = 0.5f * (s0.x + s1.x), m23 = 0.5f * (s2.x + s3.x);
        const float q01 = (s0.y + s1.y) + d01 * d01 * (0.5f * cnt);
        const float q23 = (s2.y + s3.y) + d23 * d23 * (0.5f * cnt);
        const float d = m01 - m23;
        mean[it] = 0.5f * (m01 + m23);
        rstd[it] = rsqrtf(((q01 + q23) + d * d * cnt) * inv_c + p.eps);
      }

      // ---- pass 2: normalise the values this lane wrote ----
#pragma unroll 1
      for (int j = 0; j < nch; ++j) {
        const int col = j * MU_CH + col_lane;
        mu_cp_wait1();   // everything but the newest group: chunk j has landed
        uint4 v[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) v[it] = lds128(p2_slot(j, it));
        if (j + 1 < nch) {
          issue_p2(tile, j + 2);
        } else {
          __syncwarp();   // pass-1 slots of one lane overlap pass-2 slots of another
          // both buffers are free: the next tile's first chunk (with the fused down projection the
          // staging area first serves the pre-pass of that tile)
          if (!fuse_down) issue_p1(next_tile, pl_next, 0);
        }
        const uint4 gq = lds128(gamma_u + col * 4);
        const uint4 bq = lds128(beta_u + col * 4);
        const unsigned long long g01 = f32x2_pack(__uint_as_float(gq.x), __uint_as_float(gq.y));
        const unsigned long long g23 = f32x2_pack(__uint_as_float(gq.z), __uint_as_float(gq.w));
        const unsigned long long b01 = f32x2_pack(__uint_as_float(bq.x), __uint_as_float(bq.y));
        const unsigned long long b23 = f32x2_pack(__uint_as_float(bq.z), __uint_as_float(bq.w));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          // ((v - mean) * rstd) * g + b, two columns per instruction (same roundings as rowwise.cuh)
          const unsigned long long nm = f32x2_pack(-mean[it], -mean[it]);
          const unsigned long long rs = f32x2_pack(rstd[it], rstd[it]);
          const float2 y01 = f32x2_unpack(f32x2_fma(
              f32x2_mul(f32x2_add(f32x2_pack(__uint_as_float(v[it].x), __uint_as_float(v[it].y)), nm), rs),
              g01, b01));
          const float2 y23 = f32x2_unpack(f32x2_fma(
              f32x2_mul(f32x2_add(f32x2_pack(__uint_as_float(v[it].z), __uint_as_float(v[it].w)), nm), rs),
              g23, b23));
          if ((live >> it) & 1u)
            *reinterpret_cast<uint2*>(p.ln_out + (static_cast<uint32_t>(rbase + it * 4) * p.ldn + col)) =
                make_uint2(pack_half2(y01.x, y01.y), pack_half2(y23.x, y23.y));
        }
      }
      pl = pl_next;
    }
    mu_cp_wait();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool merge_up_supported(int C, int K, long long rows) {
  // (the kernel indexes every buffer with 32-bit element offsets)
  return C % MU_CH == 0 && C >= MU_CH && C <= MU_MAXC && K >= 8 && K <= 64 && K % 8 == 0 &&
         (rows + 1) * C < (1ll << 31);
}

int merge_up(const __half* down, int ld_down, const __half* up_w, int ldw, const __half* up_b,
             float scale, int K, const float* x1, int ldx, const __half* mlp_packed, int ldm,
             const int* token_pos, int n_rows, int C, float* out, int ldo, const float* nln_w,
             const float* nln_b, float eps, __half* nln_out, int ldn, cudaStream_t stream,
             const __half* down_w, int ld_dw, const __half* down_b) {
  // down_w != nullptr: the adapter's down projection (down_w [K, C], down_b [K]) is computed inside
  // the kernel from x1 and `down` is not read
  const bool fuse_down = down_w != nullptr;
  DYT_CHECK_ARG((down || fuse_down) && up_w && x1 && mlp_packed && token_pos && out, "merge_up: null buffer");
  if (!merge_up_supported(C, K, n_rows))
    return fail(DYT_EUNSUPPORTED,
                "merge_up: needs C %% 128 == 0, C <= %d, K <= 64, rows * C < 2^31 (C=%d K=%d rows=%d)",
                MU_MAXC, C, K, n_rows);
  DYT_CHECK_ARG(ldx % 4 == 0 && ldm % 4 == 0 && ldo % 4 == 0 && (fuse_down || ld_down >= K) && ldw >= K,
                "merge_up: strides");
  DYT_CHECK_ARG(!fuse_down || (ld_dw >= C && C % 64 == 0), "merge_up: down_w stride");
  DYT_CHECK_ARG(nln_out == nullptr || (nln_w && nln_b && ldn % 4 == 0), "merge_up: next-LN args");
  DYT_CHECK_ARG(out != x1, "merge_up: out must not alias x1");
  {  // the kernel indexes with 32-bit element offsets
    const long long lim = 1ll << 31;
    const long long big = static_cast<long long>(n_rows) *
                          (ldx > ldo ? (ldx > ldn ? ldx : ldn) : (ldo > ldn ? ldo : ldn));
    if (big + C >= lim || static_cast<long long>(n_rows) * ldm + C >= lim)
      return fail(DYT_EUNSUPPORTED, "merge_up: more than 2^31 elements per buffer");
  }
  if (n_rows == 0) return DYT_OK;
  CUtensorMap ta, tw, twd;
  int s = DYT_OK;
  if (fuse_down) {
    s = make_tmap_f16_sw128(&twd, down_w, static_cast<uint64_t>(K), static_cast<uint64_t>(C),
                            static_cast<uint64_t>(ld_dw), 64);
    if (s != DYT_OK) return s;
    ta = twd;   // (unused by the kernel in this mode)
  } else {
    s = make_tmap_f16_sw128(&ta, down, static_cast<uint64_t>(n_rows), static_cast<uint64_t>(K),
                            static_cast<uint64_t>(ld_down), MU_BM);
    if (s != DYT_OK) return s;
  }
  s = make_tmap_f16_sw128(&tw, up_w, static_cast<uint64_t>(C), static_cast<uint64_t>(K),
                          static_cast<uint64_t>(ldw), MU_CH);
  if (s != DYT_OK) return s;
  static SmemAttrCache smem_cache;
  s = ensure_dyn_smem(merge_up_kernel, mu_smem_bytes(MU_MAXC), smem_cache);
  if (s != DYT_OK) return s;
  MergeUpParams p;
  p.T = n_rows; p.C = C;
  p.bias = up_b; p.scale = scale;
  p.x1 = x1; p.ldx = ldx;
  p.mlp = mlp_packed; p.ldm = ldm;
  p.token_pos = token_pos;
  p.out = out; p.ldo = ldo;
  p.ln_w = nln_w; p.ln_b = nln_b; p.eps = eps;
  p.ln_out = nln_out; p.ldn = ldn;
  p.fuse_down = fuse_down ? 1 : 0;
  p.down_bias = down_b;
  p.Kd = K;
  if (!fuse_down) twd = tw;
  const int tiles = (n_rows + MU_BM - 1) / MU_BM;
  int grid = tiles < sm_count() ? tiles : sm_count();
  // at least 120 KB of shared memory per CTA: one CTA per SM, which owns all 512 TMEM columns
  int smem = mu_smem_bytes(C);
  if (smem < 120 * 1024) smem = 120 * 1024;
  return cuda_status(launch_pdl(merge_up_kernel, dim3(grid), dim3(MU_THREADS), smem, stream, ta, tw, twd, p),
                     "merge_up_kernel launch");
}

}  // namespace dyt

extern "C" int dyt_merge_up_fwd(const void* down_f16, int ld_down, const void* up_w_f16, int ldw,
                                const void* up_b_f16, float scale, int K, const float* x1, int ldx,
                                const void* mlp_packed_f16, int ldm, const int* token_pos, int n_rows,
                                int C, float* out, int ldo, const float* next_ln_w,
                                const float* next_ln_b, float eps, void* next_ln_out_f16, int ldn,
                                void* stream) {
  return dyt::merge_up(static_cast<const __half*>(down_f16), ld_down,
                       static_cast<const __half*>(up_w_f16), ldw,
                       static_cast<const __half*>(up_b_f16), scale, K, x1, ldx,
                       static_cast<const __half*>(mlp_packed_f16), ldm, token_pos, n_rows, C, out, ldo,
                       next_ln_w, next_ln_b, eps, static_cast<__half*>(next_ln_out_f16), ldn,
                       static_cast<cudaStream_t>(stream), nullptr, 0, nullptr);
}

extern "C" int dyt_adapter_merge_fwd(const void* down_w_f16, int ld_dw, const void* down_b_f16,
                                     const void* up_w_f16, int ldw, const void* up_b_f16, float scale,
                                     int K, const float* x1, int ldx, const void* mlp_packed_f16,
                                     int ldm, const int* token_pos, int n_rows, int C, float* out,
                                     int ldo, const float* next_ln_w, const float* next_ln_b,
                                     float eps, void* next_ln_out_f16, int ldn, void* stream) {
  if (down_w_f16 == nullptr) return dyt::fail(dyt::DYT_EINVAL, "adapter_merge: null down_w");
  return dyt::merge_up(nullptr, 0, static_cast<const __half*>(up_w_f16), ldw,
                       static_cast<const __half*>(up_b_f16), scale, K, x1, ldx,
                       static_cast<const __half*>(mlp_packed_f16), ldm, token_pos, n_rows, C, out, ldo,
                       next_ln_w, next_ln_b, eps, static_cast<__half*>(next_ln_out_f16), ldn,
                       static_cast<cudaStream_t>(stream), static_cast<const __half*>(down_w_f16), ld_dw,
                       static_cast<const __half*>(down_b_f16));
}
