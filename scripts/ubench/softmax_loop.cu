// Micro-benchmark (debug aid): the attention kernel's pass-2 loop (two 32-key chunks per iteration,
// TMEM ping-pong) in isolation: clocks per chunk for 1 / 2 warps per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../dynamic-tuning_b200/csrc/ptx.cuh"
using namespace dyt;

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
// variant: interleaved (compiler free to schedule), no explicit batching
__device__ __forceinline__ void exp32b(uint32_t (&r)[32], float sl2, float mb, float (&sum)[4],
                                       uint32_t p_addr) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float e0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), sl2, -mb));
    const float e1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -mb));
    sum[j & 3] += e0 + e1;
    pk[j] = pack_half2(e0, e1);
  }
  tmem_st16(p_addr, pk);
}

template <int VARIANT>
__global__ void k(long long* out, float* sink, int units, int nfull, int mma_mode, int nw_soft, float sl2_rt, float mb_rt) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { stop = 0; mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 8) {
    // background MMA stream (one elected lane): mode 1 = PV-like TS 128x64x16 x13, mode 2 = S-like SS 128x208x16 x4, 3 = both
    if ((mma_mode & 15) != 0) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tptr, 0);
    const uint32_t sm = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    uint32_t ph = 0;
    long long n = 0;
    while (!stop) {
      if (elect_one()) {
        if (mma_mode & 2)
          for (int i = 0; i < 4; ++i) umma_ss_f16(tm + 256, umma_desc_sw128(sm) + 2 * i, umma_desc_sw128(sm + 32768) + 2 * i, umma_idesc_f16(128, 208, 0, 0), i != 0);
        if (mma_mode & 1)
          for (int i = 0; i < 13; ++i) umma_ts_f16(tm + 192, tm + 256 + i * 8, umma_desc_sw128(sm + 32768) + (i & 7) * 128, umma_idesc_f16(128, 64, 0, 1), i != 0);
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, ph); ph ^= 1; ++n;
    }
    if (lane == 0) out[32] = n;
    }
  } else if (warp < (int)gridDim.y * 0 + nw_soft) {
  const uint32_t s_addr = tptr + (uint32_t((warp & 3) * 32) << 16) + (warp >= 4 ? 256 : 0);
  const uint32_t p_base = (mma_mode & 16) ? (s_addr ^ 256u) : s_addr;
  float sum4[4] = {0, 0, 0, 0};
  const float sl2 = (mma_mode & 32) ? sl2_rt : 0.18f, mb = (mma_mode & 32) ? mb_rt * (1.0f + 1e-3f * (threadIdx.x & 31)) : 1.0f;
  long long t0 = clock64();
  for (int u = 0; u < units; ++u) {
    uint32_t ra[32], rb[32];
    tmem_ld32(s_addr, ra);
    tmem_ld_wait();
#pragma unroll 1
    for (int c = 0; c < nfull; c += 2) {
      if (c + 1 < nfull) tmem_ld32(s_addr + (c + 1) * 32, rb);
      if (VARIANT == 0) exp32(ra, sl2, mb, sum4, p_base + c * 16); else exp32b(ra, sl2, mb, sum4, p_base + c * 16);
      tmem_ld_wait();
      if (c + 1 < nfull) {
        if (c + 2 < nfull) tmem_ld32(s_addr + (c + 2) * 32, ra);
        if (VARIANT == 0) exp32(rb, sl2, mb, sum4, p_base + (c + 1) * 16); else exp32b(rb, sl2, mb, sum4, p_base + (c + 1) * 16);
        tmem_ld_wait();
      }
    }
    tmem_st_wait();
  }
  long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  if (sum4[0] + sum4[1] + sum4[2] + sum4[3] == 123.f) sink[0] = 1;
  asm volatile("bar.sync 1, %0;" :: "r"(nw_soft * 32));
  stop = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tptr, 512); }
}

int main() {
  long long* d; float* sink; cudaMalloc(&d, 64 * 8); cudaMalloc(&sink, 4);
  long long h[64];
  const int units = 50, nfull = 6;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int variant : {0, 32})
    for (int nw : {4, 8}) {
      for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 9 * 32, 70 * 1024>>>(d, sink, units, nfull, variant, nw, 0.18f, 1.0f);   // 8 softmax warps slots + MMA warp
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("bg mma mode %d, warps/SMSP %d: %.1f clk per chunk per warp, bg iters %lld (%s)\n", variant, nw / 4, 
             (double)mx / units / nfull, h[32], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
