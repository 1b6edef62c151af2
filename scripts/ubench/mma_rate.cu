// Micro-benchmark (debug aid): issue rate / throughput of single-CTA tcgen05.mma kind::f16 for the
// shapes the attention kernel uses, SS (A in smem) versus TS (A in TMEM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../dynamic-tuning_b200/csrc/ptx.cuh"
using namespace dyt;

__global__ void k(long long* out, int n_mma, int N, int ts, int b_mn, int nacc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N, 0, b_mn);
    const uint64_t a_desc = umma_desc_sw128(smem_u32(smem));
    const uint64_t b_desc = umma_desc_sw128(smem_u32(smem) + 32768);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      if (ts) umma_ts_f16(tptr + 256 + (i % nacc) * 64, tptr + (i & 7) * 8, b_desc + (b_mn ? (i & 7) * 128 : 2 * (i & 3)), idesc, i >= nacc);
      else umma_ss_f16(tptr + 256 + (i % nacc) * 64, a_desc + 2 * (i & 3), b_desc + (b_mn ? (i & 7) * 128 : 2 * (i & 3)), idesc, i >= nacc);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tptr, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[2];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  struct { int N, ts, bmn; const char* name; } cfg[] = {
    {64, 0, 1, "SS 128x64x16  (B MN-major)"}, {64, 1, 1, "TS 128x64x16  (B MN-major)  = PV"},
    {64, 0, 0, "SS 128x64x16  (B K-major)"}, {64, 1, 0, "TS 128x64x16  (B K-major)"},
    {128, 1, 1, "TS 128x128x16 (B MN-major)"}, {208, 0, 0, "SS 128x208x16 (B K-major)   = S"},
    {256, 0, 0, "SS 128x256x16 (B K-major)"}, {256, 1, 0, "TS 128x256x16 (B K-major)"}, {16, 1, 0, "TS 128x16x16"}, {16, 0, 0, "SS 128x16x16"}};
  for (auto& c : cfg)
    for (int nacc : {1, 2, 4}) {
      const int n = 104;
      if (c.N > 64 && nacc > 1) continue;
      for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 100 * 1024>>>(d, n, c.N, c.ts, c.bmn, nacc); cudaDeviceSynchronize(); }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-36s nacc=%d n=%3d: issue %6lld clk (%.1f/mma), complete %6lld clk (%.1f/mma)  %s\n", c.name, nacc, n, h[0], (double)h[0] / n, h[1], (double)h[1] / n, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
