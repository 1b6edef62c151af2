// Micro-benchmark (debug aid): do warps blocked in an mbarrier.try_wait loop steal issue slots from
// compute warps on the same SM sub-partition?  Compute warps 0..7 (2 per SMSP) run an FFMA/MUFU mix;
// warps 8..15 wait on an mbarrier that warp 0 completes at the end.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../dynamic-tuning_b200/csrc/ptx.cuh"
using namespace dyt;

__global__ void k(long long* out, float* sink, int iters, int waiters, int mode) {
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (warp < 8) {
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = -0.01f * (lane + j);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      float s = 0.f; uint32_t pk[16];
      if (mode == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = ex2_approx(fmaf(x[j], 0.18f, -1.0f));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaf(x[j], 0.18f, -1.0f);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) { s += x[2*j] + x[2*j+1]; pk[j] = pack_half2(x[2*j], x[2*j+1]); }
#pragma unroll
      for (int j = 0; j < 16; ++j) x[2*j] += __uint_as_float(pk[j] & 0x7fff) * 1e-30f;
      x[1] += s * 1e-30f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += x[j];
    if (lane == 0) out[warp] = t1 - t0;
    if (s == 123.f) sink[0] = s;
    __syncwarp();
    asm volatile("bar.sync 1, 256;");
    if (threadIdx.x == 0) mbar_arrive(&bar);
  } else if (warp < 8 + waiters) {
    if (mode >= 2) { /* sleep-poll variant */
      while (!mbar_test(&bar, 0)) __nanosleep(200);
    } else {
      mbar_wait(&bar, 0);
    }
  }
}

int main() {
  long long* d; float* sink; cudaMalloc(&d, 64 * 8); cudaMalloc(&sink, 4);
  long long h[64];
  const int iters = 300;
  for (int mode : {0, 1})
    for (int waiters : {0, 4, 8}) {
      for (int rep = 0; rep < 2; ++rep) { k<<<1, 512>>>(d, sink, iters, waiters, mode); cudaDeviceSynchronize(); }
      cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 8; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("mode %d (%s) waiters %d: %.1f clk per 32-element chunk per warp (%s)\n", mode, mode == 0 ? "ffma+ex2+sum+pack" : "ffma+sum+pack",
             waiters, (double)mx / iters, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
