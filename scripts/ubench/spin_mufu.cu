// Micro-benchmark (debug aid): does a warp that polls an mbarrier (mbar_wait / try_wait loop, as the
// idle roles of the attention kernel do) slow down MUFU-bound warps on the same SM sub-partition?
// warps 0..nw-1: softmax-like loop (ffma + ex2 + sum + pack); warps nw..nw+ns-1: spin on a barrier
// that completes only when the workers are done.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../dynamic-tuning_b200/csrc/ptx.cuh"
using namespace dyt;

__global__ void k(long long* out, float* sink, int iters, int nw, int mode) {
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, nw); fence_mbar_init(); }
  __syncthreads();
  if (warp < nw) {
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = -0.01f * (lane + j);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      float s = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = ex2_approx(fmaf(x[j], 0.18f, -1.0f));
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
    if (lane == 0) { out[warp] = t1 - t0; mbar_arrive(&bar); }
    if (s == 123.f) sink[0] = s;
  } else {
    if (mode == 0) {            // whole warp polls (issuer / softmax / output style)
      mbar_wait(&bar, 0);
    } else if (mode == 1) {     // one lane polls (TMA producer style)
      if (lane == 0) mbar_wait(&bar, 0);
    } else {                    // poll with nanosleep back-off
      while (!mbar_try_wait(&bar, 0)) __nanosleep(200);
    }
  }
}

int main() {
  long long* d; float* sink; cudaMalloc(&d, 64 * 8); cudaMalloc(&sink, 4);
  long long h[64];
  const int iters = 200;
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {4, 8})
      for (int ns : {0, 4, 8}) {
        if (ns == 0 && mode > 0) continue;
        k<<<1, (nw + ns) * 32>>>(d, sink, iters, nw, mode); cudaDeviceSynchronize();
        k<<<1, (nw + ns) * 32>>>(d, sink, iters, nw, mode); cudaDeviceSynchronize();
        cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("mode %d (%s) workers %d spinners %d: %.2f clk per warp-element per SMSP  %s\n", mode,
               mode == 0 ? "warp polls" : mode == 1 ? "lane polls" : "nanosleep poll", nw, ns,
               (double)mx / iters / 32 / (nw / 4), cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
