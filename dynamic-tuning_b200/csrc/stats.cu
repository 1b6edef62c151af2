// On-device keep-rate / FLOPs accounting of an evaluation batch.
// Replaces the per-image Python loop of block_flops_dict.batch_select_flops / select_flops
// (reference block_flops_dict.py:57-83: flops(image) = base + sum over layers of
// flops_dict[kept patch tokens + 1]) and the per-layer "tokens selected" means that
// engine_finetune.py:341-352 computes after gathering every mask to every rank with a padded
// all_gather (:446-480): here each rank keeps per-layer kept-token counters on the device and
// only the [L + 2] counters need a reduction.
#include <stdarg.h>

#include "../../include/dyt_b200.h"
#include "host_utils.h"

namespace dyt {

// one CTA per image: warp w counts layers w, w + 4, ...; thread 0 then adds the table entries in
// layer order (the same left-to-right fp32 additions as the reference loop)
__global__ void __launch_bounds__(128)
keep_stats_kernel(const float* __restrict__ token_select, int L, int Np, const float* __restrict__ table,
                  int table_len, int block_num, float base_flops, float* __restrict__ image_flops,
                  unsigned long long* __restrict__ counters) {
  __shared__ int s_count[64];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* img = token_select + static_cast<size_t>(b) * L * Np;
  for (int l = warp; l < L; l += 4) {
    float c = 0.f;
    for (int n = lane; n < Np; n += 32) c += img[static_cast<size_t>(l) * Np + n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) {
      const int ci = static_cast<int>(c);  // .sum(-1).int(): truncation, as in the reference
      s_count[l] = ci;
      if (counters != nullptr) atomicAdd(counters + l, static_cast<unsigned long long>(ci));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (image_flops != nullptr) {
      float f = base_flops;
      for (int i = 0; i < block_num - L; ++i) f += table[min(Np + 1, table_len - 1)];
      for (int l = 0; l < L; ++l) f += table[min(s_count[l] + 1, table_len - 1)];
      image_flops[b] = f;
    }
    if (counters != nullptr) atomicAdd(counters + L, 1ull);  // images seen
  }
}

}  // namespace dyt

extern "C" int dyt_keep_stats(const float* token_select, int B, int L, int Np,
                              const float* flops_table, int table_len, int block_num,
                              float base_flops, float* image_flops, unsigned long long* counters,
                              void* stream) {
  using namespace dyt;
  DYT_CHECK_ARG(token_select != nullptr, "keep_stats: null mask");
  DYT_CHECK_ARG(B >= 0 && L >= 1 && L <= 64 && Np >= 1, "keep_stats: bad sizes (1 <= L <= 64)");
  DYT_CHECK_ARG(image_flops == nullptr || (flops_table != nullptr && table_len >= Np + 2 && block_num >= L),
                "keep_stats: the FLOPs table must hold Np + 2 entries and block_num >= L");
  if (B == 0) return DYT_OK;
  keep_stats_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      token_select, L, Np, flops_table, table_len, block_num, base_flops, image_flops, counters);
  return cuda_status(cudaGetLastError(), "keep_stats_kernel launch");
}
