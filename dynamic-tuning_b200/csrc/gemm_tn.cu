// Host launcher + C-ABI entry for the tcgen05 GEMM (see gemm_tn.cuh).
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/dyt_b200.h"
#include "gemm_tn.cuh"
#include "host_utils.h"

namespace dyt {

template <int BN, int EPI, int EW, bool RS = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap (&tb)[3], const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EW, RS>;
  auto kern = gemm_tn_kernel<BN, EPI, EW, RS>;
  static SmemAttrCache smem_cache;   // one per kernel instantiation, per device inside
  {
    const int st = ensure_dyn_smem(kern, Cfg::SMEM_BYTES, smem_cache);
    if (st != DYT_OK) return st;
  }
  const int m_tiles = (p.M + Cfg::BM - 1) / Cfg::BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  int clusters = ((m_tiles + 1) / 2) * n_tiles;  // one pair of tiles per cluster of two CTAs
  const int max_clusters = p.half_grid ? sm_count() / 4 : sm_count() / 2;
  if (clusters > max_clusters) clusters = max_clusters;
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // the kernel calls pdl_wait()
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_option().load(std::memory_order_relaxed) ? 2 : 1;
  return cuda_status(cudaLaunchKernelEx(&cfg, kern, ta, tb[0], tb[1], tb[2], p), "gemm_tn_kernel launch");
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap (&tb)[3],
                        const GemmParams& p, cudaStream_t s) {
  // sixteen epilogue warps where the epilogue, not the main loop, sets the pace
  if constexpr (BN == 128 || BN == 256) {
    if (epi == EPI_BIAS_GELU) return launch_gemm<BN, EPI_BIAS_GELU, 16>(ta, tb, p, s);
    if (epi == EPI_BIAS_GELU_KEEP) return launch_gemm<BN, EPI_BIAS_GELU_KEEP, 16>(ta, tb, p, s);
    if (epi == EPI_DGELU) return launch_gemm<BN, EPI_DGELU, 16>(ta, tb, p, s);
    if (p.K <= 128) {
      switch (epi) {
        case EPI_BIAS: return launch_gemm<BN, EPI_BIAS, 16>(ta, tb, p, s);
        case EPI_BIAS_RELU: return launch_gemm<BN, EPI_BIAS_RELU, 16>(ta, tb, p, s);
        case EPI_BIAS_RESID: return launch_gemm<BN, EPI_BIAS_RESID, 16>(ta, tb, p, s);
        default: return fail(DYT_EINVAL, "unknown epilogue %d", epi);
      }
    }
  }
  if (epi == EPI_BIAS_GELU_KEEP || epi == EPI_DGELU)
    return fail(DYT_EUNSUPPORTED, "gemm: the fused GELU training epilogues need N > 64");
  if constexpr (BN == 256) {
    // residual staged one chunk ahead by cp.async: full tiles only (no tail split: the fused row-dot
    // case, i.e. the proj GEMM of the block) and a residual whose rows can take 16-byte copies
    if (epi == EPI_BIAS_RESID && p.tail_split == 0 && p.N % 256 == 0)
      return launch_gemm<BN, EPI_BIAS_RESID, 8, true>(ta, tb, p, s);
  }
  switch (epi) {
    case EPI_BIAS: return launch_gemm<BN, EPI_BIAS, 8>(ta, tb, p, s);
    case EPI_BIAS_GELU: return launch_gemm<BN, EPI_BIAS_GELU, 8>(ta, tb, p, s);
    case EPI_BIAS_RELU: return launch_gemm<BN, EPI_BIAS_RELU, 8>(ta, tb, p, s);
    case EPI_BIAS_RESID: return launch_gemm<BN, EPI_BIAS_RESID, 8>(ta, tb, p, s);
    default: return fail(DYT_EINVAL, "unknown epilogue %d", epi);
  }
}

// Internal entry used by the composite block forward as well.
int gemm_tn_dot_slices(int N) {
  const int bn = (N <= 64) ? 64 : ((N <= 128 || (N % 256 != 0 && N % 128 == 0)) ? 128 : 256);
  return ((N + bn - 1) / bn) * 2;  // RESID epilogues with K > 128 run two column slices per tile
}

int gemm_tn(const __half* a, int lda, const __half* w, int ldw, int M, int N, int K,
            const int* m_dev, int epi, const __half* bias, __half* out_h, int ldo_h, float* out_f,
            int ldo_f, const float* resid, int ld_res, float scale, cudaStream_t stream,
            const float* dot_w, float* dot_out, int dot_ld, int dot_f16, __half* aux, int ld_aux,
            int reverse_m) {
  DYT_CHECK_ARG(a != nullptr && w != nullptr, "gemm: null operand");
  DYT_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  DYT_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "gemm: N and K must be multiples of 8 (N=%d K=%d)", N, K);
  DYT_CHECK_ARG(lda >= K && ldw >= K, "gemm: leading dimension smaller than K");
  if (epi == EPI_BIAS_RESID) {
    DYT_CHECK_ARG(out_f != nullptr && resid != nullptr, "gemm: residual epilogue needs out_f/resid");
    DYT_CHECK_ARG(ldo_f % 4 == 0 && ld_res % 4 == 0, "gemm: fp32 strides must be multiples of 4");
    DYT_CHECK_ARG(out_h == nullptr || ldo_h % 4 == 0, "gemm: fp16 stride must be a multiple of 4");
  } else {
    DYT_CHECK_ARG(out_h != nullptr && ldo_h % 4 == 0, "gemm: fp16 output missing / misaligned");
  }
  if (M == 0) return DYT_OK;

  // BN: widest tile that divides the work without a mostly-empty last column tile.
  int bn = 256;
  if (N <= 64) bn = 64;
  else if (N <= 128 || (N % 256 != 0 && N % 128 == 0)) bn = 128;
  // Wave quantisation: the persistent grid runs ceil(tiles / clusters) rounds of equal-cost tiles.
  // When the row count is known on the host (no device-side count, no fused row-dot whose slice
  // layout is tied to BN = 256) and N splits into 192-wide tiles, take them if rounds x width is
  // smaller: e.g. M = 12608, N = 768: 150 tiles of 256 = 3 rounds, 200 tiles of 192 = 3 rounds of
  // 3/4 the cost.  (A 192-wide tile pair still takes in A + B/2 = 28 KB per 384 tensor clocks.)
  // The GELU epilogues and K <= 128 keep BN = 256 with sixteen epilogue warps.
  if (bn == 256 && m_dev == nullptr && dot_w == nullptr && N % 192 == 0 && K > 128 &&
      (epi == EPI_BIAS || epi == EPI_BIAS_RELU || epi == EPI_BIAS_RESID)) {
    const long pairs = (M + 255) / 256;
    const long clusters = sm_count() / 2;
    const long r256 = (pairs * ((N + 255) / 256) + clusters - 1) / clusters;
    const long r192 = (pairs * (N / 192) + clusters - 1) / clusters;
    if (r192 * 192 * 10 <= r256 * 256 * 9) bn = 192;  // only for a clear (>= 10 %) saving
  }

  CUtensorMap ta, tb[3];
  int s = make_tmap_f16_sw128(&ta, a, static_cast<uint64_t>(M), static_cast<uint64_t>(K),
                              static_cast<uint64_t>(lda), 128);
  if (s != DYT_OK) return s;
  // B box = half a tile per CTA of the pair; the 256-wide kernel also gets boxes for the half- and
  // quarter-width sub-tiles of its last round
  const bool tail_split = bn == 256 && dot_w == nullptr && tail_split_option().load(std::memory_order_relaxed) != 0;
  for (int i = 0; i < 3; ++i) {
    const int rows = (tail_split ? bn >> i : bn) / 2;
    s = make_tmap_f16_sw128(&tb[i], w, static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                            static_cast<uint64_t>(ldw), static_cast<uint32_t>(rows));
    if (s != DYT_OK) return s;
  }

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_dev = m_dev;
  p.bias = bias;
  p.out_h = out_h; p.out_f = out_f; p.resid = resid;
  p.ldo_h = ldo_h; p.ldo_f = ldo_f; p.ld_res = ld_res;
  p.scale = scale;
  p.dot_w = dot_w; p.dot_out = dot_out; p.dot_ld = dot_ld; p.dot_f16 = dot_f16;
  p.aux = aux; p.ld_aux = ld_aux;
  p.tail_split = tail_split ? 1 : 0;
  p.reverse_m = (reverse_m & GEMM_FLAG_REVERSE) ? 1 : 0;
  p.half_grid = (reverse_m & GEMM_FLAG_HALF_GRID) ? 1 : 0;
  if (epi == EPI_BIAS_GELU_KEEP || epi == EPI_DGELU) {
    DYT_CHECK_ARG(aux != nullptr && ld_aux >= N && ld_aux % 8 == 0 && N % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(aux) & 15) == 0 && ldo_h % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(out_h) & 15) == 0,
                  "gemm: the GELU training epilogues need 16-byte aligned aux / out rows");
  }
  if (dot_w != nullptr) {
    DYT_CHECK_ARG(epi == EPI_BIAS_RESID && K > 128 && dot_out != nullptr && N % 4 == 0 &&
                      dot_ld >= gemm_tn_dot_slices(N),
                  "gemm: the fused row-dot needs the residual epilogue, K > 128 and dot_ld >= slices");
  }
  p.vec8 = (out_h != nullptr && ldo_h % 8 == 0 && N % 8 == 0 &&
            (reinterpret_cast<uintptr_t>(out_h) & 15) == 0) ? 1 : 0;
  switch (bn) {
    case 64: return dispatch_epi<64>(epi, ta, tb, p, stream);
    case 128: return dispatch_epi<128>(epi, ta, tb, p, stream);
    case 192: return dispatch_epi<192>(epi, ta, tb, p, stream);
    default: return dispatch_epi<256>(epi, ta, tb, p, stream);
  }
}

}  // namespace dyt

extern "C" int dyt_linear_f16(const void* x, int ldx, const void* w, int ldw, int M, int N, int K,
                              const int* m_dev, int epilogue, const void* bias, void* out_f16,
                              int ldo_f16, float* out_f32, int ldo_f32, const float* resid,
                              int ld_resid, float scale, void* stream) {
  return dyt::gemm_tn(static_cast<const __half*>(x), ldx, static_cast<const __half*>(w), ldw, M, N,
                      K, m_dev, epilogue, static_cast<const __half*>(bias),
                      static_cast<__half*>(out_f16), ldo_f16, out_f32, ldo_f32, resid, ld_resid,
                      scale, static_cast<cudaStream_t>(stream), nullptr, nullptr, 0, 0, nullptr, 0, 0);
}

extern "C" int dyt_linear_f16_aux(const void* x, int ldx, const void* w, int ldw, int M, int N, int K,
                                  const int* m_dev, int epilogue, const void* bias, void* out_f16,
                                  int ldo_f16, void* aux_f16, int ld_aux, void* stream) {
  if (epilogue != dyt::EPI_BIAS_GELU_KEEP && epilogue != dyt::EPI_DGELU)
    return dyt::fail(dyt::DYT_EINVAL, "linear_f16_aux: epilogue must be GELU_KEEP (4) or DGELU (5)");
  return dyt::gemm_tn(static_cast<const __half*>(x), ldx, static_cast<const __half*>(w), ldw, M, N,
                      K, m_dev, epilogue, static_cast<const __half*>(bias),
                      static_cast<__half*>(out_f16), ldo_f16, nullptr, 0, nullptr, 0, 1.0f,
                      static_cast<cudaStream_t>(stream), nullptr, nullptr, 0, 0,
                      static_cast<__half*>(aux_f16), ld_aux, 0);
}
