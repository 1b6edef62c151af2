// One DyT block forward = the launch sequence of the sm_100a kernels, all on the caller's stream,
// no host synchronisation and no allocation (the caller provides the workspace).
// Mirrors Block.batch_forward of the reference (models/model_speed_test.py:274-310):
//   x1  = x + proj(attn(qkv(LN1(x))))                    :278
//   mask, logits = TokenSelect(x1)                        :284     (fused dispatcher)
//   adapt = up(relu(down(x1))) * scale                    :291     (all tokens)
//   mlp_packed = fc2(gelu(fc1(LN2(x1[kept]))))            :297-304 (kept tokens only)
//   x   = adapt + (x1 + scatter(mlp_packed))              :305-308
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/dyt_b200.h"
#include "gemm_tn.cuh"
#include "host_utils.h"
#include "internal.h"

namespace dyt {

static inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
constexpr int kMaxBatch = 16384;  // fixes the size of the dispatcher's (always-zero) workspace slot

struct BlockWorkspace {
  __half* xn;       // [T, C]   LN1(x)
  __half* attn_o;   // [T, C]   attention output
  __half* qkv;      // [T, 3C]
  float* x1;        // [T, C]   fp32 residual stream after attention
  __half* x1h;      // [T, C]   fp16 copy of x1 (adapter A operand)
  __half* packed;   // [T, C]   LN2 of kept rows (capacity T)
  __half* hidden;   // [T, hidden]
  __half* mlp;      // [T, C]   packed MLP output
  __half* down;     // [T, bottleneck_padded]
  __half* adapt;    // [T, C]
  int* packed_idx;  // [T]
  int* token_pos;   // [T]
  int* cu_seqlens;  // [B+1]
  int* n_kept;      // [1]
  float* score_part;  // [T, slices] selector score partial sums from the proj epilogue
  void* dispatch_ws;
  size_t total;
};

static BlockWorkspace carve(const dyt_block_shape* s, void* base) {
  const size_t T = static_cast<size_t>(s->B) * s->N;
  const size_t C = s->C;
  char* p = static_cast<char*>(base);
  size_t off = 0;
  BlockWorkspace w;
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += align256(bytes);
    return r;
  };
  // first, so its address is the same for every shape sharing one workspace (it must stay zeroed)
  w.dispatch_ws = take(dyt_dispatch_workspace_bytes(kMaxBatch));
  w.xn = static_cast<__half*>(take(T * C * 2));
  w.attn_o = static_cast<__half*>(take(T * C * 2));
  w.qkv = static_cast<__half*>(take(T * 3 * C * 2));
  w.x1 = static_cast<float*>(take(T * C * 4));
  w.x1h = static_cast<__half*>(take(T * C * 2));
  w.packed = static_cast<__half*>(take(T * C * 2));
  w.hidden = static_cast<__half*>(take(T * static_cast<size_t>(s->hidden) * 2));
  w.mlp = static_cast<__half*>(take(T * C * 2));
  w.down = static_cast<__half*>(take(T * static_cast<size_t>(s->bottleneck) * 2));
  w.adapt = static_cast<__half*>(take(T * C * 2));
  w.score_part = static_cast<float*>(take(T * static_cast<size_t>(gemm_tn_dot_slices(s->C)) * 4));
  w.packed_idx = static_cast<int*>(take(T * 4));
  w.token_pos = static_cast<int*>(take(T * 4));
  w.cu_seqlens = static_cast<int*>(take((static_cast<size_t>(s->B) + 1) * 4));
  w.n_kept = static_cast<int*>(take(256));
  w.total = off;
  return w;
}

// The adapter branch (down + ReLU, up * scale: two short, store-bound GEMMs) depends only on x1, like
// the dispatcher -> fc1 -> fc2 chain: it runs on a side stream forked after the proj GEMM and joined
// before the scatter-merge, so its HBM traffic overlaps the dispatcher's (fork / join by events:
// capturable into a CUDA graph like everything else).  One stream + two events per host thread and
// device, created on first use (no device memory).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool tried = false, ok = false;
};
static SideStream& side_stream() {
  static thread_local SideStream per_device[kMaxDevices];   // streams / events belong to a device
  SideStream& s = per_device[current_device()];
  if (!s.tried) {
    s.tried = true;
    s.ok = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
    // on failure the adapter branch simply stays on the caller's stream
  }
  return s;
}

static int check_shape(const dyt_block_shape* s) {
  DYT_CHECK_ARG(s != nullptr, "block: null shape");
  DYT_CHECK_ARG(s->B >= 1 && s->B <= kMaxBatch && s->N >= 2, "block: bad B=%d N=%d", s->B, s->N);
  DYT_CHECK_ARG(s->H >= 1 && s->C == s->H * 64, "block: C must be 64*H (C=%d H=%d)", s->C, s->H);
  DYT_CHECK_ARG(s->hidden % 8 == 0 && s->bottleneck % 8 == 0 && s->bottleneck >= 8,
                "block: hidden/bottleneck must be multiples of 8");
  return DYT_OK;
}

}  // namespace dyt

extern "C" size_t dyt_block_workspace_bytes(const dyt_block_shape* shape) {
  if (dyt::check_shape(shape) != 0) return 0;
  return dyt::carve(shape, nullptr).total;
}

extern "C" int dyt_block_workspace_layout(const dyt_block_shape* shape, void* workspace,
                                          dyt_block_buffers* out) {
  using namespace dyt;
  int st = check_shape(shape);
  if (st != 0) return st;
  DYT_CHECK_ARG(out != nullptr, "block: null layout output");
  BlockWorkspace w = carve(shape, workspace);
  out->xn = w.xn; out->attn_o = w.attn_o; out->qkv = w.qkv; out->x1 = w.x1; out->x1h = w.x1h;
  out->packed = w.packed; out->hidden = w.hidden; out->mlp = w.mlp; out->down = w.down;
  out->adapt = w.adapt; out->packed_idx = w.packed_idx; out->token_pos = w.token_pos;
  out->cu_seqlens = w.cu_seqlens; out->n_kept = w.n_kept;
  return DYT_OK;
}

extern "C" int dyt_block_fwd(const dyt_block_shape* shape, const dyt_block_weights* wt,
                             const dyt_block_opts* opt, float* x, float* mask_out,
                             float* logits_out, void* workspace, size_t workspace_bytes,
                             void* stream_) {
  using namespace dyt;
  int st = check_shape(shape);
  if (st != 0) return st;
  DYT_CHECK_ARG(wt && opt && x && mask_out && logits_out && workspace, "block: null argument");
  DYT_CHECK_ARG(opt->struct_size == sizeof(dyt_block_opts),
                "block: dyt_block_opts.struct_size is %zu, this library expects %zu (ABI version %d): "
                "the caller's binding does not match include/dyt_b200.h",
                opt->struct_size, sizeof(dyt_block_opts), DYT_ABI_VERSION);
  DYT_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                "block: workspace must be 256-byte aligned");
  BlockWorkspace w = carve(shape, workspace);
  DYT_CHECK_ARG(workspace_bytes >= w.total, "block: workspace too small (%zu < %zu)",
                workspace_bytes, w.total);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int B = shape->B, N = shape->N, C = shape->C, H = shape->H;
  const int T = B * N;
  const __half* h = nullptr;
  (void)h;
#define DYT_TRY(call)        \
  do {                       \
    int s_ = (call);         \
    if (s_ != 0) return s_;  \
  } while (0)
#define HP(p) static_cast<const __half*>(p)

  NvtxRange range_block("dyt_block_fwd");
  // 1. LN1 (skipped when the previous block's merge already produced it)
  if (!opt->xn_ready) {
    NvtxRange r("dyt.ln1");
    DYT_TRY(layernorm_f16(x, C, nullptr, nullptr, T, C, wt->ln1_w, wt->ln1_b, opt->eps, w.xn, C,
                          stream));
  }
  // 2. qkv
  const int tile_order = tile_order_option().load(std::memory_order_relaxed);
  { NvtxRange r("dyt.qkv");
  DYT_TRY(gemm_tn(w.xn, C, HP(wt->qkv_w), C, T, 3 * C, C, nullptr, EPI_BIAS, HP(wt->qkv_b), w.qkv,
                  3 * C, nullptr, 0, nullptr, 0, 1.0f, stream, nullptr, nullptr, 0, 0, nullptr, 0,
                  (tile_order & 1) ? GEMM_FLAG_REVERSE : 0)); }
  // 3. attention (uniform sequences of N tokens): the tcgen05 kernel up to 256 tokens; longer
  //    sequences or an additive bias (segmentation backbone, 1025 tokens) take the flash-style kernel
  { NvtxRange r("dyt.attention");
  if (N > 256 || opt->attn_bias != nullptr)
    DYT_TRY(dyt_attn_bias_fwd(w.qkv, 3 * C, opt->attn_bias, opt->attn_bias_ld, B, N, H, 64, w.attn_o, C,
                              stream_));
  else
    DYT_TRY(attn_varlen_fwd(w.qkv, 3 * C, nullptr, B, N, N, T, H, 64, w.attn_o, C, stream));
  }
  // 4. proj + residual -> x1 (fp32) and its fp16 copy
  // (the token selector's score Linear rides in the epilogue: per-row partial dot products)
  const int slices = gemm_tn_dot_slices(C);
  const bool fuse_score = C > 128;
  // the adapter's up projection rides in the merge kernel (step 10) unless unsupported / switched
  // off; with the down projection fused too, the whole adapter branch does: no fp16 copy of x1, no
  // `down` buffer, no side stream
  const bool moe = opt->moe_experts > 1;
  if (moe) {
    DYT_CHECK_ARG(opt->moe_router_w != nullptr && opt->moe_workspace != nullptr,
                  "block: MoE-adapter needs the router weights and a workspace");
  }
  const bool fuse_up = !moe && merge_up_supported(C, shape->bottleneck, T) &&
                       fuse_up_option().load(std::memory_order_relaxed) != 0;
  const bool fuse_down = fuse_up && C % 64 == 0 &&
                         fuse_down_option().load(std::memory_order_relaxed) != 0;
  { NvtxRange r("dyt.proj_residual_score");
  DYT_TRY(gemm_tn(w.attn_o, C, HP(wt->proj_w), C, T, C, C, nullptr, EPI_BIAS_RESID,
                  HP(wt->proj_b), fuse_down ? nullptr : w.x1h, C, w.x1, C, x, C, 1.0f, stream,
                  fuse_score ? wt->sel_w : nullptr, w.score_part, slices, opt->logit_fp16, nullptr, 0,
                  (tile_order & 2) ? GEMM_FLAG_REVERSE : 0)); }
  // adapter on every token (steps 8./9.), forked onto the side stream
  SideStream& ss = side_stream();
  cudaStream_t astream = stream;
  const int side_plan = side_plan_option().load(std::memory_order_relaxed);
  const bool fork = ss.ok && !fuse_down && !(side_plan & 2);
  if (fork) {   // the branch depends on proj only, wherever its kernels are launched below
    DYT_CUDA(cudaEventRecord(ss.fork, stream));
    DYT_CUDA(cudaStreamWaitEvent(ss.stream, ss.fork, 0));
    astream = ss.stream;
  }
  auto adapter_branch = [&]() -> int {
    if (moe) {
      NvtxRange r("dyt.moe_adapter");
      DYT_TRY(moe_adapter_fwd(w.x1, C, w.x1h, C, B, N, C, opt->moe_experts, shape->bottleneck,
                              opt->moe_router_w, opt->moe_router_b, HP(wt->down_w), HP(wt->down_b),
                              HP(wt->up_w), wt->adapter_scale, w.adapt, C, opt->moe_workspace,
                              opt->moe_workspace_bytes, astream));
    }
    if (!fuse_down && !moe) { NvtxRange r("dyt.adapter_down");
    DYT_TRY(gemm_tn(w.x1h, C, HP(wt->down_w), C, T, shape->bottleneck, C, nullptr, EPI_BIAS_RELU,
                    HP(wt->down_b), w.down, shape->bottleneck, nullptr, 0, nullptr, 0, 1.0f, astream, nullptr,
                    nullptr, 0, 0, nullptr, 0, (fork && (side_plan & 1)) ? GEMM_FLAG_HALF_GRID : 0)); }
    if (!fuse_up && !moe) {
      NvtxRange r("dyt.adapter_up");
      DYT_TRY(gemm_tn(w.down, shape->bottleneck, HP(wt->up_w), shape->bottleneck, T, C,
                      shape->bottleneck, nullptr, EPI_BIAS, HP(wt->up_b), w.adapt, C, nullptr, 0,
                      nullptr, 0, wt->adapter_scale, astream));
    }
    if (fork) DYT_CUDA(cudaEventRecord(ss.join, ss.stream));
    return DYT_OK;
  };
  if (!(side_plan & 4)) DYT_TRY(adapter_branch());
  // 5. dispatcher: score, gate, compaction, LN2 of kept rows
  { NvtxRange r("dyt.dispatch");
  DYT_TRY(dispatch_fwd(w.x1, C, wt->sel_w, wt->sel_b, opt->logit_fp16, opt->min_kept,
                       opt->noise1, opt->noise2, opt->tau, B, N, C, wt->ln2_w, wt->ln2_b,
                       opt->eps, opt->forced_mask, mask_out, opt->gate_out, logits_out, w.packed_idx,
                       w.token_pos, w.cu_seqlens, w.n_kept, w.packed, C, w.dispatch_ws, stream_,
                       fuse_score ? w.score_part : nullptr, slices)); }
  if (side_plan & 4) DYT_TRY(adapter_branch());   // experiment: the adapter branch is launched after the dispatcher
  // 6./7. MLP on the kept rows only (row count read from device memory)
  { NvtxRange r("dyt.mlp_kept_rows");
  DYT_TRY(gemm_tn(w.packed, C, HP(wt->fc1_w), C, T, shape->hidden, C, w.n_kept, EPI_BIAS_GELU,
                  HP(wt->fc1_b), w.hidden, shape->hidden, nullptr, 0, nullptr, 0, 1.0f, stream, nullptr,
                  nullptr, 0, 0, nullptr, 0, (tile_order & 8) ? GEMM_FLAG_REVERSE : 0));
  DYT_TRY(gemm_tn(w.hidden, shape->hidden, HP(wt->fc2_w), shape->hidden, T, C, shape->hidden,
                  w.n_kept, EPI_BIAS, HP(wt->fc2_b), w.mlp, C, nullptr, 0, nullptr, 0, 1.0f, stream,
                  nullptr, nullptr, 0, 0, nullptr, 0, (tile_order & 4) ? GEMM_FLAG_REVERSE : 0)); }
  // join the adapter branch
  if (fork) DYT_CUDA(cudaStreamWaitEvent(stream, ss.join, 0));
  // 10. scatter-merge back to [B, N, C] (in place into x), optionally with the next LayerNorm
  NvtxRange range_merge("dyt.merge");
  if (fuse_up)
    DYT_TRY(merge_up(w.down, shape->bottleneck, HP(wt->up_w), shape->bottleneck, HP(wt->up_b),
                     wt->adapter_scale, shape->bottleneck, w.x1, C, w.mlp, C, w.token_pos, T, C, x, C,
                     opt->next_ln_w, opt->next_ln_b, opt->eps, opt->next_ln_w ? w.xn : nullptr, C,
                     stream, fuse_down ? HP(wt->down_w) : nullptr, C, HP(wt->down_b)));
  else
    DYT_TRY(scatter_merge(w.x1, C, w.adapt, C, w.mlp, C, w.token_pos, T, C, x, C, opt->next_ln_w,
                          opt->next_ln_b, opt->eps, opt->next_ln_w ? w.xn : nullptr, C, stream));
#undef DYT_TRY
#undef HP
  return DYT_OK;
}
