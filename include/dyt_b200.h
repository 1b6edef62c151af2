/* dyt_b200 — C ABI of the B200-native token-dispatched ViT block (Dynamic-Tuning / DyT): the
 * inference forward (dyt_block_fwd and its parts), the backward of the block for parameter-efficient
 * fine-tuning, and the kernels of the rows either side of it (stem, heads, video pooling head,
 * long-sequence attention of the segmentation backbone, keep-rate / FLOPs accounting).
 *
 * The reference (NUS-HPC-AI-Lab/Dynamic-Tuning) has no FFI layer: its hot path is the timm-style
 * nn.Module surface of models/model_speed_test.py and models/vision_transformer_IN21K.py, which
 * bottoms out in ATen/cuBLAS calls.  This header is the boundary a maintainer binds instead of
 * those ATen calls (ctypes stub: see INTEGRATION.md).  Every entry point
 *   - takes raw DEVICE pointers, explicit sizes/strides and a cudaStream_t (as void*),
 *   - never allocates, frees or synchronises: all work is enqueued on the given stream, the caller
 *     owns every buffer including workspaces (CUDA-graph capturable),
 *   - returns an int status: 0 = ok, < 0 = argument / support error, > 0 = cudaError_t.
 *     dyt_last_error() returns a thread-local description of the last non-zero status.
 * No C++ exception crosses this ABI.
 *
 * dtype policy (what torch.cuda.amp.autocast() does to the reference, speed.py:254): GEMM and
 * attention operands fp16 with fp32 accumulation; LayerNorm/softmax statistics fp32; the residual
 * stream x stays fp32.
 */
#ifndef DYT_B200_H_
#define DYT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYT_ABI_VERSION 3

/* epilogues of dyt_linear_f16 */
#define DYT_EPI_BIAS 0       /* y = f16(x W^T + b)                                            */
#define DYT_EPI_BIAS_GELU 1  /* y = f16(gelu_erf(f16(x W^T + b)))     timm Mlp.fc1 + act      */
#define DYT_EPI_BIAS_RELU 2  /* y = f16(relu(f16(x W^T + b)))         Adapter.down_proj + ReLU */
#define DYT_EPI_BIAS_RESID 3 /* v = f16(x W^T + b); v = f16(v*scale) if scale != 1;
                                out_f32 = resid + v; optional out_f16 = f16(out_f32)          */

/* Library / ABI identification. */
int dyt_version(void);
const char* dyt_last_error(void);

/* Process-wide library options (thread-safe; take effect for subsequent launches).
 *   DYT_OPT_PDL (default 1): launch the forward-path kernels with programmatic stream serialization
 *   (cudaLaunchAttributeProgrammaticStreamSerialization): every kernel signals launch_dependents at
 *   its start and executes griddepcontrol.wait after its prologue (barrier init, TMEM allocation,
 *   tensor-map prefetch), so that prologue overlaps the previous kernel's tail.  CUDA-graph
 *   capturable.  0 = plain stream order.  Measured -0.08 ms per 9.1 ms step on B200.
 *   DYT_OPT_GEMM_TAIL_SPLIT (default 1): the persistent GEMM cuts the tiles of its last, partial
 *   round into 2 or 4 column sub-tiles when the leftover tiles number at most half / a quarter of
 *   the CTA pairs (wave quantisation: 297 tile pairs on 74 pairs = 4 rounds + 1 tile).
 *   DYT_OPT_FUSE_ADAPTER_UP (default 1): dyt_block_fwd computes the adapter's up projection inside
 *   the scatter-merge kernel (dyt_merge_up_fwd) instead of a GEMM launch whose [T, C] output makes a
 *   round trip through HBM; 0 = the separate dyt_linear_f16 + dyt_scatter_merge_fwd launches.
 *   DYT_OPT_FUSE_ADAPTER_DOWN (default 0, needs DYT_OPT_FUSE_ADAPTER_UP): 1 = dyt_block_fwd computes
 *   the adapter's down projection inside the merge kernel as well (dyt_adapter_merge_fwd): the proj
 *   GEMM no longer writes an fp16 copy of x1 and the down GEMM launch disappears (167 MB less HBM
 *   traffic per layer at 256 images; measured neutral on the step); 0 = down GEMM on a side stream.
 *   DYT_OPT_ATTN_SPLIT (default 1): dyt_attn_varlen_fwd runs uniform sequences of 161..256
 *   tokens on the four-stream kernel (query tile x key half, exact combine of the halves: 96 -> 89 us
 *   at 256 x 12 x 197 alone, -0.1 ms on the step).  The two key halves of a row share one maximum, so
 *   the probabilities are rounded to fp16 exactly as in the two-stream kernel (results differ only by
 *   the fp32 summation order of the PV product).  0 = two-stream kernel for every length.
 *   DYT_OPT_TILE_ORDER (bit mask, default 7): GEMMs of dyt_block_fwd that take their row tiles from
 *   the last to the first (1 = qkv, 2 = proj, 4 = fc2, 8 = fc1), so that each starts on the rows its
 *   producer wrote last (still in the L2) and ends on the rows the ascending kernel behind it reads
 *   first.  Bit-identical results; -0.08 ms on the 8.9 ms step.
 *   DYT_OPT_SIDE_PLAN (bit mask, default 0): how the adapter's down GEMM on the library's side stream
 *   shares the GPU with the dispatcher.  1 = the GEMM takes at most half of the SMs, 2 = no side stream,
 *   4 = branch launched after the dispatcher.  Bit-identical; measured equal on the step.
 *   DYT_OPT_SM_LIMIT (default 0 = all): every persistent grid is sized for at most this many SMs, so that
 *   two forwards on two streams can share the GPU side by side (experiments: two half batches on 74 SMs
 *   each are slower than one batch on 148). */
#define DYT_OPT_PDL 1
#define DYT_OPT_GEMM_TAIL_SPLIT 2
#define DYT_OPT_FUSE_ADAPTER_UP 3
#define DYT_OPT_ATTN_SPLIT 4
#define DYT_OPT_FUSE_ADAPTER_DOWN 5
#define DYT_OPT_TILE_ORDER 6
#define DYT_OPT_SIDE_PLAN 7
#define DYT_OPT_SM_LIMIT 8
int dyt_configure(int option, int value);

/* y = epilogue(x[M,K] * w[N,K]^T): the nn.Linear forward under fp16 autocast.
 * Replaces: attn.qkv / attn.proj (reference models/model_speed_test.py:147, :164),
 *           timm Mlp fc1/fc2 (models/model_speed_test.py:303 via timm.layers.Mlp),
 *           Adapter.down_proj / up_proj (models/model_speed_test.py:106-111).
 * x, w, bias, out_f16 are fp16; ld* are row strides in elements; N and K multiples of 8.
 * m_dev (optional) points to a device int32 holding the number of valid rows (<= M): the kernel
 * reads it on the device so a data-dependent row count (kept tokens) needs no host sync. */
int dyt_linear_f16(const void* x, int ldx, const void* w, int ldw, int M, int N, int K,
                   const int* m_dev, int epilogue, const void* bias, void* out_f16, int ldo_f16,
                   float* out_f32, int ldo_f32, const float* resid, int ld_resid, float scale,
                   void* stream);

/* dyt_linear_f16 with an auxiliary fp16 [M, N] operand, for the train-mode MLP:
 *   DYT_EPI_BIAS_GELU_KEEP: out = f16(gelu(f16(x W^T + b))) and aux = f16(x W^T + b) (written):
 *     fc1 + GELU keeping the pre-activation for the backward;
 *   DYT_EPI_DGELU: out = f16(f16(x W^T + b) * gelu'(aux)) (aux read): the fc2 data gradient
 *     (x = g_mlp, W = fc2.weight^T) times the GELU derivative at the saved pre-activation.
 * N > 64, rows of out / aux 16-byte aligned. */
#define DYT_EPI_BIAS_GELU_KEEP 4
#define DYT_EPI_DGELU 5
int dyt_linear_f16_aux(const void* x, int ldx, const void* w, int ldw, int M, int N, int K,
                       const int* m_dev, int epilogue, const void* bias, void* out_f16, int ldo_f16,
                       void* aux_f16, int ld_aux, void* stream);

/* Multi-head attention over packed variable-length sequences.
 * Replaces F.scaled_dot_product_attention(q, k, v) in Attention.forward (reference
 * models/model_speed_test.py:145-166 == models/vision_transformer_IN21K.py:54-75): non-causal,
 * scale = head_dim^-0.5, no dropout, q_norm/k_norm = Identity.
 * qkv: fp16 [total_tokens, 3, num_heads, head_dim] (the qkv Linear output, row stride ld_qkv);
 * sequence b covers rows [cu_seqlens[b], cu_seqlens[b+1]); cu_seqlens == NULL means num_seqs
 * sequences of uniform_len (== max_seqlen) tokens each.  out: fp16 [total_tokens, num_heads*head_dim]
 * (the layout of x.transpose(1, 2).reshape(B, N, C)).  head_dim must be 64, max_seqlen <= 256. */
int dyt_attn_varlen_fwd(const void* qkv, int ld_qkv, const int* cu_seqlens, int num_seqs,
                        int uniform_len, int max_seqlen, int total_tokens, int num_heads,
                        int head_dim, void* out, int ldo, void* stream);

/* Attention over sequences of any length with an additive per-head bias:
 *   out = softmax(f16(q * head_dim^-0.5) k^T + bias[h]) v.
 * Replaces the eager attention path of the segmentation backbone with its relative-position bias
 * (reference dense_tasks/Segmentation/backbone/segmentation_vision_transformer_IN21K.py:181-203;
 * 1025 tokens at 512 x 512).  qkv fp16 [num_seqs * seq_len, 3, num_heads, 64] (row stride ld_qkv),
 * uniform sequences; bias fp32 [num_heads, seq_len, ld_bias] or NULL: element (h, i, j) at
 * bias[(h * seq_len + i) * ld_bias + j], ld_bias >= seq_len (0 = seq_len, the contiguous
 * [H, N, N] tensor of the reference).  A pitch that is a multiple of 4 floats on a 16-byte aligned
 * base is read in 16-byte vectors (1025 tokens: pad the rows to 1028); any other pitch works through
 * scalar loads.  out fp16 [num_seqs * seq_len, num_heads * 64].  Rounding points of that path under
 * fp16 autocast: q * scale and q k^T fp16, bias add and softmax fp32, probabilities fp16 for the PV
 * product. */
int dyt_attn_bias_fwd(const void* qkv, int ld_qkv, const float* bias, int ld_bias, int num_seqs,
                      int seq_len, int num_heads, int head_dim, void* out, int ldo, void* stream);

/* LayerNorm(eps) over the last dim of fp32 rows, result rounded once to fp16 (the rounding the
 * consumer Linear applies under autocast).  Output row r reads input row row_idx[r] (or r when
 * row_idx is NULL); n_rows_dev optionally bounds the row count from device memory.
 * Replaces nn.LayerNorm norm1 / norm2 (reference models/vision_transformer_IN21K.py:110, :123). */
int dyt_layernorm_f16(const float* x, int ldx, const int* row_idx, const int* n_rows_dev,
                      int n_rows, int C, const float* gamma, const float* beta, float eps,
                      void* out_f16, int ldo, void* stream);

/* Fused token dispatcher: selector score + gate + stable compaction (+ LayerNorm2 of the kept rows
 * into the packed fp16 buffer).  No host synchronisation: counts stay on the device.
 * Replaces TokenSelect.forward / _gumbel_sigmoid (reference models/dynamic_adapter.py:25-77,
 * models/model_speed_test.py:27-60), `nonzero()` + row gather (models/model_speed_test.py:297-301)
 * and norm2 of the gathered rows (:303).
 *   x1 [B*N, ldx] fp32; sel_w [C], sel_b [1] fp32 (mlp_token_select.mlp_head).
 *   logit_fp16 = 1: fp16 autocast arithmetic (operands rounded to fp16, fp32 accumulate, logit
 *   rounded once to fp16); 0: plain fp32.
 *   keep(b,n) = gate_input >= min_kept, where min_kept is the smallest logit for which
 *   sigmoid(l) > threshold holds in the logit dtype (computed on the host from torch's sigmoid);
 *   gate_input = logit in eval, ((logit + noise1) - noise2) / tau in train mode (noise* = the two
 *   Gumbel draws of _gumbel_sigmoid, [B, N-1], NULL in eval).  Token 0 (cls) is always kept.
 *   forced_mask (optional, [B, N]) overrides the gate (BASELINE.json config 1).
 * Outputs: mask [B, N] (1/0, cls = 1; equals forced_mask when one is given), gate_out (optional,
 * [B, N]: the selector's own decision even when a mask is forced), logits [B, N-1], packed_idx [B*N] (first n_kept valid,
 * ascending flat token index == nonzero() order), token_pos [B*N] (packed position or -1),
 * cu_seqlens [B+1], n_kept [1], packed_f16 [B*N, ldp] (optional; needs ln_w/ln_b = norm2).
 * workspace: dyt_dispatch_workspace_bytes(B) bytes, 16-byte aligned, zero-filled once by the
 * caller before first use (the kernel restores the zeros it needs). */
size_t dyt_dispatch_workspace_bytes(int B);
int dyt_dispatch_fwd(const float* x1, int ldx, const float* sel_w, const float* sel_b,
                     int logit_fp16, float min_kept, const float* noise1, const float* noise2,
                     float tau, int B, int N, int C, const float* ln_w, const float* ln_b,
                     float eps, const float* forced_mask, float* mask, float* gate_out,
                     float* logits, int* packed_idx, int* token_pos, int* cu_seqlens, int* n_kept,
                     void* packed_f16, int ldp, void* workspace, void* stream);

/* TokenSelect.forward alone: score Linear(C -> 1) + gate, no compaction (reference
 * models/dynamic_adapter.py:70-77 / :25-54 forward value).  Same arithmetic, arguments and outputs
 * (mask [B, N] with cls = 1, logits [B, N-1]) as dyt_dispatch_fwd; one warp per token, so small
 * batches still fill the GPU.  Used by the train-mode forward, which needs no packed buffer.
 * row_of (optional, [B*N]): row_of[t] = t for kept tokens, -1 for dropped ones, i.e. the token_pos
 * argument of dyt_scatter_merge_fwd for a DENSE (unpacked) MLP output. */
int dyt_token_select_fwd(const float* x1, int ldx, const float* sel_w, const float* sel_b,
                         int logit_fp16, float min_kept, const float* noise1, const float* noise2,
                         float tau, int B, int N, int C, float* mask, float* logits, int* row_of,
                         void* stream);

/* Vectorised scatter-merge: out[t] = adapt[t] + (x1[t] + (token_pos[t] >= 0 ? mlp[token_pos[t]] : 0)).
 * Replaces torch.zeros + index_put + the two adds (reference models/model_speed_test.py:302-308).
 * Optionally also writes next_ln_out = f16(LayerNorm(out)) with the following norm's parameters. */
int dyt_scatter_merge_fwd(const float* x1, int ldx, const void* adapt_f16, int lda,
                          const void* mlp_packed_f16, int ldm, const int* token_pos, int n_rows,
                          int C, float* out, int ldo, const float* next_ln_w,
                          const float* next_ln_b, float eps, void* next_ln_out_f16, int ldn,
                          void* stream);

/* Adapter up-projection fused into the scatter-merge:
 *   adapt = f16(f16(down[T,K] * up_w[C,K]^T + up_b) * scale)        (stays on chip)
 *   out[t] = adapt[t] + (x1[t] + (token_pos[t] >= 0 ? mlp[token_pos[t]] : 0))
 *   next_ln_out = f16(LayerNorm(out))                                 (optional)
 * Replaces Adapter.up_proj * scale (reference models/model_speed_test.py:106-111, :291) together
 * with torch.zeros + index_put + the two adds (:302-308) and the next block's norm1.  Same results
 * as dyt_linear_f16(EPI_BIAS, scale) followed by dyt_scatter_merge_fwd.  Needs C % 128 == 0,
 * C <= 1024, K <= 64, K % 8 == 0 (DYT_EUNSUPPORTED otherwise); out must not alias x1. */
int dyt_merge_up_fwd(const void* down_f16, int ld_down, const void* up_w_f16, int ldw,
                     const void* up_b_f16, float scale, int K, const float* x1, int ldx,
                     const void* mlp_packed_f16, int ldm, const int* token_pos, int n_rows, int C,
                     float* out, int ldo, const float* next_ln_w, const float* next_ln_b, float eps,
                     void* next_ln_out_f16, int ldn, void* stream);

/* The whole adapter branch fused into the scatter-merge: like dyt_merge_up_fwd, and `down` itself is
 * computed inside the kernel from the tile's x1 rows,
 *   down = relu(f16(f16(x1) down_w[K,C]^T + down_b)),
 * so neither the fp16 copy of x1 nor `down` touches HBM and Adapter.down_proj + ReLU (reference
 * models/model_speed_test.py:106-108) needs no launch of its own.  Same restrictions as
 * dyt_merge_up_fwd plus C % 64 == 0. */
int dyt_adapter_merge_fwd(const void* down_w_f16, int ld_dw, const void* down_b_f16,
                          const void* up_w_f16, int ldw, const void* up_b_f16, float scale, int K,
                          const float* x1, int ldx, const void* mlp_packed_f16, int ldm,
                          const int* token_pos, int n_rows, int C, float* out, int ldo,
                          const float* next_ln_w, const float* next_ln_b, float eps,
                          void* next_ln_out_f16, int ldn, void* stream);

/* MoE-adapter branch (BASELINE configs[3]; DyT paper arXiv 2403.11808).  NOT in the reference
 * repository: no reference interface is replaced and there is no reference parity -- pinned to this
 * repository's own restatement (oracle/dyt_oracle.py moe_adapter).
 *   alpha[b] = softmax(router(mean_tokens(x1[b])));  W_mix = sum_i alpha_i W^i (down and up, biases too)
 *   adapt = f16(f16(relu(f16(x1 W_down_mix^T + b_down_mix)) W_up_mix^T + b_up_mix) * scale)
 * computed through the linearity in the weights: one GEMM over the concatenated experts, a per-image
 * mixture of the expert outputs, one GEMM with K' = round8(E * K + E).  x1 [B*N, C] fp32, x1_f16 its
 * fp16 copy, down_cat [E*K, C], down_b [E, K], up_cat [C, K'] (= [W_up^1 | .. | W_up^E | b_up^1 ..
 * b_up^E | 0]), router_w [E, C] fp32, router_b [E] fp32 or NULL; adapt [B*N, C] fp16. */
size_t dyt_moe_workspace_bytes(int B, int N, int E, int K);
int dyt_moe_adapter_fwd(const float* x1, int ldx, const void* x1_f16, int ldxh, int B, int N, int C,
                        int E, int K, const float* router_w, const float* router_b,
                        const void* down_cat_f16, const void* down_b_f16, const void* up_cat_f16,
                        float scale, void* adapt_f16, int ld_adapt, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ViT stem: x[b,0] = cls + pos[0]; x[b,1+p] = f16(patch_p . W^T + bias) + pos[1+p]  (fp32 out).
 * Replaces PatchEmbed.proj (Conv2d k = s = P) + cls concat + pos_embed add (reference
 * models/model_speed_test.py:467-472).  img [B, Cin, H, W] fp32; w_f16 [C, Cin*P*P] (the conv
 * weight flattened), bias_f16 [C]; cls [C], pos [(H/P)*(W/P)+1, C] fp32; x_out [B, L+1, C] fp32. */
size_t dyt_patch_embed_workspace_bytes(int B, int H, int W, int P, int Cin, int C);
int dyt_patch_embed_fwd(const float* img, int B, int Cin, int H, int W, int P, const void* w_f16,
                        const void* bias_f16, const float* cls, const float* pos, int C,
                        float* x_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- video pooling head (reference video_models/video_vision_transformer_IN21K.py:27-110, :474-481) ----
 * dyt_pool_layernorm_f16: out_k = f16(LN_k(LN_0(x))), out_v = f16(LN_v(LN_0(x))) per row: the model's
 * final `norm` followed by AttentiveBlock.norm_k / norm_v, one pass over the fp32 token stream.
 * dyt_query_attn_fwd: single-query cross attention per clip and head.  q fp16 [num_clips, H*64] (row
 * stride ldq; ldq = 0 broadcasts one query to every clip), already scaled by head_dim^-0.5;
 * k, v fp16 [num_clips * n_keys, H*64] (row stride ld_kv); out fp16 [num_clips, H*64].
 * Rounding points of CrossAttention.forward under fp16 autocast: fp16 scores, fp32 softmax, fp16
 * probabilities, fp32 accumulation, fp16 output. */
int dyt_pool_layernorm_f16(const float* x, int ldx, int n_rows, int C, const float* g0,
                           const float* b0, const float* gk, const float* bk, const float* gv,
                           const float* bv, float eps, void* out_k_f16, void* out_v_f16, int ldo,
                           void* stream);
int dyt_query_attn_fwd(const void* q_f16, int ldq, const void* k_f16, const void* v_f16, int ld_kv,
                       int num_clips, int n_keys, int num_heads, int head_dim, void* out_f16, int ldo,
                       void* stream);

/* ---- backward of the block for parameter-efficient fine-tuning ----
 * The backbone is frozen (main_image.py:242-256: only adaptmlp.*, mlp_token_select.* and head.* are
 * trained), so the backward needs data gradients through every frozen op and weight gradients for
 * the adapter, the selector and the head only.  Data gradients through a frozen Linear are
 * dyt_linear_f16 calls with the transposed weight (g_x = g_y W == g_y (W^T)^T).  Under fp16 autocast
 * the gradients of fp16 tensors are fp16, the gradient of the fp32 residual stream is fp32. */

/* g_x = resid + dLayerNorm(g_y) [+ row_scale[r] * axpy[:]]   (fp32), optional fp16 copy.
 * g_y fp16 [n_rows, C] (the dgrad GEMM output), x fp32 = the forward input of the LayerNorm
 * (statistics are recomputed).  With row_idx, gradient row r belongs to x / resid / out row
 * row_idx[r] (final norm on the cls rows; the kept rows of the student pass, whose count
 * n_rows_dev stays on the device).  resid may alias out.  row_scale/axpy add the selector's
 * data gradient g_logit[t] * mlp_head.weight in the same pass.
 * Backward of nn.LayerNorm norm1 / norm2 / norm (models/vision_transformer_IN21K.py:110, :123, :316). */
int dyt_layernorm_bwd(const void* gy_f16, int ldg, const float* x, int ldx, const int* row_idx,
                      const int* n_rows_dev, int n_rows, int C, const float* gamma, float eps,
                      const float* resid, int ldr,
                      const float* row_scale, const float* axpy, float* out, int ldo, void* out_f16,
                      int ldoh, void* stream);

/* Backward of  x = residual + token_select * mlp_x + adapt_x  (models/vision_transformer_IN21K.py:
 * 161-163) and of the straight-through hard gate (models/dynamic_adapter.py:46-51):
 *   g_f16        = f16(g_out)                                  gradient of adapt_x (and of mlp_x
 *                                                              when complete_model=True)
 *   g_masked_f16 = mask * g_f16                                gradient of mlp_x
 *   g_logit[t]   = (<g_f16[t], mlp_x[t]> + g_token_select[t]) * y(1-y) * (1/tau in train form)
 *                  + g_token_logits, 0 for the cls slot; y = sigmoid((l + n1 - n2)/tau) or sigmoid(l)
 * g_out fp32 [B*N, C]; mask [B*N]; logits / noise* / g_token_logits [B*(N-1)]; g_token_select [B*N]
 * (gradient arriving on the returned sub_token_select, may be NULL).  g_masked_f16 == NULL selects
 * the complete_model form (only g_f16 is written).
 * Sparse student backward (g_mlp is zero on dropped rows, so the frozen MLP's backward only needs
 * the kept rows): with token_pos (from dyt_dispatch_fwd) the masked gradient is written PACKED,
 * row token_pos[t] for kept tokens only; with gx_init (+ sel_w = mlp_head.weight [C]) the kernel
 * also writes gx_init[t] = g_out[t] + g_logit[t] * sel_w, the x1 gradient before the LayerNorm2
 * term that dyt_layernorm_bwd then adds on the kept rows (row_idx = packed_idx). */
int dyt_merge_bwd(const float* g_out, int ldg, const void* mlp_f16, int ldm, const float* mask,
                  const float* logits, const float* noise1, const float* noise2, float tau,
                  const float* g_token_select, const float* g_token_logits, int B, int N, int C,
                  void* g_f16, int ld16, void* g_masked_f16, int ldgm, float* g_logit,
                  const int* token_pos, const float* sel_w, float* gx_init, int ldgx, void* stream);

/* out[r] = g[r] * gelu'(pre[row_idx[r]]) for the first min(n_rows, *n_rows_dev) packed rows of H
 * fp16 columns: GELU backward of the kept rows, gathering the saved pre-activation. */
int dyt_gelu_bwd_rows(const void* g_f16, int ldg, const void* pre_f16, int ldp, const int* row_idx,
                      const int* n_rows_dev, int n_rows, int H, void* out_f16, int ldo, void* stream);

/* out_w[c] += sum_t s[t] * x[t, c];  out_b += sum_t s[t]: weight / bias gradient of the selector's
 * Linear(C -> 1) (TokenSelect.mlp_head, models/dynamic_adapter.py:60).  fp32 atomics: the caller
 * zero-fills (or carries) out_w / out_b. */
int dyt_rowscale_colsum(const float* s, const void* x_f16, int ldx, int T, int C, float* out_w,
                        float* out_b, void* stream);

/* Elementwise fp16 ops on contiguous, 16-byte aligned buffers of n elements (n % 8 == 0). */
#define DYT_EW_GELU_FWD 0      /* out = gelu_erf(a)                      timm Mlp.act (train path keeps the pre-activation) */
#define DYT_EW_GELU_BWD 1      /* out = a * gelu'(b)                     a = g_h, b = pre-activation */
#define DYT_EW_RELU_DROP_BWD 2 /* out = b > 0 ? a * c : 0  (c NULL = 1)  Adapter ReLU + dropout backward */
#define DYT_EW_MUL 3           /* out = a * b                            Adapter dropout forward (b = keep/(1-p)) */
int dyt_eltwise_f16(int op, const void* a, const void* b, const void* c, void* out, size_t n,
                    void* stream);

/* dW[Nout, Kin] += alpha * g[T, Nout]^T x[T, Kin],  db[Nout] += alpha * colsum(g)  (db may be NULL):
 * weight gradient of a trainable nn.Linear (Adapter.down_proj / up_proj, head).  g, x fp16; dW, db
 * fp32, accumulated with atomics (caller zero-fills). */
int dyt_wgrad_f16(const void* g_f16, int ldg, const void* x_f16, int ldx, int T, int Nout, int Kin,
                  float alpha, float* dW, int ldw, float* db, void* stream);

/* Backward of dyt_attn_varlen_fwd: d_qkv [total_tokens, 3, H, 64] fp16 from qkv, the forward output
 * `out` and its gradient d_out ([total_tokens, H*64] fp16).  Probabilities are recomputed (no
 * saved softmax statistics).  tcgen05 kernel: TMA-staged Q / K / V / dO, all five contractions on
 * tcgen05.mma with TMEM accumulators.  Sequences up to 256 tokens, head_dim 64. */
int dyt_attn_varlen_bwd(const void* qkv, int ld_qkv, const void* out, int ldo, const void* d_out,
                        int ld_do, const int* cu_seqlens, int num_seqs, int uniform_len,
                        int max_seqlen, int total_tokens, int num_heads, int head_dim, void* d_qkv,
                        int ld_dqkv, void* stream);

/* ---- evaluation analytics ----------------------------------------
 * Per-image FLOPs and per-layer kept-token counters of one batch, on the device.
 * token_select fp32 [B, L, Np] (the model's token_select output, cls stripped, 0/1).
 *   image_flops[b] = base_flops + (block_num - L) * table[Np + 1] + sum_l table[count(b, l) + 1]
 * (reference block_flops_dict.py:57-83, additions in the same order), optional;
 *   counters[l] += sum_b count(b, l) for l < L, counters[L] += B   (uint64, caller zero-fills;
 * replaces the padded all_gather of every mask, engine_finetune.py:245-252, :446-480), optional. */
int dyt_keep_stats(const float* token_select, int B, int L, int Np, const float* flops_table,
                   int table_len, int block_num, float base_flops, float* image_flops,
                   unsigned long long* counters, void* stream);

/* ---- whole block: Block.batch_forward (reference models/model_speed_test.py:274-310) ---------- */
typedef struct dyt_block_shape {
  int B;          /* images (sequences) */
  int N;          /* tokens per image incl. cls (197) */
  int C;          /* embed dim = 64 * H */
  int H;          /* heads */
  int hidden;     /* MLP hidden (4C) */
  int bottleneck; /* adapter width (tuning_config.ffn_num) */
} dyt_block_shape;

typedef struct dyt_block_weights { /* device pointers; *_w of Linears are fp16 [out,in], their
                                      biases fp16; LayerNorm and selector parameters fp32 */
  const float* ln1_w; const float* ln1_b;
  const void* qkv_w;  const void* qkv_b;
  const void* proj_w; const void* proj_b;
  const float* ln2_w; const float* ln2_b;
  const void* fc1_w;  const void* fc1_b;
  const void* fc2_w;  const void* fc2_b;
  const void* down_w; const void* down_b;
  const void* up_w;   const void* up_b;
  const float* sel_w; const float* sel_b;
  float adapter_scale;
} dyt_block_weights;

typedef struct dyt_block_opts {
  size_t struct_size;      /* = sizeof(dyt_block_opts) of the header the caller was built against;
                              dyt_block_fwd rejects any other value (a binding whose field list has
                              gone stale fails loudly instead of being read past its end) */
  float eps;               /* LayerNorm eps (1e-6) */
  int logit_fp16;          /* see dyt_dispatch_fwd */
  float min_kept;
  const float* noise1;     /* train-mode Gumbel draws or NULL */
  const float* noise2;
  float tau;
  const float* forced_mask;/* optional imposed mask [B, N] */
  float* gate_out;         /* optional [B, N]: selector decision ignoring forced_mask */
  int xn_ready;            /* 1: workspace already holds LN1(x) (written by the previous call) */
  const float* next_ln_w;  /* optional: also emit LayerNorm(out) for the next block / final norm */
  const float* next_ln_b;
  const float* attn_bias;  /* optional [H, N, attn_bias_ld] additive attention bias (relative position
                              bias of the segmentation backbone); sequences > 256 tokens or a bias
                              use dyt_attn_bias_fwd instead of dyt_attn_varlen_fwd */
  int attn_bias_ld;        /* row pitch of attn_bias in floats (0 = N), see dyt_attn_bias_fwd */
  /* MoE-adapter (moe_experts > 1; not in the reference repository, see dyt_moe_adapter_fwd): the
   * weights struct then holds down_w = the experts' down projections concatenated [E * K, C],
   * down_b = [E, K], up_w = [C, round8(E * K + E)] = [W_up^1 | .. | W_up^E | b_up^1 .. b_up^E | 0],
   * up_b unused, K = shape.bottleneck. */
  int moe_experts;
  const float* moe_router_w;   /* [E, C] fp32 */
  const float* moe_router_b;   /* [E] fp32 or NULL */
  void* moe_workspace;         /* dyt_moe_workspace_bytes(B, N, E, K) bytes, 256-byte aligned */
  size_t moe_workspace_bytes;
} dyt_block_opts;

typedef struct dyt_block_buffers { /* where dyt_block_fwd keeps its intermediates (for tests) */
  void* xn; void* attn_o; void* qkv; float* x1; void* x1h; void* packed; void* hidden; void* mlp;
  void* down; void* adapt; int* packed_idx; int* token_pos; int* cu_seqlens; int* n_kept;
} dyt_block_buffers;

size_t dyt_block_workspace_bytes(const dyt_block_shape* shape);
int dyt_block_workspace_layout(const dyt_block_shape* shape, void* workspace, dyt_block_buffers* out);
/* x [B, N, C] fp32 is updated in place; mask_out [B, N]; logits_out [B, N-1].  The workspace
 * (256-byte aligned, dyt_block_workspace_bytes) must be zero-filled once before first use. */
int dyt_block_fwd(const dyt_block_shape* shape, const dyt_block_weights* weights,
                  const dyt_block_opts* opts, float* x, float* mask_out, float* logits_out,
                  void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYT_B200_H_ */
