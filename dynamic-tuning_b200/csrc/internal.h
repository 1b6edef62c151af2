// Internal (non-ABI) entry points shared between translation units of libdyt_b200.so.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace dyt {

int gemm_tn(const __half* a, int lda, const __half* w, int ldw, int M, int N, int K,
            const int* m_dev, int epi, const __half* bias, __half* out_h, int ldo_h, float* out_f,
            int ldo_f, const float* resid, int ld_res, float scale, cudaStream_t stream,
            const float* dot_w = nullptr, float* dot_out = nullptr, int dot_ld = 0, int dot_f16 = 0,
            __half* aux = nullptr, int ld_aux = 0, int reverse_m = 0);
// reverse_m: launch flags GEMM_FLAG_REVERSE / GEMM_FLAG_HALF_GRID (host_utils.h)
// number of per-row partial sums the fused row-dot of a residual-epilogue GEMM writes (N columns)
int gemm_tn_dot_slices(int N);

int attn_varlen_fwd(const __half* qkv, int ld_qkv, const int* cu_seqlens, int num_seqs,
                    int uniform_len, int max_seqlen, int total_tokens, int num_heads, int head_dim,
                    __half* out, int ldo, cudaStream_t stream);

// four-stream attention for uniform sequences of 161..256 tokens (attn_split.cu)
bool attn_split_supported(const int* cu_seqlens, int uniform_len, int head_dim);
int attn_split_fwd(const __half* qkv, int ld_qkv, int num_seqs, int seq_len, int total_tokens,
                   int num_heads, __half* out, int ldo, cudaStream_t stream);

int layernorm_f16(const float* x, int ldx, const int* row_idx, const int* n_rows_dev, int n_rows,
                  int C, const float* gamma, const float* beta, float eps, __half* out, int ldo,
                  cudaStream_t stream);

int scatter_merge(const float* x1, int ldx, const __half* adapt, int lda, const __half* mlp_packed,
                  int ldm, const int* token_pos, int n_rows, int C, float* out, int ldo,
                  const float* nln_w, const float* nln_b, float eps, __half* nln_out, int ldn,
                  cudaStream_t stream);

// adapter up-projection + scatter-merge + next LayerNorm in one kernel (merge_up.cu)
bool merge_up_supported(int C, int K, long long rows);
int merge_up(const __half* down, int ld_down, const __half* up_w, int ldw, const __half* up_b,
             float scale, int K, const float* x1, int ldx, const __half* mlp_packed, int ldm,
             const int* token_pos, int n_rows, int C, float* out, int ldo, const float* nln_w,
             const float* nln_b, float eps, __half* nln_out, int ldn, cudaStream_t stream,
             const __half* down_w = nullptr, int ld_dw = 0, const __half* down_b = nullptr);

// MoE-adapter branch (moe.cu; not in the reference: own oracle, no reference parity)
size_t moe_workspace_bytes(int B, int N, int E, int K);
int moe_adapter_fwd(const float* x1, int ldx, const __half* x1h, int ldxh, int B, int N, int C, int E,
                    int K, const float* router_w, const float* router_b, const __half* down_cat,
                    const __half* down_b, const __half* up_cat, float scale, __half* adapt, int ld_adapt,
                    void* ws, size_t ws_bytes, cudaStream_t stream);

int dispatch_fwd(const float* x1, int ldx, const float* sel_w, const float* sel_b, int logit_fp16,
                 float min_kept, const float* noise1, const float* noise2, float tau, int B, int N,
                 int C, const float* ln_w, const float* ln_b, float eps, const float* forced_mask,
                 float* mask, float* gate_out, float* logits, int* packed_idx, int* token_pos,
                 int* cu_seqlens, int* n_kept, void* packed_f16, int ldp, void* workspace,
                 void* stream, const float* partials, int n_partials);

}  // namespace dyt
