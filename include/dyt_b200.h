/* dyt_b200 — C ABI of the B200-native token-dispatched ViT block forward (Dynamic-Tuning / DyT).
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

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYT_ABI_VERSION 1

/* epilogues of dyt_linear_f16 */
#define DYT_EPI_BIAS 0       /* y = f16(x W^T + b)                                            */
#define DYT_EPI_BIAS_GELU 1  /* y = f16(gelu_erf(f16(x W^T + b)))     timm Mlp.fc1 + act      */
#define DYT_EPI_BIAS_RELU 2  /* y = f16(relu(f16(x W^T + b)))         Adapter.down_proj + ReLU */
#define DYT_EPI_BIAS_RESID 3 /* v = f16(x W^T + b); v = f16(v*scale) if scale != 1;
                                out_f32 = resid + v; optional out_f16 = f16(out_f32)          */

/* Library / ABI identification. */
int dyt_version(void);
const char* dyt_last_error(void);

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

#ifdef __cplusplus
}
#endif
#endif /* DYT_B200_H_ */
