/*
 * acnn.h -- C ABI of libacnn.so: the B200 (sm_100a) kernels behind the assembled-ResNet hot path.
 *
 * The reference (clovaai/assembled-cnn) has no FFI layer: every op below is a TensorFlow-1.14 graph
 * op emitted by the reference's Python.  Each entry point cites the reference call site it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain C: pointers + sizes only; device pointers are caller-owned (the Python host passes
 *     torch tensors' data_ptr()); the library never allocates device memory.
 *   - every call returns 0 on success, non-zero on failure; acnn_last_error() gives the message.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*): no hidden
 *     synchronisation, CUDA-graph capturable.
 *   - activations are NHWC, bf16 unless stated; "raw" = conv output before batch-norm.
 *   - conv weights are OHWI: [Cout][kh][kw][Cin] (TF's HWIO permuted).
 *   - per-channel float vectors (scale/shift/sum/...) are fp32, length C.
 */
#ifndef ACNN_H_
#define ACNN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACNN_OK 0
#define ACNN_ERR_INVALID 1
#define ACNN_ERR_CUDA 2
#define ACNN_ERR_UNSUPPORTED 3

/* Library / environment ------------------------------------------------------------------- */
const char* acnn_last_error(void);
int acnn_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t acnn_launch_count(void);

/* Convolution geometry (correlation, no bias) -- nets/model_helper.py:67-78 conv2d_fixed_padding
 * + fixed_padding :40-64.  Ho = (H + pad_h_lo + pad_h_hi - kh) / stride + 1, same for W. */
typedef struct acnn_conv_geom {
  int32_t B, H, W, Cin;
  int32_t Cout, kh, kw, stride;
  int32_t pad_h_lo, pad_h_hi, pad_w_lo, pad_w_hi;
} acnn_conv_geom;

/* y[B,Ho,Wo,Cout] = conv(x[B,H,W,Cin], w[Cout,kh,kw,Cin])  as a tcgen05 implicit GEMM
 * (tf.layers.conv2d: nets/model_helper.py:74-78; tf.layers.dense: nets/resnet_model.py:595-597
 * when kh=kw=H=W=1).  Optional fused epilogue, applied in this order:
 *   + bias[Cout] (fp32)            -> dense bias
 *   + add_src[B,Ho,Wo,Cout] (bf16) -> gradient accumulation across consumers
 *   * (mask_src > 0)               -> ReLU backward of the tensor this gradient belongs to
 *   ch_sum/ch_sumsq[Cout] += column sums of the bf16-rounded output (batch-norm statistics,
 *                             nets/model_helper.py:34-37; buffers must be zeroed by the caller)
 * out_f32 != 0 stores y as fp32 (logits), else bf16.  Requires Cin % 16 == 0, Cout % 32 == 0. */
int acnn_conv_fprop(const acnn_conv_geom* g, const void* x, const void* w, void* y,
                    float* ch_sum, float* ch_sumsq, const void* add_src, const void* mask_src,
                    const float* bias, int out_f32, void* stream);

/* dx[B,H,W,Cin] = conv_transpose(dy[B,Ho,Wo,Cout]) for a stride-1 conv of geometry g (the backward
 * of tf.layers.conv2d the reference gets from tf.gradients, nets/optimizer_setting.py:30).
 * w_dgrad is [Cin][kh][kw][Cout] with taps already flipped (acnn_prep_weights writes it).
 * Same optional add_src / mask_src epilogue as acnn_conv_fprop (shapes of dx). */
int acnn_conv_dgrad(const acnn_conv_geom* g, const void* dy, const void* w_dgrad, void* dx,
                    const void* add_src, const void* mask_src, void* stream);

/* dw[Cout,kh,kw,Cin] (fp32) += sum_pixels x (*) dy  -- weight gradient, split-K over pixels with
 * fp32 atomics; dw must be zeroed (or hold the running sum) by the caller. */
int acnn_conv_wgrad(const acnn_conv_geom* g, const void* x, const void* dy, float* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACNN_H_ */
