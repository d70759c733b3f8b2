/* kaldi_fp16_cnn.h -- the go/kaldibridge operator surface (second CGO boundary of the reference):
 *   kaldi_*   opaque-tensor API of /root/reference/cpp/src/cgo_interface.cu (bound by go/kaldibridge/bridge.go:15-53)
 *   launch_*  conv / batch-norm launchers of /root/reference/cpp/include/cnn_fp16.h:24-166
 *             (bound by go/kaldibridge/cnn_bridge.go:13-71 and internal/gpu/backward_ops.go:19-28)
 * Same names, argument order and void / pointer return conventions; `stream` is a cudaStream_t passed as
 * void* so that the header stays plain C.  The matrix work (kaldi_gemm, conv1d forward / input gradient /
 * weight gradient, pointwise conv) runs on the tcgen05 GEMM of kaldi_fp16_fused.h; pooling kernels the
 * CNN-TDNN path does not use (avg / stats pooling, depthwise conv, layernorm, SE block) are out of scope.
 */
#ifndef KALDI_FP16_CNN_H
#define KALDI_FP16_CNN_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- cgo_interface.cu:20-31 */
const char *kaldi_get_last_error(void);
void kaldi_clear_error(void);
/* ---- cgo_interface.cu:54-76: the "cuBLAS handle" is a kfp16 context (device + stream) */
void *kaldi_cublas_create(void);
void kaldi_cublas_destroy(void *handle);
void kaldi_cublas_enable_tensor_cores(void *handle); /* no-op: tcgen05 is the only path */
/* ---- cgo_interface.cu:81-145: opaque struct {__half* data; int rows, cols; size_t size;} */
void *kaldi_tensor_create(int rows, int cols);
void *kaldi_tensor_zeros(int rows, int cols);
void *kaldi_tensor_ones(int rows, int cols);
void kaldi_tensor_free(void *tensor);
int kaldi_tensor_rows(void *tensor);
int kaldi_tensor_cols(void *tensor);
size_t kaldi_tensor_size(void *tensor);
/* fp32 host <-> fp16 device, round to nearest even on the device (cgo_interface.cu:149-200) */
void kaldi_tensor_copy_from_host_fp32(void *tensor, const float *data, size_t count);
void kaldi_tensor_copy_to_host_fp32(void *tensor, float *data, size_t count);
/* C = alpha*op(A)*op(B) + beta*C, row-major, alpha/beta rounded to fp16 as the reference does
 * (cgo_interface.cu:206-243; cublasHgemm there, FP32 accumulation here) */
void kaldi_gemm(void *handle, void *A, void *B, void *C, float alpha, float beta, int transA, int transB);
void kaldi_relu(void *tensor);
void kaldi_sigmoid(void *tensor);
void kaldi_tanh(void *tensor);
void kaldi_softmax(void *tensor); /* per row */
void kaldi_add(void *a, void *b); /* a += b */
void kaldi_scale(void *tensor, float alpha);
/* host-side dynamic loss scaler (cgo_interface.cu:405-449): x0.5 on overflow, x2 every 2000 clean steps, [1, 65536] */
void *kaldi_loss_scaler_create(float initial_scale);
void kaldi_loss_scaler_free(void *scaler);
float kaldi_loss_scaler_get_scale(void *scaler);
void kaldi_loss_scaler_update(void *scaler, int overflow);

/* ---- cnn_fp16.h:24-56: input [B,T,Cin], weight [Cout,Cin,K], bias [Cout] or NULL, output [B,Tout,Cout],
 * Tout = (T + 2*padding - dilation*(K-1) - 1)/stride + 1.  Lowered to a patch gather + one GEMM: the tcgen05 / TMA kernel
 * when Cin*K and Cout are multiples of 8, a dense SIMT GEMM otherwise (any channel count is accepted, like the reference's
 * kernels).  These are void functions: a failure leaves its message in kaldi_get_last_error(). */
void launch_conv1d_forward_fp16(const void *input, const void *weight, const void *bias, void *output,
                                int batch_size, int time_in, int in_channels, int out_channels,
                                int kernel_size, int stride, int padding, int dilation, void *stream);
/* grad_input [B,T,Cin], grad_weight [Cout,Cin,K], grad_bias [Cout]; any of the three may be NULL.
 * (The reference's weight-gradient kernel atomically adds floats into a half buffer, cnn_kernels.cu:204;
 *  this computes the mathematically intended gradient.) */
void launch_conv1d_backward_fp16(const void *input, const void *grad_output, const void *weight,
                                 void *grad_input, void *grad_weight, void *grad_bias, int batch_size,
                                 int time_in, int in_channels, int out_channels, int kernel_size,
                                 int stride, int padding, int dilation, void *stream);
/* cnn_fp16.h:62-86 (internal/gpu/backward_ops.go:19-28 links the backward one) */
void launch_maxpool1d_forward_fp16(const void *input, void *output, void *indices, int batch_size,
                                   int time_in, int channels, int kernel_size, int stride, void *stream);
void launch_maxpool1d_backward_fp16(const void *grad_output, const void *indices, void *grad_input,
                                    int batch_size, int time_in, int time_out, int channels, void *stream);
/* cnn_fp16.h:104-120: train / inference batch-norm over [B*T] per channel, all parameters FP16 */
void launch_batchnorm1d_forward_fp16(const void *input, const void *gamma, const void *beta,
                                     void *running_mean, void *running_var, void *output,
                                     void *save_mean, void *save_invstd, int batch_size, int time_steps,
                                     int channels, float momentum, float eps, bool training, void *stream);
/* cnn_fp16.h:153-163: 1x1 conv = [B*T x Cin] * [Cout x Cin]^T + bias */
void launch_pointwise_conv1d_fp16(const void *input, const void *weight, const void *bias, void *output,
                                  int batch_size, int time_steps, int in_channels, int out_channels,
                                  void *stream);

#ifdef __cplusplus
}
#endif
#endif /* KALDI_FP16_CNN_H */
