/* kaldi_fp16_ops.h -- drop-in C ABI for the reference's operator library.
 *
 * Every symbol below has the name, argument order and error convention (0 / -1, message via
 * ops_last_error) of the declaration it replaces in /root/reference/cpp/include/ops.h; the
 * citation after each prototype is that declaration.  Callers: the Go files of internal/gpu (cgo, LDFLAGS
 * -lkaldi_fp16) -- see INTEGRATION.md.  All matrices are FP16, row-major, device pointers.
 */
#ifndef KALDI_FP16_B200_OPS_H
#define KALDI_FP16_B200_OPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* handle: here a kfp16 context (device, stream, workspace) instead of a cublasHandle_t */
void *ops_cublas_create(void);          /* ops.h:24 */
void ops_cublas_destroy(void *handle);  /* ops.h:25 */

/* C[MxN] = alpha*A[MxK]*B[KxN] + beta*C, FP32 accumulate; lda/ldb/ldc are ignored exactly as in
 * the reference (dense K, N, N) -- ops.h:34-40, ops.cu:366-400 */
int ops_gemm(void *handle, int M, int N, int K, float alpha, const void *A, int lda, const void *B,
             int ldb, float beta, void *C, int ldc);
/* batched, element strides between problems -- ops.h:45-53 */
int ops_gemm_strided(void *handle, int M, int N, int K, float alpha, const void *A, int lda,
                     int64_t strideA, const void *B, int ldb, int64_t strideB, float beta, void *C,
                     int ldc, int64_t strideC, int batch_count);

/* in-place activations on `count` halves -- ops.h:59-62 */
int ops_relu(void *data, int count);
int ops_sigmoid(void *data, int count);
int ops_tanh_act(void *data, int count);
int ops_clipped_relu(void *data, int count, float ceiling);

/* per-row softmax / log-softmax, in place -- ops.h:69-70 */
int ops_softmax(void *data, int rows, int cols);
int ops_log_softmax(void *data, int rows, int cols);

/* inference batch-norm, x[T x D] in place, FP32 stats -- ops.h:82-90 */
int ops_batchnorm_forward(void *x, int T, int D, const float *mean, const float *var,
                          const float *gamma, const float *beta, float epsilon);
int ops_batchnorm_forward_rms(void *x, int T, int D, const float *mean, const float *var,
                              float target_rms, float epsilon);

/* dst = alpha*src + beta*dst ; dst += src ; dst = src ; dst = val -- ops.h:97-107 */
int ops_add_scaled(void *dst, const void *src, int count, float alpha, float beta);
int ops_add(void *dst, const void *src, int count);
int ops_copy(void *dst, const void *src, int count);
int ops_fill(void *dst, int count, float val);

/* dst[t, off:off+src_cols] = src[t,:]  -- ops.h:114-116 */
int ops_concat_cols(void *dst, int T, int dst_cols, const void *src, int src_cols,
                    int dst_col_offset);
/* dst[t,:] = src[t, off:off+dst_cols] -- ops.h:185-186 */
int ops_slice_cols(const void *src, int T, int src_cols, void *dst, int dst_cols,
                   int src_col_offset);
/* [T x (H*F1 + H*F2)] -> [T x H*(F1+F2)], in place -- ops.h:139-140 */
int ops_combine_feature_maps(void *data, int T, int total_dim, int height, int num_filters1,
                             int num_filters2);
/* dst[r,:] = src[row_offset + r*stride, :] -- ops.h:16-18 (void: reports nothing, as upstream) */
void ops_subsample_rows(void *dst, const void *src, int in_rows, int cols, int stride,
                        int row_offset);

/* backward ops -- ops.h:160-183 */
int ops_relu_backward(const void *x, void *grad, int count);
int ops_sigmoid_backward(const void *output, void *grad, int count);
int ops_tanh_backward(const void *output, void *grad, int count);
int ops_transpose(const void *src, void *dst, int M, int N);
int ops_batchnorm_backward(const void *grad_out, void *grad_in, const float *gamma,
                           const float *variance, float eps, int rows, int cols);
int ops_fp16_to_fp32(const void *src, float *dst, int count);
int ops_sgd_update(float *w_fp32, void *w_fp16, const void *grad_fp16, float *velocity, float lr,
                   float momentum, int count);

/* thread-local message of the last failure, NULL when clear -- ops.h:146-147.  Unlike the
 * reference, failures of the backward ops are reported here too (upstream writes them to a
 * buffer no getter exposes, backward_wrappers.cu:25-35). */
const char *ops_last_error(void);
void ops_clear_error(void);

#ifdef __cplusplus
}
#endif
#endif /* KALDI_FP16_B200_OPS_H */
