// Second operator surface of the reference (go/kaldibridge): kaldi_* opaque tensors and the launch_*
// conv / batch-norm entry points (include/kaldi_fp16_cnn.h), implemented on the tcgen05 GEMM.
// Reference: /root/reference/cpp/src/cgo_interface.cu, /root/reference/cpp/cuda/cnn_kernels.cu:19-320,667-828.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <string.h>

#include <map>
#include <mutex>

#include "../../include/kaldi_fp16_cnn.h"
#include "../../include/kaldi_fp16_fused.h"
#include "../../include/kaldi_fp16_ops.h"
#include "host_common.h"

using namespace kfp16;

namespace {

struct TensorFP16 {   // cgo_interface.cu:81-86
  __half* data;
  int rows;
  int cols;
  size_t size;
};

struct LossScaler {   // cgo_interface.cu:405-411
  float scale, growth_factor, backoff_factor;
  int growth_interval, steps_since_growth;
};

// context-free entry points (the reference's launch_* kernels were stateless per call) run on a context cached per
// (device, stream): concurrent callers on different streams never see each other's stream or split-K workspace
kfp16_ctx* device_ctx(cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, kfp16_ctx*> ctxs;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return nullptr; }
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(dev, stream);
  auto it = ctxs.find(key);
  if (it == ctxs.end()) {
    kfp16_ctx* c = kfp16_ctx_create(dev);
    if (!c) return nullptr;
    c->stream = stream;
    it = ctxs.emplace(key, c).first;
  }
  return it->second;
}

__global__ void f32_to_f16_k(const float* __restrict__ s, __half* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __float2half_rn(s[i]);
}
__global__ void f16_to_f32_k(const __half* __restrict__ s, float* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __half2float(s[i]);
}
__global__ void scale_k(__half* __restrict__ d, float a, size_t n) {   // cgo_interface.cu:375-380
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __float2half(__half2float(d[i]) * a);
}
__global__ void fill_k(__half* __restrict__ d, __half v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = v;
}
int blocks_for(size_t n) { size_t b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b)); }

// ---- conv1d lowering: P[(b*Tout+t), ic*K+k] = in[b, t*stride - padding + k*dilation, ic]  (zero outside)
struct Conv1dGeom { int B, T, Tout, Cin, K, stride, padding, dilation, ldp; };
__global__ void conv1d_gather_k(const __half* __restrict__ in, __half* __restrict__ P, Conv1dGeom g) {
  const size_t total = (size_t)g.B * g.Tout * g.ldp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % g.ldp);
    const size_t row = i / g.ldp;
    __half v = __float2half(0.f);
    if (col < g.Cin * g.K) {
      const int ic = col / g.K, k = col % g.K;
      const int b = (int)(row / g.Tout), t = (int)(row % g.Tout);
      const int ti = t * g.stride - g.padding + k * g.dilation;
      if (ti >= 0 && ti < g.T) v = in[((size_t)b * g.T + ti) * g.Cin + ic];
    }
    P[i] = v;
  }
}
// adjoint: gin[b, ti, ic] = sum_k dP[(b, t), ic*K+k] over the (t, k) with t*stride - padding + k*dilation == ti
__global__ void conv1d_scatter_k(const __half* __restrict__ dP, __half* __restrict__ gin, Conv1dGeom g) {
  const size_t total = (size_t)g.B * g.T * g.Cin;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ic = (int)(i % g.Cin);
    const size_t bt = i / g.Cin;
    const int b = (int)(bt / g.T), ti = (int)(bt % g.T);
    float acc = 0.f;
    for (int k = 0; k < g.K; ++k) {
      const int num = ti + g.padding - k * g.dilation;
      if (num < 0 || (num % g.stride) != 0) continue;
      const int t = num / g.stride;
      if (t >= g.Tout) continue;
      acc += __half2float(dP[((size_t)b * g.Tout + t) * g.ldp + ic * g.K + k]);
    }
    gin[i] = __float2half_rn(acc);
  }
}

// ---- batchnorm1d: per-channel statistics over rows, two passes like the reference (mean, then variance)
__global__ void bn1d_colsum_k(const __half* __restrict__ x, size_t rows, int C, const float* __restrict__ mean, float* __restrict__ out) {
  // block (32, 8): 32 channels x 8 row lanes; out[c] += sum_r (x[r,c] - mean[c])^p   (mean == nullptr: plain sum)
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  const float m = (mean && c < C) ? mean[c] : 0.f;
  if (c < C)
    for (size_t r = (size_t)blockIdx.y * 8 + threadIdx.y; r < rows; r += (size_t)gridDim.y * 8) {
      const float v = __half2float(x[r * C + c]);
      acc += mean ? (v - m) * (v - m) : v;
    }
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    atomicAdd(out + c, s);
  }
}
__global__ void bn1d_scale_k(float* v, float inv_n, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) v[c] *= inv_n;
}
// cnn_kernels.cu:281-300 (train) / 302-318 (inference)
__global__ void bn1d_finalize_k(const float* __restrict__ mean, const float* __restrict__ var, __half* running_mean,
                                __half* running_var, __half* save_mean, __half* save_invstd, float* __restrict__ use_mean,
                                float* __restrict__ use_invstd, int C, float momentum, float eps, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (training) {
    const float m = mean[c], v = var[c];
    const float invstd = rsqrtf(v + eps);
    save_mean[c] = __float2half(m);
    save_invstd[c] = __float2half(invstd);
    const float rm = __half2float(running_mean[c]), rv = __half2float(running_var[c]);
    running_mean[c] = __float2half(rm * (1 - momentum) + m * momentum);
    running_var[c] = __float2half(rv * (1 - momentum) + v * momentum);
    use_mean[c] = m; use_invstd[c] = invstd;
  } else {
    use_mean[c] = __half2float(running_mean[c]);
    use_invstd[c] = rsqrtf(__half2float(running_var[c]) + eps);
  }
}
__global__ void bn1d_apply_k(const __half* __restrict__ x, __half* __restrict__ y, size_t total, int C, const float* __restrict__ mean,
                             const float* __restrict__ invstd, const __half* __restrict__ gamma, const __half* __restrict__ beta) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float norm = (__half2float(x[i]) - mean[c]) * invstd[c];
    y[i] = __float2half(norm * __half2float(gamma[c]) + __half2float(beta[c]));
  }
}

__global__ void maxpool_fwd_k(const __half* __restrict__ in, __half* __restrict__ out, int* __restrict__ idx, int B, int T, int C,
                              int K, int stride, int Tout) {   // cnn_kernels.cu:327-362
  const size_t total = (size_t)B * Tout * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t bt = i / C;
    const int b = (int)(bt / Tout), t = (int)(bt % Tout);
    float best = -1e10f;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
      const int ti = t * stride + k;
      const float v = __half2float(in[((size_t)b * T + ti) * C + c]);
      if (v > best) { best = v; arg = ti; }
    }
    out[i] = __float2half(best);
    idx[i] = arg;
  }
}
// gather form of cnn_kernels.cu:365-386 (the reference atomically adds floats into a half buffer):
// grad_input[b, ti, c] = sum of the grad_output windows whose arg-max is ti
__global__ void maxpool_bwd_k(const __half* __restrict__ gout, const int* __restrict__ idx, __half* __restrict__ gin, int B, int T,
                              int Tout, int C) {
  const size_t total = (size_t)B * T * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t bt = i / C;
    const int b = (int)(bt / T), ti = (int)(bt % T);
    float acc = 0.f;
    for (int t = 0; t < Tout; ++t)
      if (idx[((size_t)b * Tout + t) * C + c] == ti) acc += __half2float(gout[((size_t)b * Tout + t) * C + c]);
    gin[i] = __float2half_rn(acc);
  }
}

bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

const char* kaldi_get_last_error(void) { return get_error(); }
void kaldi_clear_error(void) { clear_error(); }

void* kaldi_cublas_create(void) { return ops_cublas_create(); }
void kaldi_cublas_destroy(void* handle) { ops_cublas_destroy(handle); }
void kaldi_cublas_enable_tensor_cores(void* handle) { (void)handle; }

void* kaldi_tensor_create(int rows, int cols) {
  if (rows < 0 || cols < 0) { set_error("kaldi_tensor_create: negative shape"); return nullptr; }
  TensorFP16* t = new TensorFP16();
  t->rows = rows; t->cols = cols; t->size = (size_t)rows * cols;
  t->data = nullptr;
  if (!check_cuda(cudaMalloc(&t->data, (t->size ? t->size : 8) * sizeof(__half)), "kaldi_tensor_create")) { delete t; return nullptr; }
  return t;
}
void* kaldi_tensor_zeros(int rows, int cols) {
  TensorFP16* t = (TensorFP16*)kaldi_tensor_create(rows, cols);
  if (t && t->size) check_cuda(cudaMemsetAsync(t->data, 0, t->size * sizeof(__half), default_stream()), "kaldi_tensor_zeros");
  return t;
}
void* kaldi_tensor_ones(int rows, int cols) {
  TensorFP16* t = (TensorFP16*)kaldi_tensor_create(rows, cols);
  if (t && t->size) { fill_k<<<blocks_for(t->size), 256, 0, default_stream()>>>(t->data, __float2half(1.0f), t->size); count_launch(); check_launch("kaldi_tensor_ones"); }
  return t;
}
void kaldi_tensor_free(void* tensor) {
  if (!tensor) return;
  TensorFP16* t = (TensorFP16*)tensor;
  if (t->data) cudaFree(t->data);
  delete t;
}
int kaldi_tensor_rows(void* tensor) { return tensor ? ((TensorFP16*)tensor)->rows : 0; }
int kaldi_tensor_cols(void* tensor) { return tensor ? ((TensorFP16*)tensor)->cols : 0; }
size_t kaldi_tensor_size(void* tensor) { return tensor ? ((TensorFP16*)tensor)->size : 0; }

void kaldi_tensor_copy_from_host_fp32(void* tensor, const float* data, size_t count) {
  if (!tensor || !data) { set_error("null pointer in copy_from_host"); return; }
  TensorFP16* t = (TensorFP16*)tensor;
  if (count > t->size) count = t->size;
  if (count == 0) return;
  float* tmp = nullptr;
  cudaStream_t s = default_stream();
  if (!check_cuda(cudaMalloc(&tmp, count * sizeof(float)), "copy_from_host staging")) return;
  if (check_cuda(cudaMemcpyAsync(tmp, data, count * sizeof(float), cudaMemcpyHostToDevice, s), "copy_from_host")) {
    f32_to_f16_k<<<blocks_for(count), 256, 0, s>>>(tmp, t->data, count);
    count_launch();
    check_launch("fp32_to_fp16");
  }
  cudaStreamSynchronize(s);
  cudaFree(tmp);
}
void kaldi_tensor_copy_to_host_fp32(void* tensor, float* data, size_t count) {
  if (!tensor || !data) { set_error("null pointer in copy_to_host"); return; }
  TensorFP16* t = (TensorFP16*)tensor;
  if (count > t->size) count = t->size;
  if (count == 0) return;
  float* tmp = nullptr;
  cudaStream_t s = default_stream();
  if (!check_cuda(cudaMalloc(&tmp, count * sizeof(float)), "copy_to_host staging")) return;
  f16_to_f32_k<<<blocks_for(count), 256, 0, s>>>(t->data, tmp, count);
  count_launch();
  if (check_launch("fp16_to_fp32")) check_cuda(cudaMemcpyAsync(data, tmp, count * sizeof(float), cudaMemcpyDeviceToHost, s), "copy_to_host");
  cudaStreamSynchronize(s);
  cudaFree(tmp);
}

void kaldi_gemm(void* handle, void* A, void* B, void* C, float alpha, float beta, int transA, int transB) {
  if (!handle || !A || !B || !C) { set_error("null pointer in GEMM"); return; }
  TensorFP16 *tA = (TensorFP16*)A, *tB = (TensorFP16*)B, *tC = (TensorFP16*)C;
  const int m = transA ? tA->cols : tA->rows, k = transA ? tA->rows : tA->cols, n = transB ? tB->rows : tB->cols;
  const int kb = transB ? tB->cols : tB->rows;
  if (k != kb || tC->rows != m || tC->cols != n) { set_error("kaldi_gemm: shape mismatch (%dx%d * %dx%d -> %dx%d)", m, k, kb, n, tC->rows, tC->cols); return; }
  // the reference rounds alpha / beta to half before cublasHgemm (cgo_interface.cu:227-228)
  const float a = __half2float(__float2half(alpha)), b = __half2float(__float2half(beta));
  kfp16_gemm((kfp16_ctx*)handle, m, n, k, a, tA->data, transA, tB->data, transB, b, tC->data);
}

void kaldi_relu(void* tensor) { if (tensor) ops_relu(((TensorFP16*)tensor)->data, (int)((TensorFP16*)tensor)->size); }
void kaldi_sigmoid(void* tensor) { if (tensor) ops_sigmoid(((TensorFP16*)tensor)->data, (int)((TensorFP16*)tensor)->size); }
void kaldi_tanh(void* tensor) { if (tensor) ops_tanh_act(((TensorFP16*)tensor)->data, (int)((TensorFP16*)tensor)->size); }
void kaldi_softmax(void* tensor) { if (tensor) ops_softmax(((TensorFP16*)tensor)->data, ((TensorFP16*)tensor)->rows, ((TensorFP16*)tensor)->cols); }
void kaldi_add(void* a, void* b) {
  if (!a || !b) return;
  TensorFP16 *tA = (TensorFP16*)a, *tB = (TensorFP16*)b;
  ops_add(tA->data, tB->data, (int)(tA->size < tB->size ? tA->size : tB->size));
}
void kaldi_scale(void* tensor, float alpha) {
  if (!tensor) return;
  TensorFP16* t = (TensorFP16*)tensor;
  if (!t->size) return;
  scale_k<<<blocks_for(t->size), 256, 0, default_stream()>>>(t->data, alpha, t->size);
  count_launch();
  check_launch("kaldi_scale");
}

void* kaldi_loss_scaler_create(float initial_scale) {
  LossScaler* ls = new LossScaler();
  ls->scale = initial_scale; ls->growth_factor = 2.0f; ls->backoff_factor = 0.5f; ls->growth_interval = 2000; ls->steps_since_growth = 0;
  return ls;
}
void kaldi_loss_scaler_free(void* scaler) { delete (LossScaler*)scaler; }
float kaldi_loss_scaler_get_scale(void* scaler) { return scaler ? ((LossScaler*)scaler)->scale : 1.0f; }
void kaldi_loss_scaler_update(void* scaler, int overflow) {
  if (!scaler) return;
  LossScaler* ls = (LossScaler*)scaler;
  if (overflow) { ls->scale *= ls->backoff_factor; ls->steps_since_growth = 0; }
  else if (++ls->steps_since_growth >= ls->growth_interval) { ls->scale *= ls->growth_factor; ls->steps_since_growth = 0; }
  if (ls->scale < 1.0f) ls->scale = 1.0f;
  if (ls->scale > 65536.0f) ls->scale = 65536.0f;
}

// ----------------------------------------------------------------------------- conv1d
static bool conv1d_geom(Conv1dGeom& g, int B, int T, int Cin, int K, int stride, int padding, int dilation) {
  if (B <= 0 || T <= 0 || Cin <= 0 || K <= 0 || stride <= 0 || dilation <= 0 || padding < 0) { set_error("conv1d: bad geometry"); return false; }
  g.B = B; g.T = T; g.Cin = Cin; g.K = K; g.stride = stride; g.padding = padding; g.dilation = dilation;
  g.Tout = (T + 2 * padding - dilation * (K - 1) - 1) / stride + 1;
  g.ldp = (Cin * K + 15) & ~15;
  if (g.Tout <= 0) { set_error("conv1d: empty output"); return false; }
  return true;
}

void launch_conv1d_forward_fp16(const void* input, const void* weight, const void* bias, void* output, int batch_size,
                                int time_in, int in_channels, int out_channels, int kernel_size, int stride, int padding,
                                int dilation, void* stream) {
  if (!input || !weight || !output) { set_error("launch_conv1d_forward_fp16: null pointer"); return; }
  Conv1dGeom g;
  if (!conv1d_geom(g, batch_size, time_in, in_channels, kernel_size, stride, padding, dilation)) return;
  cudaStream_t s = (cudaStream_t)stream;
  kfp16_ctx* ctx = device_ctx(s);
  if (!ctx) return;
  const size_t rows = (size_t)g.B * g.Tout;
  const int Kd = g.Cin * g.K;
  // any channel count is accepted, as by the reference kernels (cnn_kernels.cu:19-65): shapes the TMA path cannot
  // address (Cin*K or Cout not a multiple of 8) run the same lowering on a dense patch matrix through the SIMT GEMM
  const bool tma = (Kd % 8) == 0 && (out_channels % 8) == 0 && al16(weight) && al16(output);
  if (!tma) g.ldp = Kd;
  __half* P = nullptr;
  if (!check_cuda(cudaMallocAsync(&P, rows * g.ldp * sizeof(__half), s), "conv1d patch buffer")) return;
  conv1d_gather_k<<<blocks_for(rows * g.ldp), 256, 0, s>>>((const __half*)input, P, g);
  count_launch();
  if (tma) {   // out = P * W^T (+ bias): W stored [Cout x Cin*K] is a K-major B operand
    kfp16_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = (int)rows; d.N = out_channels; d.K = g.ldp;
    d.a_major = KFP16_K_MAJOR; d.b_major = KFP16_K_MAJOR;
    d.A.ptr = P; d.A.rows = (int)rows; d.A.cols = g.ldp; d.A.ld = g.ldp;
    d.B.ptr = weight; d.B.rows = out_channels; d.B.cols = Kd; d.B.ld = Kd;   // columns [Kd, ldp) read as zeros
    d.groups = 1; d.kslabs = 1; d.kslab_len = g.ldp;
    d.D[0] = output; d.ldd = out_channels; d.alpha = 1.0f;
    if (bias) { d.flags = KFP16_EPI_BIAS; d.bias = bias; }
    kfp16_gemm_ex(ctx, &d);      // (a failure leaves its message in the thread-local error: kaldi_get_last_error)
  } else {
    // bias first (broadcast into the output rows), then C = acc + 1*C in fp32: one rounding, as the reference kernel
    if (bias && kfp16_bcast_rows(ctx, bias, out_channels, output, out_channels, 0, (int)rows, (int)rows) != 0) { cudaFreeAsync(P, s); return; }
    kfp16_gemm(ctx, (int)rows, out_channels, Kd, 1.0f, P, 0, weight, 1, bias ? 1.0f : 0.0f, output);
  }
  cudaFreeAsync(P, s);
}

void launch_conv1d_backward_fp16(const void* input, const void* grad_output, const void* weight, void* grad_input,
                                 void* grad_weight, void* grad_bias, int batch_size, int time_in, int in_channels,
                                 int out_channels, int kernel_size, int stride, int padding, int dilation, void* stream) {
  if (!grad_output) { set_error("launch_conv1d_backward_fp16: null grad_output"); return; }
  Conv1dGeom g;
  if (!conv1d_geom(g, batch_size, time_in, in_channels, kernel_size, stride, padding, dilation)) return;
  cudaStream_t s = (cudaStream_t)stream;
  kfp16_ctx* ctx = device_ctx(s);
  if (!ctx) return;
  const size_t rows = (size_t)g.B * g.Tout;
  const int Kd = g.Cin * g.K;
  const bool tma = (Kd % 8) == 0 && (out_channels % 8) == 0 && al16(grad_output) && (!weight || al16(weight)) && (!grad_weight || al16(grad_weight));
  if (!tma) g.ldp = Kd;      // dense patch matrix + SIMT GEMMs for channel counts TMA cannot address
  if (grad_bias) {   // column sum in fp32, one fp16 rounding (cnn_kernels.cu:208-229)
    float* acc = nullptr;
    if (!check_cuda(cudaMallocAsync(&acc, (size_t)out_channels * sizeof(float), s), "conv1d bias-gradient scratch")) return;
    kfp16_colsum(ctx, grad_output, out_channels, (int)rows, out_channels, acc, grad_bias);
    cudaFreeAsync(acc, s);
  }
  __half* P = nullptr;
  if (!check_cuda(cudaMallocAsync(&P, rows * g.ldp * sizeof(__half), s), "conv1d patch buffer")) return;
  if (grad_weight && input) {   // dW[oc, ic*K+k] = sum_rows gout[row, oc] * P[row, ic*K+k]
    conv1d_gather_k<<<blocks_for(rows * g.ldp), 256, 0, s>>>((const __half*)input, P, g);
    count_launch();
    if (!tma) {
      kfp16_gemm(ctx, out_channels, Kd, (int)rows, 1.0f, grad_output, 1, P, 0, 0.0f, grad_weight);
    } else {
    kfp16_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = out_channels; d.N = Kd; d.K = (int)rows;
    d.a_major = KFP16_MN_MAJOR; d.b_major = KFP16_MN_MAJOR;
    d.A.ptr = grad_output; d.A.rows = (int)rows; d.A.cols = out_channels; d.A.ld = out_channels;
    d.B.ptr = P; d.B.rows = (int)rows; d.B.cols = Kd; d.B.ld = g.ldp;
    d.groups = 1; d.kslabs = 1; d.kslab_len = (int)rows;
    d.D[0] = grad_weight; d.ldd = Kd; d.alpha = 1.0f;
    kfp16_gemm_ex(ctx, &d);
    }
  }
  if (grad_input && weight && !tma) {
    if (kfp16_gemm(ctx, (int)rows, Kd, out_channels, 1.0f, grad_output, 0, weight, 0, 0.0f, P) == 0) {
      conv1d_scatter_k<<<blocks_for((size_t)g.B * g.T * g.Cin), 256, 0, s>>>(P, (__half*)grad_input, g);
      count_launch();
      check_launch("conv1d input gradient");
    }
  } else if (grad_input && weight) {   // dP = gout * W, then the adjoint of the gather
    kfp16_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = (int)rows; d.N = Kd; d.K = out_channels;
    d.a_major = KFP16_K_MAJOR; d.b_major = KFP16_MN_MAJOR;
    d.A.ptr = grad_output; d.A.rows = (int)rows; d.A.cols = out_channels; d.A.ld = out_channels;
    d.B.ptr = weight; d.B.rows = out_channels; d.B.cols = Kd; d.B.ld = Kd;
    d.groups = 1; d.kslabs = 1; d.kslab_len = out_channels;
    d.D[0] = P; d.ldd = g.ldp; d.alpha = 1.0f;
    if (kfp16_gemm_ex(ctx, &d) == 0) {
      conv1d_scatter_k<<<blocks_for((size_t)g.B * g.T * g.Cin), 256, 0, s>>>(P, (__half*)grad_input, g);
      count_launch();
      check_launch("conv1d input gradient");
    }
  }
  cudaFreeAsync(P, s);
}

void launch_pointwise_conv1d_fp16(const void* input, const void* weight, const void* bias, void* output, int batch_size,
                                  int time_steps, int in_channels, int out_channels, void* stream) {
  launch_conv1d_forward_fp16(input, weight, bias, output, batch_size, time_steps, in_channels, out_channels, 1, 1, 0, 1, stream);
}

void launch_maxpool1d_forward_fp16(const void* input, void* output, void* indices, int batch_size, int time_in, int channels,
                                   int kernel_size, int stride, void* stream) {
  if (!input || !output || !indices || kernel_size <= 0 || stride <= 0) { set_error("launch_maxpool1d_forward_fp16: bad argument"); return; }
  const int Tout = (time_in - kernel_size) / stride + 1;
  if (Tout <= 0) return;
  maxpool_fwd_k<<<blocks_for((size_t)batch_size * Tout * channels), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)input, (__half*)output, (int*)indices, batch_size, time_in, channels, kernel_size, stride, Tout);
  count_launch();
  check_launch("launch_maxpool1d_forward_fp16");
}
void launch_maxpool1d_backward_fp16(const void* grad_output, const void* indices, void* grad_input, int batch_size, int time_in,
                                    int time_out, int channels, void* stream) {
  if (!grad_output || !indices || !grad_input) { set_error("launch_maxpool1d_backward_fp16: null pointer"); return; }
  maxpool_bwd_k<<<blocks_for((size_t)batch_size * time_in * channels), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)grad_output, (const int*)indices, (__half*)grad_input, batch_size, time_in, time_out, channels);
  count_launch();
  check_launch("launch_maxpool1d_backward_fp16");
}

void launch_batchnorm1d_forward_fp16(const void* input, const void* gamma, const void* beta, void* running_mean,
                                     void* running_var, void* output, void* save_mean, void* save_invstd, int batch_size,
                                     int time_steps, int channels, float momentum, float eps, bool training, void* stream) {
  if (!input || !gamma || !beta || !running_mean || !running_var || !output || (training && (!save_mean || !save_invstd))) {
    set_error("launch_batchnorm1d_forward_fp16: null pointer"); return;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t rows = (size_t)batch_size * time_steps;
  const int C = channels;
  if (rows == 0 || C <= 0) return;
  float* scratch = nullptr;   // mean, var, use_mean, use_invstd
  if (!check_cuda(cudaMallocAsync(&scratch, 4 * (size_t)C * sizeof(float), s), "batchnorm1d scratch")) return;
  float *mean = scratch, *var = scratch + C, *use_mean = scratch + 2 * C, *use_invstd = scratch + 3 * C;
  const int cb = (C + 255) / 256;
  if (training) {
    cudaMemsetAsync(scratch, 0, 2 * (size_t)C * sizeof(float), s);
    const dim3 blk(32, 8);
    int gy = (int)((rows + 63) / 64);
    if (gy > 148 * 4) gy = 148 * 4;
    const dim3 grd((C + 31) / 32, gy < 1 ? 1 : gy);
    bn1d_colsum_k<<<grd, blk, 0, s>>>((const __half*)input, rows, C, nullptr, mean);
    bn1d_scale_k<<<cb, 256, 0, s>>>(mean, 1.0f / (float)rows, C);
    bn1d_colsum_k<<<grd, blk, 0, s>>>((const __half*)input, rows, C, mean, var);
    bn1d_scale_k<<<cb, 256, 0, s>>>(var, 1.0f / (float)rows, C);
    count_launch(4);
  }
  bn1d_finalize_k<<<cb, 256, 0, s>>>(mean, var, (__half*)running_mean, (__half*)running_var, (__half*)save_mean,
                                     (__half*)save_invstd, use_mean, use_invstd, C, momentum, eps, training ? 1 : 0);
  bn1d_apply_k<<<blocks_for(rows * C), 256, 0, s>>>((const __half*)input, (__half*)output, rows * C, C, use_mean, use_invstd,
                                                    (const __half*)gamma, (const __half*)beta);
  count_launch(2);
  check_launch("launch_batchnorm1d_forward_fp16");
  cudaFreeAsync(scratch, s);
}

}  // extern "C"
