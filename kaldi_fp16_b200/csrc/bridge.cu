// Memory / transfer bridge: device init, allocation, pinned host buffers, H2D / D2H copies and the
// packed minibatch buffer.  Same entry points, return conventions and section layout as the
// reference's cpp/cuda/bridge.cu:30-334 (declared in cpp/include/bridge.h:12-60), which
// internal/gpu/bridge.go and tensor.go bind through cgo.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/kaldi_fp16_bridge.h"
#include "host_common.h"

namespace {

thread_local char g_bridge_err[512] = {0};

void bridge_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_bridge_err, sizeof(g_bridge_err), fmt, ap);
  va_end(ap);
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

int copy_checked(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, const char* what, size_t count) {
  if (bytes == 0) return 0;
  if (!dst || !src) { bridge_set_error("%s (%zu): null pointer", what, count); return -1; }
  // stream-ordered on the library's stream, then waited for: same blocking behaviour as the
  // reference's cudaMemcpy (bridge.cu:126-176) without serialising against other streams
  cudaStream_t s = kfp16::default_stream();
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { bridge_set_error("%s (%zu): %s", what, count, cudaGetErrorString(e)); return -1; }
  return 0;
}

__global__ void k_f16_to_f32(float* __restrict__ dst, const __half* __restrict__ src, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __half2float(src[i]);
}
__global__ void k_f32_to_f16(__half* __restrict__ dst, const float* __restrict__ src, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2half_rn(src[i]);
}
int conv_grid(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  return (int)(b ? b : 1);
}

}  // namespace

extern "C" {

const char* bridge_last_error(void) { return g_bridge_err[0] ? g_bridge_err : nullptr; }
void bridge_clear_error(void) { g_bridge_err[0] = 0; }

int bridge_gpu_init(int device_id) {
  cudaError_t e = cudaSetDevice(device_id);
  if (e != cudaSuccess) { bridge_set_error("cudaSetDevice(%d): %s", device_id, cudaGetErrorString(e)); return -1; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device_id);
  if (e != cudaSuccess) { bridge_set_error("cudaGetDeviceProperties(%d): %s", device_id, cudaGetErrorString(e)); return -1; }
  if (prop.major != 10) {
    bridge_set_error("device %d is sm_%d%d; this library only carries sm_100a code (no fallback)", device_id, prop.major, prop.minor);
    return -1;
  }
  return 0;
}
int bridge_gpu_get_free_memory(size_t* free_bytes, size_t* total_bytes) {
  cudaError_t e = cudaMemGetInfo(free_bytes, total_bytes);
  if (e != cudaSuccess) { bridge_set_error("cudaMemGetInfo: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
int bridge_gpu_sync(void) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { bridge_set_error("cudaDeviceSynchronize: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
void* bridge_gpu_malloc(size_t bytes) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) { bridge_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
  return p;
}
void bridge_gpu_free(void* ptr) { if (ptr) cudaFree(ptr); }
void* bridge_host_alloc(size_t bytes) {
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) { bridge_set_error("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
  return p;
}
void bridge_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

int bridge_transfer_fp16(void* dst_device, const uint16_t* src_host, size_t count) {
  return copy_checked(dst_device, src_host, count * sizeof(uint16_t), cudaMemcpyHostToDevice, "cudaMemcpy FP16 H2D", count);
}
int bridge_read_fp16(uint16_t* dst_host, const void* src_device, size_t count) {
  return copy_checked(dst_host, src_device, count * sizeof(uint16_t), cudaMemcpyDeviceToHost, "cudaMemcpy FP16 D2H", count);
}
int bridge_transfer_int32(void* dst_device, const int32_t* src_host, size_t count) {
  return copy_checked(dst_device, src_host, count * sizeof(int32_t), cudaMemcpyHostToDevice, "cudaMemcpy int32 H2D", count);
}
int bridge_transfer_float32(void* dst_device, const float* src_host, size_t count) {
  return copy_checked(dst_device, src_host, count * sizeof(float), cudaMemcpyHostToDevice, "cudaMemcpy float32 H2D", count);
}
int bridge_read_float32(float* dst_host, const void* src_device, size_t count) {
  return copy_checked(dst_host, src_device, count * sizeof(float), cudaMemcpyDeviceToHost, "cudaMemcpy float32 D2H", count);
}

// One allocation, six 256-byte aligned sections:
//   [features fp16 | ivectors fp16 | csr row_ptr i32 | csr col_idx i32 | csr labels i32 | csr weights f32]
int bridge_batch_alloc(int total_frames, int feat_dim, int batch_size, int ivec_dim, int num_states, int num_arcs,
                       GPUBatchPtrs* out) {
  if (!out) { bridge_set_error("bridge_batch_alloc: null output"); return -1; }
  memset(out, 0, sizeof(*out));
  if (total_frames < 0 || feat_dim < 0 || batch_size < 0 || ivec_dim < 0 || num_states < 0 || num_arcs < 0) {
    bridge_set_error("bridge_batch_alloc: negative size"); return -1;
  }
  out->features_bytes = align256((size_t)total_frames * feat_dim * sizeof(uint16_t));
  out->ivectors_bytes = align256((size_t)batch_size * ivec_dim * sizeof(uint16_t));
  out->csr_rowptr_bytes = align256((size_t)(num_states + 1) * sizeof(int32_t));
  out->csr_colidx_bytes = align256((size_t)num_arcs * sizeof(int32_t));
  out->csr_labels_bytes = align256((size_t)num_arcs * sizeof(int32_t));
  out->csr_weights_bytes = align256((size_t)num_arcs * sizeof(float));
  out->total_bytes = out->features_bytes + out->ivectors_bytes + out->csr_rowptr_bytes + out->csr_colidx_bytes +
                     out->csr_labels_bytes + out->csr_weights_bytes;
  cudaError_t e = cudaMalloc(&out->d_buffer, out->total_bytes);
  if (e != cudaSuccess) {
    bridge_set_error("cudaMalloc combined (%zu bytes): %s", out->total_bytes, cudaGetErrorString(e));
    memset(out, 0, sizeof(*out));
    return -1;
  }
  char* p = (char*)out->d_buffer;
  out->d_features = p;    p += out->features_bytes;
  out->d_ivectors = p;    p += out->ivectors_bytes;
  out->d_csr_row_ptr = p; p += out->csr_rowptr_bytes;
  out->d_csr_col_idx = p; p += out->csr_colidx_bytes;
  out->d_csr_labels = p;  p += out->csr_labels_bytes;
  out->d_csr_weights = p;
  return 0;
}
int bridge_batch_transfer(const GPUBatchPtrs* ptrs, const void* host_buf, size_t total_bytes) {
  if (!ptrs || !ptrs->d_buffer) { bridge_set_error("bridge_batch_transfer: batch not allocated"); return -1; }
  if (total_bytes > ptrs->total_bytes) {
    bridge_set_error("bridge_batch_transfer: %zu bytes do not fit the %zu-byte batch buffer", total_bytes, ptrs->total_bytes);
    return -1;
  }
  return copy_checked(ptrs->d_buffer, host_buf, total_bytes, cudaMemcpyHostToDevice, "cudaMemcpy batch", total_bytes);
}
void bridge_batch_free(GPUBatchPtrs* ptrs) {
  if (ptrs && ptrs->d_buffer) {
    cudaFree(ptrs->d_buffer);
    memset(ptrs, 0, sizeof(*ptrs));
  }
}
void bridge_gpu_memset(void* ptr, int value, size_t bytes) {
  if (ptr && bytes) cudaMemsetAsync(ptr, value, bytes, kfp16::default_stream());
}
int bridge_fp16_to_fp32_gpu(float* dst_device, const void* src_device, size_t count) {
  if (count == 0) return 0;
  if (!dst_device || !src_device) { bridge_set_error("fp16_to_fp32 kernel: null pointer"); return -1; }
  k_f16_to_f32<<<conv_grid(count), 256, 0, kfp16::default_stream()>>>(dst_device, (const __half*)src_device, count);
  kfp16::count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { bridge_set_error("fp16_to_fp32 kernel: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
int bridge_fp32_to_fp16_gpu(void* dst_device, const float* src_device, size_t count) {
  if (count == 0) return 0;
  if (!dst_device || !src_device) { bridge_set_error("fp32_to_fp16 kernel: null pointer"); return -1; }
  k_f32_to_f16<<<conv_grid(count), 256, 0, kfp16::default_stream()>>>((__half*)dst_device, src_device, count);
  kfp16::count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { bridge_set_error("fp32_to_fp16 kernel: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // extern "C"
