// HBM-bound elementwise / data-movement operators of the reference surface
// (/root/reference/cpp/cuda/ops.cu:26-320, backward_wrappers.cu:41-142) plus the fused helpers the
// layer executor uses.  The reference runs every one of them as one scalar __half element per
// thread; here each thread moves 16 bytes per access (8 halves) and the grid is sized as a
// multiple of the SM count with a grid-stride loop, so the kernels run at HBM speed.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/kaldi_fp16_fused.h"
#include "../../include/kaldi_fp16_ops.h"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace kfp16 {

constexpr int kThreads = 256;
// (griddep_wait / griddep_launch -- programmatic dependent launch -- and dropout_uniform come from sm100_ptx.cuh)

static int num_sms_cached() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// grid for `work` thread-items: enough CTAs to cover the work, capped at 8 resident CTAs per SM
static int grid_for(size_t work) {
  size_t blocks = (work + kThreads - 1) / kThreads;
  const size_t cap = (size_t)num_sms_cached() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

struct alignas(16) Half8 {
  __half2 v[4];
};

// one 16-byte access (a member-wise copy of the __half2 array compiles to four 4-byte accesses)
__device__ __forceinline__ Half8 ld8(const __half* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  Half8 r;
  *reinterpret_cast<uint4*>(&r) = u;
  return r;
}
__device__ __forceinline__ void st8(__half* p, const Half8& x) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&x); }

__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Generic unary / binary map: op(float)->float on every element, fp16 storage.
//   body:  n8 vectors of 8 halves handled with 16-byte accesses, then a scalar tail.
template <class Op>
__global__ void map1_kernel(__half* __restrict__ x, size_t n, Op op) {
  const size_t n8 = n >> 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    Half8 a = ld8(x + i * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) a.v[j] = __halves2half2(op(a.v[j].x), op(a.v[j].y));
    st8(x + i * 8, a);
  }
  for (size_t i = (n8 << 3) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = op(x[i]);
}
template <class Op>
__global__ void map1_scalar_kernel(__half* __restrict__ x, size_t n, Op op) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = op(x[i]);
}
// dst = op(dst, src)
template <class Op>
__global__ void map2_kernel(__half* __restrict__ dst, const __half* __restrict__ src, size_t n, Op op) {
  const size_t n8 = n >> 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    Half8 a = ld8(dst + i * 8);
    const Half8 b = ld8(src + i * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) a.v[j] = __halves2half2(op(a.v[j].x, b.v[j].x), op(a.v[j].y, b.v[j].y));
    st8(dst + i * 8, a);
  }
  for (size_t i = (n8 << 3) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = op(dst[i], src[i]);
}
template <class Op>
__global__ void map2_scalar_kernel(__half* __restrict__ dst, const __half* __restrict__ src, size_t n, Op op) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = op(dst[i], src[i]);
}

template <class Op>
static int run_map1(void* data, long long count, Op op, const char* what) {
  if (count <= 0) return 0;
  if (!data) { set_error("%s: null pointer", what); return -1; }
  cudaStream_t s = default_stream();
  if (al16(data)) map1_kernel<<<grid_for(((size_t)count + 7) / 8), kThreads, 0, s>>>((__half*)data, (size_t)count, op);
  else map1_scalar_kernel<<<grid_for((size_t)count), kThreads, 0, s>>>((__half*)data, (size_t)count, op);
  count_launch();
  return check_launch(what) ? 0 : -1;
}
template <class Op>
static int run_map2(void* dst, const void* src, long long count, Op op, const char* what) {
  if (count <= 0) return 0;
  if (!dst || !src) { set_error("%s: null pointer", what); return -1; }
  cudaStream_t s = default_stream();
  if (al16(dst) && al16(src))
    map2_kernel<<<grid_for(((size_t)count + 7) / 8), kThreads, 0, s>>>((__half*)dst, (const __half*)src, (size_t)count, op);
  else
    map2_scalar_kernel<<<grid_for((size_t)count), kThreads, 0, s>>>((__half*)dst, (const __half*)src, (size_t)count, op);
  count_launch();
  return check_launch(what) ? 0 : -1;
}

// ---- functors (semantics: the cited reference kernels)
struct ReluOp {   // ops.cu:26-37: only x < 0 is rewritten, so NaN and -0 pass through
  __device__ __half operator()(__half x) const { return __hlt(x, __float2half(0.f)) ? __float2half(0.f) : x; }
};
struct SigmoidOp {  // ops.cu:39-47
  __device__ __half operator()(__half x) const { return __float2half(1.0f / (1.0f + expf(-__half2float(x)))); }
};
struct TanhOp {  // ops.cu:49-57
  __device__ __half operator()(__half x) const { return __float2half(tanhf(__half2float(x))); }
};
struct ClippedReluOp {  // ops.cu:59-68
  float ceiling;
  __device__ __half operator()(__half x) const { return __float2half(fmaxf(0.0f, fminf(__half2float(x), ceiling))); }
};
struct FillOp {
  __half v;
  __device__ __half operator()(__half) const { return v; }
};
struct AddScaledOp {  // ops.cu:207-217   dst = alpha*src + beta*dst
  float alpha, beta;
  __device__ __half operator()(__half d, __half s) const {
    return __float2half(alpha * __half2float(s) + beta * __half2float(d));
  }
};
struct AddOp {  // ops.cu:219-228
  __device__ __half operator()(__half d, __half s) const { return __float2half(__half2float(d) + __half2float(s)); }
};
// backward_wrappers.cu:41-49 : dst = grad, src = saved activation
struct ReluBwdOp {
  __device__ __half operator()(__half g, __half x) const { return __hgt(x, __float2half(0.f)) ? g : __float2half(0.f); }
};
// backward_wrappers.cu:51-61 : half arithmetic, rounded after every multiply like the reference
struct SigmoidBwdOp {
  __device__ __half operator()(__half g, __half o) const {
    return __hmul(__hmul(g, o), __hsub(__float2half(1.0f), o));
  }
};
struct TanhBwdOp {  // backward_wrappers.cu:63-73
  __device__ __half operator()(__half g, __half o) const {
    return __hmul(g, __hsub(__float2half(1.0f), __hmul(o, o)));
  }
};

// ---- per-column affine transform  y = x*scale[d] + shift[d]  (batch-norm family)
// rows x cols, in place or out of place; cols % 8 == 0 and 16B alignment give the vector path.
__global__ void colscale_kernel(const __half* __restrict__ x, __half* __restrict__ y, size_t rows, int cols,
                                const float* __restrict__ mean, const float* __restrict__ var,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                float target_rms, int mode) {
  // mode 0: gamma*(x-mean)/sqrt(var+eps)+beta   (ops.cu:171-187)
  // mode 1: (x-mean)/sqrt(var+eps)*target_rms   (ops.cu:191-204)
  // mode 2: x*gamma/sqrt(var+eps)               (backward_wrappers.cu:104-115)
  const size_t total8 = rows * (size_t)cols / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
    const int d0 = (int)((i * 8) % (size_t)cols);
    Half8 a = ld8(x + i * 8);
    __half* h = reinterpret_cast<__half*>(&a);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int d = d0 + j;
      const float v = __half2float(h[j]);
      float r;
      if (mode == 0) r = gamma[d] * ((v - mean[d]) / sqrtf(var[d] + eps)) + beta[d];
      else if (mode == 1) r = ((v - mean[d]) / sqrtf(var[d] + eps)) * target_rms;
      else r = v * (gamma[d] / sqrtf(var[d] + eps));
      h[j] = __float2half(r);
    }
    st8(y + i * 8, a);
  }
}
__global__ void colscale_scalar_kernel(const __half* __restrict__ x, __half* __restrict__ y, size_t total, int cols,
                                       const float* __restrict__ mean, const float* __restrict__ var,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                       float target_rms, int mode) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int d = (int)(i % (size_t)cols);
    const float v = __half2float(x[i]);
    float r;
    if (mode == 0) r = gamma[d] * ((v - mean[d]) / sqrtf(var[d] + eps)) + beta[d];
    else if (mode == 1) r = ((v - mean[d]) / sqrtf(var[d] + eps)) * target_rms;
    else r = v * (gamma[d] / sqrtf(var[d] + eps));
    y[i] = __float2half(r);
  }
}
static int run_colscale(const void* x, void* y, long long rows, int cols, const float* mean, const float* var,
                        const float* gamma, const float* beta, float eps, float rms, int mode, const char* what) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!x || !y) { set_error("%s: null pointer", what); return -1; }
  cudaStream_t s = default_stream();
  const size_t total = (size_t)rows * cols;
  if ((cols % 8) == 0 && al16(x) && al16(y))
    colscale_kernel<<<grid_for(total / 8), kThreads, 0, s>>>((const __half*)x, (__half*)y, (size_t)rows, cols, mean, var, gamma, beta, eps, rms, mode);
  else
    colscale_scalar_kernel<<<grid_for(total), kThreads, 0, s>>>((const __half*)x, (__half*)y, total, cols, mean, var, gamma, beta, eps, rms, mode);
  count_launch();
  return check_launch(what) ? 0 : -1;
}

// ---- strided 2-D block copy: dst[r, dcol0 + c] = src[row0 + r*rstride, scol0 + c]
// covers concat_cols, slice_cols and subsample_rows (ops.cu:241-254, 290-320)
__global__ void copy2d_kernel(__half* __restrict__ dst, long long ldd, int dcol0, const __half* __restrict__ src,
                              long long lds, int scol0, long long row0, int rstride, long long rows, int cols,
                              int vec, int accumulate) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const int c8 = cols >> 3;
    const size_t total = (size_t)rows * c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const long long r = (long long)(i / c8);
      const int c = (int)(i % c8) * 8;
      Half8 v = ld8(src + (row0 + r * rstride) * lds + scol0 + c);
      if (accumulate) {
        const Half8 o = ld8(dst + r * ldd + dcol0 + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = __half22float2(v.v[j]), b = __half22float2(o.v[j]);
          v.v[j] = __floats2half2_rn(a.x + b.x, a.y + b.y);
        }
      }
      st8(dst + r * ldd + dcol0 + c, v);
    }
  } else {
    const size_t total = (size_t)rows * cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const long long r = (long long)(i / cols);
      const int c = (int)(i % cols);
      __half v = src[(row0 + r * rstride) * lds + scol0 + c];
      if (accumulate) v = __float2half_rn(__half2float(v) + __half2float(dst[r * ldd + dcol0 + c]));
      dst[r * ldd + dcol0 + c] = v;
    }
  }
}
static int run_copy2d(void* dst, long long ldd, int dcol0, const void* src, long long lds, int scol0, long long row0,
                      int rstride, long long rows, int cols, const char* what, cudaStream_t stream, int accumulate = 0) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!dst || !src) { set_error("%s: null pointer", what); return -1; }
  const int vec = (cols % 8 == 0) && (ldd % 8 == 0) && (lds % 8 == 0) && (dcol0 % 8 == 0) && (scol0 % 8 == 0) && al16(dst) && al16(src);
  const size_t work = vec ? (size_t)rows * (cols / 8) : (size_t)rows * cols;
  copy2d_kernel<<<grid_for(work), kThreads, 0, stream>>>((__half*)dst, ldd, dcol0, (const __half*)src, lds, scol0, row0, rstride, rows, cols, vec, accumulate);
  count_launch();
  return check_launch(what) ? 0 : -1;
}

// ---- combine_feature_maps (ops.cu:258-287): per row, [H*F1 | H*F2] -> H x (F1+F2); the row is
// staged in shared memory so the permutation runs in place without the reference's temp buffer.
__global__ void combine_fm_kernel(__half* __restrict__ data, int T, int total_dim, int height, int nf1, int nf2,
                                  int inverse, int rows_per_pass) {
  // rows_per_pass rows per pass (one pair of block barriers per pass instead of per row: 9984 rows of 240 halves took 10.4 us)
  extern __shared__ __half srow[];
  const int tf = nf1 + nf2;
  for (int t0 = blockIdx.x * rows_per_pass; t0 < T; t0 += gridDim.x * rows_per_pass) {
    const int nr = min(rows_per_pass, T - t0);
    __half* rows = data + (size_t)t0 * total_dim;        // the nr rows are contiguous
    for (int i = threadIdx.x; i < nr * total_dim; i += blockDim.x) srow[i] = rows[i];
    __syncthreads();
    for (int d = threadIdx.x; d < total_dim; d += blockDim.x) {
      const int h = d / tf, f = d % tf;
      const int s = (f < nf1) ? h * nf1 + f : height * nf1 + h * nf2 + (f - nf1);
      for (int r = 0; r < nr; ++r) {
        if (inverse) rows[r * total_dim + s] = srow[r * total_dim + d]; else rows[r * total_dim + d] = srow[r * total_dim + s];
      }
    }
    __syncthreads();
  }
}

// ---- softmax / log-softmax, one warp-group (128 threads) per row, fp32 math (ops.cu:70-166).
// The reference finds the row max with atomicMax on the float bit pattern, which is wrong for rows
// whose maximum is negative (the result becomes inf/NaN); this computes the true maximum.
template <bool LOG>
__global__ void softmax_kernel(__half* __restrict__ data, int rows, int cols) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    __half* row = data + (size_t)r * cols;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < cols; j += blockDim.x) m = fmaxf(m, __half2float(row[j]));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffff, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < nw; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int j = threadIdx.x; j < cols; j += blockDim.x) s += expf(__half2float(row[j]) - m);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffff, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w];
    __syncthreads();
    if (LOG) {
      const float lse = m + logf(s);
      for (int j = threadIdx.x; j < cols; j += blockDim.x) row[j] = __float2half(__half2float(row[j]) - lse);
    } else {
      // the reference rounds exp(x-max) to fp16 before normalising (ops.cu:96-110)
      const float inv = 1.0f / s;
      for (int j = threadIdx.x; j < cols; j += blockDim.x)
        row[j] = __float2half(__half2float(__float2half(expf(__half2float(row[j]) - m))) * inv);
    }
  }
}

// Same arithmetic with the row held in registers: one 16-byte read and one 16-byte write per 8 elements (rows of up
// to 8192 columns; the output layer has 6016).  256 threads x up to 4 chunks of 8 halves.
template <bool LOG>
__global__ void __launch_bounds__(256)
softmax_row_regs_kernel(__half* __restrict__ data, int rows, int cols) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = cols >> 3;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    uint4* row = reinterpret_cast<uint4*>(data + (size_t)r * cols);
    uint4 v[4];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = threadIdx.x + k * 256;
      if (j < chunks) v[k] = row[j];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (threadIdx.x + k * 256 < chunks) {
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = unpack_h2(w[e]); m = fmaxf(m, fmaxf(f.x, f.y)); }
      }
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffff, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (threadIdx.x + k * 256 < chunks) {
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = unpack_h2(w[e]); s += expf(f.x - m) + expf(f.y - m); }
      }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffff, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    __syncthreads();
    const float lse = m + logf(s), inv = 1.0f / s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = threadIdx.x + k * 256;
      if (j < chunks) {
        uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_h2(w[e]);
          if (LOG) w[e] = pack_h2(f.x - lse, f.y - lse);
          else     // the reference rounds exp(x-max) to fp16 before normalising (ops.cu:96-110)
            w[e] = pack_h2(__half2float(__float2half(expf(f.x - m))) * inv, __half2float(__float2half(expf(f.y - m))) * inv);
        }
        row[j] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

// ---- transpose dst[c*rows + r] = src[r*cols + c]  (backward_wrappers.cu:75-85), 32x32 smem tiles
__global__ void transpose_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int rows, int cols) {
  __shared__ __half tile[32][34];
  const int tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  for (long long t = blockIdx.x; t < (long long)tiles_c * tiles_r; t += gridDim.x) {
    const int tr = (int)(t / tiles_c) * 32, tc = (int)(t % tiles_c) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = tr + i, c = tc + threadIdx.x;
      if (r < rows && c < cols) tile[i][threadIdx.x] = src[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = tc + i, r = tr + threadIdx.x;
      if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
    __syncthreads();
  }
}

__global__ void f16_to_f32_kernel(const __half* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __half2float(src[i]);
}
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2half_rn(src[i]);
}

// dst = half(src * *scale_dev): the flat FP32 gradient bucket -> the FP16 gradient tensors the reference keeps
// (16-byte loads, 8-byte stores; 6 B per element against the measured HBM copy rate)
__global__ void __launch_bounds__(256)
scale_f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n4, size_t n, const float* __restrict__ scale_dev) {
  griddep_launch();
  griddep_wait();
  const float sc = scale_dev ? *scale_dev : 1.0f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = tid; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    const __half2 lo = __floats2half2_rn(v.x * sc, v.y * sc), hi = __floats2half2_rn(v.z * sc, v.w * sc);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo);
    o.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  for (size_t i = n4 * 4 + tid; i < n; i += stride) dst[i] = __float2half_rn(src[i] * sc);
}

__global__ void scale_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, float c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * c;
}
__global__ void bump_counter_kernel(uint32_t* c) { *c += 1u; }

// ---- SGD with momentum on FP32 master weights (backward_wrappers.cu:129-142):
//   g = float(grad); v = m*v + g; w32 -= lr*v; w16 = half(w32)
// One launch covers a whole flat parameter bucket (the reference launches once per tensor).
template <bool GRAD_F32>
__device__ __forceinline__ void sgd_one(float& w, float& v, float g, int round_grad, float grad_scale, float lr, float mom) {
  g *= grad_scale;
  if (GRAD_F32 && round_grad) g = __half2float(__float2half_rn(g));
  v = mom * v + g;
  w = w - lr * v;
}
template <bool GRAD_F32>
__global__ void sgd_kernel(float* __restrict__ w32, __half* __restrict__ w16, const void* __restrict__ grad,
                           int round_grad, float grad_scale, float* __restrict__ vel, float lr, float mom, size_t n,
                           const float* __restrict__ hp) {
  griddep_launch();
  griddep_wait();
  if (hp) { lr = hp[0]; mom = hp[1]; grad_scale = hp[2]; }   // hyper-parameters live in device memory (graph replays see updates)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = GRAD_F32 ? reinterpret_cast<const float*>(grad)[i] : __half2float(reinterpret_cast<const __half*>(grad)[i]);
    float v = vel[i], w = w32[i];
    sgd_one<GRAD_F32>(w, v, g, round_grad, grad_scale, lr, mom);
    vel[i] = v;
    w32[i] = w;
    w16[i] = __float2half_rn(w);
  }
}
// 4 parameters per thread and two such groups in flight: 16-byte loads/stores on the fp32 arrays (20 B/parameter total)
template <bool GRAD_F32>
__global__ void __launch_bounds__(256)
sgd_kernel_v4(float* __restrict__ w32, __half* __restrict__ w16, const void* __restrict__ grad,
              int round_grad, float grad_scale, float* __restrict__ vel, float lr, float mom, size_t n4,
              const float* __restrict__ hp) {
  griddep_launch();
  griddep_wait();
  if (hp) { lr = hp[0]; mom = hp[1]; grad_scale = hp[2]; }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    float4 w[2], v[2], g[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const size_t i = i0 + u * stride;
      ok[u] = i < n4;
      if (ok[u]) {
        w[u] = reinterpret_cast<const float4*>(w32)[i];
        v[u] = reinterpret_cast<const float4*>(vel)[i];
        if (GRAD_F32) {
          g[u] = reinterpret_cast<const float4*>(grad)[i];
        } else {
          const uint2 h = reinterpret_cast<const uint2*>(grad)[i];
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
          g[u] = make_float4(a.x, a.y, b.x, b.y);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      const size_t i = i0 + u * stride;
      sgd_one<GRAD_F32>(w[u].x, v[u].x, g[u].x, round_grad, grad_scale, lr, mom);
      sgd_one<GRAD_F32>(w[u].y, v[u].y, g[u].y, round_grad, grad_scale, lr, mom);
      sgd_one<GRAD_F32>(w[u].z, v[u].z, g[u].z, round_grad, grad_scale, lr, mom);
      sgd_one<GRAD_F32>(w[u].w, v[u].w, g[u].w, round_grad, grad_scale, lr, mom);
      reinterpret_cast<float4*>(vel)[i] = v[u];
      reinterpret_cast<float4*>(w32)[i] = w[u];
      const __half2 lo = __floats2half2_rn(w[u].x, w[u].y), hi = __floats2half2_rn(w[u].z, w[u].w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      reinterpret_cast<uint2*>(w16)[i] = o;
    }
  }
}

// ---- fused helpers for the layer executor ------------------------------------------------
__global__ void bn_fold_kernel(const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                               float target_rms, int D, float* scale, float* shift) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float inv = 1.0f / sqrtf(var[d] + eps);
  const float g = gamma ? gamma[d] : target_rms;
  const float b = beta ? beta[d] : 0.0f;
  const float s = g * inv;
  scale[d] = s;
  shift[d] = b - mean[d] * s;
}

// dZ = mask ? h(dY*scale[n]) : 0      (ops_batchnorm_backward then ops_relu_backward in one pass)
__global__ void bn_relu_bwd_kernel(const __half* __restrict__ dY, int ldy, const float* __restrict__ scale,
                                   const uint32_t* __restrict__ mask, int mask_ld, __half* __restrict__ dZ, int ldz,
                                   size_t rows, int cols) {
  const int c8 = cols >> 3;
  const size_t total = rows * (size_t)c8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / c8;
    const int c = (int)(i % c8) * 8;
    Half8 a = ld8(dY + r * ldy + c);
    __half* h = reinterpret_cast<__half*>(&a);
    uint32_t bits = 0xFFu;
    if (mask) bits = (mask[r * mask_ld + (c >> 5)] >> (c & 31)) & 0xFFu;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = __half2float(h[j]);
      if (scale) v *= scale[c + j];
      h[j] = ((bits >> j) & 1u) ? __float2half_rn(v) : __float2half(0.f);
    }
    st8(dZ + r * ldz + c, a);
  }
}

// x[t,:] = h(x[t,:] + bias)   (gpu.AddBias, internal/gpu/ops.go:335-351)
__global__ void add_bias_kernel(__half* __restrict__ x, int ld, const __half* __restrict__ bias, size_t rows, int cols) {
  const size_t total = rows * (size_t)cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / cols;
    const int c = (int)(i % cols);
    __half* p = x + r * ld + c;
    *p = __float2half_rn(__half2float(*p) + __half2float(bias[c]));
  }
}

// column sums with fp32 accumulation: out[n] = sum_t X[t,n]   (AffineBackwardBias)
// grid (col_chunks of 64 columns, row_splits); each thread owns 2 columns for a strided row set,
// block-reduces in smem and adds its partial with one atomic per column.
__global__ void colsum_kernel(const __half* __restrict__ X, int ld, size_t rows, int cols, float* __restrict__ out) {
  __shared__ float red[8][64];
  const int tx = threadIdx.x & 31;   // column pair
  const int ty = threadIdx.x >> 5;   // 8 row lanes
  const int c = blockIdx.x * 64 + tx * 2;
  float s0 = 0.f, s1 = 0.f;
  if (c < cols) {
    for (size_t r = (size_t)blockIdx.y * 8 + ty; r < rows; r += (size_t)gridDim.y * 8) {
      const __half2 v = *reinterpret_cast<const __half2*>(X + r * ld + c);
      s0 += __low2float(v);
      s1 += __high2float(v);
    }
  }
  red[ty][tx * 2] = s0;
  red[ty][tx * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < cols) atomicAdd(out + cc, s);
  }
}

// any column count / leading dimension (odd channel counts of the kaldibridge conv surface): scalar loads
__global__ void colsum_scalar_kernel(const __half* __restrict__ X, int ld, size_t rows, int cols, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (size_t r = blockIdx.y; r < rows; r += gridDim.y) s += __half2float(X[r * ld + c]);
  atomicAdd(out + c, s);
}

// End of a column-sum block (32 column lanes x 8 row lanes, 8 columns per thread): add the block's 256 column sums into
// out[col % col_mod].  Columns that fold onto the same output (narrow matrices processed k rows side by side) are combined
// in shared memory first and every group of 4 outputs goes out as ONE red.global.add.v4.f32 -- measured on B200: with one
// scalar atomic per thread the 592 blocks of a [384000 x 64] pass queue 2368 atomics on each of 64 addresses and the
// kernel takes 104 us instead of the ~10 us its 49 MB need.
__device__ __forceinline__ void colsum_flush(float (&red)[8][32][8], const float (&acc)[8], int cg, int ry, int cols, int col_mod,
                                             float* __restrict__ out) {
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ry][cg][j] = acc[j];
  __syncthreads();
  const int t = ry * 32 + cg;                 // block-local column t = column lane t >> 3, element t & 7
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += red[k][t >> 3][t & 7];
  __syncthreads();
  float* flat = &red[0][0][0];
  flat[t] = s;
  __syncthreads();
  const int col0 = blockIdx.x * 256;
  const int width = col_mod < 256 ? col_mod : 256;     // distinct outputs this block touches
  if (t < width && (t & 3) == 0 && col0 + t < cols) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = t; q < 256 && col0 + q < cols; q += width) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] += flat[q + e];
    }
    float* dst = out + (col0 + t) % col_mod;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(dst + e, v[e]);
    }
  }
}

// same column sum with the access pattern of the fused backward pass: 8 columns (16 bytes) per thread, kBnBwdRows rows in
// flight, one wave of blocks; col_mod folds a narrow dense matrix into wide rows (see bn_relu_bwd_colsum_kernel)
constexpr int kColsumRows = 4;
__global__ void __launch_bounds__(256, 4)
colsum_v2_kernel(const __half* __restrict__ X, int ld, uint32_t rows, int cols, float* __restrict__ out, int col_mod) {
  griddep_launch();
  griddep_wait();
  __shared__ float red[8][32][8];
  const int cg = threadIdx.x, ry = threadIdx.y;
  const int c = (blockIdx.x * 32 + cg) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < cols) {
    const uint32_t rows_per = (rows + gridDim.y - 1) / gridDim.y;
    const uint32_t r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, rows);
    for (uint32_t rb = r0 + ry; rb < r1; rb += 8 * kColsumRows) {
      uint4 a[kColsumRows];
#pragma unroll
      for (int k = 0; k < kColsumRows; ++k) {
        const uint32_t r = rb + k * 8;
        a[k] = r < r1 ? *reinterpret_cast<const uint4*>(X + (size_t)r * ld + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < kColsumRows; ++k) {
        const uint32_t w[4] = {a[k].x, a[k].y, a[k].z, a[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
          acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
        }
      }
    }
  }
  colsum_flush(red, acc, cg, ry, cols, col_mod, out);
}

// ---- train-mode batch-norm (batch statistics; cpp/cuda/cnn_kernels.cu:236-320 batchnorm1d_forward_fp16 training branch,
// go/gotorch/layers.go:257-300): three small passes around the producing GEMM instead of one thread per channel looping over
// the whole batch three times.
//  (1) per-column sum and sum of squares over the rows that belong to the minibatch: rows r with (r % period - lo) < len
//      (the real frames of the padded layout; period == 0: every row), fp32, accumulated into stats[0..cols) / stats[cols..2cols)
__global__ void __launch_bounds__(256, 4)
bn_stats_kernel(const __half* __restrict__ X, int ld, uint32_t rows, int cols, float* __restrict__ stats, uint32_t period, uint32_t lo,
                uint32_t len) {
  __shared__ float red[8][32][8];
  const int cg = threadIdx.x, ry = threadIdx.y;
  const int c = (blockIdx.x * 32 + cg) * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (c < cols) {
    const uint32_t rows_per = (rows + gridDim.y - 1) / gridDim.y;
    const uint32_t r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, rows);
    for (uint32_t r = r0 + ry; r < r1; r += 8) {
      if (period != 0 && (r % period - lo) >= len) continue;
      const uint4 a = *reinterpret_cast<const uint4*>(X + (size_t)r * ld + c);
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_h2(w[j]);
        s1[2 * j] += f.x; s1[2 * j + 1] += f.y;
        s2[2 * j] = fmaf(f.x, f.x, s2[2 * j]); s2[2 * j + 1] = fmaf(f.y, f.y, s2[2 * j + 1]);
      }
    }
  }
  colsum_flush(red, s1, cg, ry, cols, cols, stats);
  __syncthreads();
  colsum_flush(red, s2, cg, ry, cols, cols, stats + cols);
}
//  (2) statistics -> folded scale / shift (what the epilogues and the backward pass read) + running-statistics update.
//      mean = S1/n, var = S2/n - mean^2 (biased, as the reference), running = (1 - momentum)*running + momentum*batch
__global__ void bn_finalize_kernel(const float* __restrict__ stats, float n, int D, float* __restrict__ run_mean, float* __restrict__ run_var,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float target_rms,
                                   float momentum, float bwd_mul, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ scale_bwd) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float mean = stats[d] / n;
  const float var = fmaxf(stats[D + d] / n - mean * mean, 0.f);
  run_mean[d] = run_mean[d] * (1.f - momentum) + mean * momentum;
  run_var[d] = run_var[d] * (1.f - momentum) + var * momentum;
  const float inv = 1.0f / sqrtf(var + eps);
  const float g = gamma ? gamma[d] : target_rms;
  const float b = beta ? beta[d] : 0.0f;
  const float sc = g * inv;
  scale[d] = sc;
  shift[d] = b - mean * sc;
  if (scale_bwd) scale_bwd[d] = sc * bwd_mul;
}
//  (3) y = h(z*scale[c % col_mod] + shift[c % col_mod] (+ res_scale*R)) in place on the minibatch's rows (same row filter)
__global__ void bn_apply_kernel(__half* __restrict__ Z, int ld, const float* __restrict__ scale, const float* __restrict__ shift,
                                const __half* __restrict__ R, int ldr, float res_scale, uint32_t rows, int cols, int col_mod, uint32_t period,
                                uint32_t lo, uint32_t len) {
  const int cv = cols >> 3;
  const size_t total = (size_t)rows * cv, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint32_t r = (uint32_t)(i / cv);
    if (period != 0 && (r % period - lo) >= len) continue;
    const int c = (int)(i % cv) << 3;
    const int cm = c % col_mod;
    uint4 v = *reinterpret_cast<const uint4*>(Z + (size_t)r * ld + c);
    uint4 rv = make_uint4(0, 0, 0, 0);
    if (R) rv = *reinterpret_cast<const uint4*>(R + (size_t)r * ldr + c);
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_h2(w[j]);
      f.x = fmaf(f.x, scale[cm + 2 * j], shift[cm + 2 * j]);
      f.y = fmaf(f.y, scale[cm + 2 * j + 1], shift[cm + 2 * j + 1]);
      if (R) { const float2 q = unpack_h2(rw[j]); f.x = fmaf(res_scale, q.x, f.x); f.y = fmaf(res_scale, q.y, f.y); }
      w[j] = pack_h2(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(Z + (size_t)r * ld + c) = v;
  }
}

// replicate row 0 / row rows-1 of every sequence block into its halo rows
//   buffer rows: n_seq blocks of (seq_len + 2*halo) rows; X points at the first block's row -halo
__global__ void pad_edges_kernel(__half* __restrict__ X, int ld, int n_seq, int seq_len, int cols, int halo) {
  griddep_launch();
  griddep_wait();
  const int c8 = cols >> 3;
  const size_t per_seq = (size_t)2 * halo * c8;
  const size_t total = per_seq * n_seq;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int blk = seq_len + 2 * halo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int s = (int)(i / per_seq);
    const size_t rem = i % per_seq;
    const int hr = (int)(rem / c8);     // 0..2*halo-1
    const int c = (int)(rem % c8) * 8;
    const size_t base = (size_t)s * blk;
    const size_t dst_row = hr < halo ? base + hr : base + halo + seq_len + (hr - halo);
    const size_t src_row = hr < halo ? base + halo : base + halo + seq_len - 1;
    st8(X + dst_row * ld + c, ld8(X + src_row * ld + c));
  }
}
// adjoint of pad_edges: edge row += sum of its halo rows (fp32), halo rows = 0
__global__ void fold_edges_kernel(__half* __restrict__ G, int ld, int n_seq, int seq_len, int cols, int halo) {
  griddep_launch();
  griddep_wait();
  const int c8 = cols >> 3;
  const size_t total = (size_t)n_seq * 2 * c8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int blk = seq_len + 2 * halo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int s = (int)(i / (2 * c8));
    const int side = (int)((i / c8) & 1);
    const int c = (int)(i % c8) * 8;
    const size_t base = (size_t)s * blk;
    if (seq_len == 1 && side == 1) continue;   // a one-frame sequence has both edges on one row: side 0 folds both halos
    const size_t edge = side == 0 ? base + halo : base + halo + seq_len - 1;
    Half8 e = ld8(G + edge * ld + c);
    __half* eh = reinterpret_cast<__half*>(&e);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __half2float(eh[j]);
    Half8 z;
#pragma unroll
    for (int j = 0; j < 4; ++j) z.v[j] = __float2half2_rn(0.f);
    for (int pass = 0; pass < (seq_len == 1 ? 2 : 1); ++pass) {
      const size_t h0 = (side == 0 && pass == 0) ? base : base + halo + seq_len;
      for (int k0 = 0; k0 < halo; k0 += 4) {     // 4 independent row loads in flight per pass
        Half8 h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k0 + k < halo) h[k] = ld8(G + (h0 + k0 + k) * ld + c);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k0 + k < halo) {
            const __half* hh = reinterpret_cast<const __half*>(&h[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += __half2float(hh[j]);
            st8(G + (h0 + k0 + k) * ld + c, z);
          }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) eh[j] = __float2half_rn(acc[j]);
    st8(G + edge * ld + c, e);
  }
}


// ---- padded minibatch layout helpers ----------------------------------------------------
// dense [n_seq*L x cols] -> padded [n_seq*(L+2h) x ld]; halo rows: mode 0 = zero, 1 = replicate edge
template <int VEC>
__global__ void pack_rows_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int ld, int n_seq, int L,
                                 int halo, int cols, int mode) {
  griddep_launch();
  griddep_wait();
  const int blk = L + 2 * halo;
  const int cv = cols / VEC;
  const size_t total = (size_t)n_seq * blk * cv;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / cv;
    const int c = (int)(i % cv) * VEC;
    const int s = (int)(r / blk);
    const int t = (int)(r % blk) - halo;
    const bool real = t >= 0 && t < L;
    const size_t srow = (size_t)s * L + (real ? t : (t < 0 ? 0 : L - 1));
    if (VEC == 8) {
      Half8 v;
      if (real || mode == 1) v = ld8(src + srow * cols + c);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v.v[j] = __float2half2_rn(0.f);
      }
      st8(dst + r * ld + c, v);
    } else {
      dst[r * ld + c] = (real || mode == 1) ? src[srow * cols + c] : __float2half(0.f);
    }
  }
}
// same from FP32 rows: the RNE FP32 -> FP16 conversion of the features (internal/gpu/bridge.go:141, internal/fp16/fp16.go:13-70:
// round-to-nearest-even, overflow to Inf) done on the device while scattering into the padded layout
__global__ void pack_rows_f32_kernel(const float* __restrict__ src, __half* __restrict__ dst, int ld, int n_seq, int L,
                                     int halo, int cols, int mode) {
  griddep_launch();
  griddep_wait();
  const int blk = L + 2 * halo;
  const size_t total = (size_t)n_seq * blk * cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / cols;
    const int c = (int)(i % cols);
    const int s = (int)(r / blk);
    const int t = (int)(r % blk) - halo;
    const bool real = t >= 0 && t < L;
    const size_t srow = (size_t)s * L + (real ? t : (t < 0 ? 0 : L - 1));
    dst[r * ld + c] = (real || mode == 1) ? __float2half_rn(src[srow * cols + c]) : __float2half(0.f);
  }
}
// padded [.. x ld] (cols from col0) -> dense [n_seq*L x cols]
__global__ void unpack_rows_kernel(const __half* __restrict__ src, int ld, int col0, __half* __restrict__ dst, int n_seq,
                                   int L, int halo, int cols) {
  const int blk = L + 2 * halo;
  const size_t total = (size_t)n_seq * L * cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / cols;
    const int c = (int)(i % cols);
    const int s = (int)(r / L), t = (int)(r % L);
    dst[i] = src[((size_t)s * blk + halo + t) * ld + col0 + c];
  }
}
// dst[r, col0 + c] = src[r / blk, c]   (per-sequence vector broadcast to every frame of its block)
__global__ void bcast_rows_kernel(const __half* __restrict__ src, int cols, __half* __restrict__ dst, int ld, int col0,
                                  size_t rows, int blk) {
  // one block row per destination row (no division per element: 9984 x 200 halves took 7.8 us)
  for (size_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const __half* s = src + (r / blk) * cols;
    __half* d = dst + r * ld + col0;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) d[c] = s[c];
  }
}
// out[s, c] = h( sum over the real rows of block s of G[r, col0 + c] )   (adjoint of the broadcast)
// block = (32 columns) x (8 row lanes), grid = (sequences, column groups): row lane y sums frames y, y+8, ... (four loads in
// flight), the 8 partial sums are added in lane order -- a fixed summation order, so the result is reproducible
__global__ void __launch_bounds__(256)
seq_sum_kernel(const __half* __restrict__ G, int ld, int col0, __half* __restrict__ out, int cols, int L, int halo) {
  __shared__ float part[8][33];
  const int s = blockIdx.x;
  const int blk = L + 2 * halo;
  const int c = blockIdx.y * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols) {
    const __half* g = G + ((size_t)s * blk + halo) * ld + col0 + c;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int t = threadIdx.y;
    for (; t + 24 < L; t += 32) {
      a0 += __half2float(g[(size_t)t * ld]); a1 += __half2float(g[(size_t)(t + 8) * ld]);
      a2 += __half2float(g[(size_t)(t + 16) * ld]); a3 += __half2float(g[(size_t)(t + 24) * ld]);
    }
    for (; t < L; t += 8) a0 += __half2float(g[(size_t)t * ld]);
    acc = (a0 + a1) + (a2 + a3);
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float tot = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) tot += part[y][threadIdx.x];
    out[(size_t)s * cols + c] = __float2half_rn(tot);
  }
}
__global__ void zero_halo_kernel(__half* __restrict__ X, int ld, int n_seq, int L, int cols, int halo) {
  const size_t per_seq = (size_t)2 * halo * cols;
  const size_t total = per_seq * n_seq;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int blk = L + 2 * halo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int s = (int)(i / per_seq);
    const size_t rem = i % per_seq;
    const int hr = (int)(rem / cols), c = (int)(rem % cols);
    const size_t row = (size_t)s * blk + (hr < halo ? hr : L + hr);
    X[row * ld + c] = __float2half(0.f);
  }
}
// SpecAugment on the padded minibatch layout (go/gotorch/cnn_tdnn.go:612-668 SpecAugment.Apply: per sequence, frequency
// masks [f0, f0+f) over every frame and time masks [t0, t0+t) over every bin, masked value 0), with a counter-based
// generator instead of math/rand: draw k of sequence s is u(seed, s, k) (the hash the dropout epilogue uses), so the masks
// are a pure function of (seed, sequence) -- the backward pass and the CPU oracle rebuild them.
//   frequency mask m:  f = floor(u(4m) * (fmax + 1)),   f0 = floor(u(4m+1) * (dim - f + 1))
//   time mask m:       t = floor(u(4m+2) * (tmax + 1)), t0 = floor(u(4m+3) * (L - t + 1))
// y = x outside the masks (x == y allowed: in place); halo rows are copied through.
__global__ void spec_augment_kernel(const __half* __restrict__ x, __half* __restrict__ y, int ld, int L, int halo, int dim,
                                    int fmax, int nfreq, int tmax, int ntime, uint32_t seed, const uint32_t* __restrict__ seed_dev) {
  __shared__ int f_lo[8], f_hi[8], t_lo[8], t_hi[8];
  const int s = blockIdx.x;
  if (threadIdx.x < 8) {
    const uint32_t sd = seed ^ (seed_dev ? *seed_dev : 0u);
    const int m = threadIdx.x;
    const int f = m < nfreq ? (int)(dropout_uniform(sd, (uint32_t)s, 4u * m) * (float)(fmax + 1)) : 0;
    const int f0 = (int)(dropout_uniform(sd, (uint32_t)s, 4u * m + 1u) * (float)(dim - f + 1));
    const int t = m < ntime ? (int)(dropout_uniform(sd, (uint32_t)s, 4u * m + 2u) * (float)(tmax + 1)) : 0;
    const int t0 = (int)(dropout_uniform(sd, (uint32_t)s, 4u * m + 3u) * (float)(L - t + 1));
    f_lo[m] = f0; f_hi[m] = f0 + f; t_lo[m] = t0; t_hi[m] = t0 + t;
  }
  __syncthreads();
  const int blk = L + 2 * halo;
  const size_t base = (size_t)s * blk * ld;
  for (int i = threadIdx.x; i < blk * dim; i += blockDim.x) {
    const int r = i / dim, c = i - r * dim;
    const int tt = r - halo;
    bool masked = false;
    if (tt >= 0 && tt < L) {
#pragma unroll
      for (int m = 0; m < 8; ++m) masked = masked || (c >= f_lo[m] && c < f_hi[m]) || (tt >= t_lo[m] && tt < t_hi[m]);
    }
    y[base + (size_t)r * ld + c] = masked ? __float2half(0.f) : x[base + (size_t)r * ld + c];
  }
}
// strided-row form of scale_shift_kernel, 8 columns per thread (cols % 8 == 0, 16-byte aligned rows)
__global__ void scale_shift_ld_kernel(const __half* __restrict__ x, long long ldx, __half* __restrict__ y, long long ldy, int rows,
                                      int cols, const float* __restrict__ scale, const float* __restrict__ shift) {
  const int cv = cols >> 3;
  const size_t total = (size_t)rows * cv, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / cv;
    const int c = (int)(i % cv) << 3;
    uint4 v = *reinterpret_cast<const uint4*>(x + r * ldx + c);
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __half22float2(h[e]);
      f.x = f.x * scale[c + 2 * e] + (shift ? shift[c + 2 * e] : 0.f);
      f.y = f.y * scale[c + 2 * e + 1] + (shift ? shift[c + 2 * e + 1] : 0.f);
      h[e] = __floats2half2_rn(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(y + r * ldy + c) = v;
  }
}
// zero every row r of X[rows x cols] that is NOT of the form row0 + k*step (k >= 0): makes a row-subsampled gradient dense
__global__ void zero_rows_except_kernel(__half* __restrict__ X, int ld, int rows, int cols, int row0, int step) {
  // one block row per matrix row: the keep test is per row, the columns go out as 16-byte stores
  const int cv = cols >> 3;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    if (r >= row0 && (r - row0) % step == 0) continue;
    uint4* d = reinterpret_cast<uint4*>(X + (size_t)r * ld);
    for (int c = threadIdx.x; c < cv; c += blockDim.x) d[c] = make_uint4(0u, 0u, 0u, 0u);
  }
}
// y = h(x*scale[c] + shift[c])  (shift may be null)
__global__ void scale_shift_kernel(const __half* __restrict__ x, __half* __restrict__ y, size_t total, int cols,
                                   const float* __restrict__ scale, const float* __restrict__ shift) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % cols);
    const float v = __half2float(x[i]) * scale[c] + (shift ? shift[c] : 0.f);
    y[i] = __float2half_rn(v);
  }
}
// dY = Y on real rows, 0 on halo rows; *loss += 0.5*sum(Y^2)   (cmd/sgdtest/main.go:258-267)
template <int VEC>
__global__ void half_sq_loss_kernel(const __half* __restrict__ Y, __half* __restrict__ dY, int n_seq, int L, int halo,
                                    int cols, float* __restrict__ loss) {
  griddep_launch();
  griddep_wait();
  __shared__ float red[32];
  const int blk = L + 2 * halo;
  const int cv = cols / VEC;
  const size_t total = (size_t)n_seq * blk * cv;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / cv;
    const int t = (int)(r % blk) - halo;
    const size_t off = r * cols + (i % cv) * VEC;
    const bool real = t >= 0 && t < L;
    if (VEC == 8) {
      Half8 v;
      if (real) {
        v = ld8(Y + off);
        const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float f = __half2float(h[j]); acc += 0.5f * f * f; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v.v[j] = __float2half2_rn(0.f);
      }
      st8(dY + off, v);
    } else {
      __half v = __float2half(0.f);
      if (real) { v = Y[off]; const float f = __half2float(v); acc += 0.5f * f * f; }
      dY[off] = v;
    }
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffff, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
    if (threadIdx.x == 0) atomicAdd(loss, v);
  }
}
// dZ = mask ? h(dY * scale) : 0 and db += colsum(dZ) in one pass (the reference: ops_batchnorm_backward +
// ops_relu_backward + an M=1 GEMM, backward_wrappers.cu:41-115, backward_ops.go:228-253).
// Optionally also applies the adjoint of pad_edges to dY first (blk > 0): halo rows of every sequence block are
// added into their edge row and zeroed (dY is updated in place on those rows), so no separate fold launch is needed.
// Each thread keeps kRowsInFlight independent 16-byte row loads in flight (one per 8-row step).
constexpr int kBnBwdRows = 4;
template <bool FOLD>
__global__ void __launch_bounds__(256, 4)
bn_relu_bwd_colsum_kernel(__half* __restrict__ dY, int ldy, const float* __restrict__ scale,
                          const uint32_t* __restrict__ mask, int mask_ld, __half* __restrict__ dZ,
                          int ldz, uint32_t rows, int cols, float* __restrict__ db, uint32_t blk, uint32_t seq_len, uint32_t halo,
                          int col_mod) {
  // col_mod: a narrow dense matrix [R x C] (the conv layers: C = 64..256 filters) is processed as [R/k x k*C] so that all 32
  // column lanes of a warp are busy; the per-column vectors (scale, db) are then indexed modulo C = col_mod
  griddep_launch();
  griddep_wait();
  __shared__ float red[8][32][8];
  const int cg = threadIdx.x, ry = threadIdx.y;
  const int c = (blockIdx.x * 32 + cg) * 8;
  const int cv = c % col_mod;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < cols) {
    float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0;
    if (scale) { sc0 = *reinterpret_cast<const float4*>(scale + cv); sc1 = *reinterpret_cast<const float4*>(scale + cv + 4); }
    const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
    const uint32_t rows_per = (rows + gridDim.y - 1) / gridDim.y;
    const uint32_t r0 = blockIdx.y * rows_per;
    const uint32_t r1 = min(r0 + rows_per, rows);
    const uint32_t mshift = c & 31;
    const __half* ycol = dY + c;
    const uint32_t* mcol = mask ? mask + (c >> 5) : nullptr;
    for (uint32_t rb = r0 + ry; rb < r1; rb += 8 * kBnBwdRows) {
      uint4 a[kBnBwdRows];
      uint32_t bits[kBnBwdRows];
#pragma unroll
      for (int k = 0; k < kBnBwdRows; ++k) {     // all loads of the group first: 4 x 16 bytes in flight per thread
        const uint32_t r = rb + k * 8;
        a[k] = make_uint4(0, 0, 0, 0);
        bits[k] = 0xFFFFFFFFu;
        if (r < r1) {
          a[k] = *reinterpret_cast<const uint4*>(ycol + (size_t)r * ldy);
          if (mcol) bits[k] = __ldg(mcol + (size_t)r * mask_ld);
        }
      }
#pragma unroll
      for (int k = 0; k < kBnBwdRows; ++k) {
        const uint32_t r = rb + k * 8;
        if (r >= r1) continue;
        uint32_t w[4] = {a[k].x, a[k].y, a[k].z, a[k].w};
        const uint32_t mb = bits[k] >> mshift;
        if (FOLD) {
          const uint32_t pos = r % blk;
          if (pos < halo || pos >= halo + seq_len) {
            // halo row: contributes nothing; its dY values are consumed (and zeroed) by the edge row's thread
            *reinterpret_cast<uint4*>(dZ + (size_t)r * ldz + c) = make_uint4(0, 0, 0, 0);
            continue;
          }
          const bool left = pos == halo, right = pos == halo + seq_len - 1;
          if (left || right) {
            // edge row += its halo rows (fp32, row order), halo rows = 0; a one-frame sequence has both edges on one row
            float e[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = unpack_h2(w[j]); e[2 * j] = f.x; e[2 * j + 1] = f.y; }
            for (int side = 0; side < 2; ++side) {
              if (side == 0 ? !left : !right) continue;
              __half* hp = dY + (size_t)(side == 0 ? r - halo : r + 1) * ldy + c;
              for (uint32_t q0 = 0; q0 < halo; q0 += 4) {
                uint4 hv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (q0 + q < halo) hv[q] = *reinterpret_cast<const uint4*>(hp + (size_t)(q0 + q) * ldy);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (q0 + q < halo) {
                    const uint32_t hw[4] = {hv[q].x, hv[q].y, hv[q].z, hv[q].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 f = unpack_h2(hw[j]); e[2 * j] += f.x; e[2 * j + 1] += f.y; }
                    *reinterpret_cast<uint4*>(hp + (size_t)(q0 + q) * ldy) = make_uint4(0, 0, 0, 0);
                  }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = pack_h2(e[2 * j], e[2 * j + 1]);
            *reinterpret_cast<uint4*>(dY + (size_t)r * ldy + c) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_h2(w[j]);
          // h(dY * scale) per element, then the relu mask as a bit-select on the packed pair
          uint32_t o = pack_h2(f.x * sc[2 * j], f.y * sc[2 * j + 1]);
          const uint32_t two = (mb >> (2 * j)) & 3u;
          o &= (two & 1u) * 0xFFFFu + (two >> 1) * 0xFFFF0000u;
          w[j] = o;
          const float2 g = unpack_h2(o);
          acc[2 * j] += g.x;
          acc[2 * j + 1] += g.y;
        }
        *reinterpret_cast<uint4*>(dZ + (size_t)r * ldz + c) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  if (db) colsum_flush(red, acc, cg, ry, cols, col_mod, db);
}
// out[n] += sum_t X[t,n] without the leading memset (accumulates into the flat gradient bucket)
}  // namespace kfp16

using namespace kfp16;

extern "C" {

// ------------------------------------------------------------------ ops.h forward elementwise
int ops_relu(void* data, int count) { return run_map1(data, count, ReluOp{}, "relu kernel"); }
int ops_sigmoid(void* data, int count) { return run_map1(data, count, SigmoidOp{}, "sigmoid kernel"); }
int ops_tanh_act(void* data, int count) { return run_map1(data, count, TanhOp{}, "tanh kernel"); }
int ops_clipped_relu(void* data, int count, float ceiling) {
  return run_map1(data, count, ClippedReluOp{ceiling}, "clipped_relu kernel");
}
int ops_fill(void* dst, int count, float val) { return run_map1(dst, count, FillOp{__float2half(val)}, "fill kernel"); }
int ops_add_scaled(void* dst, const void* src, int count, float alpha, float beta) {
  return run_map2(dst, src, count, AddScaledOp{alpha, beta}, "add_scaled kernel");
}
int ops_add(void* dst, const void* src, int count) { return run_map2(dst, src, count, AddOp{}, "add kernel"); }
int ops_copy(void* dst, const void* src, int count) {
  if (count <= 0) return 0;
  return check_cuda(cudaMemcpyAsync(dst, src, (size_t)count * sizeof(__half), cudaMemcpyDeviceToDevice, default_stream()), "copy") ? 0 : -1;
}

}  // extern "C"
int kfp16::softmax_on_stream(cudaStream_t stream, void* data, int rows, int cols, bool log) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!data) { set_error("softmax: null pointer"); return -1; }
  const int grid = rows < num_sms_cached() * 16 ? rows : num_sms_cached() * 16;
  if ((cols % 8) == 0 && cols <= 8192 && al16(data)) {
    const int g2 = rows < num_sms_cached() * 8 ? rows : num_sms_cached() * 8;
    if (log) softmax_row_regs_kernel<true><<<g2, 256, 0, stream>>>((__half*)data, rows, cols);
    else softmax_row_regs_kernel<false><<<g2, 256, 0, stream>>>((__half*)data, rows, cols);
  } else if (log) softmax_kernel<true><<<grid, 128, 0, stream>>>((__half*)data, rows, cols);
  else softmax_kernel<false><<<grid, 128, 0, stream>>>((__half*)data, rows, cols);
  count_launch();
  return check_launch(log ? "log_softmax kernel" : "softmax kernel") ? 0 : -1;
}
extern "C" {
int ops_softmax(void* data, int rows, int cols) { return kfp16::softmax_on_stream(default_stream(), data, rows, cols, false); }
int ops_log_softmax(void* data, int rows, int cols) { return kfp16::softmax_on_stream(default_stream(), data, rows, cols, true); }

int ops_batchnorm_forward(void* x, int T, int D, const float* mean, const float* var, const float* gamma,
                          const float* beta, float epsilon) {
  if (T > 0 && D > 0 && (!mean || !var || !gamma || !beta)) { set_error("batchnorm kernel: null statistics"); return -1; }
  return run_colscale(x, x, T, D, mean, var, gamma, beta, epsilon, 1.0f, 0, "batchnorm kernel");
}
int ops_batchnorm_forward_rms(void* x, int T, int D, const float* mean, const float* var, float target_rms,
                              float epsilon) {
  if (T > 0 && D > 0 && (!mean || !var)) { set_error("batchnorm_rms kernel: null statistics"); return -1; }
  return run_colscale(x, x, T, D, mean, var, nullptr, nullptr, epsilon, target_rms, 1, "batchnorm_rms kernel");
}

int ops_concat_cols(void* dst, int T, int dst_cols, const void* src, int src_cols, int dst_col_offset) {
  return run_copy2d(dst, dst_cols, dst_col_offset, src, src_cols, 0, 0, 1, T, src_cols, "concat_cols kernel", default_stream());
}
int ops_slice_cols(const void* src, int T, int src_cols, void* dst, int dst_cols, int src_col_offset) {
  return run_copy2d(dst, dst_cols, 0, src, src_cols, src_col_offset, 0, 1, T, dst_cols, "slice_cols kernel", default_stream());
}
void ops_subsample_rows(void* dst, const void* src, int in_rows, int cols, int stride, int row_offset) {
  if (stride <= 0 || in_rows <= row_offset) return;
  const int out_rows = (in_rows - row_offset + stride - 1) / stride;   // ops.cu:633
  run_copy2d(dst, cols, 0, src, cols, 0, row_offset, stride, out_rows, cols, "subsample_rows kernel", default_stream());
}
int ops_combine_feature_maps(void* data, int T, int total_dim, int height, int nf1, int nf2) {
  return kfp16::ops_combine_feature_maps_on(default_stream(), data, T, total_dim, height, nf1, nf2, 0);
}
}  // extern "C"
int kfp16::ops_combine_feature_maps_on(cudaStream_t stream, void* data, int T, int total_dim, int height, int nf1, int nf2, int inverse) {
  if (T <= 0 || total_dim <= 0) return 0;
  if (!data) { set_error("combine_feature_maps: null pointer"); return -1; }
  if (height * (nf1 + nf2) != total_dim) { set_error("combine_feature_maps: height*(nf1+nf2) != total_dim (%d*(%d+%d) != %d)", height, nf1, nf2, total_dim); return -1; }
  const size_t row_bytes = (size_t)total_dim * sizeof(__half);
  if (row_bytes > 200 * 1024) { set_error("combine_feature_maps: row of %d halves exceeds shared memory", total_dim); return -1; }
  const int rows_per_pass = (int)std::max<size_t>(1, std::min<size_t>(8, (32 * 1024) / row_bytes));     // <= 32 KB per block: several blocks per SM
  const size_t smem = row_bytes * rows_per_pass;
  if (smem > 48 * 1024 &&
      !check_cuda(cudaFuncSetAttribute(combine_fm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "combine smem"))
    return -1;
  const int passes = (T + rows_per_pass - 1) / rows_per_pass;
  const int grid = passes < num_sms_cached() * 8 ? passes : num_sms_cached() * 8;
  combine_fm_kernel<<<grid, kThreads, smem, stream>>>((__half*)data, T, total_dim, height, nf1, nf2, inverse, rows_per_pass);
  count_launch();
  return check_launch("combine_feature_maps kernel") ? 0 : -1;
}
int kfp16::ops_concat_cols_on(cudaStream_t stream, void* dst, int T, int dst_cols, const void* src, int src_cols, int dst_col_offset) {
  return run_copy2d(dst, dst_cols, dst_col_offset, src, src_cols, 0, 0, 1, T, src_cols, "concat_cols kernel", stream);
}
int kfp16::ops_slice_cols_on(cudaStream_t stream, const void* src, int T, int src_cols, void* dst, int dst_cols, int src_col_offset) {
  return run_copy2d(dst, dst_cols, 0, src, src_cols, src_col_offset, 0, 1, T, dst_cols, "slice_cols kernel", stream);
}
int kfp16::ops_slice_add_on(cudaStream_t stream, const void* src, int T, int src_cols, void* dst, int dst_cols, int src_col_offset) {
  return run_copy2d(dst, dst_cols, 0, src, src_cols, src_col_offset, 0, 1, T, dst_cols, "slice_add kernel", stream, 1);
}
extern "C" {

// ------------------------------------------------------------------ ops.h backward + optimiser
int ops_relu_backward(const void* x, void* grad, int count) { return run_map2(grad, x, count, ReluBwdOp{}, "relu_backward"); }
int ops_sigmoid_backward(const void* output, void* grad, int count) { return run_map2(grad, output, count, SigmoidBwdOp{}, "sigmoid_backward"); }
int ops_tanh_backward(const void* output, void* grad, int count) { return run_map2(grad, output, count, TanhBwdOp{}, "tanh_backward"); }

int ops_transpose(const void* src, void* dst, int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  if (!src || !dst) { set_error("transpose: null pointer"); return -1; }
  const long long tiles = (long long)((M + 31) / 32) * ((N + 31) / 32);
  const long long cap = (long long)num_sms_cached() * 8;
  transpose_kernel<<<(int)(tiles < cap ? tiles : cap), dim3(32, 8), 0, default_stream()>>>((const __half*)src, (__half*)dst, M, N);
  count_launch();
  return check_launch("transpose") ? 0 : -1;
}
int ops_batchnorm_backward(const void* grad_out, void* grad_in, const float* gamma, const float* variance, float eps,
                           int rows, int cols) {
  if (rows > 0 && cols > 0 && (!gamma || !variance)) { set_error("batchnorm_backward: null statistics"); return -1; }
  return run_colscale(grad_out, grad_in, rows, cols, nullptr, variance, gamma, nullptr, eps, 1.0f, 2, "batchnorm_backward");
}
int ops_fp16_to_fp32(const void* src, float* dst, int count) {
  if (count <= 0) return 0;
  if (!src || !dst) { set_error("fp16_to_fp32: null pointer"); return -1; }
  f16_to_f32_kernel<<<grid_for((size_t)count), kThreads, 0, default_stream()>>>((const __half*)src, dst, (size_t)count);
  count_launch();
  return check_launch("fp16_to_fp32") ? 0 : -1;
}
int ops_sgd_update(float* w_fp32, void* w_fp16, const void* grad_fp16, float* velocity, float lr, float momentum,
                   int count) {
  if (count <= 0) return 0;
  if (!w_fp32 || !w_fp16 || !grad_fp16 || !velocity) { set_error("sgd_update: null pointer"); return -1; }
  sgd_kernel<false><<<grid_for((size_t)count), kThreads, 0, default_stream()>>>(w_fp32, (__half*)w_fp16, grad_fp16, 0, 1.0f, velocity, lr, momentum, (size_t)count, nullptr);
  count_launch();
  return check_launch("sgd_update") ? 0 : -1;
}

// ------------------------------------------------------------------ fused helpers (kaldi_fp16_fused.h)
static cudaStream_t ctx_stream(kfp16_ctx* ctx) { return ctx ? ctx->stream : default_stream(); }

int kfp16_bn_fold(kfp16_ctx* ctx, const float* mean, const float* var, const float* gamma, const float* beta,
                  float eps, float target_rms, int D, float* scale, float* shift) {
  if (D <= 0) return 0;
  if (!mean || !var || !scale || !shift) { set_error("kfp16_bn_fold: null pointer"); return -1; }
  bn_fold_kernel<<<(D + 255) / 256, 256, 0, ctx_stream(ctx)>>>(mean, var, gamma, beta, eps, target_rms, D, scale, shift);
  count_launch();
  return check_launch("kfp16_bn_fold") ? 0 : -1;
}
int kfp16_bn_batch_stats(kfp16_ctx* ctx, const void* X, int ld, int rows, int cols, float* stats, int period, int lo, int len) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!X || !stats || (cols % 8) || (ld % 8) || !al16(X) || !al16(stats)) { set_error("kfp16_bn_batch_stats: needs 16-byte aligned buffers and cols / ld %% 8 == 0"); return -1; }
  if (period < 0 || lo < 0 || len < 0 || (period > 0 && lo + len > period)) { set_error("kfp16_bn_batch_stats: bad row filter"); return -1; }
  cudaStream_t s = ctx_stream(ctx);
  if (!check_cuda(cudaMemsetAsync(stats, 0, (size_t)2 * cols * sizeof(float), s), "kfp16_bn_batch_stats memset")) return -1;
  const int gx = (cols + 255) / 256;
  int gy = (num_sms_cached() * 4) / gx;
  const int max_gy = (rows + 63) / 64;
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  bn_stats_kernel<<<dim3(gx, gy), dim3(32, 8), 0, s>>>((const __half*)X, ld, (uint32_t)rows, cols, stats, (uint32_t)period, (uint32_t)lo, (uint32_t)len);
  count_launch();
  return check_launch("kfp16_bn_batch_stats") ? 0 : -1;
}
int kfp16_bn_finalize(kfp16_ctx* ctx, const float* stats, double n_rows, int D, float* run_mean, float* run_var, const float* gamma,
                      const float* beta, float eps, float target_rms, float momentum, float bwd_mul, float* scale, float* shift,
                      float* scale_bwd) {
  if (D <= 0) return 0;
  if (!stats || !run_mean || !run_var || !scale || !shift || n_rows < 1) { set_error("kfp16_bn_finalize: null pointer / empty batch"); return -1; }
  bn_finalize_kernel<<<(D + 255) / 256, 256, 0, ctx_stream(ctx)>>>(stats, (float)n_rows, D, run_mean, run_var, gamma, beta, eps, target_rms, momentum,
                                                                  bwd_mul, scale, shift, scale_bwd);
  count_launch();
  return check_launch("kfp16_bn_finalize") ? 0 : -1;
}
int kfp16_bn_apply(kfp16_ctx* ctx, void* Z, int ld, const float* scale, const float* shift, const void* R, int ldr, float res_scale, int rows,
                   int cols, int col_mod, int period, int lo, int len) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!Z || !scale || !shift || (cols % 8) || (ld % 8) || !al16(Z) || (R && ((ldr % 8) || !al16(R))) || col_mod < 8 || (col_mod % 8)) {
    set_error("kfp16_bn_apply: needs 16-byte aligned buffers and cols / ld / col_mod %% 8 == 0"); return -1;
  }
  bn_apply_kernel<<<grid_for((size_t)rows * (cols / 8)), kThreads, 0, ctx_stream(ctx)>>>((__half*)Z, ld, scale, shift, (const __half*)R, ldr, res_scale,
                                                                                        (uint32_t)rows, cols, col_mod, (uint32_t)period, (uint32_t)lo, (uint32_t)len);
  count_launch();
  return check_launch("kfp16_bn_apply") ? 0 : -1;
}
int kfp16_bn_relu_backward(kfp16_ctx* ctx, const void* dY, int ldy, const float* scale, const uint32_t* mask,
                           int mask_ld, void* dZ, int ldz, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!dY || !dZ || (cols % 8) || (ldy % 8) || (ldz % 8) || !al16(dY) || !al16(dZ)) {
    set_error("kfp16_bn_relu_backward: needs 16B-aligned buffers and cols/ld %% 8 == 0"); return -1;
  }
  bn_relu_bwd_kernel<<<grid_for((size_t)rows * (cols / 8)), kThreads, 0, ctx_stream(ctx)>>>(
      (const __half*)dY, ldy, scale, mask, mask_ld, (__half*)dZ, ldz, (size_t)rows, cols);
  count_launch();
  return check_launch("kfp16_bn_relu_backward") ? 0 : -1;
}
int kfp16_add_bias(kfp16_ctx* ctx, void* x, int ld, const void* bias, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!x || !bias) { set_error("kfp16_add_bias: null pointer"); return -1; }
  add_bias_kernel<<<grid_for((size_t)rows * cols), kThreads, 0, ctx_stream(ctx)>>>((__half*)x, ld, (const __half*)bias, (size_t)rows, cols);
  count_launch();
  return check_launch("kfp16_add_bias") ? 0 : -1;
}
int kfp16_colsum(kfp16_ctx* ctx, const void* X, int ld, int rows, int cols, float* out_f32, void* out_f16) {
  if (cols <= 0) return 0;
  if (!X || !out_f32) { set_error("kfp16_colsum: null pointer"); return -1; }
  cudaStream_t s = ctx_stream(ctx);
  if (!check_cuda(cudaMemsetAsync(out_f32, 0, (size_t)cols * sizeof(float), s), "kfp16_colsum memset")) return -1;
  if (rows > 0 && ((cols % 2) || (ld % 2) || ((uintptr_t)X & 3))) {
    int gy = rows < 256 ? (int)rows : 256;
    colsum_scalar_kernel<<<dim3((cols + 127) / 128, gy), 128, 0, s>>>((const __half*)X, ld, (size_t)rows, cols, out_f32);
    count_launch();
    if (!check_launch("kfp16_colsum")) return -1;
  } else if (rows > 0) {
    const int gx = (cols + 63) / 64;
    int gy = (num_sms_cached() * 4 + gx - 1) / gx;
    const int max_gy = (rows + 63) / 64;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    colsum_kernel<<<dim3(gx, gy), kThreads, 0, s>>>((const __half*)X, ld, (size_t)rows, cols, out_f32);
    count_launch();
    if (!check_launch("kfp16_colsum")) return -1;
  }
  if (out_f16) return kfp16_f32_to_f16(ctx, out_f32, out_f16, (size_t)cols);
  return 0;
}
int kfp16_f32_to_f16(kfp16_ctx* ctx, const float* src, void* dst, size_t n) {
  if (n == 0) return 0;
  if (!src || !dst) { set_error("kfp16_f32_to_f16: null pointer"); return -1; }
  f32_to_f16_kernel<<<grid_for(n), kThreads, 0, ctx_stream(ctx)>>>(src, (__half*)dst, n);
  count_launch();
  return check_launch("kfp16_f32_to_f16") ? 0 : -1;
}
int kfp16_scale_f32(kfp16_ctx* ctx, const float* src, float* dst, int n, float c) {
  if (n <= 0) return 0;
  if (!src || !dst) { set_error("kfp16_scale_f32: null pointer"); return -1; }
  scale_f32_kernel<<<(n + 255) / 256, 256, 0, ctx_stream(ctx)>>>(src, dst, n, c);
  count_launch();
  return check_launch("kfp16_scale_f32") ? 0 : -1;
}
int kfp16_bump_counter(kfp16_ctx* ctx, uint32_t* counter_dev) {
  if (!counter_dev) { set_error("kfp16_bump_counter: null pointer"); return -1; }
  bump_counter_kernel<<<1, 1, 0, ctx_stream(ctx)>>>(counter_dev);
  count_launch();
  return check_launch("kfp16_bump_counter") ? 0 : -1;
}
int kfp16_scale_f32_to_f16(kfp16_ctx* ctx, const float* src, void* dst, size_t n, const float* scale_dev) {
  if (n == 0) return 0;
  if (!src || !dst) { set_error("kfp16_scale_f32_to_f16: null pointer"); return -1; }
  const size_t n4 = (al16(src) && ((uintptr_t)dst % 8) == 0) ? n / 4 : 0;
  launch_pdl(scale_f32_to_f16_kernel, grid_for(n4 ? n4 : n), kThreads, 0, ctx_stream(ctx), src, (__half*)dst, n4, n, scale_dev);
  count_launch();
  return check_launch("kfp16_scale_f32_to_f16") ? 0 : -1;
}
int kfp16_pad_edges(kfp16_ctx* ctx, void* X, int ld, int n_seq, int seq_len, int cols, int halo) {
  if (n_seq <= 0 || seq_len <= 0 || halo <= 0 || cols <= 0) return 0;
  if (!X || (cols % 8) || (ld % 8) || !al16(X)) { set_error("kfp16_pad_edges: needs a 16B-aligned buffer and cols/ld %% 8 == 0"); return -1; }
  launch_pdl(pad_edges_kernel, grid_for((size_t)n_seq * 2 * halo * (cols / 8)), kThreads, 0, ctx_stream(ctx), (__half*)X, ld, n_seq, seq_len, cols, halo);
  count_launch();
  return check_launch("kfp16_pad_edges") ? 0 : -1;
}
int kfp16_fold_edges(kfp16_ctx* ctx, void* G, int ld, int n_seq, int seq_len, int cols, int halo) {
  if (n_seq <= 0 || seq_len <= 0 || halo <= 0 || cols <= 0) return 0;
  if (!G || (cols % 8) || (ld % 8) || !al16(G)) { set_error("kfp16_fold_edges: needs a 16B-aligned buffer and cols/ld %% 8 == 0"); return -1; }
  launch_pdl(fold_edges_kernel, grid_for((size_t)n_seq * 2 * (cols / 8)), kThreads, 0, ctx_stream(ctx), (__half*)G, ld, n_seq, seq_len, cols, halo);
  count_launch();
  return check_launch("kfp16_fold_edges") ? 0 : -1;
}
static int sgd_flat(kfp16_ctx* ctx, float* w32, void* w16, const void* grad, int grad_is_f32, int round_grad,
                    float grad_scale, float* velocity, float lr, float momentum, size_t n, const float* hp, const char* who) {
  if (n == 0) return 0;
  if (!w32 || !w16 || !grad || !velocity) { set_error("%s: null pointer", who); return -1; }
  cudaStream_t s = ctx_stream(ctx);
  const bool aligned = al16(w32) && al16(velocity) && ((uintptr_t)w16 % 8) == 0 && ((uintptr_t)grad % (grad_is_f32 ? 16 : 8)) == 0;
  const size_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    const int grid = grid_for((n4 + 1) / 2);
    if (grad_is_f32) sgd_kernel_v4<true><<<grid, kThreads, 0, s>>>(w32, (__half*)w16, grad, round_grad, grad_scale, velocity, lr, momentum, n4, hp);
    else sgd_kernel_v4<false><<<grid, kThreads, 0, s>>>(w32, (__half*)w16, grad, 0, grad_scale, velocity, lr, momentum, n4, hp);
    count_launch();
  }
  const size_t done = n4 * 4, rest = n - done;
  if (rest > 0) {
    const char* gp = (const char*)grad + done * (grad_is_f32 ? 4 : 2);
    if (grad_is_f32) sgd_kernel<true><<<grid_for(rest), kThreads, 0, s>>>(w32 + done, (__half*)w16 + done, gp, round_grad, grad_scale, velocity + done, lr, momentum, rest, hp);
    else sgd_kernel<false><<<grid_for(rest), kThreads, 0, s>>>(w32 + done, (__half*)w16 + done, gp, 0, grad_scale, velocity + done, lr, momentum, rest, hp);
    count_launch();
  }
  return check_launch(who) ? 0 : -1;
}
int kfp16_sgd_update_flat(kfp16_ctx* ctx, float* w32, void* w16, const void* grad, int grad_is_f32, int round_grad,
                          float grad_scale, float* velocity, float lr, float momentum, size_t n) {
  return sgd_flat(ctx, w32, w16, grad, grad_is_f32, round_grad, grad_scale, velocity, lr, momentum, n, nullptr, "kfp16_sgd_update_flat");
}
// same update with {lr, momentum, grad_scale} read from a 3-float device block at kernel start: a captured CUDA graph of
// the update then follows SGDOptimizer.SetLR (internal/gpu/optimize.go:123) without re-capture
int kfp16_sgd_update_flat_hp(kfp16_ctx* ctx, float* w32, void* w16, const void* grad, int grad_is_f32, int round_grad,
                             float* velocity, const float* hyper_dev, size_t n) {
  if (!hyper_dev) { set_error("kfp16_sgd_update_flat_hp: null hyper-parameter block"); return -1; }
  return sgd_flat(ctx, w32, w16, grad, grad_is_f32, round_grad, 1.0f, velocity, 0.f, 0.f, n, hyper_dev, "kfp16_sgd_update_flat_hp");
}

int kfp16_pack_rows(kfp16_ctx* ctx, const void* src, void* dst, int ld, int n_seq, int seq_len, int halo, int cols,
                    int mode) {
  if (n_seq <= 0 || seq_len <= 0 || cols <= 0) return 0;
  if (!src || !dst) { set_error("kfp16_pack_rows: null pointer"); return -1; }
  const size_t elems = (size_t)n_seq * (seq_len + 2 * halo) * cols;
  if ((cols % 8) == 0 && (ld % 8) == 0 && al16(src) && al16(dst))
    launch_pdl(pack_rows_kernel<8>, grid_for(elems / 8), kThreads, 0, ctx_stream(ctx), (const __half*)src, (__half*)dst, ld, n_seq, seq_len, halo, cols, mode);
  else
    launch_pdl(pack_rows_kernel<1>, grid_for(elems), kThreads, 0, ctx_stream(ctx), (const __half*)src, (__half*)dst, ld, n_seq, seq_len, halo, cols, mode);
  count_launch();
  return check_launch("kfp16_pack_rows") ? 0 : -1;
}
int kfp16_pack_rows_f32(kfp16_ctx* ctx, const float* src, void* dst, int ld, int n_seq, int seq_len, int halo, int cols,
                        int mode) {
  if (n_seq <= 0 || seq_len <= 0 || cols <= 0) return 0;
  if (!src || !dst) { set_error("kfp16_pack_rows_f32: null pointer"); return -1; }
  const size_t elems = (size_t)n_seq * (seq_len + 2 * halo) * cols;
  launch_pdl(pack_rows_f32_kernel, grid_for(elems), kThreads, 0, ctx_stream(ctx), src, (__half*)dst, ld, n_seq, seq_len, halo, cols, mode);
  count_launch();
  return check_launch("kfp16_pack_rows_f32") ? 0 : -1;
}
int kfp16_unpack_rows(kfp16_ctx* ctx, const void* src, int ld, int col0, void* dst, int n_seq, int seq_len, int halo,
                      int cols) {
  if (n_seq <= 0 || seq_len <= 0 || cols <= 0) return 0;
  if (!src || !dst) { set_error("kfp16_unpack_rows: null pointer"); return -1; }
  unpack_rows_kernel<<<grid_for((size_t)n_seq * seq_len * cols), kThreads, 0, ctx_stream(ctx)>>>(
      (const __half*)src, ld, col0, (__half*)dst, n_seq, seq_len, halo, cols);
  count_launch();
  return check_launch("kfp16_unpack_rows") ? 0 : -1;
}
int kfp16_bcast_rows(kfp16_ctx* ctx, const void* src, int cols, void* dst, int ld, int col0, int rows, int blk) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!src || !dst || blk <= 0) { set_error("kfp16_bcast_rows: bad argument"); return -1; }
  bcast_rows_kernel<<<rows < 148 * 16 ? rows : 148 * 16, cols <= 128 ? 128 : kThreads, 0, ctx_stream(ctx)>>>((const __half*)src, cols, (__half*)dst, ld, col0, (size_t)rows, blk);
  count_launch();
  return check_launch("kfp16_bcast_rows") ? 0 : -1;
}
int kfp16_seq_sum(kfp16_ctx* ctx, const void* G, int ld, int col0, void* out, int cols, int n_seq, int seq_len,
                  int halo) {
  if (n_seq <= 0 || cols <= 0) return 0;
  if (!G || !out) { set_error("kfp16_seq_sum: null pointer"); return -1; }
  seq_sum_kernel<<<dim3(n_seq, (cols + 31) / 32), dim3(32, 8), 0, ctx_stream(ctx)>>>((const __half*)G, ld, col0, (__half*)out, cols, seq_len, halo);
  count_launch();
  return check_launch("kfp16_seq_sum") ? 0 : -1;
}
int kfp16_zero_halo(kfp16_ctx* ctx, void* X, int ld, int n_seq, int seq_len, int cols, int halo) {
  if (n_seq <= 0 || halo <= 0 || cols <= 0) return 0;
  if (!X) { set_error("kfp16_zero_halo: null pointer"); return -1; }
  zero_halo_kernel<<<grid_for((size_t)n_seq * 2 * halo * cols), kThreads, 0, ctx_stream(ctx)>>>((__half*)X, ld, n_seq, seq_len, cols, halo);
  count_launch();
  return check_launch("kfp16_zero_halo") ? 0 : -1;
}
int kfp16_scale_shift(kfp16_ctx* ctx, const void* x, void* y, int rows, int cols, const float* scale,
                      const float* shift) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!x || !y || !scale) { set_error("kfp16_scale_shift: null pointer"); return -1; }
  scale_shift_kernel<<<grid_for((size_t)rows * cols), kThreads, 0, ctx_stream(ctx)>>>((const __half*)x, (__half*)y, (size_t)rows * cols, cols, scale, shift);
  count_launch();
  return check_launch("kfp16_scale_shift") ? 0 : -1;
}
int kfp16_spec_augment(kfp16_ctx* ctx, const void* x, void* y, int ld, int n_seq, int seq_len, int halo, int dim, int fmax, int nfreq,
                       int tmax, int ntime, uint32_t seed, const uint32_t* seed_dev) {
  if (n_seq <= 0 || dim <= 0) return 0;
  if (!x || !y) { set_error("kfp16_spec_augment: null pointer"); return -1; }
  if (nfreq < 0 || nfreq > 8 || ntime < 0 || ntime > 8 || fmax < 0 || fmax > dim || tmax < 0 || tmax > seq_len) {
    set_error("kfp16_spec_augment: up to 8 masks of each kind, widths within the matrix (fmax %d of %d, tmax %d of %d)", fmax, dim, tmax, seq_len);
    return -1;
  }
  spec_augment_kernel<<<n_seq, 256, 0, ctx_stream(ctx)>>>((const __half*)x, (__half*)y, ld, seq_len, halo, dim, fmax, nfreq, tmax, ntime, seed, seed_dev);
  count_launch();
  return check_launch("kfp16_spec_augment") ? 0 : -1;
}
int kfp16_scale_shift_ld(kfp16_ctx* ctx, const void* x, long long ldx, void* y, long long ldy, int rows, int cols,
                         const float* scale, const float* shift) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!x || !y || !scale) { set_error("kfp16_scale_shift_ld: null pointer"); return -1; }
  if ((cols % 8) || (ldx % 8) || (ldy % 8) || !al16(x) || !al16(y)) { set_error("kfp16_scale_shift_ld: cols / ld must be multiples of 8 and the pointers 16-byte aligned"); return -1; }
  scale_shift_ld_kernel<<<grid_for((size_t)rows * (cols / 8)), kThreads, 0, ctx_stream(ctx)>>>((const __half*)x, ldx, (__half*)y, ldy, rows, cols, scale, shift);
  count_launch();
  return check_launch("kfp16_scale_shift_ld") ? 0 : -1;
}
int kfp16_zero_rows_except(kfp16_ctx* ctx, void* X, int ld, int rows, int cols, int row0, int step) {
  if (rows <= 0 || cols <= 0 || step <= 1) return 0;
  if (!X) { set_error("kfp16_zero_rows_except: null pointer"); return -1; }
  if ((cols % 8) || (ld % 8) || !al16(X) || row0 < 0) { set_error("kfp16_zero_rows_except: cols / ld must be multiples of 8, the pointer 16-byte aligned, row0 >= 0"); return -1; }
  zero_rows_except_kernel<<<rows < 148 * 16 ? rows : 148 * 16, (cols / 8) <= 64 ? 64 : 192, 0, ctx_stream(ctx)>>>((__half*)X, ld, rows, cols, row0, step);
  count_launch();
  return check_launch("kfp16_zero_rows_except") ? 0 : -1;
}
int kfp16_half_sq_loss(kfp16_ctx* ctx, const void* Y, void* dY, int n_seq, int seq_len, int halo, int cols,
                       float* loss_dev) {
  if (n_seq <= 0 || cols <= 0) return 0;
  if (!Y || !dY || !loss_dev) { set_error("kfp16_half_sq_loss: null pointer"); return -1; }
  const size_t elems = (size_t)n_seq * (seq_len + 2 * halo) * cols;
  if ((cols % 8) == 0 && al16(Y) && al16(dY))
    launch_pdl(half_sq_loss_kernel<8>, grid_for(elems / 8), kThreads, 0, ctx_stream(ctx), (const __half*)Y, (__half*)dY, n_seq, seq_len, halo, cols, loss_dev);
  else
    launch_pdl(half_sq_loss_kernel<1>, grid_for(elems), kThreads, 0, ctx_stream(ctx), (const __half*)Y, (__half*)dY, n_seq, seq_len, halo, cols, loss_dev);
  count_launch();
  return check_launch("kfp16_half_sq_loss") ? 0 : -1;
}
static int bn_relu_bwd_launch(kfp16_ctx* ctx, void* dY, int ldy, const float* scale, const uint32_t* mask,
                              int mask_ld, void* dZ, int ldz, int rows, int cols, float* db_accum, int blk, int seq_len,
                              int halo) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!dY || !dZ || (cols % 8) || (ldy % 8) || (ldz % 8) || !al16(dY) || !al16(dZ)) {
    set_error("kfp16_bn_relu_backward_bias: needs 16B-aligned buffers and cols/ld %% 8 == 0"); return -1;
  }
  if (blk > 0 && (seq_len < 1 || halo < 1 || blk != seq_len + 2 * halo || rows % blk != 0 || dY == dZ)) {
    set_error("kfp16_bn_relu_backward_bias_fold: rows must be whole sequence blocks of seq_len + 2*halo, out of place"); return -1;
  }
  // narrow dense matrices (conv layers: 64..256 filters, hundreds of thousands of rows): k rows side by side as one wide row
  int col_mod = cols;
  if (blk == 0 && cols < 256 && (cols % 32) == 0 && ldy == cols && ldz == cols && (!mask || mask_ld * 32 == cols)) {
    int k = 1;
    while (cols * k * 2 <= 256 && rows % (k * 2) == 0) k *= 2;
    rows /= k; cols *= k; ldy *= k; ldz *= k; mask_ld *= k;
  }
  const int gx = (cols + 255) / 256;
  // exactly ONE wave of resident blocks (a few blocks more than fit spill into a second, nearly empty wave that
  // doubles the kernel time: measured 25 us -> 12 us on 9984 x 1536)
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    int b0 = 0, b1 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, bn_relu_bwd_colsum_kernel<false>, 256, 0) != cudaSuccess || b0 < 1) b0 = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, bn_relu_bwd_colsum_kernel<true>, 256, 0) != cudaSuccess || b1 < 1) b1 = 1;
    blocks_per_sm = b0 < b1 ? b0 : b1;
  }
  // SMs this context may use (kfp16_ctx_set_max_ctas leaves the rest to a concurrent NCCL kernel)
  int sms = num_sms_cached();
  if (ctx && ctx->max_ctas > 0 && ctx->max_ctas < sms) sms = ctx->max_ctas;
  int gy = (sms * blocks_per_sm) / gx;
  const int max_gy = (rows + 8 * kBnBwdRows - 1) / (8 * kBnBwdRows);
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  if (blk > 0)
    launch_pdl(bn_relu_bwd_colsum_kernel<true>, dim3(gx, gy), dim3(32, 8), 0, ctx_stream(ctx), (__half*)dY, ldy, scale, mask, mask_ld,
               (__half*)dZ, ldz, (uint32_t)rows, cols, db_accum, (uint32_t)blk, (uint32_t)seq_len, (uint32_t)halo, col_mod);
  else
    launch_pdl(bn_relu_bwd_colsum_kernel<false>, dim3(gx, gy), dim3(32, 8), 0, ctx_stream(ctx), (__half*)dY, ldy, scale, mask, mask_ld,
               (__half*)dZ, ldz, (uint32_t)rows, cols, db_accum, 0u, 0u, 0u, col_mod);
  count_launch();
  return check_launch("kfp16_bn_relu_backward_bias") ? 0 : -1;
}
int kfp16_bn_relu_backward_bias(kfp16_ctx* ctx, const void* dY, int ldy, const float* scale, const uint32_t* mask,
                                int mask_ld, void* dZ, int ldz, int rows, int cols, float* db_accum) {
  return bn_relu_bwd_launch(ctx, const_cast<void*>(dY), ldy, scale, mask, mask_ld, dZ, ldz, rows, cols, db_accum, 0, 0, 0);
}
int kfp16_bn_relu_backward_bias_fold(kfp16_ctx* ctx, void* dY, int ldy, const float* scale, const uint32_t* mask,
                                     int mask_ld, void* dZ, int ldz, int n_seq, int seq_len, int halo, int cols,
                                     float* db_accum) {
  return bn_relu_bwd_launch(ctx, dY, ldy, scale, mask, mask_ld, dZ, ldz, n_seq * (seq_len + 2 * halo), cols, db_accum,
                            seq_len + 2 * halo, seq_len, halo);
}
int kfp16_colsum_accum(kfp16_ctx* ctx, const void* X, int ld, int rows, int cols, float* out_f32) {
  if (cols <= 0 || rows <= 0) return 0;
  if (!X || !out_f32 || (cols % 2) || (ld % 2)) { set_error("kfp16_colsum_accum: needs fp32 output and even cols/ld"); return -1; }
  if ((cols % 8) == 0 && (ld % 8) == 0 && al16(X)) {
    // 16-byte loads, four rows in flight per thread, one wave of blocks (measured on 9984 x 6016: 48 -> ~22 us)
    int col_mod = cols;
    if (cols < 256 && (cols % 32) == 0 && ld == cols) {       // narrow dense matrix: k rows side by side
      int k = 1;
      while (cols * k * 2 <= 256 && rows % (k * 2) == 0) k *= 2;
      rows /= k; cols *= k; ld *= k;
    }
    const int gx = (cols + 255) / 256;
    static int per_sm = 0;
    if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, colsum_v2_kernel, 256, 0) != cudaSuccess || per_sm < 1)) per_sm = 1;
    int sms = num_sms_cached();
    if (ctx && ctx->max_ctas > 0 && ctx->max_ctas < sms) sms = ctx->max_ctas;
    int gy = (sms * per_sm) / gx;
    const int max_gy = (rows + 8 * kColsumRows - 1) / (8 * kColsumRows);
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    launch_pdl(colsum_v2_kernel, dim3(gx, gy), dim3(32, 8), 0, ctx_stream(ctx), (const __half*)X, ld, (uint32_t)rows, cols, out_f32, col_mod);
    count_launch();
    return check_launch("kfp16_colsum_accum") ? 0 : -1;
  }
  const int gx = (cols + 63) / 64;
  int gy = (num_sms_cached() * 4 + gx - 1) / gx;
  const int max_gy = (rows + 63) / 64;
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  colsum_kernel<<<dim3(gx, gy), kThreads, 0, ctx_stream(ctx)>>>((const __half*)X, ld, (size_t)rows, cols, out_f32);
  count_launch();
  return check_launch("kfp16_colsum_accum") ? 0 : -1;
}

}  // extern "C"
