// Restricted self-attention (attention-relu-batchnorm-layer) on the device.
//
// The reference computes this layer on the CPU between a D2H and an H2D copy
// (/root/reference/internal/nnet/forward.go:795-909: projection GEMM on the GPU, then a Go loop nest over heads, frames and
// context positions, then upload + ReLU + batch-norm).  Per head the projection row of a frame holds
//     [ key (K) | value (V) | query key part (K) | query context part (C) ],   C = 1 + num-left + num-right,
// and output frame t attends to the C frames t + (o - num_left)*stride of its own sequence (zeros outside):
//     b[o] = q_ctx[o] + key_scale * <q_key, key[t + (o - nl)*s]>,  w = softmax(b),
//     out  = [ sum_o w[o] * value[t + (o - nl)*s]  |  w ]  ->  ReLU  ->  batch-norm.
// Here: one warp per (frame, head), lanes = context positions for the scores and the softmax, lanes = value / key columns for
// the weighted sums; FP32 arithmetic on the FP16 projection, ReLU + folded batch-norm fused into the store.  The backward pass
// is the exact transpose in two gather passes (no atomics, reproducible): (A) per query frame the score gradients db[o],
// (B) per projection row everything that flows into it -- its own query parts and the key / value contributions of the
// query frames that attended to it.  (The reference back-propagates the layer as a plain affine, network_backward.go:539-545.)
// Rows are the executor's padded layout: n_seq blocks of seq_len + 2*halo rows; halo rows are written as zeros.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/kaldi_fp16_fused.h"
#include "host_common.h"

using namespace kfp16;

namespace {

struct AttGeom {
  int n_seq, L, halo, blk;
  int H, K, V, C, nl, stride, per, out_per;   // per = 2K + V + C (projection columns per head), out_per = V + C
  float ks;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}

// one warp per (padded row, head)
__global__ void __launch_bounds__(128)
attention_fwd_kernel(const __half* __restrict__ proj, int ldp, __half* __restrict__ z, __half* __restrict__ y, int ldy,
                     const float* __restrict__ scale, const float* __restrict__ shift, AttGeom g) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)g.n_seq * g.blk * g.H;
  if (wid >= total) return;
  const int hd = (int)(wid % g.H);
  const long long r = wid / g.H;
  const int t = (int)(r % g.blk) - g.halo;
  const int oc0 = hd * g.out_per;
  if (t < 0 || t >= g.L) {      // halo row: zeros
    for (int d = lane; d < g.out_per; d += 32) {
      y[r * ldy + oc0 + d] = __float2half(0.f);
      z[r * ldy + oc0 + d] = __float2half(0.f);
    }
    return;
  }
  const __half* q = proj + r * ldp + (size_t)hd * g.per;
  // scores: lane o owns context position o
  float b = -INFINITY;
  const int tc = t + (lane - g.nl) * g.stride;
  const bool in_seq = lane < g.C && tc >= 0 && tc < g.L;
  if (lane < g.C) {
    float dot = 0.f;
    if (in_seq) {
      const __half* krow = q + (ptrdiff_t)(tc - t) * ldp;          // key part of frame tc, same head
      for (int d = 0; d < g.K; ++d) dot = fmaf(__half2float(q[g.K + g.V + d]), __half2float(krow[d]), dot);
    }
    b = __half2float(q[2 * g.K + g.V + lane]) + g.ks * dot;
  }
  const float m = warp_max(b);
  const float e = lane < g.C ? expf(b - m) : 0.f;
  const float w = e / warp_sum(e);
  // weighted values: lanes over the value columns
  for (int d0 = 0; d0 < g.V; d0 += 32) {
    const int d = d0 + lane;
    float u = 0.f;
    for (int o = 0; o < g.C; ++o) {
      const float wo = __shfl_sync(0xFFFFFFFFu, w, o);
      const int to = t + (o - g.nl) * g.stride;
      if (d < g.V && to >= 0 && to < g.L) u = fmaf(wo, __half2float(q[(ptrdiff_t)(to - t) * ldp + g.K + d]), u);
    }
    if (d < g.V) {
      const __half zr = __float2half_rn(fmaxf(u, 0.f));
      z[r * ldy + oc0 + d] = zr;
      y[r * ldy + oc0 + d] = __float2half_rn(fmaf(__half2float(zr), scale[oc0 + d], shift[oc0 + d]));
    }
  }
  if (lane < g.C) {             // the attention weights themselves are outputs too (always > 0: ReLU is the identity)
    const __half zr = __float2half_rn(w);
    z[r * ldy + oc0 + g.V + lane] = zr;
    y[r * ldy + oc0 + g.V + lane] = __float2half_rn(fmaf(__half2float(zr), scale[oc0 + g.V + lane], shift[oc0 + g.V + lane]));
  }
}

// dZ element of (row r, output column c): mask ? h(dY * scale) : 0   (ops_batchnorm_backward + ops_relu_backward)
__device__ __forceinline__ float dz_at(const __half* __restrict__ dy, const __half* __restrict__ z, const float* __restrict__ scale,
                                       long long r, int ldy, int c) {
  if (!(__half2float(z[r * ldy + c]) > 0.f)) return 0.f;
  return __half2float(__float2half_rn(__half2float(dy[r * ldy + c]) * scale[c]));
}

// pass A: db[r, head, o] = w[o] * (gw[o] - sum_o' w[o'] gw[o']),  gw[o] = dW[o] + <dU, value[t + (o - nl)*s]>
__global__ void __launch_bounds__(128)
attention_bwd_scores_kernel(const __half* __restrict__ proj, int ldp, const __half* __restrict__ z, const __half* __restrict__ dy,
                            int ldy, const float* __restrict__ scale, float* __restrict__ db, AttGeom g) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)g.n_seq * g.blk * g.H;
  if (wid >= total) return;
  const int hd = (int)(wid % g.H);
  const long long r = wid / g.H;
  const int t = (int)(r % g.blk) - g.halo;
  float* out = db + wid * 32;
  if (t < 0 || t >= g.L) { out[lane] = 0.f; return; }
  const int oc0 = hd * g.out_per;
  const __half* q = proj + r * ldp + (size_t)hd * g.per;
  float gw = 0.f, w = 0.f;
  if (lane < g.C) {
    w = __half2float(z[r * ldy + oc0 + g.V + lane]);
    gw = dz_at(dy, z, scale, r, ldy, oc0 + g.V + lane);
    const int tc = t + (lane - g.nl) * g.stride;
    if (tc >= 0 && tc < g.L) {
      const __half* vrow = q + (ptrdiff_t)(tc - t) * ldp + g.K;
      for (int d = 0; d < g.V; ++d) gw = fmaf(dz_at(dy, z, scale, r, ldy, oc0 + d), __half2float(vrow[d]), gw);
    }
  }
  const float s = warp_sum(w * gw);
  out[lane] = lane < g.C ? w * (gw - s) : 0.f;
}

// pass B: the gradient of projection row (r, head): its query parts from its own scores, its key / value parts gathered
// from the query frames t_o = t - (o - nl)*s that attended to it at context position o
__global__ void __launch_bounds__(128)
attention_bwd_proj_kernel(const __half* __restrict__ proj, int ldp, const __half* __restrict__ z, const __half* __restrict__ dy,
                          int ldy, const float* __restrict__ scale, const float* __restrict__ db, __half* __restrict__ dproj,
                          AttGeom g) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)g.n_seq * g.blk * g.H;
  if (wid >= total) return;
  const int hd = (int)(wid % g.H);
  const long long r = wid / g.H;
  const int t = (int)(r % g.blk) - g.halo;
  __half* dst = dproj + r * ldp + (size_t)hd * g.per;
  if (t < 0 || t >= g.L) {
    for (int d = lane; d < g.per; d += 32) dst[d] = __float2half(0.f);
    return;
  }
  const int oc0 = hd * g.out_per;
  const __half* q = proj + r * ldp + (size_t)hd * g.per;
  const float* dbr = db + wid * 32;
  // query context part
  if (lane < g.C) dst[2 * g.K + g.V + lane] = __float2half_rn(dbr[lane]);
  // key columns: d q_key and d key
  for (int d0 = 0; d0 < g.K; d0 += 32) {
    const int d = d0 + lane;
    float dqk = 0.f, dk = 0.f;
    for (int o = 0; o < g.C; ++o) {
      const int sh = (o - g.nl) * g.stride;
      const int tc = t + sh;                        // frame this row attended to at position o
      const int tq = t - sh;                        // query frame that attended to this row at position o
      if (d < g.K) {
        if (tc >= 0 && tc < g.L) dqk = fmaf(dbr[o], __half2float(q[(ptrdiff_t)sh * ldp + d]), dqk);
        if (tq >= 0 && tq < g.L) {
          const long long wq = wid - (long long)sh * g.H;           // same head, row r - sh
          dk = fmaf(db[wq * 32 + o], __half2float(q[-(ptrdiff_t)sh * ldp + g.K + g.V + d]), dk);
        }
      }
    }
    if (d < g.K) {
      dst[g.K + g.V + d] = __float2half_rn(g.ks * dqk);
      dst[d] = __float2half_rn(g.ks * dk);
    }
  }
  // value columns
  for (int d0 = 0; d0 < g.V; d0 += 32) {
    const int d = d0 + lane;
    float dv = 0.f;
    for (int o = 0; o < g.C; ++o) {
      const int sh = (o - g.nl) * g.stride;
      const int tq = t - sh;
      if (d < g.V && tq >= 0 && tq < g.L) {
        const long long rq = r - sh;
        const float wo = __half2float(z[rq * ldy + oc0 + g.V + o]);
        dv = fmaf(wo, dz_at(dy, z, scale, rq, ldy, oc0 + d), dv);
      }
    }
    if (d < g.V) dst[g.K + d] = __float2half_rn(dv);
  }
}

bool fill(AttGeom& g, int n_seq, int seq_len, int halo, int heads, int key_dim, int value_dim, int n_left, int n_right, int stride,
          float key_scale, const char* who) {
  if (n_seq <= 0 || seq_len <= 0 || halo < 0 || heads <= 0 || key_dim <= 0 || value_dim <= 0 || n_left < 0 || n_right < 0 || stride < 1) {
    set_error("%s: bad geometry", who); return false;
  }
  if (1 + n_left + n_right > 32) { set_error("%s: at most 32 context positions (got %d)", who, 1 + n_left + n_right); return false; }
  g.n_seq = n_seq; g.L = seq_len; g.halo = halo; g.blk = seq_len + 2 * halo;
  g.H = heads; g.K = key_dim; g.V = value_dim; g.C = 1 + n_left + n_right; g.nl = n_left; g.stride = stride;
  g.per = 2 * key_dim + value_dim + g.C; g.out_per = value_dim + g.C; g.ks = key_scale;
  return true;
}

}  // namespace

extern "C" {

int kfp16_attention_forward(kfp16_ctx* ctx, const void* proj, int ldp, void* z, void* y, int ldy, const float* scale, const float* shift,
                            int n_seq, int seq_len, int halo, int heads, int key_dim, int value_dim, int n_left, int n_right, int stride,
                            float key_scale) {
  AttGeom g;
  if (!fill(g, n_seq, seq_len, halo, heads, key_dim, value_dim, n_left, n_right, stride, key_scale, "kfp16_attention_forward")) return -1;
  if (!proj || !z || !y || !scale || !shift || ldp < heads * g.per || ldy < heads * g.out_per) { set_error("kfp16_attention_forward: null pointer / short rows"); return -1; }
  const long long warps = (long long)n_seq * g.blk * heads;
  attention_fwd_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ctx ? ctx->stream : default_stream()>>>(
      (const __half*)proj, ldp, (__half*)z, (__half*)y, ldy, scale, shift, g);
  count_launch();
  return check_launch("kfp16_attention_forward") ? 0 : -1;
}

int kfp16_attention_backward(kfp16_ctx* ctx, const void* proj, int ldp, const void* z, const void* dy, int ldy, const float* scale,
                             float* db_scratch, void* dproj, int n_seq, int seq_len, int halo, int heads, int key_dim, int value_dim,
                             int n_left, int n_right, int stride, float key_scale) {
  AttGeom g;
  if (!fill(g, n_seq, seq_len, halo, heads, key_dim, value_dim, n_left, n_right, stride, key_scale, "kfp16_attention_backward")) return -1;
  if (!proj || !z || !dy || !scale || !db_scratch || !dproj) { set_error("kfp16_attention_backward: null pointer"); return -1; }
  const long long warps = (long long)n_seq * g.blk * heads;
  cudaStream_t s = ctx ? ctx->stream : default_stream();
  attention_bwd_scores_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, s>>>((const __half*)proj, ldp, (const __half*)z, (const __half*)dy, ldy, scale,
                                                                        db_scratch, g);
  attention_bwd_proj_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, s>>>((const __half*)proj, ldp, (const __half*)z, (const __half*)dy, ldy, scale,
                                                                      db_scratch, (__half*)dproj, g);
  count_launch(2);
  return check_launch("kfp16_attention_backward") ? 0 : -1;
}

}  // extern "C"
