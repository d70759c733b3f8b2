// CNN front-end support: patch gather / scatter for the conv-relu-batchnorm layer.
//
// The reference lowers its convolution to a GEMM through an im2col done in Go on the CPU, between a
// D2H and an H2D copy (/root/reference/internal/nnet/forward.go:418-524: patches P[(t*Hout+ho),
// off*Fin+f] = X[t+dt, (ho*sub+dh)*Fin+f], zero outside).  Here the same patch matrix is built on
// the device (one HBM-bound pass) and fed to the tcgen05 GEMM with its bias/ReLU/batch-norm epilogue;
// the backward pass scatters dP back with the adjoint gather (fp32 accumulate, one fp16 rounding).
// Activations stay height-major [T x H*F] with batch-norm per filter (SURVEY Appendix A / oracle header).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "../../include/kaldi_fp16_fused.h"
#include "host_common.h"

using namespace kfp16;

namespace {

constexpr int kMaxTaps = 32;
struct Taps {
  int n;
  int8_t dt[kMaxTaps], dh[kMaxTaps];
};

struct ConvGeom {
  int n_seq, L, halo, blk;   // rows: n_seq blocks of blk = L + 2*halo
  int hin, hout, sub, fin, Kp;
};

// one thread per (row m = r*hout + ho, VEC-wide filter group), looping over the taps: the index arithmetic (32-bit) is
// paid once per row and the taps' loads are independent (9 x 16 bytes in flight per thread).  Consecutive threads
// cover the fin filters of one row, i.e. one contiguous fin*2-byte run per tap on both sides.
template <int VEC>
__global__ void im2col_kernel(const __half* __restrict__ x, __half* __restrict__ P, ConvGeom g, Taps taps) {
  const uint32_t fv = (uint32_t)(g.fin + VEC - 1) / VEC;
  const uint32_t rows = (uint32_t)g.n_seq * g.blk * g.hout;
  const uint32_t total = rows * fv;                       // < 2^32: checked by the launcher
  const uint32_t stride = gridDim.x * blockDim.x;
  const size_t ldx = (size_t)g.hin * g.fin;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint32_t f = (i % fv) * VEC;
    const uint32_t m = i / fv;
    const int ho = (int)(m % (uint32_t)g.hout);
    const uint32_t r = m / (uint32_t)g.hout;              // padded row
    const int local = (int)(r % (uint32_t)g.blk) - g.halo;
    const bool row_ok = local >= 0 && local < g.L;
    __half* dst = P + (size_t)m * g.Kp + f;
    const __half* src = x + (size_t)r * ldx + f;
    if (VEC == 8) {
#pragma unroll 3
      for (int tap = 0; tap < taps.n; ++tap) {
        const int ts = local + taps.dt[tap];
        const int hs = ho * g.sub + taps.dh[tap];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row_ok && ts >= 0 && ts < g.L && hs >= 0 && hs < g.hin)
          v = *reinterpret_cast<const uint4*>(src + (ptrdiff_t)taps.dt[tap] * (ptrdiff_t)ldx + (size_t)hs * g.fin);
        *reinterpret_cast<uint4*>(dst + (size_t)tap * g.fin) = v;
      }
    } else {
      for (int tap = 0; tap < taps.n; ++tap) {
        const int ts = local + taps.dt[tap];
        const int hs = ho * g.sub + taps.dh[tap];
        const bool ok = row_ok && ts >= 0 && ts < g.L && hs >= 0 && hs < g.hin && (int)f < g.fin;
        if ((int)f < g.fin) dst[(size_t)tap * g.fin] = ok ? src[(ptrdiff_t)taps.dt[tap] * (ptrdiff_t)ldx + (size_t)hs * g.fin] : __float2half(0.f);
      }
    }
  }
}

// zero the K padding columns [K, Kp) once per call (K = taps*fin not a multiple of 16)
__global__ void zero_pad_cols_kernel(__half* __restrict__ P, size_t rows, int K, int Kp) {
  const int padw = Kp - K;
  const size_t total = rows * padw;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    P[(i / padw) * Kp + K + (i % padw)] = __float2half(0.f);
}

// adjoint: dx[r, h, f] = sum over taps with ho*sub + dh == h of dP[((r-dt)*hout + ho), tap*fin + f]
template <int VEC>
__global__ void col2im_kernel(const __half* __restrict__ dP, __half* __restrict__ dx, ConvGeom g, Taps taps) {
  const uint32_t fv = (uint32_t)(g.fin + VEC - 1) / VEC;
  const uint32_t total = (uint32_t)g.n_seq * g.blk * g.hin * fv;    // < 2^32: checked by the launcher
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int f = (int)(i % fv) * VEC;
    const uint32_t rest = i / fv;
    const int hh = (int)(rest % (uint32_t)g.hin);
    const size_t r = rest / (uint32_t)g.hin;
    const int local = (int)((uint32_t)r % (uint32_t)g.blk) - g.halo;
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    if (local >= 0 && local < g.L) {
      // taps in groups of 3: the (up to) three 16-byte loads of a group are issued before any of them is used
      for (int tap0 = 0; tap0 < taps.n; tap0 += 3) {
        uint4 v[3];
        bool ok[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int tap = tap0 + q;
          ok[q] = false;
          v[q] = make_uint4(0, 0, 0, 0);
          if (tap < taps.n) {
            const int to = local - taps.dt[tap];     // output frame that read this input frame through `tap`
            const int hs = hh - taps.dh[tap];
            int ho = hs;
            bool hit = to >= 0 && to < g.L && hs >= 0;
            if (g.sub != 1) { hit = hit && (hs % g.sub) == 0; ho = hs / g.sub; }
            hit = hit && ho < g.hout;
            if (hit) {
              const __half* src = dP + ((r - taps.dt[tap]) * g.hout + ho) * g.Kp + (size_t)tap * g.fin + f;
              if (VEC == 8) v[q] = *reinterpret_cast<const uint4*>(src);
              else v[q].x = *reinterpret_cast<const unsigned short*>(src);
              ok[q] = true;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (!ok[q]) continue;
          if (VEC == 8) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&v[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 t = __half22float2(h2[e]); acc[2 * e] += t.x; acc[2 * e + 1] += t.y; }
          } else {
            acc[0] += __half2float(__ushort_as_half((unsigned short)v[q].x));
          }
        }
      }
    }
    __half* dst = dx + r * (size_t)(g.hin * g.fin) + (size_t)hh * g.fin + f;
    if (VEC == 8) {
      uint4 o;
      __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
      *dst = __float2half_rn(acc[0]);
    }
  }
}


// ---- few input filters (the first conv layer: 6 feature maps, K = 54 -> Kp = 64): the generic kernels above move 2 bytes
// per thread and tap and pay several integer divisions per element (measured 46 + 14 us / 44 us for 49 MB).  Here a block
// stages the frames it needs in shared memory with 16-byte global accesses and every thread assembles / sums whole units
// from 2-byte shared-memory reads with table-driven offsets (no division in the inner loops).
// Preconditions (checked by the launchers): sub == 1, hout == hin, Kp <= 64, halo >= max |dt| -- a tap then never leaves
// the sequence block of its frame, and the block's halo rows are staged as zeros = the zero padding in time.
constexpr int kSmallFrames = 8;      // frames per block (gather)
constexpr int kScatterFrames = 4;    // frames per block (scatter: 35 KB of staged patch rows per block -> 6 blocks per SM)
struct SmallLut { short off[64]; };  // im2col: k -> (dt - dtmin)*4096 + (dh + 1)*fin + f, -1 = padding column

__device__ __forceinline__ bool real_row(const ConvGeom& g, int r, int total_rows) {
  if (r < 0 || r >= total_rows) return false;
  const int local = r % g.blk - g.halo;
  return local >= 0 && local < g.L;
}

// P[(r*hout + ho), k] = x[r + dt(k), (ho + dh(k))*fin + f(k)]; block = kSmallFrames consecutive padded rows, threads (unit, height)
__global__ void __launch_bounds__(256)
im2col_small_kernel(const __half* __restrict__ x, __half* __restrict__ P, ConvGeom g, SmallLut lut, int dtmin, int dtspan) {
  extern __shared__ __align__(16) __half sx[];       // [(kSmallFrames + dtspan) frames][rowlen]: lead zeros | x row | zeros
  __shared__ int soff[64];
  __shared__ int srow_ok[kSmallFrames];
  const int ldx = g.hin * g.fin;
  const int rowlen = (ldx + 2 * g.fin + 8 + 7) & ~7;   // >= fin zeros on both sides (heights -1 and hin), 16-byte multiple
  const int lead = (g.fin + 7) & ~7;                   // the x row starts 16-byte aligned inside the shared row
  const int r0 = blockIdx.x * kSmallFrames;
  const int total_rows = g.n_seq * g.blk;
  const int nfr = kSmallFrames + dtspan;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 64) {
    const int o = lut.off[tid];
    soff[tid] = o < 0 ? -1 : (o >> 12) * rowlen + lead - g.fin + (o & 4095);
  }
  if (tid < kSmallFrames) srow_ok[tid] = real_row(g, r0 + tid, total_rows) ? 1 : 0;
  const int row_units = rowlen >> 3;
  for (int fr = threadIdx.y; fr < nfr; fr += blockDim.y) {
    const int r = r0 + dtmin + fr;
    const bool ok = real_row(g, r, total_rows);
    for (int u = threadIdx.x; u < row_units; u += blockDim.x) {
      const int xo = (u << 3) - lead;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ok && xo >= 0 && xo + 8 <= ldx) v = *reinterpret_cast<const uint4*>(x + (size_t)r * ldx + xo);
      *reinterpret_cast<uint4*>(sx + fr * rowlen + (u << 3)) = v;
    }
  }
  __syncthreads();
  const int units = g.Kp >> 3;
  const int u = threadIdx.x;                       // blockDim.x == 8 >= units
  if (u >= units) return;
  int offs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) offs[e] = soff[u * 8 + e];
  for (int fl = 0; fl < kSmallFrames; ++fl) {
    const int r = r0 + fl;
    if (r >= total_rows) break;
    const bool row_ok = srow_ok[fl] != 0;
    const __half* base = sx + fl * rowlen;
    for (int ho = threadIdx.y; ho < g.hout; ho += blockDim.y) {
      __align__(16) __half o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (row_ok && offs[e] >= 0) ? base[offs[e] + ho * g.fin] : __float2half(0.f);
      *reinterpret_cast<uint4*>(P + ((size_t)r * g.hout + ho) * g.Kp + u * 8) = *reinterpret_cast<uint4*>(o);
    }
  }
}

// dx[r, h, f] = sum over taps of dP[((r - dt)*hout + (h - dh)), tap*fin + f]; one thread per (h, f), looping over the block's frames
__global__ void __launch_bounds__(256)
col2im_small_kernel2(const __half* __restrict__ dP, __half* __restrict__ dx, ConvGeom g, Taps taps, int dtmin, int dtspan) {
  extern __shared__ __align__(16) __half sp[];       // [(kScatterFrames + dtspan) frames][hout][Kp + 8]: rows 16 bytes apart in
  const int pitch = g.Kp + 8;                        // bank phase (a 128-byte pitch put every height on the same banks)
  __shared__ int srow_ok[kScatterFrames];
  const int r0 = blockIdx.x * kScatterFrames;
  const int total_rows = g.n_seq * g.blk;
  const int nfr = kScatterFrames + dtspan;
  const int frame_units = g.hout * (g.Kp >> 3);
  const int dtmax = dtmin + dtspan;
  __shared__ int sfr_ok[kScatterFrames + 8];
  if (threadIdx.x < kScatterFrames) srow_ok[threadIdx.x] = real_row(g, r0 + threadIdx.x, total_rows) ? 1 : 0;
  if (threadIdx.x < nfr) sfr_ok[threadIdx.x] = real_row(g, r0 - dtmax + (int)threadIdx.x, total_rows) ? 1 : 0;
  __syncthreads();
  // output frame r reads patch rows of frames r - dt: stage frames r0 - dtmax .. r0 + kScatterFrames - 1 - dtmin (zeros for halo
  // rows).  Eight 16-byte loads in flight per thread: with one at a time the block spent ~12 load latencies here (measured
  // 47 us for the whole kernel, twice what the 64 MB it moves need)
  {
    // (no integer division in this loop: with one per load -- frame = i / frame_units, row = u / units-per-row -- the kernel was
    //  bound by the divisions' instruction count: 43 us for 64 MB)
    const int upr_log2 = 31 - __clz(g.Kp >> 3);       // 16-byte units per patch row: a power of two (checked by the launcher)
    const uint4* src0 = reinterpret_cast<const uint4*>(dP) + (ptrdiff_t)(r0 - dtmax) * frame_units;
    for (int fr0 = 0; fr0 < nfr; fr0 += 4) {
      for (int u0 = threadIdx.x; u0 < frame_units; u0 += 2 * blockDim.x) {
        uint4 v[4][2];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int fr = fr0 + f, u = u0 + h * (int)blockDim.x;
            v[f][h] = make_uint4(0, 0, 0, 0);
            if (fr < nfr && u < frame_units && sfr_ok[fr]) v[f][h] = src0[(ptrdiff_t)fr * frame_units + u];
          }
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int fr = fr0 + f, u = u0 + h * (int)blockDim.x;
            if (fr < nfr && u < frame_units)
              *reinterpret_cast<uint4*>(sp + (size_t)fr * g.hout * pitch + (u >> upr_log2) * pitch + (u & ((1 << upr_log2) - 1)) * 8) = v[f][h];
          }
      }
    }
  }
  __syncthreads();
  const int ldx = g.hin * g.fin;
  const int frame_halves = g.hout * pitch;
  for (int idx = threadIdx.x; idx < ldx; idx += blockDim.x) {
    const int hh = idx / g.fin, f = idx - hh * g.fin;
    float acc[kScatterFrames];
#pragma unroll
    for (int fl = 0; fl < kScatterFrames; ++fl) acc[fl] = 0.f;
    // taps outermost (their offsets are read once), the block's frames inside: each frame still sums its taps in tap order
    for (int tap = 0; tap < taps.n; ++tap) {
      const int ho = hh - taps.dh[tap];
      if (ho < 0 || ho >= g.hout) continue;
      const __half* col = sp + (dtmax - taps.dt[tap]) * frame_halves + ho * pitch + tap * g.fin + f;
#pragma unroll
      for (int fl = 0; fl < kScatterFrames; ++fl) acc[fl] += __half2float(col[fl * frame_halves]);
    }
#pragma unroll
    for (int fl = 0; fl < kScatterFrames; ++fl) {
      const int r = r0 + fl;
      if (r < total_rows) dx[(size_t)r * ldx + idx] = __float2half_rn(srow_ok[fl] ? acc[fl] : 0.f);
    }
  }
}

int grid_for_elems(size_t work) {
  size_t blocks = (work + 255) / 256;
  const size_t cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

bool fill_geom(ConvGeom& g, Taps& t, int n_seq, int seq_len, int halo, int hin, int hout, int sub, int fin, int Kp,
               int ntaps, const int* dt, const int* dh, const char* who) {
  if (ntaps < 1 || ntaps > kMaxTaps) { set_error("%s: 1..%d taps supported (got %d)", who, kMaxTaps, ntaps); return false; }
  if (n_seq <= 0 || seq_len <= 0 || halo < 0 || hin <= 0 || hout <= 0 || sub <= 0 || fin <= 0 || Kp < ntaps * fin) {
    set_error("%s: bad geometry", who); return false;
  }
  g.n_seq = n_seq; g.L = seq_len; g.halo = halo; g.blk = seq_len + 2 * halo;
  g.hin = hin; g.hout = hout; g.sub = sub; g.fin = fin; g.Kp = Kp;
  t.n = ntaps;
  for (int i = 0; i < ntaps; ++i) {
    if (dt[i] < -halo - 64 || dt[i] > 127 || dh[i] < -128 || dh[i] > 127) { set_error("%s: tap offset out of range", who); return false; }
    t.dt[i] = (int8_t)dt[i]; t.dh[i] = (int8_t)dh[i];
  }
  return true;
}

}  // namespace

extern "C" {

int kfp16_im2col(kfp16_ctx* ctx, const void* x, void* P, int Kp, int n_seq, int seq_len, int halo, int hin, int hout,
                 int sub, int fin, int ntaps, const int* dt, const int* dh) {
  if (!x || !P) { set_error("kfp16_im2col: null pointer"); return -1; }
  ConvGeom g; Taps t;
  if (!fill_geom(g, t, n_seq, seq_len, halo, hin, hout, sub, fin, Kp, ntaps, dt, dh, "kfp16_im2col")) return -1;
  cudaStream_t s = ctx ? ctx->stream : default_stream();
  const size_t rows = (size_t)n_seq * g.blk * hout;
  const int K = ntaps * fin;
  int dtmin = dt[0], dtmax = dt[0];
  for (int i = 1; i < ntaps; ++i) { dtmin = std::min(dtmin, dt[i]); dtmax = std::max(dtmax, dt[i]); }
  // few input filters (first conv layer): shared-memory staged gather, pad columns included
  if (sub == 1 && hin == hout && (fin % 8) != 0 && Kp <= 64 && (Kp % 8) == 0 && ((hin * fin) % 8) == 0 && dtmax - dtmin <= 6 &&
      halo >= std::max(dtmax, -dtmin) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)P & 15) == 0) {
    SmallLut lut;
    bool ok = true;
    for (int k = 0; k < 64; ++k) {
      lut.off[k] = -1;
      if (k >= K) continue;
      const int tap = k / fin, f = k % fin;
      const int rest = (dh[tap] + 1) * fin + f;              // relative to (height ho - 1): needs dh >= -1
      if (dh[tap] < -1 || dh[tap] > 1 || rest >= 4096) { ok = false; break; }
      lut.off[k] = (short)((dt[tap] - dtmin) * 4096 + rest);
      if ((dt[tap] - dtmin) >= 7) { ok = false; break; }      // short range
    }
    const int ldx = hin * fin;
    const int rowlen = (ldx + 2 * fin + 8 + 7) & ~7;
    const size_t smem = (size_t)(kSmallFrames + dtmax - dtmin) * rowlen * 2;
    if (ok && smem <= 48 * 1024) {
      const int blocks = (n_seq * g.blk + kSmallFrames - 1) / kSmallFrames;
      im2col_small_kernel<<<blocks, dim3(8, 32), smem, s>>>((const __half*)x, (__half*)P, g, lut, dtmin, dtmax - dtmin);
      count_launch();
      return check_launch("kfp16_im2col") ? 0 : -1;
    }
  }
  if (Kp > K) {
    zero_pad_cols_kernel<<<grid_for_elems(rows * (Kp - K)), 256, 0, s>>>((__half*)P, rows, K, Kp);
    count_launch();
  }
  const bool vec = (fin % 8) == 0 && (Kp % 8) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)P & 15) == 0;
  if (rows * (size_t)fin >= 0xFFFFFFFFull) { set_error("kfp16_im2col: more than 2^32 patch rows x filters"); return -1; }
  if (vec) im2col_kernel<8><<<grid_for_elems(rows * (fin / 8)), 256, 0, s>>>((const __half*)x, (__half*)P, g, t);
  else im2col_kernel<1><<<grid_for_elems(rows * fin), 256, 0, s>>>((const __half*)x, (__half*)P, g, t);
  count_launch();
  return check_launch("kfp16_im2col") ? 0 : -1;
}

int kfp16_col2im(kfp16_ctx* ctx, const void* dP, int Kp, void* dx, int n_seq, int seq_len, int halo, int hin, int hout,
                 int sub, int fin, int ntaps, const int* dt, const int* dh) {
  if (!dP || !dx) { set_error("kfp16_col2im: null pointer"); return -1; }
  ConvGeom g; Taps t;
  if (!fill_geom(g, t, n_seq, seq_len, halo, hin, hout, sub, fin, Kp, ntaps, dt, dh, "kfp16_col2im")) return -1;
  cudaStream_t s = ctx ? ctx->stream : default_stream();
  int dtmin = dt[0], dtmax = dt[0];
  for (int i = 1; i < ntaps; ++i) { dtmin = std::min(dtmin, dt[i]); dtmax = std::max(dtmax, dt[i]); }
  if (sub == 1 && hin == hout && (fin % 8) != 0 && Kp <= 64 && (Kp % 8) == 0 && (((Kp >> 3) & ((Kp >> 3) - 1)) == 0) && dtmax - dtmin <= 6 &&
      halo >= std::max(dtmax, -dtmin) && ((uintptr_t)dP & 15) == 0) {
    const size_t smem = (size_t)(kScatterFrames + dtmax - dtmin) * hout * (Kp + 8) * 2;
    if (smem <= 99 * 1024) {
      static bool attr = false;
      if (!attr) {
        if (!check_cuda(cudaFuncSetAttribute(col2im_small_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, 99 * 1024), "col2im smem attribute")) return -1;
        attr = true;
      }
      const int blocks = (n_seq * g.blk + kScatterFrames - 1) / kScatterFrames;
      col2im_small_kernel2<<<blocks, 256, smem, s>>>((const __half*)dP, (__half*)dx, g, t, dtmin, dtmax - dtmin);
      count_launch();
      return check_launch("kfp16_col2im") ? 0 : -1;
    }
  }
  const size_t elems = (size_t)n_seq * g.blk * hin;
  if (elems * (size_t)fin >= 0xFFFFFFFFull) { set_error("kfp16_col2im: more than 2^32 input elements"); return -1; }
  const bool vec = (fin % 8) == 0 && (Kp % 8) == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)dP & 15) == 0;
  if (vec) col2im_kernel<8><<<grid_for_elems(elems * (fin / 8)), 256, 0, s>>>((const __half*)dP, (__half*)dx, g, t);
  else col2im_kernel<1><<<grid_for_elems(elems * fin), 256, 0, s>>>((const __half*)dP, (__half*)dx, g, t);
  count_launch();
  return check_launch("kfp16_col2im") ? 0 : -1;
}

}  // extern "C"
