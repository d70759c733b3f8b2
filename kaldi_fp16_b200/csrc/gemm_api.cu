// Host side of the tcgen05 GEMM: TMA tensor maps, tile-shape dispatch, the fused C-ABI
// (include/kaldi_fp16_fused.h) and the reference's ops_gemm / ops_cublas_* entry points
// (/root/reference/cpp/cuda/ops.cu:336-433) re-implemented on top of it.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/kaldi_fp16_fused.h"
#include "../../include/kaldi_fp16_ops.h"
#include "gemm_launch.cuh"

namespace kfp16 {

// ------------------------------------------------------------------ errors / counters
static thread_local char g_err[512] = {0};
std::atomic<unsigned long long> g_launches{0};
static std::atomic<unsigned long long> g_kind_launches[EK_COUNT];   // GEMM launches per epilogue kind (tests: which bodies ran)
void count_gemm_kind(int ek) { if (ek >= 0 && ek < EK_COUNT) g_kind_launches[ek].fetch_add(1, std::memory_order_relaxed); }
void prefer_max_smem_carveout(const void* kern) {
  static std::mutex mu;
  static std::unordered_map<const void*, bool> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(kern)) return;
  done[kern] = true;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaGetLastError();   // a kernel that cannot take the hint keeps its default
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("KFP16_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
static cudaStream_t g_default_stream = nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err[0] ? g_err : nullptr; }
void clear_error() { g_err[0] = 0; }
cudaStream_t default_stream() { return g_default_stream; }

bool check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return false;
}
bool check_launch(const char* what) { return check_cuda(cudaGetLastError(), what); }

// ------------------------------------------------------------------ TMA tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D fp16 row-major matrix [outer x inner], `ld` elements between rows, 128B-swizzled boxes.
static bool make_map_2d(CUtensorMap* m, const void* base, long long inner, long long outer,
                        long long ld, int box_inner, int box_outer, const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return false; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 8) != 0 || inner < 1 || outer < 1) {
    set_error("%s: TMA needs 16-byte aligned base and ld %% 8 == 0 (ptr=%p ld=%lld inner=%lld outer=%lld)",
              what, base, ld, inner, outer);
    return false;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld box=%dx%d", what,
              (int)r, inner, outer, ld, box_inner, box_outer);
    return false;
  }
  return true;
}

// ------------------------------------------------------------------ dispatch
// one translation unit per tile width (gemm_bn*.cu)
template <int BN> bool launch_gemm_bn(kfp16_ctx* ctx, const GemmParams& p, const GemmLaunch& L);
extern template bool launch_gemm_bn<64>(kfp16_ctx*, const GemmParams&, const GemmLaunch&);
extern template bool launch_gemm_bn<128>(kfp16_ctx*, const GemmParams&, const GemmLaunch&);
extern template bool launch_gemm_bn<160>(kfp16_ctx*, const GemmParams&, const GemmLaunch&);
extern template bool launch_gemm_bn<256>(kfp16_ctx*, const GemmParams&, const GemmLaunch&);

// the specialised kind whose flag set equals `flags` exactly, else EK_GENERIC
static int pick_kind(uint32_t flags) {
  if (flags & EPI_SPLITK) return EK_SPLITK;
  for (int k = EK_PLAIN; k < EK_COUNT; ++k)
    if (k != EK_SPLITK && epi_kind_flags(k) == flags) return k;
  return EK_GENERIC;
}

static int pick_bn(int N, int K, int m_tiles, int groups, int split_k, int ctas) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  if (N <= 160) return 160;
  if (N <= 256) return 256;
  // N > 256: weigh wave quantisation (cost ~ waves * BN; a wider tile re-reads A less)
  const int cand[2] = {256, 128};
  int best = 256;
  double best_cost = 1e30;
  for (int i = 0; i < 2; ++i) {
    const int bn = cand[i];
    const long long tiles = (long long)m_tiles * ((N + bn - 1) / bn) * groups * split_k;
    const long long waves = (tiles + ctas - 1) / ctas;
    // Deep K (the main loop dominates; no layer of the network, but e.g. 4096^3): a k-block costs the larger of its MMA time
    // (~2.85 cycles per tile column at the measured 1671 TF) and its operand loads (32 KB of A + 128 B per column of B for a CTA
    // pair at ~90 B/cycle): a 128-wide pair tile is load-bound (546 vs 365 cycles), a 256-wide one balanced (730).  Measured on
    // 4096^3: 155 us with 128-wide tiles (7 waves) against the 4-wave bound of ~95 us.
    double cost = (double)waves * (bn + 24);
    if (K >= 1024) cost = (double)waves * std::max(2.85 * bn, (32768.0 + 128.0 * bn) / 90.0);
    if (cost < best_cost * 0.97) { best_cost = cost; best = bn; }
  }
  return best;
}

// Generic SIMT fallback for shapes TMA cannot address (ld % 8 != 0, K == 1 bias trick, ...).
// fp16 in, fp32 accumulate, fp16 out -- same numerics contract, negligible share of the work.
__global__ void gemm_simt_fallback(int M, int N, int K, float alpha, const __half* A, long long a_sm,
                                   long long a_sk, const __half* B, long long b_sk, long long b_sn,
                                   float beta, __half* C, int ldc) {
  __shared__ float sA[16][17], sB[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    sA[ty][tx] = (row < M && k0 + tx < K) ? __half2float(A[row * a_sm + (k0 + tx) * a_sk]) : 0.f;
    sB[ty][tx] = (k0 + ty < K && col < N) ? __half2float(B[(k0 + ty) * b_sk + col * b_sn]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sA[ty][k], sB[k][tx], acc);
    __syncthreads();
  }
  if (row < M && col < N) {
    float v = alpha * acc;
    if (beta != 0.f) v += beta * __half2float(C[(size_t)row * ldc + col]);
    C[(size_t)row * ldc + col] = __float2half_rn(v);
  }
}

// shared memory the epilogue of a kernel with epilogue kind `ek` and tile width `bn` reserves (GemmCfg::kEpiBytes)
static int epi_bytes_for(int ek, int bn) {
  const uint32_t f = epi_kind_flags(ek);
  const bool generic = ek == EK_GENERIC;
  const bool may_r = generic || (f & (EPI_RESID | EPI_BETA)) != 0;
  const bool vec = generic || (f & (EPI_BIAS | EPI_BN)) != 0;
  const int ring = ek == EK_SPLITK ? 0 : (may_r ? (bn <= 160 ? 6 : 4) : 4);
  return ring * kChunkBytes + (vec ? 2 * kVecBytes : 0);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace kfp16

using namespace kfp16;

extern "C" {

// ------------------------------------------------------------------ context
kfp16_ctx* kfp16_ctx_create(int device_id) {
  int ndev = 0;
  if (!check_cuda(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount")) return nullptr;
  if (device_id < 0 || device_id >= ndev) { set_error("invalid device %d (have %d)", device_id, ndev); return nullptr; }
  if (!check_cuda(cudaSetDevice(device_id), "cudaSetDevice")) return nullptr;
  cudaDeviceProp prop;
  if (!check_cuda(cudaGetDeviceProperties(&prop, device_id), "cudaGetDeviceProperties")) return nullptr;
  if (prop.major != 10) {
    set_error("kaldi_fp16_b200 needs an sm_100 GPU (found sm_%d%d); there is no fallback path", prop.major, prop.minor);
    return nullptr;
  }
  kfp16_ctx* c = new kfp16_ctx();
  c->device = device_id;
  c->num_sms = prop.multiProcessorCount;
  c->stream = g_default_stream;
  return c;
}
void kfp16_ctx_destroy(kfp16_ctx* ctx) {
  if (!ctx) return;
  if (ctx->ws) cudaFree(ctx->ws);
  delete ctx;
}
int kfp16_ctx_set_stream(kfp16_ctx* ctx, void* s) { if (!ctx) return -1; ctx->stream = (cudaStream_t)s; return 0; }
void* kfp16_ctx_get_stream(kfp16_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int kfp16_ctx_num_sms(kfp16_ctx* ctx) { return ctx ? ctx->num_sms : 0; }
int kfp16_ctx_set_max_ctas(kfp16_ctx* ctx, int n) { if (!ctx) return -1; ctx->max_ctas = n; return 0; }
int kfp16_ctx_set_profile(kfp16_ctx* ctx, int on) {
  if (!ctx) return -1;
  ctx->profile = on != 0;
  return 0;
}
int kfp16_ctx_profile_read(kfp16_ctx* ctx, int* launches, double* total_ms, double* total_flops) {
  if (!ctx) { set_error("kfp16_ctx_profile_read: null context"); return -1; }
  if (!check_cuda(cudaStreamSynchronize(ctx->stream), "profile sync")) return -1;
  double ms = 0, fl = 0;
  const int n = (int)ctx->prof_flops.size();
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]) == cudaSuccess) ms += t;
    fl += ctx->prof_flops[i];
    if (getenv("KFP16_PROFILE_DUMP")) fprintf(stderr, "[kfp16 gemm] %8.2f us %7.1f TF  %s\n", t * 1e3, ctx->prof_flops[i] / (t * 1e9), ctx->prof_desc[i].c_str());
    cudaEventDestroy(ctx->prof_ev[2 * i]);
    cudaEventDestroy(ctx->prof_ev[2 * i + 1]);
  }
  ctx->prof_ev.clear();
  ctx->prof_flops.clear();
  ctx->prof_desc.clear();
  if (launches) *launches = n;
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  return 0;
}
void kfp16_set_default_stream(void* s) { g_default_stream = (cudaStream_t)s; }
unsigned long long kfp16_launch_count(void) { return g_launches.load(); }
unsigned long long kfp16_gemm_kind_launches(int kind) { return (kind >= 0 && kind < EK_COUNT) ? g_kind_launches[kind].load() : 0ull; }
const char* kfp16_last_error(void) { return get_error(); }
float kfp16_dropout_uniform(uint32_t seed, uint32_t row, uint32_t col) { return dropout_uniform(seed, row, col); }

// ------------------------------------------------------------------ fused GEMM
// 4-D fp16 tensor [T][H][P][C] (C innermost), box [tbox][rows_h][1][64], 128B swizzle: the operand of an implicit-GEMM
// convolution.  Out-of-bounds box elements (negative / past-the-end coordinates) are zero-filled = zero padding.
static bool make_map_conv(CUtensorMap* m, const void* base, int C, int P, int H, long long T, int rows_h, int tbox, const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return false; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (C % 64) != 0 || P < 1 || H < 1 || T < 1 || rows_h < 1 || rows_h > 256 || tbox < 1 || tbox > 256) {
    set_error("%s: convolution operand needs a 16-byte aligned base, channels %% 64 == 0 (C=%d P=%d H=%d T=%lld rows_h=%d tbox=%d)", what, C, P, H, T, rows_h, tbox);
    return false;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)P, (cuuint64_t)H, (cuuint64_t)T};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)P * C * 2, (cuuint64_t)H * P * C * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)rows_h, (cuuint32_t)tbox};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled (4-D) failed (%d)", what, (int)r); return false; }
  return true;
}

int kfp16_gemm_ex(kfp16_ctx* ctx, const kfp16_gemm_desc* d) {
  if (!ctx || !d) { set_error("kfp16_gemm_ex: null argument"); return -1; }
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return 0;   // empty problem: nothing to do
  const int conv = d->conv.mode;
  if (conv < 0 || conv > 2) { set_error("kfp16_gemm_ex: conv.mode must be 0, 1 or 2"); return -1; }
  const int groups = d->groups < 1 ? 1 : d->groups;
  const int kslabs = conv ? 1 : (d->kslabs < 1 ? 1 : d->kslabs);
  const int kslab_len = conv == 1 ? d->conv.C : (d->kslab_len > 0 && !conv ? d->kslab_len : d->K);
  if (groups > 2 || kslabs > kMaxSlabs) { set_error("kfp16_gemm_ex: at most 2 groups / 2 K-slabs"); return -1; }
  const bool two = d->A2.ptr != nullptr;    // a second split-K problem of the same shape in this launch
  if (!conv && kslabs * kslab_len != d->K) { set_error("kfp16_gemm_ex: K (%d) != kslabs*kslab_len (%d*%d)", d->K, kslabs, kslab_len); return -1; }
  // N is the contiguous dimension of D (and of an MN-major B); K only has to be 16-byte aligned
  // where it is some operand's contiguous dimension, which the tensor-map builder checks (ld % 8)
  if (d->N % 8) { set_error("kfp16_gemm_ex: N must be a multiple of 8 (N=%d)", d->N); return -1; }
  if (kslabs > 1 && (kslab_len % 16)) {
    // a partial UMMA K step relies on TMA zero-fill past the matrix edge; between slabs there is none
    set_error("kfp16_gemm_ex: kslab_len must be a multiple of 16 when K is spliced (got %d)", kslab_len); return -1;
  }
  const bool a_mn = d->a_major == KFP16_MN_MAJOR, b_mn = d->b_major == KFP16_MN_MAJOR;
  uint32_t flags = d->flags & ~(uint32_t)KFP16_EPI_SPLITK;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.groups = two ? 2 * groups : groups; p.kslabs = kslabs; p.kslab_len = kslab_len;
  p.tile_rows = kBM;
  if (two) {
    if (!(d->split_k > 1) || kslabs != 1 || d->a_major != KFP16_MN_MAJOR || d->b_major != KFP16_MN_MAJOR || !d->B2.ptr || conv) {
      set_error("kfp16_gemm_ex: a second problem (A2/B2) needs split_k > 1, one K slab and MN-major operands"); return -1;
    }
    p.groups2_from = groups;
  }

  // ---- implicit-GEMM convolution: geometry checks and the tile shape it imposes
  int conv_tbox = 0, conv_krows = 0;
  if (conv) {
    const kfp16_conv_addr& c = d->conv;
    if (!c.x || c.ntaps < 1 || c.ntaps > kMaxTaps || groups != 1 || two || c.rows_h < 1 || c.rows_h > c.H) {
      set_error("kfp16_gemm_ex: convolution needs x, 1..%d taps, one group, 1 <= rows_h <= H", kMaxTaps); return -1;
    }
    for (int t = 0; t < c.ntaps; ++t)
      if (c.dt[t] < -127 || c.dt[t] > 127 || c.hq[t] < -127 || c.hq[t] > 127 || c.par[t] < 0 || c.par[t] >= c.P) {
        set_error("kfp16_gemm_ex: convolution tap %d out of range (dt %d hq %d par %d)", t, c.dt[t], c.hq[t], c.par[t]); return -1;
      }
    if (conv == 1) {
      if (a_mn || d->M != c.T * c.rows_h || d->K != c.ntaps * c.C || d->split_k > 1 || c.rows_h > kBM) {
        set_error("kfp16_gemm_ex: conv.mode 1 needs a K-major A, M = T*rows_h, K = ntaps*C, rows_h <= 128, no split-K"); return -1;
      }
      conv_tbox = kBM / c.rows_h;
      p.tile_rows = conv_tbox * c.rows_h;
    } else {
      if (!a_mn || !b_mn || d->M != c.ntaps * c.C || d->K != c.T * c.rows_h || !(d->split_k >= 1)) {
        set_error("kfp16_gemm_ex: conv.mode 2 needs MN-major operands, M = ntaps*C, K = T*rows_h"); return -1;
      }
      for (int tb = 80 / c.rows_h; tb >= 1; --tb)
        if ((tb * c.rows_h) % 16 == 0) { conv_tbox = tb; break; }
      if (!conv_tbox) { set_error("kfp16_gemm_ex: conv.mode 2: no k-block of <= 80 rows that is a multiple of 16 for %d heights", c.rows_h); return -1; }
      conv_krows = conv_tbox * c.rows_h;
    }
    p.conv = conv; p.conv_h = c.rows_h; p.conv_c = c.C; p.conv_taps = c.ntaps; p.conv_tbox = conv_tbox; p.conv_k16 = conv_krows / 16;
    for (int t = 0; t < c.ntaps; ++t) {
      p.conv_dt[t] = (int8_t)c.dt[t]; p.conv_hq[t] = (int8_t)c.hq[t]; p.conv_par[t] = (int8_t)c.par[t]; p.conv_brow[t] = c.brow[t];
    }
  }

  if (d->split_k > 1 || conv == 2) flags |= EPI_SPLITK;   // caller asked for fp32 accumulation into ws
  const int ek = d->force_generic ? ((flags & EPI_SPLITK) ? EK_SPLITK : EK_GENERIC) : pick_kind(flags);
  const kfp16_mat& A = d->A; const kfp16_mat& B = d->B;
  if ((!conv && !A.ptr) || !B.ptr) { set_error("kfp16_gemm_ex: null operand"); return -1; }

  // ---- kernel shape: CTA pairs (cta_group::2, 256-row tiles: each CTA stages half of B) whenever the
  // problem has at least two row tiles, and ONE shared A tile for both splice slabs when the slabs
  // are row-shifted views of the same columns (time splicing).  force_cg overrides (tests).
  p.dbg = (long long*)d->debug_clock_buf;
  int cg = d->force_cg ? d->force_cg : 2;
  if (cg != 1 && cg != 2) { set_error("kfp16_gemm_ex: force_cg must be 0, 1 or 2"); return -1; }
  if (d->M <= p.tile_rows) cg = 1;
  bool share = false;
  int span = 0, min_off = 0;
  if (!conv && cg == 2 && kslabs == 2 && groups == 1 && !a_mn && !(flags & EPI_SPLITK) && d->no_share != 1 &&
      d->a_col_off[0][0] == d->a_col_off[0][1]) {
    min_off = d->a_row_off[0][0] < d->a_row_off[0][1] ? d->a_row_off[0][0] : d->a_row_off[0][1];
    span = d->a_row_off[0][0] + d->a_row_off[0][1] - 2 * min_off;
    share = span <= 8 && gemm_variant_exists(a_mn, b_mn, ek, 2, true);
  }
  if (cg == 2 && !share && !gemm_variant_exists(a_mn, b_mn, ek, 2, false)) cg = 1;
  // merged groups: the two groups of a spliced weight gradient (dW_g = sum_k A[k + a_off_g]^T B[k + b_off_g]) differ
  // only by row shifts of <= 8 -> ONE A and ONE B tile of 64+span k-rows per k-block, read through row-shifted
  // descriptors into the two TMEM accumulator stages.  It pays in the grouped launch (kfp16_wgrad_group_*: one partial
  // per element); for a single layer's gradient the doubled fp32 reductions per unit cost more than the loads saved, so
  // here it is only taken on request (no_share == 4: the kernel's stand-alone test).
  bool merge = false;
  int ma_min = 0, mb_min = 0, ma_span = 0, mb_span = 0;
  if (!conv && !two && cg == 2 && groups == 2 && kslabs == 1 && a_mn && b_mn && (flags & EPI_SPLITK) && d->no_share == 4 &&
      d->N > 128 && d->N <= 160 && !d->force_bn &&
      d->a_col_off[0][0] == d->a_col_off[1][0] && d->b_col_off[0][0] == d->b_col_off[1][0]) {
    ma_min = d->a_row_off[0][0] < d->a_row_off[1][0] ? d->a_row_off[0][0] : d->a_row_off[1][0];
    mb_min = d->b_row_off[0][0] < d->b_row_off[1][0] ? d->b_row_off[0][0] : d->b_row_off[1][0];
    ma_span = d->a_row_off[0][0] + d->a_row_off[1][0] - 2 * ma_min;
    mb_span = d->b_row_off[0][0] + d->b_row_off[1][0] - 2 * mb_min;
    merge = ma_span <= 8 && mb_span <= 8;
  }
  const int tile_groups = merge ? 1 : (two ? 2 * groups : groups);     // groups that multiply the tile count

  // ---- convolution forward / input gradient: ONE input box per 64-channel chunk shared by all taps (MODE 5) when the
  // geometry and shared memory allow; otherwise every tap loads its own box (MODE 0 with conv addressing)
  bool cshare = false;
  int cs_sl = 0, cs_tbox = 0, cs_dtmin = 0, cs_hqmin = 0, cs_dtspan = 0, cs_planes = 0, cs_plane_par[2] = {0, 0}, cs_plane_bytes = 0, cs_nstages = 0;
  if (conv == 1 && d->conv.ntaps >= 2 && d->no_share != 1) {
    const kfp16_conv_addr& c = d->conv;
    int dtmax = c.dt[0], hqmax = c.hq[0];
    cs_dtmin = c.dt[0]; cs_hqmin = c.hq[0];
    bool planes_ok = true;
    for (int t = 0; t < c.ntaps; ++t) {
      cs_dtmin = std::min(cs_dtmin, c.dt[t]); dtmax = std::max(dtmax, c.dt[t]);
      cs_hqmin = std::min(cs_hqmin, c.hq[t]); hqmax = std::max(hqmax, c.hq[t]);
      int pl = -1;
      for (int q = 0; q < cs_planes; ++q) if (cs_plane_par[q] == c.par[t]) pl = q;
      if (pl < 0) { if (cs_planes < 2) cs_plane_par[cs_planes++] = c.par[t]; else planes_ok = false; }
    }
    cs_dtspan = dtmax - cs_dtmin;
    const int hspan = hqmax - cs_hqmin;
    cs_sl = c.rows_h + hspan;
    cs_tbox = cs_sl <= kBM ? kBM / cs_sl : 0;
    const bool kind_ok = b_mn ? (ek == EK_AFFINE || ek == EK_GENERIC) : (ek == EK_PLAIN || ek == EK_BN_GRADMASK);
    // Measured on B200 (profiles/r02_trace_cnn_*): the shared box pays where the per-tap A boxes bound the tile (N <= 128: 61 vs
    // 79 us cnn2, 73 vs 81 us cnn4, input gradients 55 vs 75 / 66 vs 79 us); 256-wide tiles are MMA-bound and lose the rows the
    // (frame, slot) index wastes on the height padding (cnn6: 88 vs 85 us, its input gradient 86 vs 79 us)
    // (Keeping the weight tiles of a narrow layer resident in shared memory instead of streaming them per M tile was measured
    //  too: cnn2 60 -> 62 us, cnn3 39 -> 44 us -- those kernels are bound by the per-tile epilogue / accumulator hand-over of
    //  their 120 x 64 tiles, not by the 73 KB of weights per tile -- and the extra branches in the issue loop cost the wider
    //  layers 10 %.  Not kept.  Four input-box buffers instead of two: no change either.  The per-role clock stamps
    //  (scripts/conv_tile_profile.py) show the MMA phase itself taking ~4.5k cycles per 240 x 64 pair tile: with N = 64 every
    //  UMMA re-reads 4 KB of A from shared memory for 64 output columns, so the tensor core is bound by its operand reads,
    //  not by its math -- inherent to 64-filter layers.)
    const bool narrow = d->N <= 128 || d->no_share == 5;
    if (planes_ok && narrow && cs_tbox >= 1 && kind_ok && cs_tbox + cs_dtspan <= 256 && cs_sl <= 256) {
      const int rows_alloc = cs_dtspan * cs_sl + hspan + kBM;           // furthest row a shifted 128-row read touches
      cs_plane_bytes = (rows_alloc * 128 + 1023) & ~1023;
      cshare = true;
    }
  }
  if (cshare) {
    int cg_s = cg;
    if (d->M <= cs_tbox * d->conv.rows_h || ek == EK_GENERIC) cg_s = 1;
    // shared-memory budget with the widest tile this N can get: two box buffers per plane + >= 3 weight-tile stages
    const int bn_w = d->force_bn ? d->force_bn : (d->N <= 64 ? 64 : d->N <= 128 ? 128 : 256);
    const int b_tile = b_mn ? ((bn_w / cg_s + 63) / 64) * 8192 : (bn_w / cg_s) * 128;
    const int fixed = 1024 + 512 + epi_bytes_for(ek, bn_w);
    if ((227 * 1024 - fixed - 2 * cs_planes * cs_plane_bytes) / b_tile < 3) cshare = false;
    else { p.tile_rows = cs_tbox * d->conv.rows_h; cg = cg_s; }
  }

  const int tile_m = p.tile_rows * cg;
  const int m_tiles = (d->M + tile_m - 1) / tile_m;
  const int kb_total = conv == 2 ? (d->conv.T + conv_tbox - 1) / conv_tbox
                                 : (conv == 1 ? d->conv.ntaps : (share ? 1 : kslabs)) * ((kslab_len + kBK - 1) / kBK);
  int split_k = d->split_k > 1 ? d->split_k : 1;
  if (merge) split_k *= 2;     // same number of work items as the two separate groups had
  if (split_k > kb_total) split_k = kb_total;
  if (split_k > 1) {   // make every split non-empty
    const int per = (kb_total + split_k - 1) / split_k;
    split_k = (kb_total + per - 1) / per;
  }
  p.split_k = split_k;
  if (conv == 1) p.kslabs = d->conv.ntaps;     // the kernel walks taps as K slabs (offsets come from the conv arrays)

  int ctas = ctx->num_sms;
  if (ctx->max_ctas > 0 && ctx->max_ctas < ctas) ctas = ctx->max_ctas;
  int units = ctas / cg;                 // CTAs or CTA pairs that can be resident
  if (units < 1) { units = 1; }
  int bn = merge ? 160 : (d->force_bn ? d->force_bn : pick_bn(d->N, d->K, m_tiles, tile_groups, split_k, units));
  if (bn != 64 && bn != 128 && bn != 160 && bn != 256) { set_error("kfp16_gemm_ex: unsupported tile width %d", bn); return -1; }
  if (conv == 2 && bn == 160) bn = 256;
  if (cshare) {
    if (bn == 160) bn = 256;
    const int b_tile = b_mn ? ((bn / cg + 63) / 64) * 8192 : (bn / cg) * 128;
    const int fixed = 1024 + 512 + epi_bytes_for(ek, bn);
    cs_nstages = (227 * 1024 - fixed - 2 * cs_planes * cs_plane_bytes) / b_tile;
    if (cs_nstages > 8) cs_nstages = 8;
    if (cs_nstages < 3) { set_error("internal: shared convolution box does not fit shared memory (bn %d)", bn); return -1; }   // excluded above
  }
  if (!(flags & EPI_SPLITK) && (bn % 64) != 0 && d->N > bn) {
    // the last 64-wide store chunk of a 160-wide tile would spill into the next tile
    set_error("kfp16_gemm_ex: tile width %d needs N <= %d unless split-K", bn, bn); return -1;
  }

  // operand maps (halo rows are part of the mapped tensor so spliced reads can address them)
  const __half* b_base = (const __half*)B.ptr - (long long)B.halo * B.ld;
  const int b_chunks = (bn / cg + 63) / 64;
  if (conv) {
    const kfp16_conv_addr& c = d->conv;
    if (!make_map_conv(&p.tmA, c.x, c.C, c.P, c.H, c.T, c.rows_h, conv_tbox, "A (convolution)")) return -1;
    if (cshare) {
      // the box covers every frame and height any tap of the tile reads: (tbox + time span) frames x conv_sl heights
      if (!make_map_conv(&p.tmA, c.x, c.C, c.P, c.H, c.T, cs_sl, cs_tbox + cs_dtspan, "A (convolution, shared box)")) return -1;
      if (!make_map_2d(&p.tmB, b_base, B.cols, (long long)B.rows + 2 * B.halo, B.ld, 64, b_mn ? 64 : bn / cg, "B")) return -1;
      p.conv_sl = cs_sl; p.conv_tbox = cs_tbox;
      p.conv_box_t0 = cs_dtmin; p.conv_box_h0 = cs_hqmin;
      p.conv_box_bytes = (cs_tbox + cs_dtspan) * cs_sl * 128;
      p.conv_plane_bytes = cs_plane_bytes; p.conv_planes = cs_planes;
      p.conv_plane_par[0] = cs_plane_par[0]; p.conv_plane_par[1] = cs_plane_par[1];
      p.conv_nstages = cs_nstages;
      for (int t = 0; t < c.ntaps; ++t) {
        const int pl = (cs_planes == 2 && c.par[t] == cs_plane_par[1]) ? 1 : 0;
        p.conv_aoff[t] = (uint32_t)(pl * (cs_plane_bytes >> 4) + ((c.dt[t] - cs_dtmin) * cs_sl + (c.hq[t] - cs_hqmin)) * 8);
      }
    } else if (conv == 1) {
      if (!make_map_2d(&p.tmB, b_base, B.cols, (long long)B.rows + 2 * B.halo, B.ld, 64, b_mn ? 64 : bn / cg, "B")) return -1;
      p.conv_tx = p.tile_rows * 128 + (b_mn ? b_chunks * 8192 : (bn / cg) * 128);
    } else {
      if (!make_map_2d(&p.tmB, b_base, B.cols, (long long)B.rows + 2 * B.halo, B.ld, 64, conv_krows, "B")) return -1;
      p.conv_tx = (2 + b_chunks) * conv_krows * 128;
    }
  } else {
    const __half* a_base = (const __half*)A.ptr - (long long)A.halo * A.ld;
    const int a_box_rows = a_mn ? (merge ? 64 + ma_span : 64) : (share ? kBM + span : kBM);
    if (!make_map_2d(&p.tmA, a_base, A.cols, (long long)A.rows + 2 * A.halo, A.ld, 64, a_box_rows, "A")) return -1;
    if (!make_map_2d(&p.tmB, b_base, B.cols, (long long)B.rows + 2 * B.halo, B.ld, 64, b_mn ? (merge ? 64 + mb_span : 64) : bn / cg, "B")) return -1;
    for (int g = 0; g < groups; ++g)
      for (int s = 0; s < kslabs; ++s) {
        p.a_row_off[g][s] = d->a_row_off[g][s] + A.halo; p.a_col_off[g][s] = d->a_col_off[g][s];
        p.b_row_off[g][s] = d->b_row_off[g][s] + B.halo; p.b_col_off[g][s] = d->b_col_off[g][s];
      }
    if (share) p.a_box_bytes = a_box_rows * kBK * 2;
  }
  if (two) {
    const kfp16_mat& A2 = d->A2; const kfp16_mat& B2 = d->B2;
    const __half* a2_base = (const __half*)A2.ptr - (long long)A2.halo * A2.ld;
    const __half* b2_base = (const __half*)B2.ptr - (long long)B2.halo * B2.ld;
    if (!make_map_2d(&p.tmA2, a2_base, A2.cols, (long long)A2.rows + 2 * A2.halo, A2.ld, 64, 64, "A2")) return -1;
    if (!make_map_2d(&p.tmB2, b2_base, B2.cols, (long long)B2.rows + 2 * B2.halo, B2.ld, 64, 64, "B2")) return -1;
    for (int g = 0; g < groups; ++g) {
      p.a_row_off[groups + g][0] = d->a2_row_off[g] + A2.halo;
      p.b_row_off[groups + g][0] = d->b2_row_off[g] + B2.halo;
    }
    p.ws_ld2 = d->ws2_ld;
    p.ws_transposed2 = d->ws2_transposed;
  }
  if (merge) {
    p.groups = 1;
    p.a_shift[0] = d->a_row_off[0][0] - ma_min; p.a_shift[1] = d->a_row_off[1][0] - ma_min;
    p.b_shift[0] = d->b_row_off[0][0] - mb_min; p.b_shift[1] = d->b_row_off[1][0] - mb_min;
    p.a_row_off[0][0] = ma_min + A.halo;
    p.b_row_off[0][0] = mb_min + B.halo;
    p.merge_tx = 2 * (64 + ma_span) * 128 + b_chunks * (64 + mb_span) * 128;
  }
  if (share) {
    p.a_shift[0] = d->a_row_off[0][0] - min_off;
    p.a_shift[1] = d->a_row_off[0][1] - min_off;
    p.a_row_off[0][0] = min_off + A.halo;       // the one A box starts at the earlier slab's row
  }

  p.flags = flags;
  p.alpha = d->alpha; p.beta = d->beta; p.res_scale = d->res_scale;
  p.bias = (const __half*)d->bias; p.bn_scale = d->bn_scale; p.bn_shift = d->bn_shift;
  p.vec_gstride = d->vec_gstride;
  p.mask_out = d->mask_out; p.mask_in = d->mask_in; p.mask_ld = d->mask_ld;
  p.ws_ld = d->ws_ld;
  p.ws_transposed = d->ws_transposed;
  p.drop_p = d->drop_p; p.drop_seed = d->drop_seed; p.drop_seed_dev = d->drop_seed_dev;
  if (d->zero_row_period > 0) {
    if (d->zero_row_lo < 0 || d->zero_row_hi < d->zero_row_lo || d->zero_row_hi > d->zero_row_period) { set_error("kfp16_gemm_ex: zero_row window must satisfy 0 <= lo <= hi <= period"); return -1; }
    p.zero_period = (uint32_t)d->zero_row_period; p.zero_lo = (uint32_t)d->zero_row_lo; p.zero_len = (uint32_t)(d->zero_row_hi - d->zero_row_lo);
  }
  if ((flags & EPI_DROPOUT) && !(d->drop_p >= 0.0f && d->drop_p < 1.0f)) { set_error("kfp16_gemm_ex: dropout probability must be in [0, 1)"); return -1; }
  if ((flags & EPI_BIAS) && !p.bias) { set_error("kfp16_gemm_ex: EPI_BIAS without bias"); return -1; }
  if ((flags & EPI_BN) && (!p.bn_scale || !p.bn_shift)) { set_error("kfp16_gemm_ex: EPI_BN without scale/shift"); return -1; }
  if ((flags & EPI_MASK) && !p.mask_out) { set_error("kfp16_gemm_ex: EPI_MASK without mask_out"); return -1; }
  if ((flags & EPI_GRADMASK) && !p.mask_in) { set_error("kfp16_gemm_ex: EPI_GRADMASK without mask_in"); return -1; }
  if (conv && (flags & (EPI_RESID | EPI_BETA))) { set_error("kfp16_gemm_ex: residual / beta epilogues are not available on convolution tiles"); return -1; }

  if (two) {
    for (int g = 0; g < groups; ++g) {
      if (d->ws2_transposed ? (!d->ws2[g] || d->ws2_ld < d->M) : (!d->ws2[g] || d->ws2_ld < d->N || (d->ws2_ld % 4) || !aligned16(d->ws2[g]))) {
        set_error("kfp16_gemm_ex: the second problem needs a 16B-aligned fp32 workspace with ws2_ld >= N, ws2_ld %% 4 == 0"); return -1;
      }
      p.ws[groups + g] = d->ws2[g];
    }
  }
  for (int g = 0; g < groups; ++g) {
    if (flags & EPI_SPLITK) {
      if (d->ws_transposed ? (!d->ws[g] || d->ws_ld < d->M) : (!d->ws[g] || d->ws_ld < d->N || (d->ws_ld % 4) || !aligned16(d->ws[g]))) {
        set_error("kfp16_gemm_ex: split-K needs a 16B-aligned fp32 workspace with ws_ld >= N, ws_ld %% 4 == 0"); return -1;
      }
      p.ws[g] = d->ws[g];
    } else {
      if (!d->D[g]) { set_error("kfp16_gemm_ex: null output"); return -1; }
      if (!make_map_2d(&p.tmD[g], d->D[g], d->N, d->M, d->ldd, 64, p.tile_rows, "D")) return -1;
      if (flags & (EPI_RESID | EPI_BETA)) {
        if (!d->R[g]) { set_error("kfp16_gemm_ex: residual / beta requested without R"); return -1; }
        if (!make_map_2d(&p.tmR[g], d->R[g], d->N, d->M, d->ldr, 64, kBM, "R")) return -1;
      }
    }
  }

  const long long tiles = (long long)m_tiles * ((d->N + bn - 1) / bn) * tile_groups * split_k;
  const int grid = cg * (int)(tiles < units ? tiles : units);
  if (!check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ctx->profile) {
    if (!check_cuda(cudaEventCreate(&ev0), "cudaEventCreate") || !check_cuda(cudaEventCreate(&ev1), "cudaEventCreate")) return -1;
    cudaEventRecord(ev0, ctx->stream);
  }
  bool ok = false;
  GemmLaunch L;
  L.bn = bn; L.a_mn = a_mn; L.b_mn = b_mn; L.ek = ek; L.cg = cg; L.share = cshare ? 5 : (conv == 2 ? 4 : (merge ? 3 : (share ? 1 : 0))); L.grid = grid;
  switch (bn) {
    case 64: ok = launch_gemm_bn<64>(ctx, p, L); break;
    case 128: ok = launch_gemm_bn<128>(ctx, p, L); break;
    case 160: ok = launch_gemm_bn<160>(ctx, p, L); break;
    case 256: ok = launch_gemm_bn<256>(ctx, p, L); break;
  }
  if (ctx->profile) {
    cudaEventRecord(ev1, ctx->stream);
    ctx->prof_ev.push_back(ev0);
    ctx->prof_ev.push_back(ev1);
    ctx->prof_flops.push_back(2.0 * d->M * d->N * d->K * groups * (two ? 2 : 1));
    char desc[192];
    snprintf(desc, sizeof(desc), "M=%d N=%d K=%d g=%d slabs=%d split=%d bn=%d A%s B%s flags=0x%x ek=%d cg=%d mode=%d conv=%d grid=%d", d->M, d->N, d->K,
             groups, kslabs, split_k, bn, a_mn ? "mn" : "k", b_mn ? "mn" : "k", flags, ek, cg, L.share, conv, grid);
    ctx->prof_desc.push_back(desc);
  }
  return ok ? 0 : -1;
}

// ------------------------------------------------------------------ grouped weight gradients
struct kfp16_wgrad_group {
  int M, N, K, count, split_k, bn, grid;
  bool merged = false;      // MODE 3 kernels: one tile pass per problem serves both row-shifted groups
  GroupProb* dev = nullptr;
};

kfp16_wgrad_group* kfp16_wgrad_group_create(kfp16_ctx* ctx, int M, int N, int K, const kfp16_wgrad_prob* probs, int count) {
  if (!ctx || !probs || count < 1 || M <= 0 || N <= 0 || K <= 0 || (N % 8)) { set_error("kfp16_wgrad_group_create: bad argument"); return nullptr; }
  if (M <= kBM) { set_error("kfp16_wgrad_group_create: needs M > 128 (CTA-pair tiles)"); return nullptr; }
  // merged groups (one A and one B tile of 64 + span k-rows per k-block for both groups of a problem): in a grouped
  // launch the main loop is bound by the L2->SM fabric (~6.3 KB/cycle), which the merge nearly halves; it needs the
  // 160-wide kernel and row shifts of at most 8.
  bool merged = N > 128 && N <= 160;
  for (int i = 0; i < count && merged; ++i) {
    const int sa = abs(probs[i].a_row_off[0] - probs[i].a_row_off[1]), sb = abs(probs[i].b_row_off[0] - probs[i].b_row_off[1]);
    if (sa > 8 || sb > 8) merged = false;
  }
  std::vector<GroupProb> host(count);
  for (int i = 0; i < count; ++i) {
    const kfp16_wgrad_prob& q = probs[i];
    GroupProb& g = host[i];
    memset(&g, 0, sizeof(g));
    if (!q.A.ptr || !q.B.ptr) { set_error("kfp16_wgrad_group_create: null operand in problem %d", i); return nullptr; }
    const __half* a_base = (const __half*)q.A.ptr - (long long)q.A.halo * q.A.ld;
    const __half* b_base = (const __half*)q.B.ptr - (long long)q.B.halo * q.B.ld;
    const int a_min = std::min(q.a_row_off[0], q.a_row_off[1]), b_min = std::min(q.b_row_off[0], q.b_row_off[1]);
    const int a_span = merged ? abs(q.a_row_off[0] - q.a_row_off[1]) : 0, b_span = merged ? abs(q.b_row_off[0] - q.b_row_off[1]) : 0;
    if (!make_map_2d(&g.tmA, a_base, q.A.cols, (long long)q.A.rows + 2 * q.A.halo, q.A.ld, 64, 64 + a_span, "A (grouped)")) return nullptr;
    if (!make_map_2d(&g.tmB, b_base, q.B.cols, (long long)q.B.rows + 2 * q.B.halo, q.B.ld, 64, 64 + b_span, "B (grouped)")) return nullptr;
    if (merged) {
      const int b_chunks = (160 / 2 + 63) / 64;
      g.merge_tx = 2 * (64 + a_span) * 128 + b_chunks * (64 + b_span) * 128;
    }
    for (int k = 0; k < 2; ++k) {
      g.a_shift[k] = q.a_row_off[k] - a_min;
      g.b_shift[k] = q.b_row_off[k] - b_min;
      g.a_row_off[k] = (merged ? a_min : q.a_row_off[k]) + q.A.halo;
      g.b_row_off[k] = (merged ? b_min : q.b_row_off[k]) + q.B.halo;
      if (q.ws_transposed ? (!q.ws[k] || q.ws_ld < M) : (!q.ws[k] || q.ws_ld < N || (q.ws_ld % 4) || !aligned16(q.ws[k]))) {
        set_error("kfp16_wgrad_group_create: problem %d needs 16B-aligned fp32 targets with ws_ld >= N, ws_ld %% 4 == 0", i); return nullptr;
      }
      g.ws[k] = q.ws[k];
    }
    g.ws_ld = q.ws_ld; g.ws_transposed = q.ws_transposed;
  }
  kfp16_wgrad_group* grp = new kfp16_wgrad_group();
  grp->M = M; grp->N = N; grp->K = K; grp->count = count;
  grp->bn = N <= 64 ? 64 : N <= 128 ? 128 : N <= 160 ? 160 : 256;
  int sms = ctx->num_sms;
  if (ctx->max_ctas > 0 && ctx->max_ctas < sms) sms = ctx->max_ctas;
  const int units = std::max(1, sms / 2);
  grp->merged = merged;
  const long long tiles = (long long)((M + 255) / 256) * ((N + grp->bn - 1) / grp->bn) * (merged ? 1 : 2) * count;
  const int kb = (K + kBK - 1) / kBK;
  // split so that the rounds of work items waste little of the last round; every extra split costs another fp32
  // reduction pass over the tiles (L2 atomics at ~3.5 TB/s), weighed as 8 k-blocks
  int best = 1; double best_cost = 1e30;
  for (int sp = 1; sp <= 8 && sp <= kb / 4; ++sp) {
    const long long rounds = (tiles * sp + units - 1) / units;
    // (merged tiles: the epilogue covers both accumulator stages and does not overlap the next tile's main loop --
    // measured on 32 problems of 1536 x 160 x 9984: split 1 / 2 / 3 / 4 = 271 / 298 / 290 / 321 us)
    const double cost = (double)rounds * ((kb + sp - 1) / sp + (merged ? 24.0 : 8.0));
    if (cost < best_cost * 0.99) { best_cost = cost; best = sp; }
  }
  grp->split_k = best < 2 && tiles < units ? 2 : best;
  { const int per = (kb + grp->split_k - 1) / grp->split_k; grp->split_k = (kb + per - 1) / per; }
  const long long items = tiles * grp->split_k;
  grp->grid = 2 * (int)(items < units ? items : units);
  if (!check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice") ||
      !check_cuda(cudaMalloc((void**)&grp->dev, sizeof(GroupProb) * count), "cudaMalloc (grouped problem table)") ||
      !check_cuda(cudaMemcpy(grp->dev, host.data(), sizeof(GroupProb) * count, cudaMemcpyHostToDevice), "grouped problem table upload")) {
    if (grp->dev) cudaFree(grp->dev);
    delete grp;
    return nullptr;
  }
  return grp;
}
int kfp16_wgrad_group_launch(kfp16_ctx* ctx, kfp16_wgrad_group* grp) {
  if (!ctx || !grp) { set_error("kfp16_wgrad_group_launch: null argument"); return -1; }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.probs = grp->dev;
  p.M = grp->M; p.N = grp->N; p.K = grp->K;
  p.groups = (grp->merged ? 1 : 2) * grp->count; p.kslabs = 1; p.kslab_len = grp->K;
  p.split_k = grp->split_k;
  p.flags = EPI_SPLITK;
  p.alpha = 1.0f;
  p.tile_rows = kBM;
  if (!check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ctx->profile) {
    if (!check_cuda(cudaEventCreate(&ev0), "cudaEventCreate") || !check_cuda(cudaEventCreate(&ev1), "cudaEventCreate")) return -1;
    cudaEventRecord(ev0, ctx->stream);
  }
  GemmLaunch L;
  L.bn = grp->bn; L.a_mn = true; L.b_mn = true; L.ek = EK_SPLITK; L.cg = 2; L.share = grp->merged ? 3 : 0; L.grid = grp->grid;
  bool ok = false;
  switch (grp->bn) {
    case 64: ok = launch_gemm_bn<64>(ctx, p, L); break;
    case 128: ok = launch_gemm_bn<128>(ctx, p, L); break;
    case 160: ok = launch_gemm_bn<160>(ctx, p, L); break;
    case 256: ok = launch_gemm_bn<256>(ctx, p, L); break;
  }
  if (ctx->profile) {
    cudaEventRecord(ev1, ctx->stream);
    ctx->prof_ev.push_back(ev0);
    ctx->prof_ev.push_back(ev1);
    ctx->prof_flops.push_back(2.0 * grp->M * grp->N * grp->K * 2 * grp->count);
    char desc[160];
    snprintf(desc, sizeof(desc), "grouped wgrad x%d M=%d N=%d K=%d split=%d bn=%d merged=%d grid=%d", grp->count, grp->M, grp->N, grp->K, grp->split_k, grp->bn, (int)grp->merged, grp->grid);
    ctx->prof_desc.push_back(desc);
  }
  return ok ? 0 : -1;
}
void kfp16_wgrad_group_destroy(kfp16_wgrad_group* grp) {
  if (!grp) return;
  if (grp->dev) cudaFree(grp->dev);
  delete grp;
}

int kfp16_gemm(kfp16_ctx* ctx, int M, int N, int K, float alpha, const void* A, int transA,
               const void* B, int transB, float beta, void* C) {
  if (!ctx) { set_error("kfp16_gemm: null context (create one with ops_cublas_create / kfp16_ctx_create)"); return -1; }
  if (M <= 0 || N <= 0) return 0;
  if (!A || !B || !C) { set_error("kfp16_gemm: null pointer"); return -1; }
  const int lda = transA ? M : K, ldb = transB ? K : N;
  const bool tma_ok = K > 0 && (N % 8) == 0 && (K % 8) == 0 && (lda % 8) == 0 && (ldb % 8) == 0 &&
                      aligned16(A) && aligned16(B) && aligned16(C);
  if (tma_ok) {
    kfp16_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = M; d.N = N; d.K = K;
    // long-reduction shapes with few output tiles (weight gradients through the plain API): split K over the
    // idle SMs into the context's fp32 workspace, then round once to fp16
    const int bn_est = N <= 64 ? 64 : N <= 128 ? 128 : N <= 160 ? 160 : 256;
    const long long tiles = (long long)((M + 255) / 256) * ((N + bn_est - 1) / bn_est);
    const int pairs = ctx->num_sms / 2;
    int split = (int)(pairs / (tiles > 0 ? tiles : 1));
    if (split > K / 256) split = K / 256;
    if (beta == 0.0f && split >= 2 && (N % 4) == 0) {
      const size_t need = (size_t)M * N * sizeof(float);
      if (ctx->ws_bytes < need) {
        if (ctx->ws) cudaFree(ctx->ws);
        ctx->ws = nullptr; ctx->ws_bytes = 0;
        if (!check_cuda(cudaMalloc(&ctx->ws, need), "cudaMalloc (split-K workspace)")) return -1;
        ctx->ws_bytes = need;
      }
      if (!check_cuda(cudaMemsetAsync(ctx->ws, 0, need, ctx->stream), "split-K workspace memset")) return -1;
      d.a_major = transA ? KFP16_MN_MAJOR : KFP16_K_MAJOR;
      d.b_major = transB ? KFP16_K_MAJOR : KFP16_MN_MAJOR;
      d.A.ptr = A; d.A.rows = transA ? K : M; d.A.cols = transA ? M : K; d.A.ld = lda;
      d.B.ptr = B; d.B.rows = transB ? N : K; d.B.cols = transB ? K : N; d.B.ld = ldb;
      d.groups = 1; d.kslabs = 1; d.kslab_len = K;
      d.alpha = alpha;
      d.split_k = split; d.ws[0] = ctx->ws; d.ws_ld = N;
      if (kfp16_gemm_ex(ctx, &d) != 0) return -1;
      return kfp16_f32_to_f16(ctx, ctx->ws, C, (size_t)M * N);
    }
    d.a_major = transA ? KFP16_MN_MAJOR : KFP16_K_MAJOR;
    d.b_major = transB ? KFP16_K_MAJOR : KFP16_MN_MAJOR;
    d.A.ptr = A; d.A.rows = transA ? K : M; d.A.cols = transA ? M : K; d.A.ld = lda;
    d.B.ptr = B; d.B.rows = transB ? N : K; d.B.cols = transB ? K : N; d.B.ld = ldb;
    d.groups = 1; d.kslabs = 1; d.kslab_len = K;
    d.D[0] = C; d.ldd = N; d.R[0] = C; d.ldr = N;
    d.alpha = alpha; d.beta = beta;
    d.flags = (beta != 0.0f) ? KFP16_EPI_BETA : 0;
    return kfp16_gemm_ex(ctx, &d);
  }
  if (!check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
  dim3 block(16, 16), grid((N + 15) / 16, (M + 15) / 16);
  gemm_simt_fallback<<<grid, block, 0, ctx->stream>>>(
      M, N, K, alpha, (const __half*)A, transA ? 1 : K, transA ? M : 1, (const __half*)B,
      transB ? 1 : N, transB ? K : 1, beta, (__half*)C, N);
  count_launch();
  return check_launch("gemm_simt_fallback") ? 0 : -1;
}

// ------------------------------------------------------------------ reference surface (ops.h)
void* ops_cublas_create(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return (void*)kfp16_ctx_create(dev);
}
void ops_cublas_destroy(void* handle) { kfp16_ctx_destroy((kfp16_ctx*)handle); }

// C = alpha*A*B + beta*C, row-major, dense: like the reference the lda/ldb/ldc arguments are
// accepted but the leading dimensions are K, N, N (ops.cu:381-392 hard-wires them).
int ops_gemm(void* handle, int M, int N, int K, float alpha, const void* A, int lda, const void* B,
             int ldb, float beta, void* C, int ldc) {
  (void)lda; (void)ldb; (void)ldc;
  if (!handle) { set_error("ops_gemm: null handle"); return -1; }
  if (kfp16_gemm((kfp16_ctx*)handle, M, N, K, alpha, A, 0, B, 0, beta, C) != 0) {
    char tmp[400];
    snprintf(tmp, sizeof(tmp), "%s", get_error() ? get_error() : "?");
    set_error("ops_gemm failed: %s (M=%d N=%d K=%d)", tmp, M, N, K);
    return -1;
  }
  return 0;
}

int ops_gemm_strided(void* handle, int M, int N, int K, float alpha, const void* A, int lda,
                     int64_t strideA, const void* B, int ldb, int64_t strideB, float beta, void* C,
                     int ldc, int64_t strideC, int batch_count) {
  for (int b = 0; b < batch_count; ++b) {
    if (ops_gemm(handle, M, N, K, alpha, (const __half*)A + b * strideA, lda,
                 (const __half*)B + b * strideB, ldb, beta, (__half*)C + b * strideC, ldc) != 0)
      return -1;
  }
  return 0;
}

const char* ops_last_error(void) { return get_error(); }
void ops_clear_error(void) { clear_error(); }

}  // extern "C"
