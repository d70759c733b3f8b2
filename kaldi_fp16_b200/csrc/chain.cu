// Chain LF-MMI objective, batched over the sequences of a minibatch (include/kaldi_fp16_chain.h).
//
// Replaces the reference's per-sequence chain_compute_loss (/root/reference/cpp/cuda/chain.cu:80-352,475: one kernel launch
// per arc set per frame, a binary search per arc, atomic log-adds, cudaMalloc / cudaMemcpy / host sync per sequence) and
// the Go loop around it (internal/nnet/chain_loss.go:221-294: ops_subsample_rows + an FST upload per sequence).  Same
// arithmetic: log-semiring forward-backward over the numerator and the denominator FST,
//   loss = -(num_logprob - den_logprob),  grad = clamp((den_post - num_post) * weight, +-30) as FP16.
//
// ONE launch for the whole minibatch: a CTA per sequence walks the frames itself (block barriers between frames, no host
// involvement; numerator and denominator advance in the same frame loop), states are spread over the threads, each state reduces its incoming (forward) / outgoing (backward) arcs in a fixed order -- no atomics on
// alpha / beta, deterministic -- and the posteriors of a frame are accumulated in shared memory during the backward step
// of that frame, so the gradient row is written once.  The x3 output-row subsampling (ops_subsample_rows, ops.cu:290-304)
// is a row stride of the read; the gradient lands on the same rows.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/kaldi_fp16_chain.h"
#include "host_common.h"
#include "sm100_ptx.cuh"

using namespace kfp16;

namespace {

constexpr float kLogZero = -1.0e+30f;   // chain.cu:20
constexpr int kChainThreads = 1024;

struct DevFst {
  // outgoing arcs (CSR by source, as given) and incoming arcs (CSR by destination, built at upload)
  int *out_ptr = nullptr, *out_dst = nullptr, *out_pdf = nullptr;
  float* out_w = nullptr;
  int *in_ptr = nullptr, *in_src = nullptr, *in_pdf = nullptr, *in_id = nullptr;   // in_id: index of the arc in outgoing order
  float* in_w = nullptr;
  int *final_state = nullptr;
  float* final_w = nullptr;
  int S = 0, A = 0, F = 0, start = 0;
};

// one sequence's view of an FST inside a concatenated batch (numerators) or the shared denominator graph
struct FstRef {
  const int *out_ptr, *out_dst, *out_pdf;
  const float* out_w;
  const int *in_ptr, *in_src, *in_pdf, *in_id;
  const float* in_w;
  const int* final_state;
  const float* final_w;
  int S, F, start;
};

struct ChainArgs {
  const __half* nnet;      // row of (sequence s, output frame t) = nnet + ((size_t)s * seq_rows + row0 + t * row_step) * ld
  __half* grad;            // same addressing (may be null)
  int ld, seq_rows, row0, row_step;
  int T, P;
  float weight;
  float* alpha_num; float* alpha_den;   // [n_seq][(T + 1) * Smax]
  float* beta;                          // [n_seq][2 * (Snum_max + Sden)]
  int snum_max, sden;
  float* result;                        // [n_seq][4]: num_logprob, den_logprob, loss, 0
  float* loss_accum;                    // += loss of every sequence (may be null)
  long long* dbg;                       // profiling: clock64 stamps of CTA 0's phases (may be null)
};

__device__ __forceinline__ float log_add(float a, float b) {   // chain.cu:44-66 without the CAS loop
  if (b <= kLogZero) return a;
  if (a <= kLogZero) return b;
  const float mx = fmaxf(a, b), mn = fminf(a, b);
  return mx + log1pf(expf(mn - mx));
}

// one forward step of one FST: alpha[t+1] from alpha[t]; the frame's network outputs are read straight from global
// memory (a few arcs per state: no staging needed in this direction)
__device__ __forceinline__ void chain_forward_step(const FstRef& f, const ChainArgs& a, const __half* __restrict__ nrow,
                                                   const float* at, float* an) {   // (alpha is written by this kernel: no read-only path)
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int d = tid; d < f.S; d += nt) {
    float acc = kLogZero;
    for (int e = f.in_ptr[d]; e < f.in_ptr[d + 1]; ++e) {
      const int pdf = f.in_pdf[e];
      if (pdf <= 0 || pdf > a.P) continue;           // epsilon arcs are skipped (chain.cu:118-121)
      const float sa = at[f.in_src[e]];
      if (sa <= kLogZero) continue;
      acc = log_add(acc, sa + __half2float(nrow[pdf - 1]) + f.in_w[e]);
    }
    an[d] = acc;
  }
}

// total over the final states, computed by ONE WARP (lane-strided running max / scaled sum, merged by shuffles): the
// sequential form below costs a dependent global load + log1p/exp per final state (256 finals: 35 us of a 300 us kernel)
__device__ float chain_total_warp(const FstRef& f, const float* alphaT) {
  const int lane = threadIdx.x & 31;
  float mx = kLogZero, sum = 0.f;
  for (int i = lane; i < f.F; i += 32) {
    const float v = alphaT[__ldg(f.final_state + i)] + __ldg(f.final_w + i);
    if (v <= kLogZero) continue;
    if (v > mx) { sum = sum * __expf(mx - v) + 1.0f; mx = v; }
    else sum += __expf(v - mx);
  }
  for (int o = 16; o; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o), osum = __shfl_xor_sync(0xffffffffu, sum, o);
    const float nmx = fmaxf(mx, omx);
    if (nmx > kLogZero) sum = sum * __expf(mx - nmx) + osum * __expf(omx - nmx);
    mx = nmx;
  }
  return mx > kLogZero ? mx + __logf(sum) : kLogZero;
}

// total = log-sum over the final states of alpha[T][s] + final weight (chain.cu kernel_total_logprob, same order)
__device__ float chain_total(const FstRef& f, const float* alphaT) {
  float total = kLogZero;
  for (int i = 0; i < f.F; ++i) {
    const float v = alphaT[f.final_state[i]] + f.final_w[i];
    if (total <= kLogZero) total = v;
    else if (v > kLogZero) {
      const float mx = fmaxf(total, v), mn = fminf(total, v);
      total = mx + log1pf(expf(mn - mx));
    }
  }
  return total;
}

// one backward step of one FST at frame t: beta_t[s] from beta_n (= beta[t+1]); posteriors added into post[] with `sign`
__device__ void chain_backward_step(const FstRef& f, const ChainArgs& a, const float* alpha_t, const float* beta_n,
                                    float* beta_t, const float* row, float* post, float total, float sign) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int s = tid; s < f.S; s += nt) {
    float acc = kLogZero;
    const float as = alpha_t[s];
    for (int e = f.out_ptr[s]; e < f.out_ptr[s + 1]; ++e) {
      const int pdf = f.out_pdf[e];
      if (pdf <= 0 || pdf > a.P) continue;
      const float b = beta_n[f.out_dst[e]];
      if (b <= kLogZero) continue;
      const float v = b + row[pdf - 1] + f.out_w[e];
      acc = log_add(acc, v);
      if (post != nullptr && as > kLogZero) {           // chain.cu kernel_chain_posteriors
        float lp = as + v - total;
        if (lp > 0.0f) lp = 0.0f;
        atomicAdd(&post[pdf - 1], sign * expf(lp));
      }
    }
    beta_t[s] = acc;
  }
}

__global__ void __launch_bounds__(kChainThreads)
chain_loss_kernel(ChainArgs a, const FstRef* __restrict__ nums, FstRef den) {
  extern __shared__ float smem_f[];
  float* row = smem_f;              // [P] network output of the current frame, FP32
  float* post = smem_f + a.P;       // [P] den_post - num_post of the current frame
  __shared__ float totals[2];
  const int seq = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const FstRef num = nums[seq];
  const __half* nnet_seq = a.nnet + ((size_t)seq * a.seq_rows + a.row0) * a.ld;
  float* alpha_num = a.alpha_num + (size_t)seq * (a.T + 1) * a.snum_max;
  float* alpha_den = a.alpha_den + (size_t)seq * (a.T + 1) * a.sden;
  float* beta = a.beta + (size_t)seq * 2 * (a.snum_max + a.sden);
  float* bnum[2] = {beta, beta + a.snum_max};
  float* bden[2] = {beta + 2 * a.snum_max, beta + 2 * a.snum_max + a.sden};

  // forward: numerator and denominator advance together, one block barrier per frame
  for (int q = tid; q < num.S; q += nt) alpha_num[q] = q == num.start ? 0.0f : kLogZero;
  for (int q = tid; q < den.S; q += nt) alpha_den[q] = q == den.start ? 0.0f : kLogZero;
  for (int t = 0; t < a.T; ++t) {
    __syncthreads();                                   // alpha[t] complete
    const __half* nrow = nnet_seq + (size_t)t * a.row_step * a.ld;
    chain_forward_step(num, a, nrow, alpha_num + (size_t)t * num.S, alpha_num + (size_t)(t + 1) * num.S);
    chain_forward_step(den, a, nrow, alpha_den + (size_t)t * den.S, alpha_den + (size_t)(t + 1) * den.S);
  }
  __syncthreads();
  if (tid == 0) totals[0] = chain_total(num, alpha_num + (size_t)a.T * num.S);
  if (tid == 32) totals[1] = chain_total(den, alpha_den + (size_t)a.T * den.S);
  // beta[T]: final weights, log-zero elsewhere (chain.cu kernel_set_finals)
  for (int s = tid; s < num.S; s += nt) bnum[a.T & 1][s] = kLogZero;
  for (int s = tid; s < den.S; s += nt) bden[a.T & 1][s] = kLogZero;
  __syncthreads();
  if (tid == 0) for (int i = 0; i < num.F; ++i) bnum[a.T & 1][num.final_state[i]] = num.final_w[i];
  if (tid == 32) for (int i = 0; i < den.F; ++i) bden[a.T & 1][den.final_state[i]] = den.final_w[i];
  const float tot_num = totals[0], tot_den = totals[1];
  if (tid == 0) {
    float* r = a.result + (size_t)seq * 4;
    r[0] = tot_num; r[1] = tot_den; r[2] = -(tot_num - tot_den); r[3] = 0.f;
    if (a.loss_accum) atomicAdd(a.loss_accum, -(tot_num - tot_den));
  }
  if (a.grad == nullptr) return;
  __half* grad_seq = a.grad + ((size_t)seq * a.seq_rows + a.row0) * a.ld;
  for (int t = a.T - 1; t >= 0; --t) {
    const __half* src_row = nnet_seq + (size_t)t * a.row_step * a.ld;
    __syncthreads();                                   // beta[t+1] complete; previous frame's gradient row written
    for (int p = tid; p < a.P; p += nt) { row[p] = __half2float(src_row[p]); post[p] = 0.f; }
    __syncthreads();
    chain_backward_step(num, a, alpha_num + (size_t)t * num.S, bnum[(t + 1) & 1], bnum[t & 1], row, post, tot_num, -1.0f);
    chain_backward_step(den, a, alpha_den + (size_t)t * den.S, bden[(t + 1) & 1], bden[t & 1], row, post, tot_den, 1.0f);
    __syncthreads();
    __half* dst_row = grad_seq + (size_t)t * a.row_step * a.ld;
    for (int p = tid; p < a.P; p += nt) {              // chain.cu kernel_chain_gradient
      float g = post[p] * a.weight;
      g = fmaxf(-30.0f, fminf(30.0f, g));
      dst_row[p] = __float2half(g);
    }
  }
}

// ---- fast path (graphs of up to kFastArcs * 1024 arcs and kFastQ * 1024 states per sequence pair).
// A frame step is latency-bound (one CTA, ~1000 arcs), so everything that does not change from frame to frame lives in
// REGISTERS -- each thread owns up to kFastArcs arcs (source / destination state, weight, pdf) and up to kFastQ states
// (their arc ranges) for the whole kernel -- and the per-frame data in shared memory: the running alpha / beta vectors,
// the per-arc path values, the posterior row, and (when it fits) the network outputs of ALL frames gathered per arc by one
// fully parallel pass up front, so that the frame loops never wait for global memory.  Per frame: an arc-parallel pass
// (path value, and in the backward direction the arc's posterior), a barrier, a state-parallel log-sum-exp (max, scaled
// sum, one logarithm) together with the gradient-row write, a barrier.
// Measured on the benchmark's shape (64 x 50 frames x 6016 pdfs, 256-state / 1024-arc denominator), us per minibatch:
// global-memory kernel 436, first shared-memory version 302 (a state-parallel loop over dependent shared-memory reads per
// arc), register-resident arcs with the smallest <arcs, states> per thread: 190 (profiles/r02_chain_kernel.txt).  What
// bounds it now is instruction issue: ~1000 arcs x ~100 instructions per frame and direction on one SM.
constexpr int kFastQMax = 4;                     // states per thread (template argument: the smallest that covers the graphs)
constexpr int kFastArcsMax = 4;                  // arcs per thread
constexpr int kFastStates = kFastQMax * kChainThreads;
constexpr size_t kChainSmemMax = 220 * 1024;     // of the 227 KB a CTA may take

struct FastSmem {
  int p_pad, stot_max, atot_max, gathered;
  size_t rowh, post, ab, alpha_s, arc_v, total;
};
__host__ __device__ inline FastSmem fast_layout(int P, int stot_max, int atot_max, int T) {
  FastSmem L;
  L.p_pad = (P + 7) & ~7; L.stot_max = stot_max; L.atot_max = atot_max;
  const size_t fixed = (size_t)L.p_pad * sizeof(float) + (size_t)4 * stot_max * sizeof(float) + (size_t)atot_max * sizeof(float) + 64;
  const size_t vals = (((size_t)T * atot_max * sizeof(__half)) + 15) & ~(size_t)15;
  L.gathered = fixed + vals <= kChainSmemMax ? 1 : 0;
  size_t o = 0;
  L.rowh = o; o += L.gathered ? vals : (size_t)2 * L.p_pad * sizeof(__half);
  L.post = o; o += (size_t)L.p_pad * sizeof(float);
  L.ab = o; o += (size_t)2 * stot_max * sizeof(float);
  L.alpha_s = o; o += (size_t)2 * stot_max * sizeof(float);
  L.arc_v = o; o += (size_t)atot_max * sizeof(float);
  L.total = (o + 15) & ~(size_t)15;
  return L;
}

template <int kFastArcs, int kFastQ>
__global__ void __launch_bounds__(kChainThreads)
chain_loss_fast_kernel(ChainArgs a, const FstRef* __restrict__ nums, FstRef den, int stot_max, int atot_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FastSmem L = fast_layout(a.P, stot_max, atot_max, a.T);
  const bool gathered = L.gathered != 0;
  __half* rowh = reinterpret_cast<__half*>(smem_raw + L.rowh);
  float* post = reinterpret_cast<float*>(smem_raw + L.post);
  float* ab = reinterpret_cast<float*>(smem_raw + L.ab);            // alpha (forward) / beta (backward), double-buffered
  float* alpha_s = reinterpret_cast<float*>(smem_raw + L.alpha_s);  // backward: alpha[t], staged one frame ahead
  float* arc_v = reinterpret_cast<float*>(smem_raw + L.arc_v);      // per-arc path value of the current frame
  __shared__ float totals[2];
  const int seq = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const FstRef num = nums[seq];
  const int nS = num.S, dS = den.S, stot = nS + dS;
  const int nA = __ldg(num.out_ptr + nS), dA = __ldg(den.out_ptr + dS), atot = nA + dA;
  const __half* nnet_seq = a.nnet + ((size_t)seq * a.seq_rows + a.row0) * a.ld;
  float* alpha_num = a.alpha_num + (size_t)seq * (a.T + 1) * a.snum_max;
  float* alpha_den = a.alpha_den + (size_t)seq * (a.T + 1) * a.sden;
  const int pv = L.p_pad / 8;                     // 16-byte chunks per row (P % 8 == 0 on this path)
  const size_t row_stride = (size_t)a.row_step * a.ld;
  auto alpha_at = [&](int t, int q) -> float* { return q < nS ? alpha_num + (size_t)t * nS + q : alpha_den + (size_t)t * dS + (q - nS); };

  // ---- this thread's arcs and states (combined numbering: numerator first), loaded once per direction
  int e_st[kFastArcs], e_src[kFastArcs], e_pdf[kFastArcs], e_col[kFastArcs];
  float e_w[kFastArcs];
  int q_e0[kFastQ], q_e1[kFastQ];
  auto load_mine = [&](bool incoming) {
#pragma unroll
    for (int k = 0; k < kFastArcs; ++k) {
      const int e = tid + k * nt;
      e_st[k] = 0; e_src[k] = 0; e_pdf[k] = 0; e_col[k] = 0; e_w[k] = 0.f;
      if (e >= atot) continue;
      const bool isn = e < nA;
      const int le = isn ? e : e - nA, base = isn ? 0 : nS;
      const FstRef& f = isn ? num : den;
      e_st[k] = base + __ldg((incoming ? f.in_src : f.out_dst) + le);
      const int pdf = __ldg((incoming ? f.in_pdf : f.out_pdf) + le);
      e_pdf[k] = (pdf > 0 && pdf <= a.P) ? pdf : 0;
      e_w[k] = __ldg((incoming ? f.in_w : f.out_w) + le);
      e_col[k] = (isn ? 0 : nA) + (incoming ? __ldg(f.in_id + le) : le);   // column of the gathered table (outgoing order)
    }
#pragma unroll
    for (int k = 0; k < kFastQ; ++k) {
      const int q = tid + k * nt;
      q_e0[k] = q_e1[k] = 0;
      if (q >= stot) continue;
      const bool isn = q < nS;
      const int* p = isn ? (incoming ? num.in_ptr : num.out_ptr) : (incoming ? den.in_ptr : den.out_ptr);
      const int lq = isn ? q : q - nS, off = isn ? 0 : nA;
      q_e0[k] = off + __ldg(p + lq); q_e1[k] = off + __ldg(p + lq + 1);
    }
  };
  // network output arc k of this thread reads at frame t
  auto arc_val = [&](int t, int k) -> float {
    return __half2float(gathered ? rowh[(size_t)t * atot + e_col[k]] : rowh[(t & 1) * L.p_pad + e_pdf[k] - 1]);
  };
  // log-sum-exp over arc_v[e0, e1): max, scaled sum, one logarithm (dead arcs hold log-zero: exp underflows to 0)
  auto lse = [&](int e0, int e1) -> float {
    float mx = kLogZero;
    for (int e = e0; e < e1; ++e) mx = fmaxf(mx, arc_v[e]);
    if (mx <= kLogZero) return kLogZero;
    float sum = 0.f;
    for (int e = e0; e < e1; ++e) sum += __expf(arc_v[e] - mx);
    return mx + __logf(sum);
  };

  if (gathered) {   // vals[t][arc in outgoing order] = nnet[t][pdf(arc) - 1]: one arc per thread and pass, 10 frames in flight
    const unsigned short* nn16 = reinterpret_cast<const unsigned short*>(nnet_seq);
    unsigned short* vals = reinterpret_cast<unsigned short*>(rowh);
    for (int e = tid; e < atot; e += nt) {
      const int pdf = e < nA ? __ldg(num.out_pdf + e) : __ldg(den.out_pdf + (e - nA));
      const bool ok = pdf > 0 && pdf <= a.P;
      const unsigned short* col = nn16 + (ok ? pdf - 1 : 0);
#pragma unroll 10
      for (int t = 0; t < a.T; ++t) vals[(size_t)t * atot + e] = ok ? __ldg(col + (size_t)t * row_stride) : (unsigned short)0;
    }
  }
  if (a.dbg && seq == 0 && tid == 0) a.dbg[0] = clock64();

  // ------------------------------------------------------------------ forward
  load_mine(true);
#pragma unroll
  for (int k = 0; k < kFastQ; ++k) {
    const int q = tid + k * nt;
    if (q >= stot) continue;
    const float v = (q < nS ? q == num.start : (q - nS) == den.start) ? 0.0f : kLogZero;
    ab[q] = v;
    *alpha_at(0, q) = v;
  }
  if (!gathered && tid < pv) reinterpret_cast<uint4*>(rowh)[tid] = __ldg(reinterpret_cast<const uint4*>(nnet_seq) + tid);
  __syncthreads();
  if (a.dbg && seq == 0 && tid == 0) a.dbg[1] = clock64();
  for (int t = 0; t < a.T; ++t) {
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (!gathered && t + 1 < a.T && tid < pv) nxt = __ldg(reinterpret_cast<const uint4*>(nnet_seq + (size_t)(t + 1) * row_stride) + tid);
    const float* cur = ab + (t & 1) * stot_max;
    float* nx = ab + ((t + 1) & 1) * stot_max;
#pragma unroll
    for (int k = 0; k < kFastArcs; ++k) {                // per arc: v = alpha[t][src] + network output + weight
      const int e = tid + k * nt;
      if (e >= atot) continue;
      const float sa = cur[e_st[k]];
      arc_v[e] = (e_pdf[k] == 0 || sa <= kLogZero) ? kLogZero : sa + arc_val(t, k) + e_w[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFastQ; ++k) {                   // per state: alpha[t+1] = log-sum-exp over its incoming arcs
      const int q = tid + k * nt;
      if (q >= stot) continue;
      const float acc = lse(q_e0[k], q_e1[k]);
      nx[q] = acc;
      *alpha_at(t + 1, q) = acc;
    }
    if (!gathered && tid < pv) reinterpret_cast<uint4*>(rowh + ((t + 1) & 1) * L.p_pad)[tid] = nxt;
    __syncthreads();
  }
  if (a.dbg && seq == 0 && tid == 0) a.dbg[2] = clock64();
  {
    const float* aT = ab + (a.T & 1) * stot_max;
    if (tid < 32) { const float v = chain_total_warp(num, aT); if (tid == 0) totals[0] = v; }
    else if (tid < 64) { const float v = chain_total_warp(den, aT + nS); if (tid == 32) totals[1] = v; }
  }
  __syncthreads();
  const float tot_num = totals[0], tot_den = totals[1];
  if (tid == 0) {
    float* r = a.result + (size_t)seq * 4;
    r[0] = tot_num; r[1] = tot_den; r[2] = -(tot_num - tot_den); r[3] = 0.f;
    if (a.loss_accum) atomicAdd(a.loss_accum, -(tot_num - tot_den));
  }
  if (a.grad == nullptr) return;

  // ------------------------------------------------------------------ backward + posteriors + gradient rows
  __half* grad_seq = a.grad + ((size_t)seq * a.seq_rows + a.row0) * a.ld;
  load_mine(false);
  // source state of each of this thread's (outgoing-order) arcs: needed for the posteriors
#pragma unroll
  for (int k = 0; k < kFastArcs; ++k) {
    const int e = tid + k * nt;
    if (e >= atot) continue;
    const bool isn = e < nA;
    const int* p = isn ? num.out_ptr : den.out_ptr;
    const int le = isn ? e : e - nA, S = isn ? nS : dS;
    int lo = 0, hi = S - 1;                               // largest s with out_ptr[s] <= le (once per kernel)
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (__ldg(p + mid) <= le) lo = mid; else hi = mid - 1; }
    e_src[k] = (isn ? 0 : nS) + lo;
  }
#pragma unroll
  for (int k = 0; k < kFastQ; ++k) {
    const int q = tid + k * nt;
    if (q >= stot) continue;
    ab[(a.T & 1) * stot_max + q] = kLogZero;
    alpha_s[((a.T - 1) & 1) * stot_max + q] = *alpha_at(a.T - 1, q);
  }
  for (int p = tid; p < L.p_pad; p += nt) post[p] = 0.f;
  __syncthreads();
  if (tid == 0) for (int i = 0; i < num.F; ++i) ab[(a.T & 1) * stot_max + num.final_state[i]] = num.final_w[i];
  if (tid == 32) for (int i = 0; i < den.F; ++i) ab[(a.T & 1) * stot_max + nS + den.final_state[i]] = den.final_w[i];
  if (!gathered && tid < pv) reinterpret_cast<uint4*>(rowh + ((a.T - 1) & 1) * L.p_pad)[tid] = __ldg(reinterpret_cast<const uint4*>(nnet_seq + (size_t)(a.T - 1) * row_stride) + tid);
  __syncthreads();
  if (a.dbg && seq == 0 && tid == 0) a.dbg[3] = clock64();
  for (int t = a.T - 1; t >= 0; --t) {
    // next frame's alpha (and row): issued now, stored to shared memory at the end of this frame
    uint4 nxt = make_uint4(0, 0, 0, 0);
    float a_pre[kFastQ];
    if (t > 0) {
      if (!gathered && tid < pv) nxt = __ldg(reinterpret_cast<const uint4*>(nnet_seq + (size_t)(t - 1) * row_stride) + tid);
#pragma unroll
      for (int k = 0; k < kFastQ; ++k) { const int q = tid + k * nt; a_pre[k] = q < stot ? *alpha_at(t - 1, q) : kLogZero; }
    }
    const float* bn = ab + ((t + 1) & 1) * stot_max;
    float* bt = ab + (t & 1) * stot_max;
    const float* al = alpha_s + (t & 1) * stot_max;
    const bool stamp = a.dbg && seq == 0 && tid == 0 && t == 10;
    if (stamp) a.dbg[8] = clock64();
#pragma unroll
    for (int k = 0; k < kFastArcs; ++k) {                // per arc: path value and posterior (chain.cu kernel_chain_posteriors)
      const int e = tid + k * nt;
      if (e >= atot) continue;
      const float b = bn[e_st[k]];
      float v = kLogZero;
      if (e_pdf[k] != 0 && b > kLogZero) {
        v = b + arc_val(t, k) + e_w[k];
        const float as = al[e_src[k]];
        if (as > kLogZero) {
          const bool isn = e < nA;
          atomicAdd(&post[e_pdf[k] - 1], (isn ? -1.0f : 1.0f) * __expf(fminf(as + v - (isn ? tot_num : tot_den), 0.0f)));
        }
      }
      arc_v[e] = v;
    }
    if (stamp) a.dbg[9] = clock64();
    __syncthreads();                                     // path values and the posteriors of frame t complete
    if (stamp) a.dbg[10] = clock64();
#pragma unroll
    for (int k = 0; k < kFastQ; ++k) {                   // per state: beta[t]
      const int q = tid + k * nt;
      if (q >= stot) continue;
      bt[q] = lse(q_e0[k], q_e1[k]);
    }
    if (stamp) a.dbg[11] = clock64();
    if (tid < pv) {                                      // gradient row t: 8 pdfs per thread, one 16-byte store (kernel_chain_gradient)
      __half2 h[4];
      float4* pp = reinterpret_cast<float4*>(post + tid * 8);       // 16-byte shared-memory accesses (scalar ones at this
      const float4 p0 = pp[0], p1 = pp[1];                          // stride are 8-way bank conflicts)
      pp[0] = make_float4(0.f, 0.f, 0.f, 0.f); pp[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float pv8[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g0 = fmaxf(-30.0f, fminf(30.0f, pv8[2 * j] * a.weight));
        const float g1 = fmaxf(-30.0f, fminf(30.0f, pv8[2 * j + 1] * a.weight));
        h[j] = __floats2half2_rn(g0, g1);
      }
      *reinterpret_cast<uint4*>(grad_seq + (size_t)t * row_stride + tid * 8) = *reinterpret_cast<const uint4*>(h);
      if (!gathered && t > 0) reinterpret_cast<uint4*>(rowh + ((t - 1) & 1) * L.p_pad)[tid] = nxt;
    }
    if (stamp) a.dbg[12] = clock64();
    if (t > 0) {
#pragma unroll
      for (int k = 0; k < kFastQ; ++k) { const int q = tid + k * nt; if (q < stot) alpha_s[((t - 1) & 1) * stot_max + q] = a_pre[k]; }
    }
    if (stamp) a.dbg[13] = clock64();
    __syncthreads();                                     // beta[t], cleared posteriors, next alpha / row staged
    if (stamp) a.dbg[14] = clock64();
  }
  if (a.dbg && seq == 0 && tid == 0) a.dbg[4] = clock64();
}

bool upload(void** dev, const void* host, size_t bytes) {
  if (bytes == 0) bytes = 4;
  if (!check_cuda(cudaMalloc(dev, bytes), "cudaMalloc (chain)")) return false;
  return host == nullptr || check_cuda(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice), "chain upload");
}

// CSR by source -> also CSR by destination; validates indices
bool build_fst(DevFst& f, const kfp16_chain_fst& h, const char* what) {
  if (h.num_states < 1 || h.num_arcs < 0 || !h.row_ptr || (h.num_arcs && (!h.col_idx || !h.labels || !h.weights)) ||
      h.start_state < 0 || h.start_state >= h.num_states || h.num_final < 0 || (h.num_final && (!h.final_states || !h.final_weights))) {
    set_error("%s: malformed FST (states %d arcs %d finals %d start %d)", what, h.num_states, h.num_arcs, h.num_final, h.start_state);
    return false;
  }
  const int S = h.num_states, A = h.num_arcs;
  if (h.row_ptr[0] != 0 || h.row_ptr[S] != A) { set_error("%s: row_ptr must run from 0 to num_arcs", what); return false; }
  std::vector<int> in_ptr(S + 1, 0), in_src(A), in_pdf(A), in_id(A);
  std::vector<float> in_w(A);
  for (int s = 0; s < S; ++s) {
    if (h.row_ptr[s + 1] < h.row_ptr[s]) { set_error("%s: row_ptr not monotone", what); return false; }
    for (int e = h.row_ptr[s]; e < h.row_ptr[s + 1]; ++e) {
      if (h.col_idx[e] < 0 || h.col_idx[e] >= S) { set_error("%s: arc %d points to state %d of %d", what, e, h.col_idx[e], S); return false; }
      in_ptr[h.col_idx[e] + 1]++;
    }
  }
  for (int s = 0; s < S; ++s) in_ptr[s + 1] += in_ptr[s];
  std::vector<int> fill(in_ptr.begin(), in_ptr.end() - 1);
  for (int s = 0; s < S; ++s)
    for (int e = h.row_ptr[s]; e < h.row_ptr[s + 1]; ++e) {
      const int k = fill[h.col_idx[e]]++;
      in_src[k] = s; in_pdf[k] = h.labels[e]; in_w[k] = h.weights[e]; in_id[k] = e;
    }
  for (int i = 0; i < h.num_final; ++i)
    if (h.final_states[i] < 0 || h.final_states[i] >= S) { set_error("%s: final state out of range", what); return false; }
  f.S = S; f.A = A; f.F = h.num_final; f.start = h.start_state;
  return upload((void**)&f.out_ptr, h.row_ptr, (S + 1) * 4) && upload((void**)&f.out_dst, h.col_idx, (size_t)A * 4) &&
         upload((void**)&f.out_pdf, h.labels, (size_t)A * 4) && upload((void**)&f.out_w, h.weights, (size_t)A * 4) &&
         upload((void**)&f.in_ptr, in_ptr.data(), (S + 1) * 4) && upload((void**)&f.in_src, in_src.data(), (size_t)A * 4) &&
         upload((void**)&f.in_pdf, in_pdf.data(), (size_t)A * 4) && upload((void**)&f.in_w, in_w.data(), (size_t)A * 4) &&
         upload((void**)&f.in_id, in_id.data(), (size_t)A * 4) &&
         upload((void**)&f.final_state, h.final_states, (size_t)h.num_final * 4) && upload((void**)&f.final_w, h.final_weights, (size_t)h.num_final * 4);
}
void free_fst(DevFst& f) {
  for (void* p : {(void*)f.out_ptr, (void*)f.out_dst, (void*)f.out_pdf, (void*)f.out_w, (void*)f.in_ptr, (void*)f.in_src,
                  (void*)f.in_pdf, (void*)f.in_id, (void*)f.in_w, (void*)f.final_state, (void*)f.final_w})
    if (p) cudaFree(p);
  f = DevFst();
}
FstRef ref_of(const DevFst& f) {
  return FstRef{f.out_ptr, f.out_dst, f.out_pdf, f.out_w, f.in_ptr, f.in_src, f.in_pdf, f.in_id, f.in_w, f.final_state, f.final_w, f.S, f.F, f.start};
}

}  // namespace

struct kfp16_chain {
  kfp16_ctx* ctx = nullptr;
  int num_pdfs = 0, n_seq = 0, frames = 0;
  DevFst den;
  std::vector<DevFst> nums;
  FstRef* nums_dev = nullptr;
  int snum_max = 0, anum_max = 0;
  long long* dbg = nullptr;
  bool force_general = false;   // tests: run the global-memory kernel even when the shared-memory one would fit
  float *alpha_num = nullptr, *alpha_den = nullptr, *beta = nullptr, *result = nullptr;
  size_t alpha_num_elems = 0;
};

// One CTA per sequence.  With an even sequence count the CTAs are launched as clusters of two -- not to cooperate (the kernels
// never look at their cluster rank) but so that the block scheduler packs them two to a TPC: 64 sequences then hold 32 whole
// TPCs and leave the other 42 free for the CTA-pair GEMMs the network executor runs beside the objective
// (nnet.cu run_phases: the xent branch's forward pass overlaps the chain kernel).
template <typename Kern, typename... Args>
static void launch_chain(Kern kern, int grid, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = (grid % 2) == 0 ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);
}


extern "C" {

kfp16_chain* kfp16_chain_create(kfp16_ctx* ctx, int num_pdfs, int n_seq, int frames_per_seq, const kfp16_chain_fst* den) {
  if (!ctx || !den || num_pdfs < 1 || n_seq < 1 || frames_per_seq < 1) { set_error("kfp16_chain_create: bad argument"); return nullptr; }
  if ((size_t)num_pdfs * 8 > 200 * 1024) { set_error("kfp16_chain_create: %d pdfs exceed the shared-memory row buffers", num_pdfs); return nullptr; }
  if (!check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice")) return nullptr;
  kfp16_chain* c = new kfp16_chain();
  c->ctx = ctx; c->num_pdfs = num_pdfs; c->n_seq = n_seq; c->frames = frames_per_seq;
  if (!build_fst(c->den, *den, "kfp16_chain_create (denominator)") ||
      !upload((void**)&c->alpha_den, nullptr, (size_t)n_seq * (frames_per_seq + 1) * c->den.S * 4) ||
      !upload((void**)&c->result, nullptr, (size_t)n_seq * 4 * 4) ||
      !upload((void**)&c->nums_dev, nullptr, (size_t)n_seq * sizeof(FstRef))) {
    kfp16_chain_destroy(c);
    return nullptr;
  }
  c->nums.resize(n_seq);
  return c;
}

void kfp16_chain_destroy(kfp16_chain* c) {
  if (!c) return;
  free_fst(c->den);
  for (DevFst& f : c->nums) free_fst(f);
  for (void* p : {(void*)c->nums_dev, (void*)c->alpha_num, (void*)c->alpha_den, (void*)c->beta, (void*)c->result})
    if (p) cudaFree(p);
  delete c;
}

// numerator FSTs of the current minibatch (per-sequence supervision, batch.PerSeqCSRs in train_step.go:183-193)
int kfp16_chain_set_numerators(kfp16_chain* c, const kfp16_chain_fst* nums, int n_seq) {
  if (!c || !nums || n_seq != c->n_seq) { set_error("kfp16_chain_set_numerators: expected %d numerator FSTs", c ? c->n_seq : 0); return -1; }
  if (!check_cuda(cudaSetDevice(c->ctx->device), "cudaSetDevice") || !check_cuda(cudaStreamSynchronize(c->ctx->stream), "sync")) return -1;
  std::vector<FstRef> refs(n_seq);
  int smax = 1, amax = 0;
  for (int i = 0; i < n_seq; ++i) {
    free_fst(c->nums[i]);
    if (!build_fst(c->nums[i], nums[i], "kfp16_chain_set_numerators")) return -1;
    refs[i] = ref_of(c->nums[i]);
    smax = std::max(smax, c->nums[i].S);
    amax = std::max(amax, c->nums[i].A);
  }
  c->anum_max = amax;
  if (!check_cuda(cudaMemcpy(c->nums_dev, refs.data(), sizeof(FstRef) * n_seq, cudaMemcpyHostToDevice), "numerator table upload")) return -1;
  const size_t need = (size_t)n_seq * (c->frames + 1) * smax;
  if (need > c->alpha_num_elems || smax != c->snum_max) {
    if (c->alpha_num) cudaFree(c->alpha_num);
    if (c->beta) cudaFree(c->beta);
    c->alpha_num = c->beta = nullptr;
    if (!upload((void**)&c->alpha_num, nullptr, need * 4) ||
        !upload((void**)&c->beta, nullptr, (size_t)n_seq * 2 * (smax + c->den.S) * 4)) return -1;
    c->alpha_num_elems = need;
    c->snum_max = smax;
  }
  return 0;
}

int kfp16_chain_loss(kfp16_chain* c, const void* nnet_out, void* grad_out, int ld, int seq_rows, int row0, int row_step,
                     float supervision_weight, float* loss_accum_dev) {
  if (!c || !nnet_out) { set_error("kfp16_chain_loss: null argument"); return -1; }
  if (!c->alpha_num) { set_error("kfp16_chain_loss: no numerator FSTs set (kfp16_chain_set_numerators)"); return -1; }
  if (ld < c->num_pdfs || row_step < 1 || row0 < 0 || row0 + (c->frames - 1) * row_step >= seq_rows) {
    set_error("kfp16_chain_loss: %d output frames at rows %d + k*%d do not fit a sequence block of %d rows (ld %d, %d pdfs)",
              c->frames, row0, row_step, seq_rows, ld, c->num_pdfs);
    return -1;
  }
  if (!check_cuda(cudaSetDevice(c->ctx->device), "cudaSetDevice")) return -1;
  ChainArgs a;
  a.nnet = (const __half*)nnet_out; a.grad = (__half*)grad_out;
  a.ld = ld; a.seq_rows = seq_rows; a.row0 = row0; a.row_step = row_step;
  a.T = c->frames; a.P = c->num_pdfs; a.weight = supervision_weight;
  a.alpha_num = c->alpha_num; a.alpha_den = c->alpha_den; a.beta = c->beta;
  a.snum_max = c->snum_max; a.sden = c->den.S;
  a.result = c->result; a.loss_accum = loss_accum_dev; a.dbg = c->dbg;
  static bool attr_done = false;
  if (!attr_done) {
    if (!check_cuda(cudaFuncSetAttribute(chain_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "cudaFuncSetAttribute(chain)") ||
        !check_cuda(cudaFuncSetAttribute(chain_loss_fast_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemMax), "cudaFuncSetAttribute(chain fast)") ||
        !check_cuda(cudaFuncSetAttribute(chain_loss_fast_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemMax), "cudaFuncSetAttribute(chain fast)") ||
        !check_cuda(cudaFuncSetAttribute(chain_loss_fast_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemMax), "cudaFuncSetAttribute(chain fast)") ||
        !check_cuda(cudaFuncSetAttribute(chain_loss_fast_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemMax), "cudaFuncSetAttribute(chain fast)")) return -1;
    attr_done = true;
  }
  // everything in shared memory when it fits (small graphs: the usual numerator, a compact denominator)
  const int stot_max = c->snum_max + c->den.S, atot_max = c->anum_max + c->den.A;
  const FastSmem L = fast_layout(c->num_pdfs, stot_max, atot_max, c->frames);
  const bool aligned = (c->num_pdfs % 8) == 0 && (ld % 8) == 0 && ((uintptr_t)nnet_out & 15) == 0 && (!grad_out || ((uintptr_t)grad_out & 15) == 0);
  if (aligned && stot_max <= kFastStates && atot_max <= kFastArcsMax * kChainThreads && L.total <= kChainSmemMax && !c->force_general) {
    // the smallest per-thread arc / state counts that cover the graphs (less unrolled code, fewer registers)
    const int ar = (atot_max + kChainThreads - 1) / kChainThreads, qr = (stot_max + kChainThreads - 1) / kChainThreads;
    const FstRef dref = ref_of(c->den);
    if (ar <= 1 && qr <= 1) launch_chain(chain_loss_fast_kernel<1, 1>, c->n_seq, L.total, c->ctx->stream, a, (const FstRef*)c->nums_dev, dref, stot_max, atot_max);
    else if (ar <= 2 && qr <= 1) launch_chain(chain_loss_fast_kernel<2, 1>, c->n_seq, L.total, c->ctx->stream, a, (const FstRef*)c->nums_dev, dref, stot_max, atot_max);
    else if (qr <= 2) launch_chain(chain_loss_fast_kernel<4, 2>, c->n_seq, L.total, c->ctx->stream, a, (const FstRef*)c->nums_dev, dref, stot_max, atot_max);
    else launch_chain(chain_loss_fast_kernel<4, 4>, c->n_seq, L.total, c->ctx->stream, a, (const FstRef*)c->nums_dev, dref, stot_max, atot_max);
  } else {
    const size_t smem = (size_t)c->num_pdfs * 2 * sizeof(float);
    launch_chain(chain_loss_kernel, c->n_seq, smem, c->ctx->stream, a, (const FstRef*)c->nums_dev, ref_of(c->den));
  }
  count_launch();
  return check_launch("kfp16_chain_loss") ? 0 : -1;
}

int kfp16_chain_read_results(kfp16_chain* c, float* host, int n_seq) {
  if (!c || !host || n_seq != c->n_seq) { set_error("kfp16_chain_read_results: bad argument"); return -1; }
  if (!check_cuda(cudaStreamSynchronize(c->ctx->stream), "sync")) return -1;
  return check_cuda(cudaMemcpy(host, c->result, (size_t)n_seq * 4 * sizeof(float), cudaMemcpyDeviceToHost), "chain results download") ? 0 : -1;
}

int kfp16_chain_set_debug(kfp16_chain* c, void* dev_i64x8) { if (!c) return -1; c->dbg = (long long*)dev_i64x8; return 0; }
int kfp16_chain_force_general(kfp16_chain* c, int on) { if (!c) return -1; c->force_general = on != 0; return 0; }
int kfp16_chain_num_sequences(const kfp16_chain* c) { return c ? c->n_seq : 0; }
int kfp16_chain_frames(const kfp16_chain* c) { return c ? c->frames : 0; }

}  // extern "C"
