// The reference's own chain entry points (cpp/include/chain.h:47-160: chain_forward_backward, chain_compute_posteriors,
// chain_compute_loss, chain_workspace_bytes, chain_last_error, chain_clear_error) on top of this library, so that
// internal/nnet/chain_loss.go (ComputeChainLoss :172-219, ForwardBackward :300-348) links against libkaldi_fp16.so alone.
// Same contract: FST arrays are DEVICE pointers in CSR form (ChainFstGPU), nnet_output is FP16 [T x num_pdfs], labels are
// 1-indexed pdf-ids (0 = epsilon, skipped), everything in the log semiring, results returned synchronously.
//
// Not a translation of cpp/cuda/chain.cu (one launch per frame and arc kernel with atomic log-adds and a binary search per arc):
//   * chain_compute_loss runs the batched objective kernel of csrc/chain.cu for ONE sequence (kfp16_chain_loss);
//   * chain_forward_backward is one single-CTA launch that walks all frames: forward over per-state INCOMING arc lists
//     (built here from the CSR), backward over the outgoing lists -- gathers, no atomics, reproducible;
//   * chain_compute_posteriors is one launch over (frame, arc) with fp32 atomicAdd, as the reference.
// The FST arrays are small (the shim copies them to the host once per call to build the incoming lists / the host CSR
// kfp16_chain takes); this is the compatibility path -- the training step uses kfp16_net_set_chain.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include <vector>

#include "../../include/kaldi_fp16_chain.h"
#include "host_common.h"

using namespace kfp16;

namespace {

constexpr float kLogZeroC = -1e30f;
thread_local char g_chain_err[512];
thread_local bool g_chain_has_err = false;

void chain_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_chain_err, sizeof(g_chain_err), fmt, ap);
  va_end(ap);
  g_chain_has_err = true;
}
bool ck(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  chain_set_error("%s: %s", what, cudaGetErrorString(e));
  return false;
}

__device__ __forceinline__ float log_add(float a, float b) {
  if (a <= kLogZeroC) return b;
  if (b <= kLogZeroC) return a;
  const float mx = fmaxf(a, b), mn = fminf(a, b);
  return mx + log1pf(expf(mn - mx));
}

// host copy of a ChainFstGPU + incoming-arc lists
struct HostFst {
  std::vector<int32_t> row_ptr, col_idx, labels, final_states;
  std::vector<float> weights, final_weights;
  int S = 0, A = 0, F = 0, start = 0;
  // incoming lists: arcs sorted by destination; in_ptr[S+1], in_arc[A] = original arc index, in_src[A] = its source state
  std::vector<int32_t> in_ptr, in_arc, in_src;
};

bool download(const ChainFstGPU* f, HostFst& h, bool incoming) {
  if (!f || f->num_states <= 0 || f->num_arcs < 0 || f->num_final < 0) { chain_set_error("bad ChainFstGPU"); return false; }
  h.S = f->num_states; h.A = f->num_arcs; h.F = f->num_final; h.start = f->start_state;
  h.row_ptr.resize(h.S + 1); h.col_idx.resize(h.A); h.labels.resize(h.A); h.weights.resize(h.A);
  h.final_states.resize(h.F); h.final_weights.resize(h.F);
  if (!ck(cudaMemcpy(h.row_ptr.data(), f->row_ptr, (h.S + 1) * 4, cudaMemcpyDeviceToHost), "row_ptr download")) return false;
  if (h.A && (!ck(cudaMemcpy(h.col_idx.data(), f->col_idx, h.A * 4, cudaMemcpyDeviceToHost), "col_idx download") ||
              !ck(cudaMemcpy(h.labels.data(), f->labels, h.A * 4, cudaMemcpyDeviceToHost), "labels download") ||
              !ck(cudaMemcpy(h.weights.data(), f->weights, h.A * 4, cudaMemcpyDeviceToHost), "weights download"))) return false;
  if (h.F && (!ck(cudaMemcpy(h.final_states.data(), f->final_states, h.F * 4, cudaMemcpyDeviceToHost), "final_states download") ||
              !ck(cudaMemcpy(h.final_weights.data(), f->final_weights, h.F * 4, cudaMemcpyDeviceToHost), "final_weights download"))) return false;
  if (h.row_ptr[0] != 0 || h.row_ptr[h.S] != h.A) { chain_set_error("ChainFstGPU: row_ptr does not span the %d arcs", h.A); return false; }
  if (incoming) {
    h.in_ptr.assign(h.S + 1, 0);
    for (int a = 0; a < h.A; ++a) {
      if (h.col_idx[a] < 0 || h.col_idx[a] >= h.S) { chain_set_error("ChainFstGPU: arc %d leads to state %d of %d", a, h.col_idx[a], h.S); return false; }
      h.in_ptr[h.col_idx[a] + 1]++;
    }
    for (int s = 0; s < h.S; ++s) h.in_ptr[s + 1] += h.in_ptr[s];
    h.in_arc.resize(h.A); h.in_src.resize(h.A);
    std::vector<int32_t> fill(h.in_ptr.begin(), h.in_ptr.end() - 1);
    for (int s = 0; s < h.S; ++s)
      for (int a = h.row_ptr[s]; a < h.row_ptr[s + 1]; ++a) {
        const int pos = fill[h.col_idx[a]]++;
        h.in_arc[pos] = a; h.in_src[pos] = s;
      }
  }
  return true;
}

// whole forward-backward of one FST in one CTA: alpha / beta are the caller's [(T+1) x S] fp32 workspaces
__global__ void __launch_bounds__(1024)
fst_forward_backward_kernel(const __half* __restrict__ nnet, int T, int P, int S, const int32_t* __restrict__ row_ptr,
                            const int32_t* __restrict__ col_idx, const int32_t* __restrict__ labels, const float* __restrict__ weights,
                            const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_arc, const int32_t* __restrict__ in_src,
                            const int32_t* __restrict__ final_states, const float* __restrict__ final_weights, int num_final, int start,
                            float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ total) {
  for (int i = threadIdx.x; i < (T + 1) * S; i += blockDim.x) { alpha[i] = kLogZeroC; beta[i] = kLogZeroC; }
  __syncthreads();
  if (threadIdx.x == 0) alpha[start] = 0.f;
  for (int i = threadIdx.x; i < num_final; i += blockDim.x) beta[(size_t)T * S + final_states[i]] = final_weights[i];
  __syncthreads();
  for (int t = 0; t < T; ++t) {       // alpha[t+1][d] = logsum over incoming arcs (s -> d, pdf, w) of alpha[t][s] + nnet[t][pdf-1] + w
    const float* a0 = alpha + (size_t)t * S;
    for (int d = threadIdx.x; d < S; d += blockDim.x) {
      float acc = kLogZeroC;
      for (int k = in_ptr[d]; k < in_ptr[d + 1]; ++k) {
        const int arc = in_arc[k], pdf = labels[arc];
        const float sa = a0[in_src[k]];
        if (pdf <= 0 || pdf > P || sa <= kLogZeroC) continue;
        acc = log_add(acc, sa + __half2float(nnet[(size_t)t * P + pdf - 1]) + weights[arc]);
      }
      alpha[(size_t)(t + 1) * S + d] = acc;
    }
    __syncthreads();
  }
  for (int t = T - 1; t >= 0; --t) {  // beta[t][s] = logsum over outgoing arcs of beta[t+1][d] + nnet[t][pdf-1] + w
    const float* b1 = beta + (size_t)(t + 1) * S;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      float acc = kLogZeroC;
      for (int a = row_ptr[s]; a < row_ptr[s + 1]; ++a) {
        const int pdf = labels[a];
        const float db = b1[col_idx[a]];
        if (pdf <= 0 || pdf > P || db <= kLogZeroC) continue;
        acc = log_add(acc, db + __half2float(nnet[(size_t)t * P + pdf - 1]) + weights[a]);
      }
      beta[(size_t)t * S + s] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float tot = kLogZeroC;
    for (int i = 0; i < num_final; ++i) tot = log_add(tot, alpha[(size_t)T * S + final_states[i]] + final_weights[i]);
    *total = tot;
  }
}

// posteriors[t][pdf-1] += exp(alpha[t][src] + nnet + w + beta[t+1][dst] - total), one thread per (frame, outgoing arc slot)
__global__ void fst_posteriors_kernel(float* __restrict__ post, const float* __restrict__ alpha, const float* __restrict__ beta,
                                      const __half* __restrict__ nnet, const int32_t* __restrict__ arc_src, const int32_t* __restrict__ col_idx,
                                      const int32_t* __restrict__ labels, const float* __restrict__ weights, int S, int P, int A, int T,
                                      float total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)T * A) return;
  const int t = (int)(i / A), a = (int)(i % A);
  const int pdf = labels[a];
  if (pdf <= 0 || pdf > P) return;
  const float av = alpha[(size_t)t * S + arc_src[a]], bv = beta[(size_t)(t + 1) * S + col_idx[a]];
  if (av <= kLogZeroC || bv <= kLogZeroC) return;
  const float lp = fminf(av + __half2float(nnet[(size_t)t * P + pdf - 1]) + weights[a] + bv - total, 0.f);
  atomicAdd(&post[(size_t)t * P + pdf - 1], expf(lp));
}

struct DevArr {
  void* p = nullptr;
  ~DevArr() { if (p) cudaFree(p); }
  bool up(const void* host, size_t bytes) {
    if (!ck(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc (chain compat)")) return false;
    return bytes == 0 || ck(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice), "upload (chain compat)");
  }
};

// one context per device for chain_compute_loss (the reference's entry point carries none)
kfp16_ctx* compat_ctx() {
  static std::mutex mu;
  static std::unordered_map<int, kfp16_ctx*> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) return it->second;
  kfp16_ctx* c = kfp16_ctx_create(dev);
  cache[dev] = c;
  return c;
}

}  // namespace

extern "C" {

size_t chain_workspace_bytes(int T, int num_states) { return 2 * (size_t)(T + 1) * (size_t)num_states * sizeof(float); }

const char* chain_last_error(void) { return g_chain_has_err ? g_chain_err : nullptr; }
void chain_clear_error(void) { g_chain_has_err = false; g_chain_err[0] = 0; }

int chain_forward_backward(const void* nnet_output, const ChainFstGPU* fst, int T, int num_pdfs, float* alpha, float* beta,
                           float* total_logprob) {
  if (!nnet_output || !alpha || !beta || !total_logprob || T <= 0 || num_pdfs <= 0) { chain_set_error("chain_forward_backward: bad argument"); return -1; }
  HostFst h;
  if (!download(fst, h, true)) return -1;
  DevArr in_ptr, in_arc, in_src, tot;
  if (!in_ptr.up(h.in_ptr.data(), h.in_ptr.size() * 4) || !in_arc.up(h.in_arc.data(), h.in_arc.size() * 4) ||
      !in_src.up(h.in_src.data(), h.in_src.size() * 4) || !tot.up(nullptr, 0)) return -1;
  fst_forward_backward_kernel<<<1, 1024>>>((const __half*)nnet_output, T, num_pdfs, h.S, fst->row_ptr, fst->col_idx, fst->labels, fst->weights,
                                          (const int32_t*)in_ptr.p, (const int32_t*)in_arc.p, (const int32_t*)in_src.p, fst->final_states,
                                          fst->final_weights, h.F, h.start, alpha, beta, (float*)tot.p);
  count_launch();
  if (!ck(cudaGetLastError(), "chain_forward_backward launch")) return -1;
  return ck(cudaMemcpy(total_logprob, tot.p, sizeof(float), cudaMemcpyDeviceToHost), "chain_forward_backward") ? 0 : -1;
}

int chain_compute_posteriors(const void* nnet_output, const ChainFstGPU* fst, int T, int num_pdfs, const float* alpha, const float* beta,
                             float total_logprob, float* posteriors) {
  if (!nnet_output || !alpha || !beta || !posteriors || T <= 0 || num_pdfs <= 0) { chain_set_error("chain_compute_posteriors: bad argument"); return -1; }
  HostFst h;
  if (!download(fst, h, false)) return -1;
  std::vector<int32_t> src(h.A);
  for (int s = 0; s < h.S; ++s)
    for (int a = h.row_ptr[s]; a < h.row_ptr[s + 1]; ++a) src[a] = s;
  DevArr arc_src;
  if (!arc_src.up(src.data(), src.size() * 4)) return -1;
  if (!ck(cudaMemset(posteriors, 0, (size_t)T * num_pdfs * sizeof(float)), "posteriors clear")) return -1;
  const long long work = (long long)T * h.A;
  if (work > 0) {
    fst_posteriors_kernel<<<(unsigned)((work + 255) / 256), 256>>>(posteriors, alpha, beta, (const __half*)nnet_output, (const int32_t*)arc_src.p,
                                                                   fst->col_idx, fst->labels, fst->weights, h.S, num_pdfs, h.A, T, total_logprob);
    count_launch();
  }
  return ck(cudaDeviceSynchronize(), "chain_compute_posteriors") ? 0 : -1;
}

int chain_compute_loss(const void* nnet_output, const ChainFstGPU* num_fst, const ChainFstGPU* den_fst, int T, int num_pdfs, void* grad_output,
                       ChainLossResult* result) {
  if (!nnet_output || !result || T <= 0 || num_pdfs <= 0) { chain_set_error("chain_compute_loss: bad argument"); return -1; }
  HostFst hn, hd;
  if (!download(num_fst, hn, false) || !download(den_fst, hd, false)) return -1;
  kfp16_ctx* ctx = compat_ctx();
  if (!ctx) { chain_set_error("chain_compute_loss: no sm_100 device context (%s)", kfp16_last_error() ? kfp16_last_error() : "?"); return -1; }
  auto as_fst = [](const HostFst& h) {
    kfp16_chain_fst f;
    f.row_ptr = h.row_ptr.data(); f.col_idx = h.col_idx.data(); f.labels = h.labels.data(); f.weights = h.weights.data();
    f.final_states = h.final_states.data(); f.final_weights = h.final_weights.data();
    f.num_states = h.S; f.num_arcs = h.A; f.num_final = h.F; f.start_state = h.start;
    return f;
  };
  const kfp16_chain_fst den = as_fst(hd), num = as_fst(hn);
  kfp16_chain* c = kfp16_chain_create(ctx, num_pdfs, 1, T, &den);
  if (!c) { chain_set_error("chain_compute_loss: %s", kfp16_last_error() ? kfp16_last_error() : "kfp16_chain_create failed"); return -1; }
  int rc = kfp16_chain_set_numerators(c, &num, 1);
  // the reference's gradient buffer is dense [T x num_pdfs]; rows = frames, no subsampling, supervision weight 1
  if (!rc) rc = kfp16_chain_loss(c, nnet_output, grad_output, num_pdfs, T, 0, 1, 1.0f, nullptr);
  float res[4] = {0, 0, 0, 0};
  if (!rc) rc = kfp16_chain_read_results(c, res, 1);
  if (rc) chain_set_error("chain_compute_loss: %s", kfp16_last_error() ? kfp16_last_error() : "failed");
  kfp16_chain_destroy(c);
  if (rc) return -1;
  result->num_logprob = res[0]; result->den_logprob = res[1]; result->loss = res[2];
  return 0;
}

}  // extern "C"
