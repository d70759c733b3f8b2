// Network executor: xconfig -> fixed launch plan of fused tcgen05 GEMMs (include/kaldi_fp16_nnet.h).
//
// Replaces the per-layer op sequences the reference issues from Go
// (/root/reference/internal/nnet/forward.go:148-1001, network_backward.go:94-656,
//  train_step.go:142-283, model.go, layers.go, xconfig.go).  Layer semantics follow SURVEY.md
// Appendix A; the backward pass is the exact transpose of the forward (the reference's backward
// ignores splicing -- quirk Q2 -- and is only meaningful for time-stride 0, where both agree).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/kaldi_fp16_nnet.h"
#include "host_common.h"

using namespace kfp16;

namespace {

enum LType { L_INPUT, L_IDCT, L_LINEAR, L_BATCHNORM, L_SPECAUG, L_COMBINE, L_CONV, L_TDNNF, L_PREFINAL, L_OUTPUT, L_ATTENTION };
enum HaloMode { HALO_NONE = 0, HALO_ZERO = 1, HALO_REPL = 2 };

struct Buf {
  __half* p = nullptr;
  int rows = 0, cols = 0;
  int ld = 0;                   // elements between rows; 0 = cols (dense).  Only row-subsampled VIEWS set it.
  size_t bytes() const { return (size_t)rows * cols * sizeof(__half); }
  int LD() const { return ld ? ld : cols; }
};

struct BNorm {
  bool present = false;
  int dim = 0;
  float eps = 0.001f;          // forward.go:1185, weight_loader.go:447-450
  float target_rms = 1.0f;
  bool rms_only = false;       // batchnorm-component with target-rms != 1 (forward.go:349-374)
  float *mean = nullptr, *var = nullptr, *gamma = nullptr, *beta = nullptr;   // fp32 [dim]
  float *scale = nullptr, *shift = nullptr, *zero = nullptr;                  // folded; zero = [dim] of 0
  float *scale_bwd = nullptr;   // scale * bwd_mul: the backward pass's factor (dropout folds 1/(1-p) in here)
  float *one = nullptr;         // [dim] of 1: identity scale for the producing epilogue when the batch-norm runs in train mode
  float bwd_mul = 1.0f;
  // conv layers (per-filter batch-norm): scale repeated for every output height, and as many zeros -- the vectors a consumer
  // that sees the layer output as [frames x heights*filters] passes to its input-gradient epilogue
  float *scale_tiled = nullptr, *zero_tiled = nullptr;
  int tiles = 0;
};

struct Param {
  std::string name;
  int rows = 0, cols = 0;
  size_t off = 0;
  bool is_bias = false;
};

struct Layer {
  LType type = L_INPUT;
  std::string name, type_name;
  std::map<std::string, std::string> kv;
  std::vector<int> in;          // producer layer indices
  bool in_replace = false;      // ReplaceIndex(x, t, 0): consumes a per-sequence input
  int in_dim = 0, out_dim = 0;
  bool per_seq = false;         // rows = n_seq (ivector branch) instead of padded frames
  bool needs_grad = false;      // lies on a path from a parameter to the chain output
  bool wants_dx = false;        // some producer needs the gradient wrt this layer's input
  int halo_mode = HALO_NONE;    // how consumers expect the halo rows of `out`
  int n_grad_consumers = 0;
  Buf in_cat, d_in_cat;         // Append(...) materialised
  Buf out, dout;                // activation / gradient wrt activation
  Buf tmp_dx;                   // scratch for accumulating into a multi-consumer producer
  // parameters (index into Net::params, -1 = none)
  int pW = -1, pB = -1, pLin = -1, pAff = -1, pAffB = -1, pBig = -1, pBigB = -1, pSmall = -1;
  BNorm bn, bn2;
  // tdnnf
  int stride = 0, bott_dim = 0;
  float dropout_p = 0.f;        // dropout-proportion (training only): inverted dropout after the batch-norm
  float bypass = 0.f;
  bool use_bypass = false;
  Buf bott, dbott, dz;
  uint32_t* mask = nullptr;
  int mask_ld = 0;
  // prefinal
  int big_dim = 0, small_dim = 0;
  Buf big, dbig, dys;
  // idct
  __half* idct_mat = nullptr;
  double lifter = 22.0;
  // output
  bool log_softmax = false;
  // combine-feature-maps
  int height = 0, nf1 = 1, nf2 = 1;
  // input prefetch (double-buffered async H2D): staged dense rows + events
  __half* pf_buf[2] = {nullptr, nullptr};
  cudaEvent_t pf_copied[2] = {nullptr, nullptr}, pf_packed[2] = {nullptr, nullptr};
  int pf_rows = 0, pf_cols = 0, pf_slot = 0, pf_ready = -1;
  bool pf_f32[2] = {false, false};
  size_t pf_cap[2] = {0, 0};
  // conv-relu-batchnorm
  int hin = 0, hout = 0, hsub = 1, fin = 0, fout = 0, convK = 0, convKp = 0;
  __half* convP = nullptr;   // this layer's patch matrix: kept from the forward pass for the weight gradient (train mode)
  std::vector<int> tap_dt, tap_dh;
  bool conv_implicit = false;   // implicit GEMM over 4-D TMA boxes (no patch matrix); else im2col + GEMM + col2im
  int grads_seen = 0;
  // rows of `dout` that carry the gradient in the current backward pass: row g_row0 + k*g_sub (k >= 0).  g_sub > 1 after
  // an objective that is evaluated on subsampled output frames (chain, frame-subsampling-factor 3): the other rows are
  // zero by definition and are neither written nor read by the row-wise layers behind the output
  int g_sub = 1, g_row0 = 0;
  // conv layers: the consumer's input-gradient epilogue already applied this layer's batch-norm scale and ReLU mask, so
  // `dout` holds dZ (set per backward pass by the consumer, cleared when the pass begins)
  bool dz_in_dout = false;
  // training step with a subsampled objective: this layer's output is only read on rows f_row0 + k*f_sub (it is row-wise
  // and feeds nothing but the objective's output layer through row-wise layers), so the step computes just those rows
  int f_sub = 1, f_row0 = 0;
  // attention-relu-batchnorm-layer (internal/nnet/layers.go:298-318): heads, key / value dims, context; `big` holds the projection
  // [rows x att_affine], `dz` the ReLU'd pre-batch-norm output, `dbig` the projection gradient, att_db the score gradients
  int att_heads = 0, att_key = 0, att_value = 0, att_left = 0, att_right = 0, att_stride = 1, att_affine = 0;
  float att_scale = 1.0f;
  float* att_db = nullptr;
  // spec-augment-layer (training, kfp16_net_set_spec_augment): mask geometry derived from the xconfig values
  int sa_fmax = 0, sa_nfreq = 0, sa_tmax = 0, sa_ntime = 0;
  double fl_fwd = 0, fl_bwd = 0;   // GEMM flops of this layer over all rows (row-wise layer types only)
};

}  // namespace

struct WgradArgs { const Buf* X; const Buf* dY; int param, groups, off0, off1; };

struct kfp16_net {
  kfp16_ctx* ctx = nullptr;
  kfp16_net_opts opts{};
  int halo = 0, blk = 0, Tp = 0, T = 0;
  std::vector<Layer> layers;
  std::vector<Param> params;
  size_t bucket = 0;
  __half* w16 = nullptr;
  float *w32 = nullptr, *vel = nullptr, *g32 = nullptr;
  __half* g16 = nullptr;                   // FP16 copy of the (scaled) gradient bucket: what the data-parallel exchange carries
  uint32_t* seed_dev = nullptr;            // per-step dropout seed word (bumped at the start of every training step)
  float* hp_dev = nullptr;                 // {lr, momentum, grad_scale}: read by the SGD kernel at run time (graph-safe SetLR)
  float hp_host[3] = {0.f, 0.f, 1.f};
  double flops_bwd = 0;
  float* loss_dev = nullptr;
  float* loss_pinned = nullptr;            // 2 pinned slots for kfp16_net_read_loss_async / kfp16_net_wait_loss
  cudaEvent_t loss_ev[2] = {nullptr, nullptr};
  std::vector<void*> allocs;
  __half *conv_P = nullptr, *conv_dP = nullptr, *conv_dz = nullptr;   // shared conv scratch (patches, patch grads, dZ)
  size_t conv_P_elems = 0, conv_dz_elems = 0;
  cudaStream_t copy_stream = nullptr;   // H2D prefetch of the next minibatch (kfp16_net_prefetch_input)
  cudaStream_t copy_extra[3] = {nullptr, nullptr, nullptr};   // large inputs are copied in 4 parts on 4 streams (one DMA
  cudaEvent_t copy_part[3] = {nullptr, nullptr, nullptr};      // engine each: a single stream does not saturate the host link)
  __half* stage_in = nullptr;   // dense staging for host uploads / downloads
  size_t stage_bytes = 0;
  double flops_fwd = 0;
  std::map<int, cudaGraphExec_t> graph;     // keyed by the phases bitmask
  // the step graph cut into segments along the backward pass (gradient all-reduce overlapped bucket by bucket)
  // spliced weight gradients deferred to one grouped launch per backward pass / segment (kfp16_wgrad_group)
  std::vector<WgradArgs> deferred;
  std::map<std::vector<int>, kfp16_wgrad_group*> wg_groups;
  std::vector<cudaGraphExec_t> seg_graph;
  std::vector<int> seg_launches, seg_lo;            // per segment: kernels, first layer index it back-propagates
  std::vector<size_t> seg_off, seg_cnt;              // gradient-bucket range completed by the segment (elements)
  bool seg_export_f16 = false;                       // each segment ends by exporting its gradient range to the FP16 bucket
  std::map<int, int> graph_launches;
  int out_layer = -1;
  // objective of the captured step: the chain LF-MMI loss when set, else 0.5*||out||^2
  kfp16_chain* chain = nullptr;
  int chain_sub = 3, chain_left = 0;
  float chain_weight = 1.0f;
  bool sparse_out_grad = true;          // kfp16_net_set_sparse_output_grad
  double flops_fwd_skipped = 0, flops_bwd_skipped = 0;   // of flops_fwd / flops_bwd, not executed by the last training step (rows outside the objective's frames)
  // the objective runs on a second stream beside the forward pass of the layers it does not depend on (xent branch)
  // train-mode batch-norm (kfp16_net_set_train_batchnorm): batch statistics instead of the stored running statistics
  bool train_bn = false;
  float bn_momentum = 0.1f;
  int bn_world = 1;                     // ranks whose statistics the hook sums (global row count = local rows * bn_world)
  kfp16_bn_stats_hook bn_hook = nullptr;
  void* bn_hook_user = nullptr;
  float* bn_stats = nullptr;            // [2 x widest batch-norm] fp32 scratch
  size_t bn_stats_dim = 0;
  bool spec_augment = false;            // kfp16_net_set_spec_augment: spec-augment-layer masks in training (off = the reference's pass-through)
  bool overlap_loss = true;             // kfp16_net_set_overlap_loss
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool fwd_rows_now = false;            // inside the training step: honour Layer::f_sub in forward_layer
  bool fuse_conv_bwd = true;            // conv producers get dZ from their consumer's input-gradient epilogue (KFP16_FUSE_CONV_BWD=0: off)
  int out_g_sub = 1, out_g_row0 = 0;   // gradient rows the last objective call wrote on the output layer (Layer::g_sub)
};

namespace {

// ------------------------------------------------------------------------------ utilities
std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

// split on whitespace outside parentheses: "input=Append(a, b)" stays one token
std::vector<std::string> tokens(const std::string& line) {
  std::vector<std::string> out;
  std::string cur;
  int depth = 0;
  for (char c : line) {
    if (c == '(') ++depth;
    if (c == ')') --depth;
    if ((c == ' ' || c == '\t') && depth == 0) {
      if (!cur.empty()) { out.push_back(cur); cur.clear(); }
    } else {
      cur.push_back(c);
    }
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

int kv_int(const Layer& l, const char* k, int def) {
  auto it = l.kv.find(k);
  if (it == l.kv.end()) return def;
  char* e = nullptr;
  long v = strtol(it->second.c_str(), &e, 10);
  return (e && *e == 0) ? (int)v : def;
}
double kv_float(const Layer& l, const char* k, double def) {
  auto it = l.kv.find(k);
  if (it == l.kv.end()) return def;
  char* e = nullptr;
  double v = strtod(it->second.c_str(), &e);
  return (e && *e == 0) ? v : def;
}
bool kv_bool(const Layer& l, const char* k, bool def) {
  auto it = l.kv.find(k);
  if (it == l.kv.end()) return def;
  std::string v = it->second;
  std::transform(v.begin(), v.end(), v.begin(), ::tolower);
  if (v == "true" || v == "1" || v == "yes") return true;
  if (v == "false" || v == "0" || v == "no") return false;
  return def;
}
std::vector<int> kv_ints(const Layer& l, const char* k) {
  std::vector<int> out;
  auto it = l.kv.find(k);
  if (it == l.kv.end()) return out;
  std::stringstream ss(it->second);
  std::string part;
  while (std::getline(ss, part, ',')) {
    part = trim(part);
    if (!part.empty()) out.push_back(atoi(part.c_str()));
  }
  return out;
}

// internal/gpu/tensor.go:158-174 float32ToFP16Bits: truncating, flush-to-zero, exp>15 -> Inf
uint16_t f32_to_f16_trunc(float f) {
  uint32_t bits;
  memcpy(&bits, &f, 4);
  const uint16_t sign = (uint16_t)((bits >> 16) & 0x8000);
  const int exp = (int)((bits >> 23) & 0xFF) - 127;
  const uint32_t frac = bits & 0x7FFFFF;
  if (exp > 15) return sign | 0x7C00;
  if (exp < -14) return sign;
  return (uint16_t)(sign | (uint16_t)((exp + 15) << 10) | (uint16_t)(frac >> 13));
}
float f16_bits_to_f32(uint16_t h) {
  __half_raw r;
  r.x = h;
  return __half2float(__half(r));
}

bool dev_alloc(kfp16_net* n, void** p, size_t bytes, bool zero = true) {
  if (bytes == 0) bytes = 16;
  if (!check_cuda(cudaMalloc(p, bytes), "cudaMalloc (network buffer)")) return false;
  n->allocs.push_back(*p);
  if (zero && !check_cuda(cudaMemsetAsync(*p, 0, bytes, n->ctx->stream), "cudaMemset")) return false;
  return true;
}
bool alloc_buf(kfp16_net* n, Buf& b, int rows, int cols) {
  b.rows = rows;
  b.cols = cols;
  return dev_alloc(n, (void**)&b.p, b.bytes());
}

int find_layer(const kfp16_net* n, const std::string& name) {
  for (size_t i = 0; i < n->layers.size(); ++i)
    if (n->layers[i].name == name) return (int)i;
  // "tdnnf7" also resolves "tdnnf7.xyz" style sub-names to the latest match (layers.go:357-370)
  int best = -1;
  for (size_t i = 0; i < n->layers.size(); ++i) {
    const std::string& ln = n->layers[i].name;
    if (ln.size() > name.size() && ln.compare(0, name.size(), name) == 0 && ln[name.size()] == '.') best = (int)i;
  }
  return best;
}
int find_param(const kfp16_net* n, const std::string& name) {
  for (size_t i = 0; i < n->params.size(); ++i)
    if (n->params[i].name == name) return (int)i;
  return -1;
}

int add_param(kfp16_net* n, const std::string& name, int rows, int cols, bool is_bias = false) {
  Param p;
  p.name = name;
  p.rows = rows;
  p.cols = cols;
  p.is_bias = is_bias;
  p.off = n->bucket;
  n->bucket += ((size_t)rows * cols + 127) & ~(size_t)127;   // 256-byte aligned fp16 sections
  n->params.push_back(p);
  return (int)n->params.size() - 1;
}

bool make_bn(kfp16_net* n, BNorm& bn, int dim, float target_rms, bool rms_only) {
  bn.present = true;
  bn.dim = dim;
  bn.target_rms = target_rms;
  bn.rms_only = rms_only;
  float* base = nullptr;
  if (!dev_alloc(n, (void**)&base, (size_t)dim * 9 * sizeof(float))) return false;
  bn.mean = base;
  bn.var = base + dim;
  bn.gamma = base + 2 * dim;
  bn.beta = base + 3 * dim;
  bn.scale = base + 4 * dim;
  bn.shift = base + 5 * dim;
  bn.zero = base + 6 * dim;
  bn.scale_bwd = base + 7 * dim;
  bn.one = base + 8 * dim;
  {
    std::vector<float> ones((size_t)dim, 1.f);
    if (!check_cuda(cudaMemcpyAsync(bn.one, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice, n->ctx->stream), "bn ones upload") ||
        !check_cuda(cudaStreamSynchronize(n->ctx->stream), "bn ones sync")) return false;
  }
  if ((size_t)dim > n->bn_stats_dim) n->bn_stats_dim = (size_t)dim;
  std::vector<float> h((size_t)dim * 4, 0.f);   // identity: mean 0, var 1, gamma 1, beta 0
  for (int i = 0; i < dim; ++i) { h[dim + i] = 1.f; h[2 * dim + i] = 1.f; }
  if (!check_cuda(cudaMemcpyAsync(base, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, n->ctx->stream), "bn upload")) return false;
  if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "bn upload sync")) return false;
  return kfp16_bn_fold(n->ctx, bn.mean, bn.var, rms_only ? nullptr : bn.gamma, rms_only ? nullptr : bn.beta, bn.eps,
                       target_rms, dim, bn.scale, bn.shift) == 0 &&
         kfp16_scale_f32(n->ctx, bn.scale, bn.scale_bwd, dim, bn.bwd_mul) == 0;
}

// conv layers: (re)build the per-height repetition of the folded batch-norm scale
bool retile_bn(kfp16_net* n, BNorm& bn, int tiles) {
  if (tiles <= 0) return true;
  if (!bn.scale_tiled) {
    if (!dev_alloc(n, (void**)&bn.scale_tiled, (size_t)bn.dim * tiles * 2 * sizeof(float))) return false;
    bn.zero_tiled = bn.scale_tiled + (size_t)bn.dim * tiles;
    bn.tiles = tiles;
    if (!check_cuda(cudaMemsetAsync(bn.zero_tiled, 0, (size_t)bn.dim * tiles * sizeof(float), n->ctx->stream), "bn tiled zeros")) return false;
  }
  for (int t = 0; t < bn.tiles; ++t)
    if (!check_cuda(cudaMemcpyAsync(bn.scale_tiled + (size_t)t * bn.dim, bn.scale, (size_t)bn.dim * sizeof(float), cudaMemcpyDeviceToDevice, n->ctx->stream), "bn tiled scale")) return false;
  return true;
}

// ------------------------------------------------------------------------------ xconfig
bool parse_input_spec(kfp16_net* n, Layer& l, int idx) {
  auto it = l.kv.find("input");
  if (it == l.kv.end() || it->second.empty()) {
    if (l.type == L_INPUT) return true;
    if (idx == 0) { set_error("layer %s has no input", l.name.c_str()); return false; }
    l.in.push_back(idx - 1);
    return true;
  }
  std::string spec = trim(it->second);
  auto resolve = [&](const std::string& nm) -> int {
    int r = find_layer(n, trim(nm));
    if (r < 0) set_error("layer %s: input \"%s\" not found", l.name.c_str(), trim(nm).c_str());
    return r;
  };
  if (spec.compare(0, 7, "Append(") == 0 && spec.back() == ')') {
    std::string inner = spec.substr(7, spec.size() - 8);
    // split on top-level commas
    int depth = 0;
    std::string cur;
    std::vector<std::string> parts;
    for (char c : inner) {
      if (c == '(') ++depth;
      if (c == ')') --depth;
      if (c == ',' && depth == 0) { parts.push_back(cur); cur.clear(); } else cur.push_back(c);
    }
    if (!trim(cur).empty()) parts.push_back(cur);
    for (auto& p : parts) {
      std::string nm = trim(p);
      if (nm.compare(0, 13, "ReplaceIndex(") == 0) nm = trim(nm.substr(13, nm.find(',') - 13));
      int r = resolve(nm);
      if (r < 0) return false;
      l.in.push_back(r);
    }
    return true;
  }
  if (spec.compare(0, 13, "ReplaceIndex(") == 0) {
    std::string nm = trim(spec.substr(13, spec.find(',') - 13));
    int r = resolve(nm);
    if (r < 0) return false;
    l.in.push_back(r);
    l.in_replace = true;
    return true;
  }
  int r = resolve(spec);
  if (r < 0) return false;
  l.in.push_back(r);
  return true;
}

bool parse_xconfig(kfp16_net* n, const char* text) {
  static const std::map<std::string, LType> types = {
      {"input", L_INPUT}, {"idct-layer", L_IDCT}, {"linear-component", L_LINEAR},
      {"batchnorm-component", L_BATCHNORM}, {"spec-augment-layer", L_SPECAUG},
      {"combine-feature-maps-layer", L_COMBINE}, {"conv-relu-batchnorm-layer", L_CONV},
      {"tdnnf-layer", L_TDNNF}, {"prefinal-layer", L_PREFINAL}, {"output-layer", L_OUTPUT},
      {"attention-relu-batchnorm-layer", L_ATTENTION}};
  std::stringstream ss(text);
  std::string line;
  int lineno = 0;
  while (std::getline(ss, line)) {
    ++lineno;
    size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    line = trim(line);
    if (line.empty()) continue;
    std::vector<std::string> tk = tokens(line);
    Layer l;
    l.type_name = tk[0];
    auto tt = types.find(tk[0]);
    if (tt == types.end()) {
      set_error("xconfig line %d: unsupported layer type \"%s\"", lineno, tk[0].c_str());
      return false;
    }
    l.type = tt->second;
    for (size_t i = 1; i < tk.size(); ++i) {
      size_t eq = tk[i].find('=');
      if (eq == std::string::npos) continue;
      l.kv[tk[i].substr(0, eq)] = tk[i].substr(eq + 1);
    }
    l.name = l.kv.count("name") ? l.kv["name"] : "";
    if (l.name.empty()) { set_error("xconfig line %d: layer without name", lineno); return false; }
    if (find_layer(n, l.name) >= 0 && n->layers[find_layer(n, l.name)].name == l.name) {
      set_error("xconfig line %d: duplicate layer name %s", lineno, l.name.c_str());
      return false;
    }
    if (!parse_input_spec(n, l, (int)n->layers.size())) return false;
    n->layers.push_back(l);
  }
  if (n->layers.empty()) { set_error("xconfig: no layers"); return false; }
  return true;
}

// ------------------------------------------------------------------------------ plan
bool resolve_dims(kfp16_net* n) {
  // per-sequence inputs: consumed only through ReplaceIndex(x, t, 0) (the ivector, quirk Q6)
  for (size_t i = 0; i < n->layers.size(); ++i) {
    Layer& l = n->layers[i];
    if (l.type != L_INPUT) continue;
    bool any = false, all_replace = true;
    for (auto& c : n->layers)
      for (int src : c.in)
        if (src == (int)i) { any = true; if (!c.in_replace) all_replace = false; }
    l.per_seq = any && all_replace;
  }
  int max_halo = 0;
  for (size_t i = 0; i < n->layers.size(); ++i) {
    Layer& l = n->layers[i];
    int in_dim = 0;
    bool all_seq = !l.in.empty();
    for (int src : l.in) { in_dim += n->layers[src].out_dim; all_seq = all_seq && n->layers[src].per_seq; }
    l.in_dim = in_dim;
    if (l.type != L_INPUT) l.per_seq = all_seq;
    switch (l.type) {
      case L_INPUT:
        l.out_dim = kv_int(l, "dim", 0);
        if (l.out_dim <= 0) { set_error("input layer %s missing dim", l.name.c_str()); return false; }
        l.in_dim = l.out_dim;
        break;
      case L_IDCT:
        l.out_dim = kv_int(l, "dim", l.in_dim);
        l.lifter = kv_float(l, "cepstral-lifter", 22.0);
        if (l.out_dim != l.in_dim) { set_error("idct-layer %s: dim %d != input dim %d", l.name.c_str(), l.out_dim, l.in_dim); return false; }
        break;
      case L_LINEAR:
        l.out_dim = kv_int(l, "dim", 0);
        if (l.out_dim <= 0) { set_error("linear-component %s missing dim", l.name.c_str()); return false; }
        break;
      case L_BATCHNORM:
        l.out_dim = l.in_dim;
        break;
      case L_SPECAUG: {   // internal/nnet/layers.go:231-240: freq-max-proportion 0.5, time-zeroed-proportion 0, time-mask-max-frames 20
        l.out_dim = l.in_dim;
        const double fprop = kv_float(l, "freq-max-proportion", 0.5), tprop = kv_float(l, "time-zeroed-proportion", 0.0);
        const int tmax = std::min(kv_int(l, "time-mask-max-frames", 20), n->opts.seq_len);
        l.sa_fmax = std::max(0, std::min(l.in_dim, (int)(fprop * l.in_dim)));
        l.sa_nfreq = l.sa_fmax > 0 ? 1 : 0;
        l.sa_tmax = std::max(0, tmax);
        // masks of mean width tmax/2 until the requested proportion of the frames is zeroed on average
        l.sa_ntime = (tprop > 0 && tmax > 0) ? std::min(8, std::max(1, (int)(tprop * n->opts.seq_len / (0.5 * tmax) + 0.5))) : 0;
        break;
      }
      case L_COMBINE:
        l.out_dim = l.in_dim;
        l.height = kv_int(l, "height", 0);
        l.nf1 = kv_int(l, "num-filters1", 1);
        l.nf2 = kv_int(l, "num-filters2", 1);
        if (l.height * (l.nf1 + l.nf2) != l.in_dim) {
          set_error("combine-feature-maps-layer %s: height*(nf1+nf2)=%d != input dim %d", l.name.c_str(), l.height * (l.nf1 + l.nf2), l.in_dim);
          return false;
        }
        break;
      case L_CONV: {   // forward.go:418-524; dims as layers.go resolves them
        l.hin = kv_int(l, "height-in", 0);
        l.hout = kv_int(l, "height-out", l.hin);
        l.hsub = kv_int(l, "height-subsample-out", 1);
        l.fout = kv_int(l, "num-filters-out", 0);
        if (l.hin <= 0 || l.hout <= 0 || l.hsub <= 0 || l.fout <= 0 || (l.in_dim % l.hin)) {
          set_error("conv-relu-batchnorm-layer %s: needs height-in dividing the input dim (%d), height-out, num-filters-out", l.name.c_str(), l.in_dim);
          return false;
        }
        l.fin = l.in_dim / l.hin;
        std::vector<int> to = kv_ints(l, "time-offsets"), ho = kv_ints(l, "height-offsets");
        if (to.empty()) to.push_back(0);
        if (ho.empty()) ho.push_back(0);
        if (n->opts.conv_cartesian) {     // Kaldi: time x height Cartesian product (quirk Q4 fixed)
          for (int dt : to) for (int dh : ho) { l.tap_dt.push_back(dt); l.tap_dh.push_back(dh); }
        } else {                          // the reference pairs them (forward.go:426-446)
          for (size_t k = 0; k < to.size(); ++k) { l.tap_dt.push_back(to[k]); l.tap_dh.push_back(ho[k < ho.size() ? k : ho.size() - 1]); }
        }
        if (l.tap_dt.size() > 32) { set_error("conv-relu-batchnorm-layer %s: more than 32 taps", l.name.c_str()); return false; }
        for (int dt : l.tap_dt) if (abs(dt) >= n->opts.seq_len) { set_error("conv layer %s: time offset %d exceeds the sequence length", l.name.c_str(), dt); return false; }
        for (int dt : l.tap_dt) max_halo = std::max(max_halo, abs(dt));   // zero-padded halo rows keep the sequences apart
        l.convK = (int)l.tap_dt.size() * l.fin;
        l.convKp = (l.convK + 15) & ~15;
        l.out_dim = l.hout * l.fout;
        break;
      }
      case L_TDNNF:
        l.out_dim = kv_int(l, "dim", 0);
        l.bott_dim = kv_int(l, "bottleneck-dim", 0);
        if (l.out_dim <= 0 || l.bott_dim <= 0) { set_error("tdnnf-layer %s missing dim or bottleneck-dim", l.name.c_str()); return false; }
        l.stride = kv_int(l, "time-stride", 3);
        l.bypass = (float)kv_float(l, "bypass-scale", 0.66);
        l.use_bypass = l.bypass > 0 && l.in_dim == l.out_dim;   // forward.go:688
        l.dropout_p = n->opts.train ? (float)kv_float(l, "dropout-proportion", 0.0) : 0.f;
        if (!(l.dropout_p >= 0.f && l.dropout_p < 1.f)) { set_error("tdnnf-layer %s: dropout-proportion must be in [0, 1)", l.name.c_str()); return false; }
        if (l.stride < 0) { set_error("tdnnf-layer %s: negative time-stride", l.name.c_str()); return false; }
        max_halo = std::max(max_halo, l.stride);
        break;
      case L_PREFINAL:
        l.small_dim = kv_int(l, "small-dim", 0);
        l.big_dim = kv_int(l, "big-dim", 0);
        if (l.small_dim <= 0 || l.big_dim <= 0) { set_error("prefinal-layer %s missing small-dim or big-dim", l.name.c_str()); return false; }
        l.out_dim = l.small_dim;
        break;
      case L_ATTENTION: {   // internal/nnet/layers.go:298-318; key-scale default 1/sqrt(key-dim) (weight_loader.go:267-271)
        l.att_heads = kv_int(l, "num-heads", 1);
        l.att_value = kv_int(l, "value-dim", 0);
        l.att_key = kv_int(l, "key-dim", 0);
        l.att_left = kv_int(l, "num-left-inputs", 0);
        l.att_right = kv_int(l, "num-right-inputs", 0);
        l.att_stride = kv_int(l, "time-stride", 1);
        if (l.att_heads <= 0 || l.att_value <= 0 || l.att_key <= 0 || l.att_left < 0 || l.att_right < 0 || l.att_stride < 1) {
          set_error("attention-relu-batchnorm-layer %s: num-heads, value-dim, key-dim must be positive, time-stride >= 1", l.name.c_str()); return false;
        }
        const int ctx_dim = 1 + l.att_left + l.att_right;
        if (ctx_dim > 32) { set_error("attention-relu-batchnorm-layer %s: at most 32 context positions", l.name.c_str()); return false; }
        l.att_scale = (float)kv_float(l, "key-scale", 1.0 / sqrt((double)l.att_key));
        l.att_affine = l.att_heads * (2 * l.att_key + l.att_value + ctx_dim);
        l.out_dim = l.att_heads * (l.att_value + ctx_dim);
        break;
      }
      case L_OUTPUT:
        l.out_dim = kv_int(l, "dim", 0);
        if (l.out_dim <= 0) { set_error("output-layer %s missing dim", l.name.c_str()); return false; }
        l.log_softmax = kv_bool(l, "include-log-softmax", true);
        break;
    }
    if (l.type != L_INPUT && l.in.empty()) { set_error("layer %s has no input", l.name.c_str()); return false; }
  }
  n->halo = max_halo;
  n->blk = n->opts.seq_len + 2 * n->halo;
  n->Tp = n->opts.n_seq * n->blk;
  n->T = n->opts.n_seq * n->opts.seq_len;
  // halo mode expected by consumers
  for (auto& c : n->layers) {
    if (c.type == L_TDNNF && c.stride > 0)
      for (int src : c.in) n->layers[src].halo_mode = HALO_REPL;
  }
  // Convolutions run as implicit GEMMs (taps = shifted 4-D TMA boxes of the layer input, zero-filled outside it) when the
  // geometry allows: 64-channel multiples on both sides, height subsampling 1 or 2, <= 16 taps, and a producer whose halo
  // rows can be kept at zero (the per-sequence zero padding in time).  Anything else (the first layer's 6 input filters)
  // gathers a patch matrix on the device and runs the same GEMM on it.
  static const bool force_im2col = getenv("KFP16_CONV_IM2COL") && atoi(getenv("KFP16_CONV_IM2COL")) != 0;   // A/B measurements
  for (auto& c : n->layers) {
    if (c.type != L_CONV) continue;
    bool kblock = false;
    for (int tb = 80 / std::max(1, c.hout); tb >= 1; --tb) kblock = kblock || (tb * c.hout) % 16 == 0;
    const bool geom = (c.hsub == 1 ? c.hin == c.hout : (c.hsub == 2 && c.hin == 2 * c.hout));
    Layer& src = n->layers[c.in[0]];
    c.conv_implicit = !force_im2col && c.in.size() == 1 && (c.fin % 64) == 0 && (c.fout % 64) == 0 && geom && c.hout <= 128 && kblock &&
                      c.tap_dt.size() <= 16 && src.halo_mode != HALO_REPL && !src.per_seq;
    if (c.conv_implicit) src.halo_mode = HALO_ZERO;
  }
  // output layer + gradient reachability (Backward seeds only the chain output, network_backward.go:104-107)
  n->out_layer = find_layer(n, "output");
  if (n->out_layer < 0)
    for (size_t i = 0; i < n->layers.size(); ++i)
      if (n->layers[i].type == L_OUTPUT) { n->out_layer = (int)i; break; }
  if (n->out_layer < 0) n->out_layer = (int)n->layers.size() - 1;
  if (n->opts.train) {
    std::vector<int> stack{n->out_layer};
    while (!stack.empty()) {
      int i = stack.back();
      stack.pop_back();
      if (n->layers[i].needs_grad) continue;
      n->layers[i].needs_grad = true;
      for (int src : n->layers[i].in) stack.push_back(src);
    }
    // a layer needs the gradient wrt its input only if something upstream has parameters
    std::vector<bool> has_params_upstream(n->layers.size(), false);
    for (size_t i = 0; i < n->layers.size(); ++i) {
      const Layer& l = n->layers[i];
      bool up = false;
      for (int src : l.in) {
        const Layer& s = n->layers[src];
        bool own = s.type == L_LINEAR || s.type == L_TDNNF || s.type == L_PREFINAL || s.type == L_OUTPUT || s.type == L_CONV || s.type == L_ATTENTION;
        up = up || own || has_params_upstream[src];
      }
      has_params_upstream[i] = up;
      n->layers[i].wants_dx = up;
    }
    for (auto& l : n->layers)
      if (l.needs_grad && l.wants_dx)
        for (int src : l.in) n->layers[src].n_grad_consumers++;
  }
  return true;
}

bool check_tma_dim(const Layer& l, int dim, const char* what) {
  if (dim % 8) { set_error("layer %s: %s = %d must be a multiple of 8 (TMA rows are 16-byte aligned)", l.name.c_str(), what, dim); return false; }
  return true;
}

bool build_plan(kfp16_net* n) {
  const bool train = n->opts.train != 0;
  // 1. parameters
  for (auto& l : n->layers) {
    switch (l.type) {
      case L_LINEAR: l.pW = add_param(n, l.name + ".W", l.in_dim, l.out_dim); break;
      case L_TDNNF: {
        const int sp = l.stride > 0 ? 2 : 1;   // true spliced shapes (quirk Q1)
        l.pLin = add_param(n, l.name + ".LinearW", sp * l.in_dim, l.bott_dim);
        l.pAff = add_param(n, l.name + ".AffineW", sp * l.bott_dim, l.out_dim);
        l.pAffB = add_param(n, l.name + ".AffineBias", 1, l.out_dim, true);
        break;
      }
      case L_PREFINAL:
        l.pBig = add_param(n, l.name + ".BigW", l.in_dim, l.big_dim);
        l.pBigB = add_param(n, l.name + ".BigBias", 1, l.big_dim, true);
        l.pSmall = add_param(n, l.name + ".SmallW", l.big_dim, l.small_dim);
        break;
      case L_OUTPUT:
        l.pW = add_param(n, l.name + ".W", l.in_dim, l.out_dim);
        l.pB = add_param(n, l.name + ".Bias", 1, l.out_dim, true);
        break;
      case L_ATTENTION:   // the projection's true shape [in x heads*(2*key + value + context)] (forward.go:803-812; quirk Q1)
        l.pW = add_param(n, l.name + ".W", l.in_dim, l.att_affine);
        l.pB = add_param(n, l.name + ".Bias", 1, l.att_affine, true);
        break;
      case L_CONV:
        l.pW = add_param(n, l.name + ".W", l.convK, l.fout);
        l.pB = add_param(n, l.name + ".Bias", 1, l.fout, true);
        break;
      default: break;
    }
  }
  if (!dev_alloc(n, (void**)&n->w16, std::max<size_t>(n->bucket, 8) * sizeof(__half))) return false;
  if (train) {
    if (!dev_alloc(n, (void**)&n->w32, std::max<size_t>(n->bucket, 8) * sizeof(float))) return false;
    if (!dev_alloc(n, (void**)&n->vel, std::max<size_t>(n->bucket, 8) * sizeof(float))) return false;
    if (!dev_alloc(n, (void**)&n->g32, std::max<size_t>(n->bucket, 8) * sizeof(float))) return false;
    if (!dev_alloc(n, (void**)&n->g16, std::max<size_t>(n->bucket, 8) * sizeof(__half))) return false;
    if (!dev_alloc(n, (void**)&n->hp_dev, 64)) return false;
    if (!dev_alloc(n, (void**)&n->seed_dev, 64)) return false;
    n->hp_host[0] = n->opts.lr; n->hp_host[1] = n->opts.momentum;
    n->hp_host[2] = n->opts.grad_scale != 0.f ? n->opts.grad_scale : 1.0f;
    {
      float h[8] = {n->hp_host[0], n->hp_host[1], n->hp_host[2], 0.f, n->hp_host[0], n->hp_host[1], 1.0f, 0.f};
      if (!check_cuda(cudaMemcpyAsync(n->hp_dev, h, sizeof(h), cudaMemcpyHostToDevice, n->ctx->stream), "hyper-parameter upload") ||
          !check_cuda(cudaStreamSynchronize(n->ctx->stream), "hyper-parameter upload sync")) return false;
    }
  }
  if (!dev_alloc(n, (void**)&n->loss_dev, 256)) return false;

  // 2. activations
  size_t max_dense = 16;
  for (auto& l : n->layers) {
    const int rows = l.per_seq ? n->opts.n_seq : n->Tp;
    // input layers may have any dim (the 100-dim ivector): their rows are stored 8-column aligned with zero
    // padding and a linear consumer reads the padded width (the weight rows past in_dim are TMA-out-of-bounds zeros)
    const int store_dim = l.type == L_INPUT ? ((l.out_dim + 7) & ~7) : l.out_dim;
    if (l.type != L_INPUT && !check_tma_dim(l, l.out_dim, "output dim")) return false;
    if (l.type == L_INPUT && store_dim != l.out_dim) {
      for (auto& c : n->layers)
        for (int src : c.in)
          if (&n->layers[src] == &l && (c.in.size() != 1 || c.type != L_LINEAR)) {
            set_error("input %s: dim %d is not a multiple of 8; only a direct linear-component consumer is supported", l.name.c_str(), l.out_dim);
            return false;
          }
    }
    if (l.in.size() > 1) {
      if (!alloc_buf(n, l.in_cat, rows, l.in_dim)) return false;
      if (train && l.needs_grad && l.wants_dx && !alloc_buf(n, l.d_in_cat, rows, l.in_dim)) return false;
    }
    if (!alloc_buf(n, l.out, rows, store_dim)) return false;
    max_dense = std::max(max_dense, (size_t)n->T * l.out_dim * (l.type == L_INPUT ? sizeof(float) : sizeof(__half)));   // FP32 feature uploads
    if (train && l.needs_grad) {
      if (!alloc_buf(n, l.dout, rows, l.out_dim)) return false;
      if (l.wants_dx && !alloc_buf(n, l.tmp_dx, rows, l.in_dim)) return false;
    }
    const double M = l.per_seq ? n->opts.n_seq : n->T;
    switch (l.type) {
      case L_IDCT: {
        // makeIDCTMatrix (forward.go:1190-1210), through the truncating converter
        const int D = l.out_dim;
        std::vector<uint16_t> h((size_t)D * D);
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            double v = cos(M_PI * j * (i + 0.5) / D) * (j == 0 ? sqrt(1.0 / D) : sqrt(2.0 / D));
            if (l.lifter > 0 && j > 0) v *= 1.0 + (l.lifter / 2.0) * sin(M_PI * j / l.lifter);
            h[(size_t)i * D + j] = f32_to_f16_trunc((float)v);
          }
        if (!dev_alloc(n, (void**)&l.idct_mat, h.size() * 2)) return false;
        if (!check_cuda(cudaMemcpy(l.idct_mat, h.data(), h.size() * 2, cudaMemcpyHostToDevice), "idct upload")) return false;
        n->flops_fwd += 2.0 * M * D * D;
        if ((train && l.needs_grad) && l.wants_dx) n->flops_bwd += 2.0 * M * D * D;
        break;
      }
      case L_LINEAR:
        l.fl_fwd = 2.0 * M * l.in_dim * l.out_dim;
        if (train && l.needs_grad) l.fl_bwd = (l.wants_dx ? 2 : 1) * 2.0 * M * l.in_dim * l.out_dim;
        n->flops_fwd += l.fl_fwd; n->flops_bwd += l.fl_bwd;
        break;
      case L_BATCHNORM: {
        const float rms = (float)kv_float(l, "target-rms", 1.0);
        if (!make_bn(n, l.bn, l.in_dim, rms, rms != 1.0f)) return false;
        break;
      }
      case L_TDNNF: {
        if (!check_tma_dim(l, l.in_dim, "input dim") || !check_tma_dim(l, l.bott_dim, "bottleneck-dim")) return false;
        if (l.stride > 0 && ((l.in_dim % 16) || (l.bott_dim % 16))) {
          set_error("tdnnf-layer %s: spliced dims must be multiples of 16 (in %d, bottleneck %d)", l.name.c_str(), l.in_dim, l.bott_dim);
          return false;
        }
        if (l.per_seq) { set_error("tdnnf-layer %s on a per-sequence input", l.name.c_str()); return false; }
        const int sp = l.stride > 0 ? 2 : 1;
        if (!alloc_buf(n, l.bott, n->Tp, l.bott_dim)) return false;
        l.bn.bwd_mul = l.dropout_p > 0.f ? 1.0f / (1.0f - l.dropout_p) : 1.0f;
        if (!make_bn(n, l.bn, l.out_dim, 1.0f, false)) return false;
        l.mask_ld = (l.out_dim + 31) / 32;
        if (!dev_alloc(n, (void**)&l.mask, (size_t)n->Tp * l.mask_ld * 4)) return false;
        if (train && l.needs_grad) {
          if (!alloc_buf(n, l.dbott, n->Tp, l.bott_dim)) return false;
          if (!alloc_buf(n, l.dz, n->Tp, l.out_dim)) return false;
        }
        n->flops_fwd += 2.0 * M * (sp * l.in_dim) * l.bott_dim + 2.0 * M * (sp * l.bott_dim) * l.out_dim;
        if (train && l.needs_grad) n->flops_bwd += (l.wants_dx ? 2 : 1) * 2.0 * M * (sp * l.in_dim) * l.bott_dim + 2 * 2.0 * M * (sp * l.bott_dim) * l.out_dim;
        break;
      }
      case L_ATTENTION: {
        if (l.per_seq) { set_error("attention layer %s on a per-sequence input", l.name.c_str()); return false; }
        if (!check_tma_dim(l, l.in_dim, "input dim") || !check_tma_dim(l, l.att_affine, "heads*(2*key-dim + value-dim + context)") ||
            !check_tma_dim(l, l.out_dim, "heads*(value-dim + context)")) return false;
        if (!alloc_buf(n, l.big, n->Tp, l.att_affine)) return false;         // projection
        if (!alloc_buf(n, l.dz, n->Tp, l.out_dim)) return false;             // ReLU'd pre-batch-norm output (mask + attention weights)
        if (!make_bn(n, l.bn, l.out_dim, 1.0f, false)) return false;
        if (train && l.needs_grad) {
          if (!alloc_buf(n, l.dbig, n->Tp, l.att_affine)) return false;
          if (!dev_alloc(n, (void**)&l.att_db, (size_t)n->Tp * l.att_heads * 32 * sizeof(float))) return false;
        }
        l.fl_fwd = 2.0 * M * l.in_dim * l.att_affine;
        if (train && l.needs_grad) l.fl_bwd = (l.wants_dx ? 2 : 1) * 2.0 * M * l.in_dim * l.att_affine;
        n->flops_fwd += l.fl_fwd; n->flops_bwd += l.fl_bwd;
        break;
      }
      case L_PREFINAL: {
        if (!check_tma_dim(l, l.in_dim, "input dim") || !check_tma_dim(l, l.big_dim, "big-dim")) return false;
        const int rows2 = l.per_seq ? n->opts.n_seq : n->Tp;
        if (!alloc_buf(n, l.big, rows2, l.big_dim)) return false;
        if (!make_bn(n, l.bn, l.big_dim, 1.0f, false)) return false;
        if (!make_bn(n, l.bn2, l.small_dim, 1.0f, false)) return false;
        l.mask_ld = (l.big_dim + 31) / 32;
        if (!dev_alloc(n, (void**)&l.mask, (size_t)rows2 * l.mask_ld * 4)) return false;
        if (train && l.needs_grad) {
          if (!alloc_buf(n, l.dbig, rows2, l.big_dim)) return false;
          if (!alloc_buf(n, l.dys, rows2, l.small_dim)) return false;
        }
        l.fl_fwd = 2.0 * M * l.in_dim * l.big_dim + 2.0 * M * l.big_dim * l.small_dim;
        if (train && l.needs_grad) l.fl_bwd = (l.wants_dx ? 2 : 1) * 2.0 * M * l.in_dim * l.big_dim + 2 * 2.0 * M * l.big_dim * l.small_dim;
        n->flops_fwd += l.fl_fwd; n->flops_bwd += l.fl_bwd;
        break;
      }
      case L_OUTPUT:
        if (!check_tma_dim(l, l.in_dim, "input dim")) return false;
        l.fl_fwd = 2.0 * M * l.in_dim * l.out_dim;
        if (train && l.needs_grad) l.fl_bwd = (l.wants_dx ? 2 : 1) * 2.0 * M * l.in_dim * l.out_dim;
        n->flops_fwd += l.fl_fwd; n->flops_bwd += l.fl_bwd;
        break;
      case L_CONV: {
        if (l.per_seq) { set_error("conv layer %s on a per-sequence input", l.name.c_str()); return false; }
        if (!check_tma_dim(l, l.fout, "num-filters-out")) return false;
        if (!make_bn(n, l.bn, l.fout, 1.0f, false)) return false;   // per filter
        if (!retile_bn(n, l.bn, l.hout)) return false;
        const size_t mrows = (size_t)n->Tp * l.hout;
        l.mask_ld = (l.fout + 31) / 32;
        if (!dev_alloc(n, (void**)&l.mask, mrows * l.mask_ld * 4)) return false;
        // training keeps every layer's patch matrix resident between forward and backward (1.8 GB for the cnn_tdnn_1a
        // front end: HBM is 180 GB) instead of gathering it a second time; inference shares one scratch buffer
        if (!l.conv_implicit) {
          if (train) { if (!dev_alloc(n, (void**)&l.convP, mrows * l.convKp * 2, false)) return false; }
          n->conv_P_elems = std::max(n->conv_P_elems, mrows * l.convKp);
        }
        n->conv_dz_elems = std::max(n->conv_dz_elems, mrows * l.fout);
        n->flops_fwd += 2.0 * M * l.hout * l.convK * l.fout;
        if (train && l.needs_grad) n->flops_bwd += (l.wants_dx ? 2 : 1) * 2.0 * M * l.hout * l.convK * l.fout;
        break;
      }
      default: break;
    }
  }
  if (n->conv_P_elems) {
    if (!train && !dev_alloc(n, (void**)&n->conv_P, n->conv_P_elems * 2, false)) return false;
    if (train && !dev_alloc(n, (void**)&n->conv_dP, n->conv_P_elems * 2, false)) return false;
  }
  if (train && n->conv_dz_elems && !dev_alloc(n, (void**)&n->conv_dz, n->conv_dz_elems * 2, false)) return false;
  n->stage_bytes = max_dense;
  if (!dev_alloc(n, (void**)&n->stage_in, n->stage_bytes)) return false;
  return check_cuda(cudaStreamSynchronize(n->ctx->stream), "plan sync");
}

// ------------------------------------------------------------------------------ GEMM helpers
kfp16_gemm_desc mk_desc(int M, int N, int K) {
  kfp16_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M; d.N = N; d.K = K;
  d.groups = 1; d.kslabs = 1; d.kslab_len = K;
  d.alpha = 1.0f;
  d.a_major = KFP16_K_MAJOR;
  d.b_major = KFP16_MN_MAJOR;
  return d;
}
void set_A(kfp16_gemm_desc& d, const void* p, int rows, int cols) { d.A.ptr = p; d.A.rows = rows; d.A.cols = cols; d.A.ld = cols; d.A.halo = 0; }
void set_B(kfp16_gemm_desc& d, const void* p, int rows, int cols) { d.B.ptr = p; d.B.rows = rows; d.B.cols = cols; d.B.ld = cols; d.B.halo = 0; }
void set_A(kfp16_gemm_desc& d, const Buf& b) { set_A(d, b.p, b.rows, b.cols); d.A.ld = b.LD(); }
// rows row0 + k*sub of a [rows x cols] buffer as a matrix of its own (sub == 1: the buffer itself)
Buf row_view(const Buf& b, int row0, int sub) {
  if (sub <= 1) return b;
  Buf v;
  v.p = b.p + (size_t)row0 * b.LD();
  v.rows = (b.rows - 1 - row0) / sub + 1;
  v.cols = b.cols;
  v.ld = b.LD() * sub;
  return v;
}

// tap list of a conv layer as implicit-GEMM addressing of its input x[Tp][hin][fin] (height subsampling 2 = parity planes)
void conv_fwd_addr(const kfp16_net* n, const Layer& l, const void* x, int mode, kfp16_conv_addr& c) {
  memset(&c, 0, sizeof(c));
  c.mode = mode; c.x = x;
  c.T = n->Tp; c.P = l.hsub; c.H = l.hin / l.hsub; c.C = l.fin;
  c.rows_h = l.hout;
  c.ntaps = (int)l.tap_dt.size();
  for (int t = 0; t < c.ntaps; ++t) {
    const int dh = l.tap_dh[t];
    const int par = l.hsub == 2 ? ((dh % 2) + 2) % 2 : 0;
    c.dt[t] = l.tap_dt[t];
    c.par[t] = par;
    c.hq[t] = l.hsub == 2 ? (dh - par) / 2 : dh;
    c.brow[t] = t * l.fin;
  }
}

int pick_split_k(const kfp16_net* n, int M, int N, int groups, int K) {
  // work items = tiles * splits should fill ONE wave of CTA pairs (256-row tiles) as evenly as possible:
  // every extra split costs another fp32 red pass over the tile, every partial wave idles SMs
  const int cg = M > 128 ? 2 : 1;
  const int m_tiles = (M + 128 * cg - 1) / (128 * cg);
  const int bn = N <= 64 ? 64 : N <= 128 ? 128 : N <= 160 ? 160 : 256;
  const int n_tiles = (N + bn - 1) / bn;
  const int tiles = m_tiles * n_tiles * groups;
  int sms = n->ctx->num_sms;
  if (n->ctx->max_ctas > 0 && n->ctx->max_ctas < sms) sms = n->ctx->max_ctas;   // SMs left to the compute kernels
  const int units = std::max(1, sms / cg);
  const int kb = (K + 63) / 64;
  int split = units / tiles;                       // largest split that still fits one wave
  if (split < 1) split = 1;
  split = std::min(split, std::max(1, kb / 4));    // keep >= 4 k-blocks per item
  return std::max(split, 2);   // >= 2 selects the fp32 accumulate path into the gradient bucket
}

__half* W16(kfp16_net* n, int p) { return n->w16 + n->params[p].off; }
float* G32(kfp16_net* n, int p) { return n->g32 + n->params[p].off; }

// dW[in x out] (+)= X^T * dY, fp32 into the gradient bucket.  Up to two row-shifted groups (splice).
int wgrad(kfp16_net* n, const Buf& X, const Buf& dY, int param, int groups, int off0, int off1) {
  const int in = n->params[param].rows / groups, out = dY.cols;   // <= X.cols (padded input storage)
  // Orientation: the UMMA M dimension should be the LONG side of dW.  For a narrow `in` (the 160-wide
  // TDNN-F bottleneck, 256-wide prefinal) compute dW^T = dY^T * X instead -- M = out fills whole 256-row pair
  // tiles, N = in is one tile wide -- and accumulate transposed into the same [in x out] bucket section.
  const bool transposed = out > in && in <= 256 && (in % 8) == 0 && X.cols == in;
  kfp16_gemm_desc d = transposed ? mk_desc(out, in, X.rows) : mk_desc(in, out, X.rows);
  d.a_major = KFP16_MN_MAJOR;
  d.groups = groups;
  if (transposed) {
    set_A(d, dY.p, dY.rows, dY.cols);
    set_B(d, X.p, X.rows, X.cols);
    d.A.ld = dY.LD(); d.B.ld = X.LD();
    d.b_row_off[0][0] = off0;
    d.b_row_off[1][0] = off1;
    d.ws_transposed = 1;
  } else {
    set_A(d, X.p, X.rows, X.cols);
    set_B(d, dY.p, dY.rows, dY.cols);
    d.A.ld = X.LD(); d.B.ld = dY.LD();
    d.a_row_off[0][0] = off0;
    d.a_row_off[1][0] = off1;
  }
  d.split_k = pick_split_k(n, d.M, d.N, groups, X.rows);
  d.ws[0] = G32(n, param);
  d.ws[1] = G32(n, param) + (size_t)in * out;
  d.ws_ld = out;
  return kfp16_gemm_ex(n->ctx, &d);
}

// Two weight gradients of the same shape in ONE launch (the affine and the linear half of a TDNN-F layer: both are
// [1536 x 160] x 2 groups reduced over the frames): half the split count, so half the fp32 reduction traffic, and one
// launch less.  Falls back to two calls when the shapes / orientations do not allow it.
static bool wgrad_fill(kfp16_net* n, const WgradArgs& a, kfp16_mat& A, kfp16_mat& B, int a_off[2], int b_off[2], float* ws[2],
                       int& ws_ld, int& transposed, int& M, int& N) {
  const int in = n->params[a.param].rows / a.groups, out = a.dY->cols;
  const bool tr = out > in && in <= 256 && (in % 8) == 0 && a.X->cols == in;
  M = tr ? out : in; N = tr ? in : out;
  const Buf& am = tr ? *a.dY : *a.X;
  const Buf& bm = tr ? *a.X : *a.dY;
  A.ptr = am.p; A.rows = am.rows; A.cols = am.cols; A.ld = am.LD(); A.halo = 0;
  B.ptr = bm.p; B.rows = bm.rows; B.cols = bm.cols; B.ld = bm.LD(); B.halo = 0;
  a_off[0] = tr ? 0 : a.off0; a_off[1] = tr ? 0 : a.off1;
  b_off[0] = tr ? a.off0 : 0; b_off[1] = tr ? a.off1 : 0;
  ws[0] = G32(n, a.param); ws[1] = G32(n, a.param) + (size_t)in * out;
  ws_ld = out; transposed = tr ? 1 : 0;
  return true;
}
int wgrad2(kfp16_net* n, const WgradArgs& w1, const WgradArgs& w2) {
  kfp16_gemm_desc d;
  int M1, N1, M2, N2, a1[2], b1[2], a2[2], b2[2], ld1, ld2, t1, t2;
  float *ws1[2], *ws2[2];
  memset(&d, 0, sizeof(d));
  wgrad_fill(n, w1, d.A, d.B, a1, b1, ws1, ld1, t1, M1, N1);
  wgrad_fill(n, w2, d.A2, d.B2, a2, b2, ws2, ld2, t2, M2, N2);
  if (M1 != M2 || N1 != N2 || w1.X->rows != w2.X->rows || w1.groups != 2 || w2.groups != 2) {
    if (wgrad(n, *w1.X, *w1.dY, w1.param, w1.groups, w1.off0, w1.off1)) return -1;
    return wgrad(n, *w2.X, *w2.dY, w2.param, w2.groups, w2.off0, w2.off1);
  }
  d.M = M1; d.N = N1; d.K = w1.X->rows;
  d.groups = 2; d.kslabs = 1; d.kslab_len = d.K;
  d.alpha = 1.0f;
  d.a_major = KFP16_MN_MAJOR; d.b_major = KFP16_MN_MAJOR;
  for (int g = 0; g < 2; ++g) {
    d.a_row_off[g][0] = a1[g]; d.b_row_off[g][0] = b1[g];
    d.a2_row_off[g] = a2[g]; d.b2_row_off[g] = b2[g];
    d.ws[g] = ws1[g]; d.ws2[g] = ws2[g];
  }
  d.ws_ld = ld1; d.ws_transposed = t1;
  d.ws2_ld = ld2; d.ws2_transposed = t2;
  d.split_k = pick_split_k(n, d.M, d.N, 4, d.K);
  return kfp16_gemm_ex(n->ctx, &d);
}

// Defer a layer's pair of spliced weight gradients to the grouped launch at the end of the backward pass (or of the
// current graph segment): all TDNN-F layers of a stack have the same gradient shape, so the whole set runs as ONE
// persistent split-K kernel (kfp16_wgrad_group_*).  Falls back to the per-layer launch when shapes differ.
int defer_wgrads(kfp16_net* n, const WgradArgs& w1, const WgradArgs& w2) {
  auto dims = [&](const WgradArgs& a, int& M, int& N) {
    kfp16_mat A, B; int ao[2], bo[2], ld, tr; float* ws[2];
    wgrad_fill(n, a, A, B, ao, bo, ws, ld, tr, M, N);
  };
  int M1, N1, M2, N2;
  dims(w1, M1, N1); dims(w2, M2, N2);
  bool ok = M1 == M2 && N1 == N2 && M1 > 128 && w1.groups == 2 && w2.groups == 2 && w1.X->rows == w2.X->rows;
  if (ok && !n->deferred.empty()) {
    int M0, N0;
    dims(n->deferred[0], M0, N0);
    ok = M0 == M1 && N0 == N1 && n->deferred[0].X->rows == w1.X->rows;
  }
  if (!ok) return wgrad2(n, w1, w2);
  n->deferred.push_back(w1);
  n->deferred.push_back(w2);
  return 0;
}
int flush_wgrads(kfp16_net* n) {
  if (n->deferred.empty()) return 0;
  std::vector<WgradArgs> list;
  list.swap(n->deferred);
  if (list.size() == 2) return wgrad2(n, list[0], list[1]);
  std::vector<int> key;
  for (const WgradArgs& a : list) key.push_back(a.param);
  auto it = n->wg_groups.find(key);
  if (it == n->wg_groups.end()) {
    std::vector<kfp16_wgrad_prob> probs(list.size());
    int M = 0, N = 0;
    for (size_t i = 0; i < list.size(); ++i) {
      kfp16_wgrad_prob& q = probs[i];
      wgrad_fill(n, list[i], q.A, q.B, q.a_row_off, q.b_row_off, q.ws, q.ws_ld, q.ws_transposed, M, N);
    }
    kfp16_wgrad_group* g = kfp16_wgrad_group_create(n->ctx, M, N, list[0].X->rows, probs.data(), (int)probs.size());
    if (!g) return -1;
    it = n->wg_groups.emplace(key, g).first;
  }
  return kfp16_wgrad_group_launch(n->ctx, it->second);
}

// route a freshly computed input-gradient into the producer(s) of layer l
// make a gradient buffer that only carries rows row0 + k*sub dense: the other rows become zeros
int densify_rows(kfp16_net* n, const Buf& b, int row0, int sub) {
  return sub > 1 ? kfp16_zero_rows_except(n->ctx, b.p, b.LD(), b.rows, b.cols, row0, sub) : 0;
}

int deliver_dx(kfp16_net* n, Layer& l, const Buf& dx, int sub = 1, int row0 = 0) {
  // dx: [rows x in_dim]; single producer: dx IS producer.dout when it was written in place.
  // sub > 1: only rows row0 + k*sub of dx were written (the gradient is zero elsewhere)
  int col = 0;
  bool dx_dense = sub <= 1;
  for (int src : l.in) {
    Layer& s = n->layers[src];
    if (!s.needs_grad || !s.dout.p) { col += s.out_dim; continue; }
    const bool first = s.grads_seen == 0;
    s.grads_seen++;
    if (l.in.size() == 1 && dx.p == s.dout.p) {   // written in place: the producer inherits the row pattern
      s.g_sub = sub; s.g_row0 = row0;
      col += s.out_dim; continue;
    }
    if (!dx_dense) { if (densify_rows(n, dx, row0, sub)) return -1; dx_dense = true; }
    if (!first && s.g_sub > 1) {                  // an earlier consumer left a row-subsampled gradient there
      if (densify_rows(n, s.dout, s.g_row0, s.g_sub)) return -1;
      s.g_sub = 1; s.g_row0 = 0;
    }
    if (s.per_seq && !l.per_seq) {
      // adjoint of the per-sequence broadcast
      if (first) {
        if (kfp16_seq_sum(n->ctx, dx.p, dx.cols, col, s.dout.p, s.out_dim, n->opts.n_seq, n->opts.seq_len, n->halo)) return -1;
      } else {
        set_error("layer %s: a per-sequence producer with several gradient consumers is not supported", s.name.c_str());
        return -1;
      }
    } else if (first) {
      if (ops_slice_cols_on(n->ctx->stream, dx.p, dx.rows, dx.cols, s.dout.p, s.out_dim, col)) return -1;
    } else {
      if (ops_slice_add_on(n->ctx->stream, dx.p, dx.rows, dx.cols, s.dout.p, s.out_dim, col)) return -1;
    }
    col += s.out_dim;
  }
  return 0;
}

// where layer l should write its input gradient so that no copy is needed
Buf dx_target(kfp16_net* n, Layer& l) {
  if (l.in.size() == 1) {
    Layer& s = n->layers[l.in[0]];
    if (s.needs_grad && s.dout.p && s.grads_seen == 0 && s.per_seq == l.per_seq) return s.dout;
    return l.tmp_dx;
  }
  return l.d_in_cat;
}

const Buf& layer_input(kfp16_net* n, Layer& l) { return l.in.size() > 1 ? l.in_cat : n->layers[l.in[0]].out; }

// Train-mode batch-norm around a producing GEMM: Z (the GEMM's output through an identity batch-norm, i.e. ReLU(XW + b) or
// XW) -> batch statistics over the minibatch's rows -> [data-parallel hook: sum over ranks] -> folded scale / shift + running
// statistics -> Z = Z*scale + shift (+ bypass*R) in place.  hmul = heights per frame of a conv activation (1 otherwise);
// the backward pass then uses the batch-derived scale exactly as the reference's (simplified) BatchNorm backward does
// (go/gotorch/layers.go:302-330: gradInput = gradOutput * gamma * invStd).
bool train_bn_on(const kfp16_net* n) { return n->train_bn && n->g32 != nullptr; }
int train_bn_pass(kfp16_net* n, BNorm& bn, __half* Z, int ld, int rows, int cols, int hmul, const __half* R, int ldr, float res_scale, bool apply) {
  kfp16_ctx* ctx = n->ctx;
  if (!n->bn_stats) { set_error("internal: no batch-norm statistics buffer"); return -1; }
  const int period = n->blk * hmul, lo = n->halo * hmul, len = n->opts.seq_len * hmul;
  // a conv activation [frames*heights x filters] has one statistic per filter; its rows are folded so that the kernel sees
  // whole 16-byte column groups of at least 8 filters (cols == bn.dim there, so nothing to fold)
  if (kfp16_bn_batch_stats(ctx, Z, ld, rows, cols, n->bn_stats, period, lo, len)) return -1;
  if (n->bn_hook && n->bn_hook(n->bn_hook_user, n->bn_stats, 2 * cols, (void*)ctx->stream)) { set_error("batch-norm statistics hook failed"); return -1; }
  const double n_rows = (double)n->opts.n_seq * n->opts.seq_len * hmul * n->bn_world;
  if (kfp16_bn_finalize(ctx, n->bn_stats, n_rows, cols, bn.mean, bn.var, bn.rms_only ? nullptr : bn.gamma, bn.rms_only ? nullptr : bn.beta, bn.eps,
                        bn.target_rms, n->bn_momentum, bn.bwd_mul, bn.scale, bn.shift, bn.scale_bwd)) return -1;
  if (apply && kfp16_bn_apply(ctx, Z, ld, bn.scale, bn.shift, R, ldr, res_scale, rows, cols, cols, period, lo, len)) return -1;
  return 0;
}

// ------------------------------------------------------------------------------ forward
int fix_halo(kfp16_net* n, Layer& l, Buf& b, int mode) {
  if (l.per_seq || n->halo == 0) return 0;
  if (mode == HALO_REPL) return kfp16_pad_edges(n->ctx, b.p, b.cols, n->opts.n_seq, n->opts.seq_len, b.cols, n->halo);
  if (mode == HALO_ZERO) return kfp16_zero_halo(n->ctx, b.p, b.cols, n->opts.n_seq, n->opts.seq_len, b.cols, n->halo);
  return 0;
}

int forward_layer(kfp16_net* n, Layer& l) {
  kfp16_ctx* ctx = n->ctx;
  const uint32_t rr = n->opts.ref_round ? KFP16_EPI_REF_ROUND : 0;
  if (l.type == L_INPUT) return 0;
  const int rows = l.per_seq ? n->opts.n_seq : n->Tp;
  if (l.in.size() > 1) {   // Append(...) (forward.go:264-310), per-sequence members broadcast (Q6)
    int col = 0;
    for (int src : l.in) {
      Layer& s = n->layers[src];
      if (s.per_seq && !l.per_seq) {
        if (kfp16_bcast_rows(ctx, s.out.p, s.out_dim, l.in_cat.p, l.in_dim, col, rows, n->blk)) return -1;
      } else {
        if (ops_concat_cols_on(ctx->stream, l.in_cat.p, rows, l.in_dim, s.out.p, s.out_dim, col)) return -1;
      }
      col += s.out_dim;
    }
  }
  const Buf& X = layer_input(n, l);
  // rows this layer has to produce (see Layer::f_sub): every buffer below is addressed through the same row view
  const int fsub = (n->fwd_rows_now && l.f_sub > 1 && l.in.size() == 1 && !l.per_seq) ? l.f_sub : 1, fr0 = fsub > 1 ? l.f_row0 : 0;
  const Buf Xv = row_view(X, fr0, fsub), Yv = row_view(l.out, fr0, fsub);
  const int vrows = fsub > 1 ? Yv.rows : rows;
  switch (l.type) {
    case L_IDCT: {   // Y = h(X*M)  forward.go:317-330
      kfp16_gemm_desc d = mk_desc(rows, l.out_dim, l.in_dim);
      set_A(d, X.p, rows, l.in_dim);
      set_B(d, l.idct_mat, l.in_dim, l.out_dim);
      d.D[0] = l.out.p; d.ldd = l.out_dim;
      if (kfp16_gemm_ex(ctx, &d)) return -1;
      break;
    }
    case L_LINEAR: {   // Y = h(X*W)  forward.go:333-346
      kfp16_gemm_desc d = mk_desc(vrows, l.out_dim, X.cols);   // X.cols >= in_dim: zero-padded storage width
      set_A(d, Xv);
      set_B(d, W16(n, l.pW), l.in_dim, l.out_dim);
      d.D[0] = Yv.p; d.ldd = Yv.LD();
      if (kfp16_gemm_ex(ctx, &d)) return -1;
      break;
    }
    case L_BATCHNORM:   // forward.go:349-374
      if (train_bn_on(n) && !l.per_seq && train_bn_pass(n, l.bn, X.p, X.cols, rows, l.out_dim, 1, nullptr, 0, 0.f, false)) return -1;
      if (kfp16_scale_shift(ctx, X.p, l.out.p, rows, l.out_dim, l.bn.scale, l.bn.shift)) return -1;
      break;
    case L_SPECAUG:     // pass-through in the reference (forward.go:377-383, a TODO); masks when switched on for training
      if (n->spec_augment && n->g32 && !l.per_seq) {
        if (kfp16_spec_augment(ctx, X.p, l.out.p, l.out_dim, n->opts.n_seq, n->opts.seq_len, n->halo, l.out_dim, l.sa_fmax, l.sa_nfreq, l.sa_tmax,
                               l.sa_ntime, (uint32_t)(&l - n->layers.data()) * 0x9E3779B9u, n->seed_dev)) return -1;
      } else if (!check_cuda(cudaMemcpyAsync(l.out.p, X.p, l.out.bytes(), cudaMemcpyDeviceToDevice, ctx->stream), "spec-augment copy")) return -1;
      break;
    case L_COMBINE:     // forward.go:386-405
      if (!check_cuda(cudaMemcpyAsync(l.out.p, X.p, l.out.bytes(), cudaMemcpyDeviceToDevice, ctx->stream), "combine copy")) return -1;
      if (ops_combine_feature_maps_on(ctx->stream, l.out.p, rows, l.out_dim, l.height, l.nf1, l.nf2, 0)) return -1;
      break;
    case L_TDNNF: {     // forward.go:589-695
      const int s = l.stride, sp = s > 0 ? 2 : 1;
      {  // bottleneck = [X(t-s) | X(t)] * Wlin
        kfp16_gemm_desc d = mk_desc(rows, l.bott_dim, sp * l.in_dim);
        set_A(d, X.p, rows, l.in_dim);
        set_B(d, W16(n, l.pLin), sp * l.in_dim, l.bott_dim);
        d.kslabs = sp; d.kslab_len = l.in_dim;
        if (sp == 2) { d.a_row_off[0][0] = -s; d.a_row_off[0][1] = 0; d.b_row_off[0][0] = 0; d.b_row_off[0][1] = l.in_dim; }
        d.D[0] = l.bott.p; d.ldd = l.bott_dim;
        if (kfp16_gemm_ex(ctx, &d)) return -1;
        if (s > 0 && fix_halo(n, l, l.bott, HALO_REPL)) return -1;
      }
      {  // Y = BN(ReLU([B(t) | B(t+s)] * Waff + b)) (+ bypass*X)
        kfp16_gemm_desc d = mk_desc(rows, l.out_dim, sp * l.bott_dim);
        set_A(d, l.bott.p, rows, l.bott_dim);
        set_B(d, W16(n, l.pAff), sp * l.bott_dim, l.out_dim);
        d.kslabs = sp; d.kslab_len = l.bott_dim;
        if (sp == 2) { d.a_row_off[0][0] = 0; d.a_row_off[0][1] = s; d.b_row_off[0][0] = 0; d.b_row_off[0][1] = l.bott_dim; }
        d.D[0] = l.out.p; d.ldd = l.out_dim;
        d.flags = KFP16_EPI_BIAS | KFP16_EPI_RELU | KFP16_EPI_BN | KFP16_EPI_MASK | rr;
        d.bias = W16(n, l.pAffB);
        d.bn_scale = l.bn.scale; d.bn_shift = l.bn.shift;
        d.mask_out = l.mask; d.mask_ld = l.mask_ld;
        const bool tbn = train_bn_on(n);
        if (tbn && l.dropout_p > 0.f) { set_error("layer %s: dropout-proportion together with train-mode batch-norm is not supported", l.name.c_str()); return -1; }
        if (tbn) {   // identity batch-norm in the epilogue (Z = ReLU(.)), statistics + normalisation + bypass in train_bn_pass
          d.bn_scale = l.bn.one; d.bn_shift = l.bn.zero;
          if (kfp16_gemm_ex(ctx, &d)) return -1;
          if (train_bn_pass(n, l.bn, l.out.p, l.out_dim, rows, l.out_dim, 1, l.use_bypass ? X.p : nullptr, l.in_dim, l.bypass, true)) return -1;
          break;
        }
        if (l.use_bypass) { d.flags |= KFP16_EPI_RESID; d.R[0] = X.p; d.ldr = l.in_dim; d.res_scale = l.bypass; }
        if (l.dropout_p > 0.f) {   // training: inverted dropout after the batch-norm (go/gotorch/layers.go:365-399), per-step seed word
          d.flags |= KFP16_EPI_DROPOUT;
          d.drop_p = l.dropout_p;
          d.drop_seed = (uint32_t)(&l - n->layers.data()) * 0x9E3779B9u;
          d.drop_seed_dev = n->seed_dev;
        }
        if (kfp16_gemm_ex(ctx, &d)) return -1;
      }
      break;
    }
    case L_PREFINAL: {   // forward.go:912-968: affine(big) -> ReLU -> BN1 -> linear(small) -> BN2
      const Buf bigv = row_view(l.big, fr0, fsub);
      kfp16_gemm_desc d = mk_desc(vrows, l.big_dim, l.in_dim);
      set_A(d, Xv);
      set_B(d, W16(n, l.pBig), l.in_dim, l.big_dim);
      d.D[0] = bigv.p; d.ldd = bigv.LD();
      d.flags = KFP16_EPI_BIAS | KFP16_EPI_RELU | KFP16_EPI_BN | KFP16_EPI_MASK | rr;
      d.bias = W16(n, l.pBigB);
      d.bn_scale = l.bn.scale; d.bn_shift = l.bn.shift;
      d.mask_out = l.mask + (size_t)fr0 * l.mask_ld; d.mask_ld = l.mask_ld * fsub;
      const bool tbn = train_bn_on(n) && fsub == 1;      // (plan_forward_rows leaves every row in place in train-BN mode)
      if (tbn) { d.bn_scale = l.bn.one; d.bn_shift = l.bn.zero; }
      if (kfp16_gemm_ex(ctx, &d)) return -1;
      if (tbn && train_bn_pass(n, l.bn, l.big.p, l.big_dim, rows, l.big_dim, 1, nullptr, 0, 0.f, true)) return -1;
      kfp16_gemm_desc e = mk_desc(vrows, l.small_dim, l.big_dim);
      set_A(e, bigv);
      set_B(e, W16(n, l.pSmall), l.big_dim, l.small_dim);
      e.D[0] = Yv.p; e.ldd = Yv.LD();
      e.flags = KFP16_EPI_BN | rr;
      e.bn_scale = l.bn2.scale; e.bn_shift = l.bn2.shift;
      if (tbn) { e.bn_scale = l.bn2.one; e.bn_shift = l.bn2.zero; }
      if (kfp16_gemm_ex(ctx, &e)) return -1;
      if (tbn && train_bn_pass(n, l.bn2, l.out.p, l.small_dim, rows, l.small_dim, 1, nullptr, 0, 0.f, true)) return -1;
      break;
    }
    case L_CONV: {       // forward.go:418-524 with the im2col on the device; Z = BN(ReLU(P*W + b)) per filter
      const int mrows = rows * l.hout;
      kfp16_gemm_desc d = mk_desc(mrows, l.fout, l.conv_implicit ? l.convK : l.convKp);
      if (l.conv_implicit) {   // taps = shifted 4-D boxes of X (its halo rows are zero: per-sequence zero padding)
        conv_fwd_addr(n, l, X.p, 1, d.conv);
      } else {
        __half* P = l.convP ? l.convP : n->conv_P;
        if (kfp16_im2col(ctx, X.p, P, l.convKp, n->opts.n_seq, n->opts.seq_len, n->halo, l.hin, l.hout, l.hsub, l.fin,
                         (int)l.tap_dt.size(), l.tap_dt.data(), l.tap_dh.data())) return -1;
        set_A(d, P, mrows, l.convKp);
      }
      set_B(d, W16(n, l.pW), l.convK, l.fout);          // rows [convK, convKp) read as zeros (TMA bounds)
      d.D[0] = l.out.p; d.ldd = l.fout;
      d.flags = KFP16_EPI_BIAS | KFP16_EPI_RELU | KFP16_EPI_BN | KFP16_EPI_MASK | rr;
      d.bias = W16(n, l.pB);
      d.bn_scale = l.bn.scale; d.bn_shift = l.bn.shift;
      d.mask_out = l.mask; d.mask_ld = l.mask_ld;
      const bool tbn = train_bn_on(n);
      if (tbn) { d.bn_scale = l.bn.one; d.bn_shift = l.bn.zero; }
      if (n->halo > 0) {   // halo rows of the output (and their mask bits) are written as zeros by the epilogue: they are the
        // next convolution's zero padding in time, and a zero mask keeps their gradients out of the backward pass
        d.zero_row_period = n->blk * l.hout; d.zero_row_lo = n->halo * l.hout; d.zero_row_hi = (n->halo + n->opts.seq_len) * l.hout;
      }
      if (kfp16_gemm_ex(ctx, &d)) return -1;
      // per-filter statistics over every (frame, height) of the minibatch; the halo rows stay zero
      if (tbn && train_bn_pass(n, l.bn, l.out.p, l.fout, mrows, l.fout, l.hout, nullptr, 0, 0.f, true)) return -1;
      break;
    }
    case L_ATTENTION: {  // forward.go:795-909: projection GEMM, then the attention + ReLU + batch-norm in one kernel
      kfp16_gemm_desc d = mk_desc(rows, l.att_affine, l.in_dim);
      set_A(d, X.p, rows, l.in_dim);
      set_B(d, W16(n, l.pW), l.in_dim, l.att_affine);
      d.D[0] = l.big.p; d.ldd = l.att_affine;
      d.flags = KFP16_EPI_BIAS | rr;
      d.bias = W16(n, l.pB);
      if (kfp16_gemm_ex(ctx, &d)) return -1;
      if (train_bn_on(n)) { set_error("layer %s: train-mode batch-norm is not implemented for the attention layer", l.name.c_str()); return -1; }
      if (kfp16_attention_forward(ctx, l.big.p, l.att_affine, l.dz.p, l.out.p, l.out_dim, l.bn.scale, l.bn.shift, n->opts.n_seq, n->opts.seq_len,
                                  n->halo, l.att_heads, l.att_key, l.att_value, l.att_left, l.att_right, l.att_stride, l.att_scale)) return -1;
      break;
    }
    case L_OUTPUT: {     // forward.go:971-1001
      const bool view = fsub > 1 && !l.log_softmax;
      kfp16_gemm_desc d = mk_desc(view ? vrows : rows, l.out_dim, l.in_dim);
      if (view) set_A(d, Xv); else set_A(d, X.p, rows, l.in_dim);
      set_B(d, W16(n, l.pW), l.in_dim, l.out_dim);
      d.D[0] = view ? Yv.p : l.out.p; d.ldd = view ? Yv.LD() : l.out_dim;
      d.flags = KFP16_EPI_BIAS | rr;
      d.bias = W16(n, l.pB);
      if (kfp16_gemm_ex(ctx, &d)) return -1;
      if (l.log_softmax && softmax_on_stream(ctx->stream, l.out.p, rows, l.out_dim, true)) return -1;
      break;
    }
    default: break;
  }
  if (l.type == L_CONV && l.halo_mode == HALO_ZERO) return 0;    // its epilogue already stored zeros there
  if (l.halo_mode != HALO_NONE && fix_halo(n, l, l.out, l.halo_mode)) return -1;
  return 0;
}

// ------------------------------------------------------------------------------ backward
// When layer l writes its input gradient straight into the gradient buffer of a conv-relu-batchnorm producer that nobody
// else feeds, the GEMM's epilogue applies that producer's batch-norm scale and ReLU mask on the way out (EK_BN_GRADMASK):
// the producer finds dZ in its `dout` and skips its own elementwise pass over the (frames x heights x filters) tensor.
// ncols = columns of the GEMM output: the producer's filter count (conv consumer) or heights*filters (a dense consumer).
// row_mul / row_add: GEMM output row r is row r*row_mul + row_add of the producer's output (height-subsampled consumers).
bool fuse_conv_dz(kfp16_net* n, Layer& l, const Buf& dx, kfp16_gemm_desc& d, int ncols, int row_mul = 1, int row_add = 0) {
  if (!n->fuse_conv_bwd || l.in.size() != 1) return false;
  Layer& s = n->layers[l.in[0]];
  if (s.type != L_CONV || !s.needs_grad || s.n_grad_consumers != 1 || dx.p != s.dout.p || !s.mask) return false;
  const float *scale, *zero;
  if (ncols == s.fout) { scale = s.bn.scale; zero = s.bn.zero; }
  else if (ncols == s.fout * s.bn.tiles && s.bn.scale_tiled && !train_bn_on(n)) { scale = s.bn.scale_tiled; zero = s.bn.zero_tiled; }   // (the tiled copy follows kfp16_net_set_bn only)
  else return false;
  d.flags |= KFP16_EPI_BN | KFP16_EPI_GRADMASK | (n->opts.ref_round ? KFP16_EPI_REF_ROUND : 0);
  d.bn_scale = scale; d.bn_shift = zero;
  // mask words per GEMM row: the producer's mask is [frames*heights x fout/32]; a dense consumer sees heights*fout/32 per frame
  const int words = (ncols / s.fout) * s.mask_ld;
  d.mask_in = s.mask + (size_t)row_add * words; d.mask_ld = words * row_mul;
  return true;
}

int backward_layer(kfp16_net* n, Layer& l) {
  kfp16_ctx* ctx = n->ctx;
  if (!l.needs_grad || l.type == L_INPUT) return 0;
  const int rows_all = l.per_seq ? n->opts.n_seq : n->Tp;
  // rows of dout that carry the gradient (Layer::g_sub): layers that treat every row on its own work on exactly those
  // rows -- a third of the GEMM work behind a chain objective with frame-subsampling-factor 3; every other layer first
  // gets the dense form (zeros on the rows the objective never wrote)
  int sub = l.g_sub, r0 = l.g_row0;
  const bool rowwise = !l.per_seq && l.halo_mode == HALO_NONE && l.in.size() == 1 && !n->layers[l.in[0]].per_seq &&
                       (l.type == L_OUTPUT || l.type == L_PREFINAL || l.type == L_LINEAR || l.type == L_IDCT || l.type == L_BATCHNORM);
  if (sub > 1 && !rowwise) {
    if (densify_rows(n, l.dout, r0, sub)) return -1;
    sub = 1; r0 = 0; l.g_sub = 1; l.g_row0 = 0;
  }
  const int rows = sub > 1 ? (rows_all - 1 - r0) / sub + 1 : rows_all;
  if (sub > 1) n->flops_bwd_skipped += l.fl_bwd * (1.0 - (double)rows / rows_all);
  // adjoint of the halo fix-up applied to this layer's output in the forward pass
  if (!l.per_seq && n->halo > 0) {
    if (l.halo_mode == HALO_REPL && kfp16_fold_edges(ctx, l.dout.p, l.out_dim, n->opts.n_seq, n->opts.seq_len, l.out_dim, n->halo)) return -1;
    // (a convolution's halo rows carry a zero ReLU mask from its forward epilogue: dZ is zero there whatever dY holds)
    if (l.halo_mode == HALO_ZERO && l.type != L_CONV && kfp16_zero_halo(ctx, l.dout.p, l.out_dim, n->opts.n_seq, n->opts.seq_len, l.out_dim, n->halo)) return -1;
  }
  const Buf& X = layer_input(n, l);
  Buf dx = l.wants_dx ? dx_target(n, l) : Buf();
  // the same rows of every buffer the row-wise layers touch (sub == 1: the buffers themselves)
  const Buf Xv = row_view(X, r0, sub), dYv = row_view(l.dout, r0, sub), dxv = l.wants_dx ? row_view(dx, r0, sub) : Buf();
  switch (l.type) {
    case L_IDCT:
    case L_LINEAR: {
      const __half* W = l.type == L_IDCT ? l.idct_mat : W16(n, l.pW);
      if (l.wants_dx) {   // dX = dY * W^T   (backward_ops.go:162-192)
        kfp16_gemm_desc d = mk_desc(rows, l.in_dim, l.out_dim);
        set_A(d, dYv);
        d.b_major = KFP16_K_MAJOR;
        set_B(d, W, l.in_dim, l.out_dim);
        d.D[0] = dxv.p; d.ldd = dxv.LD();
        if (kfp16_gemm_ex(ctx, &d)) return -1;
      }
      if (l.type == L_LINEAR && wgrad(n, Xv, dYv, l.pW, 1, 0, 0)) return -1;   // backward_ops.go:195-225
      break;
    }
    case L_BATCHNORM:   // dX = dY * gamma/sqrt(var+eps)  (backward_wrappers.cu:104-115)
      if (l.wants_dx && sub > 1 && kfp16_scale_shift_ld(ctx, dYv.p, dYv.LD(), dxv.p, dxv.LD(), rows, l.out_dim, l.bn.scale, nullptr)) return -1;
      if (l.wants_dx && sub <= 1 && kfp16_scale_shift(ctx, l.dout.p, dx.p, rows, l.out_dim, l.bn.scale, nullptr)) return -1;
      break;
    case L_SPECAUG:     // the same masks on the gradient
      if (l.wants_dx && n->spec_augment && !l.per_seq) {
        if (kfp16_spec_augment(ctx, l.dout.p, dx.p, l.out_dim, n->opts.n_seq, n->opts.seq_len, n->halo, l.out_dim, l.sa_fmax, l.sa_nfreq, l.sa_tmax,
                               l.sa_ntime, (uint32_t)(&l - n->layers.data()) * 0x9E3779B9u, n->seed_dev)) return -1;
      } else if (l.wants_dx && !check_cuda(cudaMemcpyAsync(dx.p, l.dout.p, l.dout.bytes(), cudaMemcpyDeviceToDevice, ctx->stream), "spec-augment grad")) return -1;
      break;
    case L_COMBINE:     // inverse permutation
      if (l.wants_dx) {
        if (!check_cuda(cudaMemcpyAsync(dx.p, l.dout.p, l.dout.bytes(), cudaMemcpyDeviceToDevice, ctx->stream), "combine grad")) return -1;
        if (ops_combine_feature_maps_on(ctx->stream, dx.p, rows, l.out_dim, l.height, l.nf1, l.nf2, 1)) return -1;
      }
      break;
    case L_TDNNF: {
      const int s = l.stride, sp = s > 0 ? 2 : 1;
      // dZ = mask ? h(dY * bn_scale) : 0 ; db += colsum(dZ)
      // (with dropout the mask bit is ReLU-active AND kept, and the factor 1/(1-p) is folded into scale_bwd)
      if (kfp16_bn_relu_backward_bias(ctx, l.dout.p, l.out_dim, l.bn.scale_bwd, l.mask, l.mask_ld, l.dz.p, l.out_dim, rows, l.out_dim, G32(n, l.pAffB))) return -1;
      // with a splice both weight gradients of the layer are deferred to the grouped launch at the end of the pass
      const bool both = sp == 2;
      if (!both && wgrad(n, l.bott, l.dz, l.pAff, sp, 0, s)) return -1;
      {  // dB(r) = dZ(r)*Waff[0:bn]^T + dZ(r-s)*Waff[bn:2bn]^T
        kfp16_gemm_desc d = mk_desc(rows, l.bott_dim, sp * l.out_dim);
        set_A(d, l.dz.p, rows, l.out_dim);
        d.b_major = KFP16_K_MAJOR;
        set_B(d, W16(n, l.pAff), sp * l.bott_dim, l.out_dim);
        d.kslabs = sp; d.kslab_len = l.out_dim;
        if (sp == 2) { d.a_row_off[0][0] = 0; d.a_row_off[0][1] = -s; d.b_row_off[0][0] = 0; d.b_row_off[0][1] = l.bott_dim; }
        d.D[0] = l.dbott.p; d.ldd = l.bott_dim;
        if (kfp16_gemm_ex(ctx, &d)) return -1;
        if (s > 0 && kfp16_fold_edges(ctx, l.dbott.p, l.bott_dim, n->opts.n_seq, n->opts.seq_len, l.bott_dim, n->halo)) return -1;
      }
      // dWlin = [X(t-s) | X(t)]^T * dB
      if (both) {
        const WgradArgs wa{&l.bott, &l.dz, l.pAff, sp, 0, s}, wl{&X, &l.dbott, l.pLin, sp, -s, 0};
        if (defer_wgrads(n, wa, wl)) return -1;
      } else if (wgrad(n, X, l.dbott, l.pLin, sp, -s, 0)) return -1;
      if (l.wants_dx) {   // dX(r) = dB(r+s)*Wlin[0:in]^T + dB(r)*Wlin[in:2in]^T (+ bypass*dY)
        kfp16_gemm_desc d = mk_desc(rows, l.in_dim, sp * l.bott_dim);
        set_A(d, l.dbott.p, rows, l.bott_dim);
        d.b_major = KFP16_K_MAJOR;
        set_B(d, W16(n, l.pLin), sp * l.in_dim, l.bott_dim);
        d.kslabs = sp; d.kslab_len = l.bott_dim;
        if (sp == 2) { d.a_row_off[0][0] = s; d.a_row_off[0][1] = 0; d.b_row_off[0][0] = 0; d.b_row_off[0][1] = l.in_dim; }
        d.D[0] = dx.p; d.ldd = l.in_dim;
        if (l.use_bypass) { d.flags |= KFP16_EPI_RESID; d.R[0] = l.dout.p; d.ldr = l.out_dim; d.res_scale = l.bypass; }
        else if (sp == 1 && fuse_conv_dz(n, l, dx, d, l.in_dim)) n->layers[l.in[0]].dz_in_dout = true;   // first TDNN-F layer behind the CNN
        if (kfp16_gemm_ex(ctx, &d)) return -1;
      }
      break;
    }
    case L_PREFINAL: {
      const Buf dysv = row_view(l.dys, r0, sub), dbigv = row_view(l.dbig, r0, sub), bigv = row_view(l.big, r0, sub);
      // dYs = dY * bn2_scale
      if (kfp16_scale_shift_ld(ctx, dYv.p, dYv.LD(), dysv.p, dysv.LD(), rows, l.small_dim, l.bn2.scale, nullptr)) return -1;
      {  // dG = mask ? h((dYs * Wsmall^T) * bn1_scale) : 0
        kfp16_gemm_desc d = mk_desc(rows, l.big_dim, l.small_dim);
        set_A(d, dysv);
        d.b_major = KFP16_K_MAJOR;
        set_B(d, W16(n, l.pSmall), l.big_dim, l.small_dim);
        d.D[0] = dbigv.p; d.ldd = dbigv.LD();
        d.flags = KFP16_EPI_BN | KFP16_EPI_GRADMASK | (n->opts.ref_round ? KFP16_EPI_REF_ROUND : 0);
        d.bn_scale = l.bn.scale; d.bn_shift = l.bn.zero;
        d.mask_in = l.mask + (size_t)(sub > 1 ? r0 : 0) * l.mask_ld; d.mask_ld = l.mask_ld * sub;
        if (kfp16_gemm_ex(ctx, &d)) return -1;
      }
      if (wgrad(n, bigv, dysv, l.pSmall, 1, 0, 0)) return -1;
      if (kfp16_colsum_accum(ctx, dbigv.p, dbigv.LD(), rows, l.big_dim, G32(n, l.pBigB))) return -1;
      if (l.wants_dx) {
        kfp16_gemm_desc d = mk_desc(rows, l.in_dim, l.big_dim);
        set_A(d, dbigv);
        d.b_major = KFP16_K_MAJOR;
        set_B(d, W16(n, l.pBig), l.in_dim, l.big_dim);
        d.D[0] = dxv.p; d.ldd = dxv.LD();
        if (kfp16_gemm_ex(ctx, &d)) return -1;
      }
      if (wgrad(n, Xv, dbigv, l.pBig, 1, 0, 0)) return -1;
      break;
    }
    case L_CONV: {     // transpose of the forward (the reference treats the conv as a dense affine: quirk Q2)
      const int mrows = rows * l.hout;
      // gradients on halo rows are not part of the minibatch: the forward epilogue left a zero ReLU mask on them, so dZ is
      // zero there and neither the weight gradient (a sum over ALL padded rows) nor the input gradient of neighbouring
      // real frames sees them
      __half* dz = n->conv_dz;
      if (l.dz_in_dout) {   // the consumer's epilogue already produced dZ (fuse_conv_dz): only the bias gradient is left
        dz = l.dout.p;
        if (kfp16_colsum_accum(ctx, dz, l.fout, mrows, l.fout, G32(n, l.pB))) return -1;
      } else if (kfp16_bn_relu_backward_bias(ctx, l.dout.p, l.fout, l.bn.scale, l.mask, l.mask_ld, dz, l.fout, mrows, l.fout, G32(n, l.pB))) return -1;
      if (l.conv_implicit) {
        {  // dW[(tap, f), fo] = sum over (t, h) of X[t+dt, h*sub+dh, f] * dZ[(t, h), fo]
          kfp16_gemm_desc d = mk_desc(l.convK, l.fout, mrows);
          d.a_major = KFP16_MN_MAJOR;
          conv_fwd_addr(n, l, X.p, 2, d.conv);
          set_B(d, dz, mrows, l.fout);
          d.split_k = pick_split_k(n, l.convK, l.fout, 1, mrows);
          d.ws[0] = G32(n, l.pW); d.ws_ld = l.fout;
          if (kfp16_gemm_ex(ctx, &d)) return -1;
        }
        if (l.wants_dx) {
          // dX[t, hi, f] = sum over taps of dZ[t-dt, (hi-dh)/sub, :] * W[(tap, f), :]^T: again a convolution, of dZ with the
          // mirrored taps; with height subsampling 2 the even and the odd input heights take different tap subsets and are
          // written as two interleaved row sets of dX
          bool fused_dz = false;
          for (int par = 0; par < l.hsub; ++par) {
            kfp16_gemm_desc d = mk_desc(rows * l.hout, l.fin, 0);
            kfp16_conv_addr& c = d.conv;
            c.mode = 1; c.x = dz;
            c.T = rows; c.H = l.hout; c.P = 1; c.C = l.fout; c.rows_h = l.hout;
            for (size_t t = 0; t < l.tap_dt.size(); ++t) {
              const int dh = l.tap_dh[t];
              if (((par - dh) % l.hsub) != 0) continue;
              c.dt[c.ntaps] = -l.tap_dt[t];
              c.hq[c.ntaps] = (par - dh) / l.hsub;
              c.par[c.ntaps] = 0;
              c.brow[c.ntaps] = (int)t * l.fin;
              ++c.ntaps;
            }
            __half* dst = dx.p + (size_t)par * l.fin;
            const int ldd = l.hsub * l.fin;
            if (c.ntaps == 0) {   // no tap reaches this height parity
              if (!check_cuda(cudaMemset2DAsync(dst, (size_t)ldd * 2, 0, (size_t)l.fin * 2, (size_t)rows * l.hout, ctx->stream), "conv input-gradient clear")) return -1;
              continue;
            }
            d.K = c.ntaps * l.fout; d.kslab_len = d.K;
            d.b_major = KFP16_K_MAJOR;
            set_B(d, W16(n, l.pW), l.convK, l.fout);
            d.D[0] = dst; d.ldd = ldd;
            // (rows of this launch are the producer's rows (t, hsub*ho + par): every hsub-th mask row from row `par` on)
            if (fuse_conv_dz(n, l, dx, d, l.fin, l.hsub, par)) fused_dz = true;
            if (kfp16_gemm_ex(ctx, &d)) return -1;
          }
          if (fused_dz) n->layers[l.in[0]].dz_in_dout = true;
        }
        break;
      }
      // patch-matrix path: the patch matrix of the forward pass is still resident (l.convP)
      {  // dW[K x fout] = P^T dZ
        kfp16_gemm_desc d = mk_desc(l.convK, l.fout, mrows);
        d.a_major = KFP16_MN_MAJOR;
        set_A(d, l.convP, mrows, l.convKp);
        set_B(d, dz, mrows, l.fout);
        d.split_k = pick_split_k(n, l.convK, l.fout, 1, mrows);
        d.ws[0] = G32(n, l.pW); d.ws_ld = l.fout;
        if (kfp16_gemm_ex(ctx, &d)) return -1;
      }
      if (l.wants_dx) {   // dP = dZ W^T, then the adjoint of the patch gather
        kfp16_gemm_desc d = mk_desc(mrows, l.convKp, l.fout);
        set_A(d, dz, mrows, l.fout);
        d.b_major = KFP16_K_MAJOR;
        set_B(d, W16(n, l.pW), l.convK, l.fout);
        d.D[0] = n->conv_dP; d.ldd = l.convKp;
        if (kfp16_gemm_ex(ctx, &d)) return -1;
        if (kfp16_col2im(ctx, n->conv_dP, l.convKp, dx.p, n->opts.n_seq, n->opts.seq_len, n->halo, l.hin, l.hout, l.hsub, l.fin,
                         (int)l.tap_dt.size(), l.tap_dt.data(), l.tap_dh.data())) return -1;
      }
      break;
    }
    case L_ATTENTION: {   // exact transpose of the forward (the reference treats the layer as a plain affine: quirk Q2)
      if (kfp16_attention_backward(ctx, l.big.p, l.att_affine, l.dz.p, l.dout.p, l.out_dim, l.bn.scale, l.att_db, l.dbig.p, n->opts.n_seq,
                                   n->opts.seq_len, n->halo, l.att_heads, l.att_key, l.att_value, l.att_left, l.att_right, l.att_stride, l.att_scale)) return -1;
      if (kfp16_colsum_accum(ctx, l.dbig.p, l.att_affine, rows, l.att_affine, G32(n, l.pB))) return -1;
      if (wgrad(n, X, l.dbig, l.pW, 1, 0, 0)) return -1;
      if (l.wants_dx) {
        kfp16_gemm_desc d = mk_desc(rows, l.in_dim, l.att_affine);
        set_A(d, l.dbig.p, rows, l.att_affine);
        d.b_major = KFP16_K_MAJOR;
        set_B(d, W16(n, l.pW), l.in_dim, l.att_affine);
        d.D[0] = dx.p; d.ldd = l.in_dim;
        if (kfp16_gemm_ex(ctx, &d)) return -1;
      }
      break;
    }
    case L_OUTPUT: {   // log-softmax Jacobian is not applied (network_backward.go:243-244)
      if (l.wants_dx) {
        kfp16_gemm_desc d = mk_desc(rows, l.in_dim, l.out_dim);
        set_A(d, dYv);
        d.b_major = KFP16_K_MAJOR;
        set_B(d, W16(n, l.pW), l.in_dim, l.out_dim);
        d.D[0] = dxv.p; d.ldd = dxv.LD();
        if (kfp16_gemm_ex(ctx, &d)) return -1;
      }
      if (wgrad(n, Xv, dYv, l.pW, 1, 0, 0)) return -1;
      if (kfp16_colsum_accum(ctx, dYv.p, dYv.LD(), rows, l.out_dim, G32(n, l.pB))) return -1;
      break;
    }
    default: break;
  }
  if (l.wants_dx && deliver_dx(n, l, dx, sub, r0)) return -1;
  return 0;
}

// can the chain objective's gradient be handled as the rows row0 + k*sub of the padded matrix (see kfp16_net_loss_chain)?
bool chain_by_rows(const kfp16_net* n, int layer, int sub, int frames) {
  const Layer& l = n->layers[layer];
  return layer == n->out_layer && n->sparse_out_grad && sub > 1 && n->blk % sub == 0 && n->blk / sub - frames <= 4 && (l.out_dim % 8) == 0;
}

// Layer::f_sub for the training step: walk back from the objective's output layer through row-wise layers whose every
// consumer is already restricted to the objective's rows
void plan_forward_rows(kfp16_net* n) {
  for (auto& l : n->layers) { l.f_sub = 1; l.f_row0 = 0; }
  n->flops_fwd_skipped = 0;
  if (train_bn_on(n)) return;      // batch statistics are taken over every real row of the minibatch
  if (!n->chain || n->out_layer < 0 || !chain_by_rows(n, n->out_layer, n->chain_sub, kfp16_chain_frames(n->chain))) return;
  const int L = (int)n->layers.size();
  std::vector<int> consumers(L, 0), restricted(L, 0);
  for (const auto& l : n->layers) for (int src : l.in) consumers[src]++;
  std::vector<bool> mark(L, false);
  mark[n->out_layer] = true;
  for (int i = n->out_layer; i >= 0; --i) {
    Layer& l = n->layers[i];
    if (!mark[i]) continue;
    const bool rowwise = !l.per_seq && l.halo_mode == HALO_NONE && l.in.size() == 1 && !n->layers[l.in[0]].per_seq &&
                         ((l.type == L_OUTPUT && !l.log_softmax) || l.type == L_PREFINAL || l.type == L_LINEAR);
    if (!rowwise) continue;
    l.f_sub = n->chain_sub; l.f_row0 = n->halo + n->chain_left;
    n->flops_fwd_skipped += l.fl_fwd * (1.0 - (double)((n->Tp - 1 - l.f_row0) / l.f_sub + 1) / n->Tp);
    const int src = l.in[0];
    if (++restricted[src] == consumers[src]) mark[src] = true;
  }
}

// Forward pass + objective of one step (after the gradient accumulators were cleared, before the backward pass): shared by
// the whole-step graph (run_phases) and by segment 0 of the segmented step (run_segment).
int forward_and_objective(kfp16_net* n) {
  // a chain objective on subsampled output frames: the row-wise layers that feed only the objective compute just its rows
  plan_forward_rows(n);
  // Layers the objective does not depend on (the xent branch: prefinal-xent, output-xent + log-softmax -- computed by the
  // reference's Forward as well, read by nobody in the step) run AFTER the objective has been queued on a second stream:
  // the chain kernel holds one CTA per sequence (64 of 148 SMs, packed two to a TPC) for ~190 us, the GEMMs beside it
  // are limited to the remaining SMs.  Both join before the backward pass.
  const int L = (int)n->layers.size();
  std::vector<char> anc(L, 0);
  int n_rest = 0;
  if (n->out_layer >= 0) {
    anc[n->out_layer] = 1;
    for (int i = n->out_layer; i >= 0; --i)
      if (anc[i]) for (int src : n->layers[i].in) anc[src] = 1;
    for (int i = 0; i < L; ++i) if (!anc[i] && n->layers[i].type != L_INPUT) ++n_rest;
  }
  const int sms = n->ctx->num_sms, held = (n->opts.n_seq + 1) & ~1;
  const bool overlap = n->chain && n->overlap_loss && n->ctx->stream && n_rest > 0 && held + 16 <= sms && (n->ctx->max_ctas == 0 || n->ctx->max_ctas >= sms);
  if (overlap && !n->side_stream) {
    if (!check_cuda(cudaStreamCreateWithFlags(&n->side_stream, cudaStreamNonBlocking), "side stream") ||
        !check_cuda(cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming), "fork event") ||
        !check_cuda(cudaEventCreateWithFlags(&n->ev_join, cudaEventDisableTiming), "join event")) return -1;
  }
  n->fwd_rows_now = true;
  int frc = 0;
  for (int i = 0; i < L && !frc; ++i)
    if (!overlap || anc[i]) frc = forward_layer(n, n->layers[i]);
  if (!frc && overlap) {
    cudaStream_t main_stream = n->ctx->stream;
    frc = !check_cuda(cudaEventRecord(n->ev_fork, main_stream), "fork record") ||
          !check_cuda(cudaStreamWaitEvent(n->side_stream, n->ev_fork, 0), "fork wait");
    if (!frc) {
      n->ctx->stream = n->side_stream;
      frc = kfp16_net_loss_chain(n, "", n->chain, n->chain_sub, n->chain_left, n->chain_weight);
      n->ctx->stream = main_stream;
    }
    if (!frc) frc = !check_cuda(cudaEventRecord(n->ev_join, n->side_stream), "join record");
    const int saved_max = n->ctx->max_ctas;
    n->ctx->max_ctas = (sms - held - 4) & ~1;            // whole TPCs, a little slack for the block scheduler
    for (int i = 0; i < L && !frc; ++i)
      if (!anc[i]) frc = forward_layer(n, n->layers[i]);
    n->ctx->max_ctas = saved_max;
    if (!frc) frc = !check_cuda(cudaStreamWaitEvent(main_stream, n->ev_join, 0), "join wait");
  }
  n->fwd_rows_now = false;
  if (frc) return -1;
  if (!overlap && (n->chain ? kfp16_net_loss_chain(n, "", n->chain, n->chain_sub, n->chain_left, n->chain_weight) : kfp16_net_loss_half_sq(n, ""))) return -1;
  return 0;
}

int run_phases(kfp16_net* n, int phases) {
  if (phases & 1) {
    if (kfp16_bump_counter(n->ctx, n->seed_dev)) return -1;     // a new dropout mask per step, graph replays included
    if (kfp16_net_zero_grads(n)) return -1;
    if (forward_and_objective(n)) return -1;
    if (kfp16_net_backward(n)) return -1;
  }
  if (phases & 4) {
    if (kfp16_net_grads_to_f16(n)) return -1;
  }
  if (phases & 2) {
    if (kfp16_net_sgd_step(n, n->hp_host[2], n->opts.round_grad)) return -1;
  }
  if (phases & 8) {
    if (kfp16_net_sgd_step_f16(n)) return -1;
  }
  return 0;
}

}  // namespace

// =================================================================================== C ABI
extern "C" {

kfp16_net* kfp16_net_create(kfp16_ctx* ctx, const char* xconfig, const kfp16_net_opts* opts) {
  if (!ctx || !xconfig || !opts) { set_error("kfp16_net_create: null argument"); return nullptr; }
  if (opts->n_seq <= 0 || opts->seq_len <= 0) { set_error("kfp16_net_create: n_seq and seq_len must be positive"); return nullptr; }
  if (!check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice")) return nullptr;
  std::unique_ptr<kfp16_net> n(new kfp16_net());
  n->ctx = ctx;
  n->opts = *opts;
  bool ok = parse_xconfig(n.get(), xconfig) && resolve_dims(n.get()) && build_plan(n.get());
  if (!ok) {
    char msg[512];
    snprintf(msg, sizeof(msg), "%s", get_error() ? get_error() : "?");
    kfp16_net_destroy(n.release());
    set_error("%s", msg);
    return nullptr;
  }
  if (kfp16_net_init_random(n.get(), 42) != 0) { kfp16_net_destroy(n.release()); return nullptr; }
  return n.release();
}

void kfp16_net_destroy(kfp16_net* n) {
  if (!n) return;
  for (auto& g : n->graph)
    if (g.second) cudaGraphExecDestroy(g.second);
  for (void* p : n->allocs) cudaFree(p);
  for (auto& l : n->layers)
    for (int b = 0; b < 2; ++b) {
      if (l.pf_copied[b]) cudaEventDestroy(l.pf_copied[b]);
      if (l.pf_packed[b]) cudaEventDestroy(l.pf_packed[b]);
    }
  if (n->copy_stream) cudaStreamDestroy(n->copy_stream);
  for (int k = 0; k < 3; ++k) { if (n->copy_extra[k]) cudaStreamDestroy(n->copy_extra[k]); if (n->copy_part[k]) cudaEventDestroy(n->copy_part[k]); }
  for (auto& kv : n->wg_groups) kfp16_wgrad_group_destroy(kv.second);
  for (cudaGraphExec_t g : n->seg_graph) if (g) cudaGraphExecDestroy(g);
  if (n->side_stream) cudaStreamDestroy(n->side_stream);
  if (n->ev_fork) cudaEventDestroy(n->ev_fork);
  if (n->ev_join) cudaEventDestroy(n->ev_join);
  if (n->loss_pinned) cudaFreeHost(n->loss_pinned);
  for (cudaEvent_t e : n->loss_ev) if (e) cudaEventDestroy(e);
  delete n;
}

int kfp16_net_num_layers(const kfp16_net* n) { return n ? (int)n->layers.size() : 0; }
const char* kfp16_net_layer_name(const kfp16_net* n, int i) { return (n && i >= 0 && i < (int)n->layers.size()) ? n->layers[i].name.c_str() : nullptr; }
const char* kfp16_net_layer_type(const kfp16_net* n, int i) { return (n && i >= 0 && i < (int)n->layers.size()) ? n->layers[i].type_name.c_str() : nullptr; }
int kfp16_net_layer_dim(const kfp16_net* n, int i) { return (n && i >= 0 && i < (int)n->layers.size()) ? n->layers[i].out_dim : 0; }
int kfp16_net_padded_rows(const kfp16_net* n) { return n ? n->Tp : 0; }
int kfp16_net_halo(const kfp16_net* n) { return n ? n->halo : 0; }
double kfp16_net_flops_forward(const kfp16_net* n) { return n ? n->flops_fwd : 0.0; }
double kfp16_net_flops_backward(const kfp16_net* n) { return n ? n->flops_bwd : 0.0; }
double kfp16_net_flops_skipped(const kfp16_net* n) { return n ? n->flops_fwd_skipped + n->flops_bwd_skipped : 0.0; }

int kfp16_net_num_params(const kfp16_net* n) { return n ? (int)n->params.size() : 0; }
const char* kfp16_net_param_name(const kfp16_net* n, int i) { return (n && i >= 0 && i < (int)n->params.size()) ? n->params[i].name.c_str() : nullptr; }
int kfp16_net_param_shape(const kfp16_net* n, int i, int* rows, int* cols) {
  if (!n || i < 0 || i >= (int)n->params.size()) { set_error("kfp16_net_param_shape: bad index"); return -1; }
  if (rows) *rows = n->params[i].rows;
  if (cols) *cols = n->params[i].cols;
  return 0;
}
size_t kfp16_net_param_offset(const kfp16_net* n, int i) { return (n && i >= 0 && i < (int)n->params.size()) ? n->params[i].off : 0; }
size_t kfp16_net_bucket_size(const kfp16_net* n) { return n ? n->bucket : 0; }
void* kfp16_net_params_f16(kfp16_net* n) { return n ? n->w16 : nullptr; }
float* kfp16_net_params_f32(kfp16_net* n) { return n ? n->w32 : nullptr; }
float* kfp16_net_velocity(kfp16_net* n) { return n ? n->vel : nullptr; }
float* kfp16_net_grads_f32(kfp16_net* n) { return n ? n->g32 : nullptr; }

int kfp16_net_set_param(kfp16_net* n, const char* name, const float* host, int rows, int cols) {
  if (!n || !name || !host) { set_error("kfp16_net_set_param: null argument"); return -1; }
  const int i = find_param(n, name);
  if (i < 0) { set_error("kfp16_net_set_param: no parameter named %s", name); return -1; }
  const Param& P = n->params[i];
  if (P.rows != rows || P.cols != cols) { set_error("kfp16_net_set_param: %s is [%d x %d], got [%d x %d]", name, P.rows, P.cols, rows, cols); return -1; }
  const size_t cnt = (size_t)rows * cols;
  std::vector<uint16_t> h16(cnt);
  std::vector<float> h32(cnt);
  for (size_t k = 0; k < cnt; ++k) { h16[k] = f32_to_f16_trunc(host[k]); h32[k] = f16_bits_to_f32(h16[k]); }
  if (!check_cuda(cudaMemcpy(n->w16 + P.off, h16.data(), cnt * 2, cudaMemcpyHostToDevice), "param upload")) return -1;
  if (n->w32) {
    if (!check_cuda(cudaMemcpy(n->w32 + P.off, h32.data(), cnt * 4, cudaMemcpyHostToDevice), "master upload")) return -1;
    if (!check_cuda(cudaMemset(n->vel + P.off, 0, cnt * 4), "velocity reset")) return -1;
  }
  return 0;
}
// idct-layer: replace the computed matrix (makeIDCTMatrix, forward.go:1190-1210) by a loaded one -- Kaldi's `idct` component
// through the truncating converter, as LoadWeights does (weight_loader.go:766-776).  host: fp32 [dim x dim], [in x out].
int kfp16_net_set_idct(kfp16_net* n, const char* layer, const float* host, int dim) {
  if (!n || !layer || !host) { set_error("kfp16_net_set_idct: null argument"); return -1; }
  const int i = find_layer(n, layer);
  if (i < 0 || n->layers[i].type != L_IDCT) { set_error("kfp16_net_set_idct: no idct-layer named %s", layer); return -1; }
  Layer& l = n->layers[i];
  if (dim != l.in_dim) { set_error("kfp16_net_set_idct: %s is [%d x %d], got dim %d", layer, l.in_dim, l.out_dim, dim); return -1; }
  std::vector<uint16_t> h16((size_t)dim * dim);
  for (size_t k = 0; k < h16.size(); ++k) h16[k] = f32_to_f16_trunc(host[k]);
  if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "sync")) return -1;
  return check_cuda(cudaMemcpy(l.idct_mat, h16.data(), h16.size() * 2, cudaMemcpyHostToDevice), "idct upload") ? 0 : -1;
}
int kfp16_net_get_param(kfp16_net* n, const char* name, uint16_t* host, int rows, int cols) {
  if (!n || !name || !host) { set_error("kfp16_net_get_param: null argument"); return -1; }
  const int i = find_param(n, name);
  if (i < 0) { set_error("kfp16_net_get_param: no parameter named %s", name); return -1; }
  const Param& P = n->params[i];
  if (P.rows != rows || P.cols != cols) { set_error("kfp16_net_get_param: shape mismatch for %s", name); return -1; }
  if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "sync")) return -1;
  return check_cuda(cudaMemcpy(host, n->w16 + P.off, (size_t)rows * cols * 2, cudaMemcpyDeviceToHost), "param download") ? 0 : -1;
}

int kfp16_net_set_bn(kfp16_net* n, const char* layer, const char* which, const float* mean, const float* var,
                     const float* gamma, const float* beta, float eps, int dim) {
  if (!n || !layer || !mean || !var) { set_error("kfp16_net_set_bn: null argument"); return -1; }
  const int i = find_layer(n, layer);
  if (i < 0) { set_error("kfp16_net_set_bn: no layer %s", layer); return -1; }
  Layer& l = n->layers[i];
  BNorm* bn = &l.bn;
  if (which && (!strcmp(which, "BN") || !strcmp(which, "bn2")) && l.type == L_PREFINAL) bn = &l.bn2;
  if (!bn->present || bn->dim != dim) { set_error("kfp16_net_set_bn: layer %s has no batch-norm of dim %d", layer, dim); return -1; }
  bn->eps = eps;
  std::vector<float> ones(dim, 1.f), zeros(dim, 0.f);
  const size_t b = (size_t)dim * sizeof(float);
  if (!check_cuda(cudaMemcpy(bn->mean, mean, b, cudaMemcpyHostToDevice), "bn mean") ||
      !check_cuda(cudaMemcpy(bn->var, var, b, cudaMemcpyHostToDevice), "bn var") ||
      !check_cuda(cudaMemcpy(bn->gamma, gamma ? gamma : ones.data(), b, cudaMemcpyHostToDevice), "bn gamma") ||
      !check_cuda(cudaMemcpy(bn->beta, beta ? beta : zeros.data(), b, cudaMemcpyHostToDevice), "bn beta"))
    return -1;
  const bool rms = bn->rms_only;
  if (kfp16_bn_fold(n->ctx, bn->mean, bn->var, rms ? nullptr : bn->gamma, rms ? nullptr : bn->beta, eps, bn->target_rms, dim, bn->scale, bn->shift)) return -1;
  if (kfp16_scale_f32(n->ctx, bn->scale, bn->scale_bwd, dim, bn->bwd_mul)) return -1;
  if (!retile_bn(n, *bn, bn->tiles)) return -1;
  return check_cuda(cudaStreamSynchronize(n->ctx->stream), "bn fold sync") ? 0 : -1;
}

int kfp16_net_init_random(kfp16_net* n, uint64_t seed) {
  if (!n) { set_error("kfp16_net_init_random: null network"); return -1; }
  std::mt19937_64 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<uint16_t> h16(std::max<size_t>(n->bucket, 1), 0);
  for (const Param& P : n->params) {
    if (P.is_bias) continue;
    const float scale = sqrtf(2.0f / (float)(P.rows + P.cols));   // forward.go:1161-1168
    for (size_t k = 0; k < (size_t)P.rows * P.cols; ++k) h16[P.off + k] = f32_to_f16_trunc(nd(rng) * scale);
  }
  if (!check_cuda(cudaMemcpy(n->w16, h16.data(), n->bucket * 2, cudaMemcpyHostToDevice), "param upload")) return -1;
  if (n->w32) {
    std::vector<float> h32(n->bucket);
    for (size_t k = 0; k < n->bucket; ++k) h32[k] = f16_bits_to_f32(h16[k]);
    if (!check_cuda(cudaMemcpy(n->w32, h32.data(), n->bucket * 4, cudaMemcpyHostToDevice), "master upload")) return -1;
    if (!check_cuda(cudaMemset(n->vel, 0, n->bucket * 4), "velocity reset")) return -1;
  }
  return 0;
}

static int set_input_common(kfp16_net* n, const char* input_name, const void* dense_dev, int rows, int cols, bool f32 = false) {
  const int i = find_layer(n, input_name);
  if (i < 0 || n->layers[i].type != L_INPUT) { set_error("kfp16_net_set_input: no input layer named %s", input_name); return -1; }
  Layer& l = n->layers[i];
  const int want_rows = l.per_seq ? n->opts.n_seq : n->T;
  if (rows != want_rows || cols != l.out_dim) {
    set_error("kfp16_net_set_input: %s expects [%d x %d], got [%d x %d]", input_name, want_rows, l.out_dim, rows, cols);
    return -1;
  }
  // halo rows: replicate for spliced consumers, zero otherwise (finite values either way)
  const int mode = l.halo_mode == HALO_ZERO ? 0 : 1;
  if (f32) {   // FP32 rows: RNE conversion on the device while scattering (bridge.go:141 does it on the CPU)
    if (l.per_seq) return kfp16_pack_rows_f32(n->ctx, (const float*)dense_dev, l.out.p, l.out.cols, rows, 1, 0, cols, 0);
    return kfp16_pack_rows_f32(n->ctx, (const float*)dense_dev, l.out.p, l.out.cols, n->opts.n_seq, n->opts.seq_len, n->halo, cols, mode);
  }
  if (l.per_seq)
    return check_cuda(cudaMemcpy2DAsync(l.out.p, (size_t)l.out.cols * 2, dense_dev, (size_t)cols * 2, (size_t)cols * 2, rows,
                                        cudaMemcpyDeviceToDevice, n->ctx->stream), "input copy") ? 0 : -1;
  return kfp16_pack_rows(n->ctx, dense_dev, l.out.p, l.out.cols, n->opts.n_seq, n->opts.seq_len, n->halo, cols, mode);
}
int kfp16_net_set_input_device(kfp16_net* n, const char* input_name, const void* dev, int rows, int cols) {
  if (!n || !input_name || !dev) { set_error("kfp16_net_set_input_device: null argument"); return -1; }
  return set_input_common(n, input_name, dev, rows, cols);
}
static int set_input_host(kfp16_net* n, const char* input_name, const void* host, int rows, int cols, bool f32) {
  if (!n || !input_name || !host) { set_error("kfp16_net_set_input: null argument"); return -1; }
  const size_t bytes = (size_t)rows * cols * (f32 ? 4 : 2);
  if (bytes > n->stage_bytes) { set_error("kfp16_net_set_input: [%d x %d] exceeds the staging buffer", rows, cols); return -1; }
  // the staging buffer is reused by the next call: keep the copy and the scatter ordered on the stream
  if (!check_cuda(cudaMemcpyAsync(n->stage_in, host, bytes, cudaMemcpyHostToDevice, n->ctx->stream), "input upload")) return -1;
  if (set_input_common(n, input_name, n->stage_in, rows, cols, f32)) return -1;
  return check_cuda(cudaStreamSynchronize(n->ctx->stream), "input sync") ? 0 : -1;
}
int kfp16_net_set_input(kfp16_net* n, const char* input_name, const uint16_t* host, int rows, int cols) {
  return set_input_host(n, input_name, host, rows, cols, false);
}
int kfp16_net_set_input_f32(kfp16_net* n, const char* input_name, const float* host, int rows, int cols) {
  return set_input_host(n, input_name, host, rows, cols, true);
}

// Asynchronous input path: the NEXT minibatch's rows travel host -> device on a copy stream while the current
// step computes; kfp16_net_commit_input then scatters the staged rows into the padded layout on the main
// stream (converting FP32 -> FP16 there when the rows were FP32).  `host` must be pinned (bridge_host_alloc) and stay
// untouched until the matching commit.
static int prefetch_common(kfp16_net* n, const char* input_name, const void* host, int rows, int cols, bool f32) {
  if (!n || !input_name || !host) { set_error("kfp16_net_prefetch_input: null argument"); return -1; }
  const int i = find_layer(n, input_name);
  if (i < 0 || n->layers[i].type != L_INPUT) { set_error("kfp16_net_prefetch_input: no input layer named %s", input_name); return -1; }
  Layer& l = n->layers[i];
  const int want_rows = l.per_seq ? n->opts.n_seq : n->T;
  if (rows != want_rows || cols != l.out_dim) { set_error("kfp16_net_prefetch_input: %s expects [%d x %d], got [%d x %d]", input_name, want_rows, l.out_dim, rows, cols); return -1; }
  if (!n->copy_stream && !check_cuda(cudaStreamCreateWithFlags(&n->copy_stream, cudaStreamNonBlocking), "copy stream")) return -1;
  const size_t bytes = (size_t)rows * cols * (f32 ? 4 : 2);
  const int b = l.pf_slot;
  if (l.pf_cap[b] < bytes) {      // first use of this slot (or a wider element type than before): FP32-sized from then on
    const size_t cap = (size_t)rows * cols * 4;
    if (l.pf_buf[b] && !check_cuda(cudaStreamSynchronize(n->ctx->stream), "prefetch regrow sync")) return -1;
    if (!dev_alloc(n, (void**)&l.pf_buf[b], cap, false)) return -1;
    l.pf_cap[b] = cap;
    if (!l.pf_copied[b] && (!check_cuda(cudaEventCreateWithFlags(&l.pf_copied[b], cudaEventDisableTiming), "prefetch event") ||
                            !check_cuda(cudaEventCreateWithFlags(&l.pf_packed[b], cudaEventDisableTiming), "prefetch event"))) return -1;
  } else if (!check_cuda(cudaStreamWaitEvent(n->copy_stream, l.pf_packed[b], 0), "prefetch wait")) {
    return -1;   // the scatter that last read this staging slot must have run
  }
  const int parts = bytes >= ((size_t)4 << 20) ? 4 : 1;
  const size_t chunk = ((bytes / parts) + 255) & ~(size_t)255;
  if (!check_cuda(cudaMemcpyAsync(l.pf_buf[b], host, parts > 1 ? std::min(chunk, bytes) : bytes, cudaMemcpyHostToDevice, n->copy_stream), "prefetch copy")) return -1;
  for (int k = 1; k < parts; ++k) {       // parts 1.. on their own streams (joined into the copy stream), part 0 above
    if (!n->copy_extra[k - 1]) {
      if (!check_cuda(cudaStreamCreateWithFlags(&n->copy_extra[k - 1], cudaStreamNonBlocking), "copy stream") ||
          !check_cuda(cudaEventCreateWithFlags(&n->copy_part[k - 1], cudaEventDisableTiming), "copy event")) return -1;
    }
    const size_t off = chunk * k;
    if (off >= bytes) break;
    const size_t len = std::min(chunk, bytes - off);
    cudaStream_t cs = n->copy_extra[k - 1];
    if (l.pf_packed[b] && !check_cuda(cudaStreamWaitEvent(cs, l.pf_packed[b], 0), "prefetch wait")) return -1;
    if (!check_cuda(cudaMemcpyAsync((char*)l.pf_buf[b] + off, (const char*)host + off, len, cudaMemcpyHostToDevice, cs), "prefetch copy") ||
        !check_cuda(cudaEventRecord(n->copy_part[k - 1], cs), "prefetch part record") ||
        !check_cuda(cudaStreamWaitEvent(n->copy_stream, n->copy_part[k - 1], 0), "prefetch part join")) return -1;
  }
  if (!check_cuda(cudaEventRecord(l.pf_copied[b], n->copy_stream), "prefetch record")) return -1;
  l.pf_rows = rows; l.pf_cols = cols;
  l.pf_f32[b] = f32;
  l.pf_ready = b;
  l.pf_slot = b ^ 1;
  return 0;
}
int kfp16_net_prefetch_input(kfp16_net* n, const char* input_name, const uint16_t* host, int rows, int cols) {
  return prefetch_common(n, input_name, host, rows, cols, false);
}
int kfp16_net_prefetch_input_f32(kfp16_net* n, const char* input_name, const float* host, int rows, int cols) {
  return prefetch_common(n, input_name, host, rows, cols, true);
}
int kfp16_net_commit_input(kfp16_net* n, const char* input_name) {
  if (!n || !input_name) { set_error("kfp16_net_commit_input: null argument"); return -1; }
  const int i = find_layer(n, input_name);
  if (i < 0 || n->layers[i].type != L_INPUT || n->layers[i].pf_ready < 0) { set_error("kfp16_net_commit_input: nothing prefetched for %s", input_name ? input_name : "?"); return -1; }
  Layer& l = n->layers[i];
  const int b = l.pf_ready;
  l.pf_ready = -1;
  if (!check_cuda(cudaStreamWaitEvent(n->ctx->stream, l.pf_copied[b], 0), "commit wait")) return -1;
  if (set_input_common(n, input_name, l.pf_buf[b], l.pf_rows, l.pf_cols, l.pf_f32[b])) return -1;
  return check_cuda(cudaEventRecord(l.pf_packed[b], n->ctx->stream), "commit record") ? 0 : -1;
}

int kfp16_net_set_input_compressed(kfp16_net* n, const char* input_name, const void* payload_host, size_t payload_size,
                                   const kfp16_cm_desc* descs, int count) {
  if (!n || !input_name || !payload_host || !descs) { set_error("kfp16_net_set_input_compressed: null argument"); return -1; }
  const int i = find_layer(n, input_name);
  if (i < 0 || n->layers[i].type != L_INPUT) { set_error("kfp16_net_set_input_compressed: no input layer named %s", input_name); return -1; }
  Layer& l = n->layers[i];
  if (l.per_seq || count != n->opts.n_seq) { set_error("kfp16_net_set_input_compressed: %s takes one matrix per sequence (%d)", input_name, n->opts.n_seq); return -1; }
  if (payload_size > n->stage_bytes) { set_error("kfp16_net_set_input_compressed: payload of %zu bytes exceeds the staging buffer", payload_size); return -1; }
  std::vector<kfp16_cm_desc> d(descs, descs + count);
  for (int q = 0; q < count; ++q) {
    if (d[q].rows != n->opts.seq_len || d[q].cols != l.out_dim) {
      set_error("kfp16_net_set_input_compressed: sequence %d is [%d x %d], expected [%d x %d]", q, d[q].rows, d[q].cols, n->opts.seq_len, l.out_dim);
      return -1;
    }
    d[q].dst_row = q * n->blk + n->halo;      // the sequence's real rows inside the padded layout
  }
  cudaStream_t st = n->ctx->stream;
  if (!check_cuda(cudaMemcpyAsync(n->stage_in, payload_host, payload_size, cudaMemcpyHostToDevice, st), "payload upload")) return -1;
  if (kfp16_decode_matrices(n->ctx, n->stage_in, payload_size, d.data(), count, l.out.p, l.out.cols, n->Tp)) return -1;
  // halo rows: replicate for spliced consumers, zero otherwise (as kfp16_net_set_input does)
  if (n->halo > 0) {
    const int rc = l.halo_mode == HALO_ZERO ? kfp16_zero_halo(n->ctx, l.out.p, l.out.cols, n->opts.n_seq, n->opts.seq_len, l.out.cols, n->halo)
                                            : kfp16_pad_edges(n->ctx, l.out.p, l.out.cols, n->opts.n_seq, n->opts.seq_len, l.out.cols, n->halo);
    if (rc) return -1;
  }
  return check_cuda(cudaStreamSynchronize(st), "input sync") ? 0 : -1;
}

int kfp16_net_forward(kfp16_net* n) {
  if (!n) { set_error("kfp16_net_forward: null network"); return -1; }
  for (auto& l : n->layers)
    if (forward_layer(n, l)) return -1;
  return 0;
}

static int read_dense(kfp16_net* n, const Layer& l, const Buf& b, uint16_t* host, int rows, int cols) {
  const int want_rows = l.per_seq ? n->opts.n_seq : n->T;
  if (!b.p) { set_error("layer %s: buffer not allocated (train = 0 or no gradient path)", l.name.c_str()); return -1; }
  if (rows != want_rows || cols != b.cols) { set_error("layer %s is [%d x %d], got [%d x %d]", l.name.c_str(), want_rows, b.cols, rows, cols); return -1; }
  const size_t bytes = (size_t)rows * cols * 2;
  if (l.per_seq) {
    if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "sync")) return -1;
    return check_cuda(cudaMemcpy(host, b.p, bytes, cudaMemcpyDeviceToHost), "download") ? 0 : -1;
  }
  if (bytes > n->stage_bytes) { set_error("layer %s: output exceeds the staging buffer", l.name.c_str()); return -1; }
  if (kfp16_unpack_rows(n->ctx, b.p, b.cols, 0, n->stage_in, n->opts.n_seq, n->opts.seq_len, n->halo, cols)) return -1;
  if (!check_cuda(cudaMemcpyAsync(host, n->stage_in, bytes, cudaMemcpyDeviceToHost, n->ctx->stream), "download")) return -1;
  return check_cuda(cudaStreamSynchronize(n->ctx->stream), "download sync") ? 0 : -1;
}
int kfp16_net_get_output(kfp16_net* n, const char* layer, uint16_t* host, int rows, int cols) {
  if (!n || !layer || !host) { set_error("kfp16_net_get_output: null argument"); return -1; }
  const int i = (layer[0] == 0) ? n->out_layer : find_layer(n, layer);
  if (i < 0) { set_error("kfp16_net_get_output: no layer %s", layer); return -1; }
  return read_dense(n, n->layers[i], n->layers[i].out, host, rows, cols);
}
int kfp16_net_get_grad(kfp16_net* n, const char* layer, uint16_t* host, int rows, int cols) {
  if (!n || !layer || !host) { set_error("kfp16_net_get_grad: null argument"); return -1; }
  const int i = (layer[0] == 0) ? n->out_layer : find_layer(n, layer);
  if (i < 0) { set_error("kfp16_net_get_grad: no layer %s", layer); return -1; }
  Layer& l = n->layers[i];
  // a gradient that was only written on the objective's output frames is read back in its dense form
  if (i == n->out_layer && n->out_g_sub > 1) {
    if (densify_rows(n, l.dout, n->out_g_row0, n->out_g_sub)) return -1;
    n->out_g_sub = 1; n->out_g_row0 = 0; l.g_sub = 1; l.g_row0 = 0;
  } else if (l.g_sub > 1) {
    if (densify_rows(n, l.dout, l.g_row0, l.g_sub)) return -1;
    l.g_sub = 1; l.g_row0 = 0;
  }
  return read_dense(n, l, l.dout, host, rows, cols);
}

// ReLU mask of a tdnnf / prefinal layer as one byte per element, dense real rows (tests: lets the
// oracle's backward use the same masks, so that only rounding-level differences remain)
int kfp16_net_get_mask(kfp16_net* n, const char* layer, uint8_t* host, int rows, int cols) {
  if (!n || !layer || !host) { set_error("kfp16_net_get_mask: null argument"); return -1; }
  const int i = find_layer(n, layer);
  if (i < 0 || !n->layers[i].mask) { set_error("kfp16_net_get_mask: layer %s has no ReLU mask", layer); return -1; }
  const Layer& l = n->layers[i];
  const int dim = l.type == L_PREFINAL ? l.big_dim : l.out_dim;
  const int prow = l.per_seq ? n->opts.n_seq : n->Tp;
  const int want_rows = l.per_seq ? n->opts.n_seq : n->T;
  if (rows != want_rows || cols != dim) { set_error("kfp16_net_get_mask: %s mask is [%d x %d]", layer, want_rows, dim); return -1; }
  // conv layers keep one mask row per (frame, height): [Tp*hout x fout] == [Tp x hout*fout] in memory
  const int sub_rows = l.type == L_CONV ? l.hout : 1;
  const int sub_cols = l.type == L_CONV ? l.fout : cols;
  std::vector<uint32_t> words((size_t)prow * sub_rows * l.mask_ld);
  if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "sync")) return -1;
  if (!check_cuda(cudaMemcpy(words.data(), l.mask, words.size() * 4, cudaMemcpyDeviceToHost), "mask download")) return -1;
  for (int r = 0; r < rows; ++r) {
    const int pr = l.per_seq ? r : (r / n->opts.seq_len) * n->blk + n->halo + (r % n->opts.seq_len);
    for (int c = 0; c < cols; ++c) {
      const size_t mrow = (size_t)pr * sub_rows + c / sub_cols;
      const int mc = c % sub_cols;
      host[(size_t)r * cols + c] = (words[mrow * l.mask_ld + (mc >> 5)] >> (mc & 31)) & 1u;
    }
  }
  return 0;
}

int kfp16_net_zero_grads(kfp16_net* n) {
  if (!n || !n->g32) { set_error("kfp16_net_zero_grads: network was created with train = 0"); return -1; }
  return check_cuda(cudaMemsetAsync(n->g32, 0, n->bucket * sizeof(float), n->ctx->stream), "zero grads") ? 0 : -1;
}

int kfp16_net_loss_half_sq(kfp16_net* n, const char* layer) {
  if (!n) { set_error("kfp16_net_loss_half_sq: null network"); return -1; }
  const int i = (!layer || layer[0] == 0) ? n->out_layer : find_layer(n, layer);
  if (i < 0) { set_error("kfp16_net_loss_half_sq: no such layer"); return -1; }
  Layer& l = n->layers[i];
  if (!l.dout.p) { set_error("kfp16_net_loss_half_sq: layer %s has no gradient buffer (train = 0?)", l.name.c_str()); return -1; }
  n->out_g_sub = 1; n->out_g_row0 = 0;
  if (l.per_seq) return kfp16_half_sq_loss(n->ctx, l.out.p, l.dout.p, n->opts.n_seq, 1, 0, l.out_dim, n->loss_dev);
  return kfp16_half_sq_loss(n->ctx, l.out.p, l.dout.p, n->opts.n_seq, n->opts.seq_len, n->halo, l.out_dim, n->loss_dev);
}

int kfp16_net_loss_chain(kfp16_net* n, const char* layer, kfp16_chain* chain, int subsampling, int left_context, float weight) {
  if (!n || !chain) { set_error("kfp16_net_loss_chain: null argument"); return -1; }
  const int i = (!layer || layer[0] == 0) ? n->out_layer : find_layer(n, layer);
  if (i < 0) { set_error("kfp16_net_loss_chain: no such layer"); return -1; }
  Layer& l = n->layers[i];
  if (!l.dout.p || l.per_seq) { set_error("kfp16_net_loss_chain: layer %s has no per-frame gradient buffer (train = 0?)", l.name.c_str()); return -1; }
  if (kfp16_chain_num_sequences(chain) != n->opts.n_seq) { set_error("kfp16_net_loss_chain: the chain object is for %d sequences, the network for %d", kfp16_chain_num_sequences(chain), n->opts.n_seq); return -1; }
  if (subsampling < 1 || left_context < 0 || left_context + (kfp16_chain_frames(chain) - 1) * subsampling >= n->opts.seq_len) {
    set_error("kfp16_net_loss_chain: %d output frames at %d + t*%d exceed the %d frames of a sequence", kfp16_chain_frames(chain), left_context, subsampling, n->opts.seq_len);
    return -1;
  }
  // The gradient is zero on every row that is not an output frame (and on the halo rows).  With a subsampling factor
  // that divides the rows of a sequence block, the rows row0 + k*subsampling of the whole padded matrix are exactly the
  // output frames plus a few rows per sequence behind its last frame: only those few are cleared, and the backward pass is
  // told which rows carry the gradient (Layer::g_sub) -- the output / prefinal layers then back-propagate a third of the
  // rows and nobody reads the rest.  Otherwise: clear everything, dense backward.
  const int row0 = n->halo + left_context, frames = kfp16_chain_frames(chain);
  const int per_blk = n->blk / subsampling;
  const bool by_rows = chain_by_rows(n, i, subsampling, frames);
  n->out_g_sub = 1; n->out_g_row0 = 0;
  if (by_rows) {
    for (int t = frames; t < per_blk; ++t) {
      const int rb = row0 + t * subsampling;                 // row inside the block (may fall into the next block's head halo)
      const int count = rb < n->blk ? n->opts.n_seq : n->opts.n_seq - 1;
      if (count > 0 && !check_cuda(cudaMemset2DAsync(l.dout.p + (size_t)rb * l.out_dim, (size_t)n->blk * l.out_dim * 2, 0, (size_t)l.out_dim * 2,
                                                       (size_t)count, n->ctx->stream), "chain gradient clear (rows behind the last output frame)")) return -1;
    }
    n->out_g_sub = subsampling; n->out_g_row0 = row0;
  } else if (!check_cuda(cudaMemsetAsync(l.dout.p, 0, l.dout.bytes(), n->ctx->stream), "chain gradient clear")) return -1;
  return kfp16_chain_loss(chain, l.out.p, l.dout.p, l.out_dim, n->blk, row0, subsampling, weight, n->loss_dev);
}
int kfp16_net_set_train_batchnorm(kfp16_net* n, int on, float momentum) {
  if (!n) { set_error("kfp16_net_set_train_batchnorm: null network"); return -1; }
  if (on && !n->g32) { set_error("kfp16_net_set_train_batchnorm: network was created with train = 0"); return -1; }
  if (on && !(momentum >= 0.f && momentum <= 1.f)) { set_error("kfp16_net_set_train_batchnorm: momentum must be in [0, 1]"); return -1; }
  if (on && !n->bn_stats) {
    if (!check_cuda(cudaMalloc((void**)&n->bn_stats, std::max<size_t>(n->bn_stats_dim, 8) * 2 * sizeof(float)), "batch-norm statistics buffer")) return -1;
    n->allocs.push_back(n->bn_stats);
  }
  const bool was_on = n->train_bn;
  n->train_bn = on != 0;
  if (on) n->bn_momentum = momentum;
  if (was_on && !on) {
    // back to the stored statistics: the folded scale / shift still hold the last minibatch's, so fold the (updated) running
    // statistics again -- what an inference pass or the fused training epilogues read from now on
    for (auto& l : n->layers)
      for (BNorm* bn : {&l.bn, &l.bn2}) {
        if (!bn->present) continue;
        const bool rms = bn->rms_only;
        if (kfp16_bn_fold(n->ctx, bn->mean, bn->var, rms ? nullptr : bn->gamma, rms ? nullptr : bn->beta, bn->eps, bn->target_rms, bn->dim, bn->scale, bn->shift)) return -1;
        if (kfp16_scale_f32(n->ctx, bn->scale, bn->scale_bwd, bn->dim, bn->bwd_mul)) return -1;
        if (!retile_bn(n, *bn, bn->tiles)) return -1;
      }
    if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "bn refold sync")) return -1;
  }
  return 0;
}
int kfp16_net_set_bn_stats_hook(kfp16_net* n, kfp16_bn_stats_hook hook, void* user, int world) {
  if (!n || world < 1) { set_error("kfp16_net_set_bn_stats_hook: bad argument"); return -1; }
  n->bn_hook = hook; n->bn_hook_user = user; n->bn_world = hook ? world : 1;
  return 0;
}
int kfp16_net_get_bn(kfp16_net* n, const char* layer, const char* which, float* mean, float* var, int dim) {
  if (!n || !layer || !mean || !var) { set_error("kfp16_net_get_bn: null argument"); return -1; }
  const int i = find_layer(n, layer);
  if (i < 0) { set_error("kfp16_net_get_bn: no layer %s", layer); return -1; }
  Layer& l = n->layers[i];
  BNorm* bn = &l.bn;
  if (which && (!strcmp(which, "BN") || !strcmp(which, "bn2")) && l.type == L_PREFINAL) bn = &l.bn2;
  if (!bn->present || bn->dim != dim) { set_error("kfp16_net_get_bn: layer %s has no batch-norm of dim %d", layer, dim); return -1; }
  if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "sync")) return -1;
  return check_cuda(cudaMemcpy(mean, bn->mean, (size_t)dim * 4, cudaMemcpyDeviceToHost), "bn mean download") &&
         check_cuda(cudaMemcpy(var, bn->var, (size_t)dim * 4, cudaMemcpyDeviceToHost), "bn var download") ? 0 : -1;
}
int kfp16_net_set_spec_augment(kfp16_net* n, int on) {
  if (!n) { set_error("kfp16_net_set_spec_augment: null network"); return -1; }
  if (on && !n->g32) { set_error("kfp16_net_set_spec_augment: network was created with train = 0"); return -1; }
  n->spec_augment = on != 0;
  return 0;
}
int kfp16_net_set_overlap_loss(kfp16_net* n, int on) {
  if (!n) { set_error("kfp16_net_set_overlap_loss: null network"); return -1; }
  n->overlap_loss = on != 0;
  return 0;
}
int kfp16_net_set_fuse_conv_backward(kfp16_net* n, int on) {
  if (!n) { set_error("kfp16_net_set_fuse_conv_backward: null network"); return -1; }
  n->fuse_conv_bwd = on != 0;
  return 0;
}
int kfp16_net_set_sparse_output_grad(kfp16_net* n, int on) {
  if (!n) { set_error("kfp16_net_set_sparse_output_grad: null network"); return -1; }
  n->sparse_out_grad = on != 0;
  return 0;
}
int kfp16_net_set_chain(kfp16_net* n, kfp16_chain* chain, int subsampling, int left_context, float weight) {
  if (!n) { set_error("kfp16_net_set_chain: null network"); return -1; }
  n->chain = chain; n->chain_sub = subsampling; n->chain_left = left_context; n->chain_weight = weight;
  return 0;
}

int kfp16_net_set_output_grad(kfp16_net* n, const char* layer, const uint16_t* host, int rows, int cols) {
  if (!n || !host) { set_error("kfp16_net_set_output_grad: null argument"); return -1; }
  const int i = (!layer || layer[0] == 0) ? n->out_layer : find_layer(n, layer);
  if (i < 0) { set_error("kfp16_net_set_output_grad: no such layer"); return -1; }
  Layer& l = n->layers[i];
  if (!l.dout.p) { set_error("kfp16_net_set_output_grad: layer %s has no gradient buffer", l.name.c_str()); return -1; }
  n->out_g_sub = 1; n->out_g_row0 = 0;
  const int want_rows = l.per_seq ? n->opts.n_seq : n->T;
  if (rows != want_rows || cols != l.out_dim) { set_error("kfp16_net_set_output_grad: shape mismatch"); return -1; }
  const size_t bytes = (size_t)rows * cols * 2;
  if (bytes > n->stage_bytes) { set_error("kfp16_net_set_output_grad: exceeds the staging buffer"); return -1; }
  if (!check_cuda(cudaMemcpyAsync(n->stage_in, host, bytes, cudaMemcpyHostToDevice, n->ctx->stream), "grad upload")) return -1;
  int rc;
  if (l.per_seq) rc = check_cuda(cudaMemcpyAsync(l.dout.p, n->stage_in, bytes, cudaMemcpyDeviceToDevice, n->ctx->stream), "grad copy") ? 0 : -1;
  else rc = kfp16_pack_rows(n->ctx, n->stage_in, l.dout.p, l.out_dim, n->opts.n_seq, n->opts.seq_len, n->halo, cols, 0);
  if (rc) return -1;
  return check_cuda(cudaStreamSynchronize(n->ctx->stream), "grad sync") ? 0 : -1;
}

// back-propagate through layers [lo, hi) in reverse order; `begin` starts a new backward pass
static int backward_range(kfp16_net* n, int hi, int lo, bool begin) {
  if (begin) {
    for (auto& l : n->layers) { l.grads_seen = 0; l.g_sub = 1; l.g_row0 = 0; l.dz_in_dout = false; }
    n->flops_bwd_skipped = 0;
    n->layers[n->out_layer].grads_seen = 1;
    n->layers[n->out_layer].g_sub = n->out_g_sub;
    n->layers[n->out_layer].g_row0 = n->out_g_row0;
  }
  for (int i = hi - 1; i >= lo; --i) {
    Layer& l = n->layers[i];
    if (!l.needs_grad || l.type == L_INPUT) continue;
    if (l.grads_seen == 0) continue;   // nothing flowed into this layer
    if (backward_layer(n, l)) return -1;
  }
  return 0;
}

int kfp16_net_backward(kfp16_net* n) {
  if (!n || !n->g32) { set_error("kfp16_net_backward: network was created with train = 0"); return -1; }
  if (backward_range(n, (int)n->layers.size(), 0, true)) return -1;
  if (flush_wgrads(n)) return -1;
  return 0;
}

int kfp16_net_sgd_step(kfp16_net* n, float grad_scale, int round_grad) {
  if (!n || !n->g32) { set_error("kfp16_net_sgd_step: network was created with train = 0"); return -1; }
  // the usual case (the scale the network was created with): every hyper-parameter is read from the device block, so a
  // captured graph of this launch follows kfp16_net_set_lr / set_momentum; any other scale is passed by value
  if (grad_scale == n->hp_host[2])
    return kfp16_sgd_update_flat_hp(n->ctx, n->w32, n->w16, n->g32, 1, round_grad, n->vel, n->hp_dev, n->bucket);
  return kfp16_sgd_update_flat(n->ctx, n->w32, n->w16, n->g32, 1, round_grad, grad_scale, n->vel, n->hp_host[0], n->hp_host[1], n->bucket);
}
// g16 = half(g32 * grad_scale): the FP16 gradient tensors of the reference (AffineBackwardWeights returns FP16,
// internal/gpu/backward_ops.go:195-225) as ONE flat bucket -- what the data-parallel exchange all-reduces (half the bytes
// of the FP32 bucket).  kfp16_net_sgd_step_f16 then applies ops_sgd_update's arithmetic to it (scale 1).
int kfp16_net_grads_to_f16(kfp16_net* n) {
  if (!n || !n->g32) { set_error("kfp16_net_grads_to_f16: network was created with train = 0"); return -1; }
  return kfp16_scale_f32_to_f16(n->ctx, n->g32, n->g16, n->bucket, n->hp_dev + 2);
}
void* kfp16_net_grads_f16(kfp16_net* n) { return n ? (void*)n->g16 : nullptr; }
int kfp16_net_sgd_step_f16(kfp16_net* n) {
  if (!n || !n->g16) { set_error("kfp16_net_sgd_step_f16: network was created with train = 0"); return -1; }
  return kfp16_sgd_update_flat_hp(n->ctx, n->w32, n->w16, n->g16, 0, 0, n->vel, n->hp_dev + 4, n->bucket);
}
static int upload_hp(kfp16_net* n) {
  // hp_dev[0..2] = {lr, momentum, grad_scale}; hp_dev[4..6] = {lr, momentum, 1} for the FP16-gradient update
  float h[8] = {n->hp_host[0], n->hp_host[1], n->hp_host[2], 0.f, n->hp_host[0], n->hp_host[1], 1.0f, 0.f};
  // stream-ordered: takes effect for every update queued after this call, captured graphs included
  return check_cuda(cudaMemcpyAsync(n->hp_dev, h, sizeof(h), cudaMemcpyHostToDevice, n->ctx->stream), "hyper-parameter upload") &&
         check_cuda(cudaStreamSynchronize(n->ctx->stream), "hyper-parameter upload sync") ? 0 : -1;
}
int kfp16_net_set_lr(kfp16_net* n, float lr) {
  if (!n) { set_error("kfp16_net_set_lr: null network"); return -1; }
  n->opts.lr = lr;
  if (!n->hp_dev) return 0;
  n->hp_host[0] = lr;
  return upload_hp(n);
}
int kfp16_net_set_momentum(kfp16_net* n, float momentum) {
  if (!n) { set_error("kfp16_net_set_momentum: null network"); return -1; }
  n->opts.momentum = momentum;
  if (!n->hp_dev) return 0;
  n->hp_host[1] = momentum;
  return upload_hp(n);
}
float kfp16_net_get_lr(const kfp16_net* n) { return n ? n->opts.lr : 0.f; }
// dropout: the mask of layer i at element (padded row, col) is kfp16_dropout_uniform(seed ^ i*0x9E3779B9, row, col) > p
int kfp16_net_set_dropout_seed(kfp16_net* n, uint32_t seed) {
  if (!n || !n->seed_dev) { set_error("kfp16_net_set_dropout_seed: network was created with train = 0"); return -1; }
  return check_cuda(cudaMemcpyAsync(n->seed_dev, &seed, 4, cudaMemcpyHostToDevice, n->ctx->stream), "seed upload") &&
         check_cuda(cudaStreamSynchronize(n->ctx->stream), "seed upload sync") ? 0 : -1;
}
int kfp16_net_get_dropout_seed(kfp16_net* n, uint32_t* seed) {
  if (!n || !n->seed_dev || !seed) { set_error("kfp16_net_get_dropout_seed: bad argument"); return -1; }
  if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "sync")) return -1;
  return check_cuda(cudaMemcpy(seed, n->seed_dev, 4, cudaMemcpyDeviceToHost), "seed download") ? 0 : -1;
}
int kfp16_net_read_loss(kfp16_net* n, float* loss) {
  if (!n || !loss) { set_error("kfp16_net_read_loss: null argument"); return -1; }
  if (!check_cuda(cudaMemcpyAsync(loss, n->loss_dev, 4, cudaMemcpyDeviceToHost, n->ctx->stream), "loss download")) return -1;
  if (!check_cuda(cudaMemsetAsync(n->loss_dev, 0, 4, n->ctx->stream), "loss reset")) return -1;
  return check_cuda(cudaStreamSynchronize(n->ctx->stream), "loss sync") ? 0 : -1;
}

// Pipelined form of kfp16_net_read_loss: the copy of the loss accumulated so far (and its reset) is queued behind the
// work already in the stream and lands in pinned slot `slot` (0/1); kfp16_net_wait_loss blocks on that copy only, so a
// training loop can queue minibatch i+1 before it looks at the loss of minibatch i and the GPU never drains.
int kfp16_net_read_loss_async(kfp16_net* n, int slot) {
  if (!n || slot < 0 || slot > 1) { set_error("kfp16_net_read_loss_async: slot must be 0 or 1"); return -1; }
  if (!n->loss_pinned) {
    if (!check_cuda(cudaMallocHost((void**)&n->loss_pinned, 2 * sizeof(float)), "cudaMallocHost (loss slots)")) return -1;
    n->loss_pinned[0] = n->loss_pinned[1] = 0.f;
    for (int i = 0; i < 2; ++i)
      if (!check_cuda(cudaEventCreateWithFlags(&n->loss_ev[i], cudaEventDisableTiming), "loss event")) return -1;
  }
  if (!check_cuda(cudaMemcpyAsync(n->loss_pinned + slot, n->loss_dev, 4, cudaMemcpyDeviceToHost, n->ctx->stream), "loss download")) return -1;
  if (!check_cuda(cudaMemsetAsync(n->loss_dev, 0, 4, n->ctx->stream), "loss reset")) return -1;
  return check_cuda(cudaEventRecord(n->loss_ev[slot], n->ctx->stream), "loss event record") ? 0 : -1;
}
int kfp16_net_wait_loss(kfp16_net* n, int slot, float* loss) {
  if (!n || !loss || slot < 0 || slot > 1 || !n->loss_pinned) { set_error("kfp16_net_wait_loss: no read queued for this slot"); return -1; }
  if (!check_cuda(cudaEventSynchronize(n->loss_ev[slot]), "loss event sync")) return -1;
  *loss = n->loss_pinned[slot];
  return 0;
}

int kfp16_net_capture(kfp16_net* n, int phases) {
  if (!n || phases < 1 || phases > 15) { set_error("kfp16_net_capture: phases is a bitmask of 1 (step), 2 (SGD), 4 (FP16 gradient export), 8 (SGD from FP16 gradients)"); return -1; }
  if (!n->ctx->stream) { set_error("kfp16_net_capture: graph capture needs a non-default stream (kfp16_ctx_set_stream)"); return -1; }
  if ((phases & ~1) && !n->g32) { set_error("kfp16_net_capture: network was created with train = 0"); return -1; }
  if ((phases & 1) && n->train_bn && n->bn_hook) { set_error("kfp16_net_capture: a host hook sums the batch-norm statistics over the ranks between kernels -- run the step eagerly (kfp16_net_forward / loss / backward)"); return -1; }
  cudaStream_t st = n->ctx->stream;
  // One eager pass first: kernel attributes (dynamic shared memory opt-in) and the grouped weight-gradient tables are
  // set up outside the capture.  Its side effects on the training state are undone: master / FP16 weights, velocities,
  // both gradient buckets and the loss accumulator are snapshotted before and restored after, so capturing neither
  // takes an optimiser step nor disturbs a pending loss read (activations are scratch and simply recomputed).
  struct Snap { void* dst; const void* src; size_t bytes; };
  std::vector<Snap> snaps;
  char* save = nullptr;
  if (n->g32) {
    const size_t b = n->bucket;
    const size_t total = b * (4 + 4 + 4 + 2 + 2) + 256;
    if (!check_cuda(cudaMalloc((void**)&save, total), "cudaMalloc (capture snapshot)")) return -1;
    char* q = save;
    auto add = [&](void* live, size_t bytes) { snaps.push_back({live, q, bytes}); q += bytes; };
    add(n->w32, b * 4); add(n->vel, b * 4); add(n->g32, b * 4); add(n->w16, b * 2); add(n->g16, b * 2); add(n->loss_dev, 4);
    for (const Snap& sn : snaps)
      if (!check_cuda(cudaMemcpyAsync(const_cast<void*>(sn.src), sn.dst, sn.bytes, cudaMemcpyDeviceToDevice, st), "capture snapshot")) { cudaFree(save); return -1; }
  }
  int rc = run_phases(n, phases);
  for (const Snap& sn : snaps)
    if (!check_cuda(cudaMemcpyAsync(sn.dst, sn.src, sn.bytes, cudaMemcpyDeviceToDevice, st), "capture restore")) rc = -1;
  const bool synced = check_cuda(cudaStreamSynchronize(st), "pre-capture sync");
  if (save) cudaFree(save);
  if (rc || !synced) return -1;
  const unsigned long long before = kfp16_launch_count();
  if (!check_cuda(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture")) return -1;
  rc = run_phases(n, phases);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &g);
  if (rc) { if (g) cudaGraphDestroy(g); return -1; }
  if (!check_cuda(e, "cudaStreamEndCapture")) return -1;
  if (const char* dot = getenv("KFP16_GRAPH_DOT")) {   // debugging aid: <path>.<phases>
    char path[512];
    snprintf(path, sizeof(path), "%s.%d", dot, phases);
    cudaGraphDebugDotPrint(g, path, 0);
  }
  if (n->graph.count(phases) && n->graph[phases]) cudaGraphExecDestroy(n->graph[phases]);
  cudaGraphExec_t ge = nullptr;
  const bool ok = check_cuda(cudaGraphInstantiate(&ge, g, 0), "cudaGraphInstantiate");
  cudaGraphDestroy(g);
  n->graph[phases] = ge;
  n->graph_launches[phases] = (int)(kfp16_launch_count() - before);
  return ok ? 0 : -1;
}
// ---- segmented step: segment 0 = zero grads + forward + objective + backward of the top layers, segment k > 0 =
// backward of the next group of layers.  When segment k has run, the gradient-bucket range kfp16_net_segment_grads(k)
// is final, so a data-parallel loop can all-reduce it while segment k+1 computes (parameters sit in the bucket in
// layer order and the backward pass walks the layers in reverse).
static int first_param_of_layer(const Layer& l) {
  int m = -1;
  for (int p : {l.pW, l.pB, l.pLin, l.pAff, l.pAffB, l.pBig, l.pBigB, l.pSmall})
    if (p >= 0 && (m < 0 || p < m)) m = p;
  return m;
}
static int run_segment(kfp16_net* n, int seg) {
  const int hi = seg == 0 ? (int)n->layers.size() : n->seg_lo[seg - 1];
  if (seg == 0) {      // exactly what the whole-step graph starts with
    if (kfp16_bump_counter(n->ctx, n->seed_dev) || kfp16_net_zero_grads(n) || forward_and_objective(n)) return -1;
  }
  if (backward_range(n, hi, n->seg_lo[seg], seg == 0)) return -1;
  if (flush_wgrads(n)) return -1;      // the segment's gradient range must be complete when it ends
  if (n->seg_export_f16 && n->seg_cnt[seg] > 0)
    return kfp16_scale_f32_to_f16(n->ctx, n->g32 + n->seg_off[seg], n->g16 + n->seg_off[seg], n->seg_cnt[seg], n->hp_dev + 2);
  return 0;
}
// cut_layers: NULL = cut at equal shares of the parameter count; else a comma-separated list of layer names, top of the
// network first: segment k back-propagates down to (and including) the k-th named layer, the last segment takes the rest.
// tail_max_ctas > 0: segments 1.. are captured with their persistent grids limited to that many CTAs (the all-reduce
// of the previous segment's gradients runs beside them and needs SMs of its own).
int kfp16_net_capture_segments_ex(kfp16_net* n, int nseg, const char* cut_layers, int export_f16, int tail_max_ctas) {
  if (!n || !n->g32 || nseg < 1) { set_error("kfp16_net_capture_segments: needs a training network and nseg >= 1"); return -1; }
  if (!n->ctx->stream) { set_error("kfp16_net_capture_segments: graph capture needs a non-default stream"); return -1; }
  const int L = (int)n->layers.size();
  std::vector<size_t> first(L + 1, n->bucket);
  for (int i = L - 1; i >= 0; --i) {
    const int p = first_param_of_layer(n->layers[i]);
    first[i] = p >= 0 ? n->params[p].off : first[i + 1];
  }
  for (cudaGraphExec_t g : n->seg_graph) if (g) cudaGraphExecDestroy(g);
  n->seg_graph.clear(); n->seg_launches.clear(); n->seg_lo.clear(); n->seg_off.clear(); n->seg_cnt.clear();
  n->seg_export_f16 = export_f16 != 0;
  std::vector<int> cuts;     // explicit lower bounds (layer indices), top first
  if (cut_layers && cut_layers[0]) {
    std::stringstream ss(cut_layers);
    std::string nm;
    while (std::getline(ss, nm, ',')) {
      nm = trim(nm);
      if (nm.empty()) continue;
      const int li = find_layer(n, nm);
      if (li < 0) { set_error("kfp16_net_capture_segments: no layer named %s", nm.c_str()); return -1; }
      if (!cuts.empty() && li >= cuts.back()) { set_error("kfp16_net_capture_segments: cut layers must be listed top of the network first"); return -1; }
      cuts.push_back(li);
    }
    nseg = (int)cuts.size() + 1;
  }
  int hi = L;
  for (int k = 0; k < nseg && hi > 0; ++k) {
    int lo = 0;
    if (!cuts.empty()) {
      lo = k < (int)cuts.size() ? cuts[k] : 0;
    } else if (k + 1 < nseg) {
      // cut points: equal shares of the parameter count, counted from the top of the network
      const size_t target = (size_t)((double)n->bucket * (nseg - 1 - k) / nseg);   // bucket offset where this segment should start
      lo = hi - 1;
      while (lo > 0 && first[lo] > target) --lo;
    }
    if (first[lo] == first[hi]) {          // no parameters in [lo, hi)
      if (lo > 0 && cuts.empty()) continue; // try again with the next (lower) target
      if (cuts.empty()) break;              // only parameter-free layers are left: the last segment takes them (below)
    }
    n->seg_lo.push_back(lo);
    n->seg_off.push_back(first[lo]);
    n->seg_cnt.push_back(first[hi] - first[lo]);
    hi = lo;
  }
  if (n->seg_lo.empty()) {
    n->seg_lo.push_back(0); n->seg_off.push_back(first[0]); n->seg_cnt.push_back(first[L] - first[0]);
  } else if (n->seg_lo.back() != 0) {       // the last segment back-propagates down to the first layer
    n->seg_cnt.back() += n->seg_off.back() - first[0];
    n->seg_off.back() = first[0];
    n->seg_lo.back() = 0;
  }
  const int S = (int)n->seg_lo.size();
  const int saved_max = n->ctx->max_ctas;
  auto limit = [&](int k) { n->ctx->max_ctas = (k > 0 && tail_max_ctas > 0) ? tail_max_ctas : saved_max; };
  // eager pass: kernel attributes are set outside the capture (this pass leaves a gradient in the buckets and adds to
  // the loss accumulator; weights are untouched)
  for (int k = 0; k < S; ++k) {
    limit(k);
    if (run_segment(n, k)) { n->ctx->max_ctas = saved_max; return -1; }
  }
  n->ctx->max_ctas = saved_max;
  if (!check_cuda(cudaStreamSynchronize(n->ctx->stream), "pre-capture sync")) return -1;
  for (int k = 0; k < S; ++k) {
    const unsigned long long before = kfp16_launch_count();
    limit(k);
    if (!check_cuda(cudaStreamBeginCapture(n->ctx->stream, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture")) { n->ctx->max_ctas = saved_max; return -1; }
    const int rc = run_segment(n, k);
    n->ctx->max_ctas = saved_max;
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(n->ctx->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return -1; }
    if (!check_cuda(e, "cudaStreamEndCapture")) return -1;
    cudaGraphExec_t ge = nullptr;
    const bool ok = check_cuda(cudaGraphInstantiate(&ge, g, 0), "cudaGraphInstantiate");
    cudaGraphDestroy(g);
    if (!ok) return -1;
    n->seg_graph.push_back(ge);
    n->seg_launches.push_back((int)(kfp16_launch_count() - before));
  }
  return S;
}
int kfp16_net_capture_segments(kfp16_net* n, int nseg) { return kfp16_net_capture_segments_ex(n, nseg, nullptr, 0, 0); }
int kfp16_net_launch_segment(kfp16_net* n, int seg) {
  if (!n || seg < 0 || seg >= (int)n->seg_graph.size()) { set_error("kfp16_net_launch_segment: segment %d not captured", seg); return -1; }
  if (!check_cuda(cudaGraphLaunch(n->seg_graph[seg], n->ctx->stream), "cudaGraphLaunch")) return -1;
  count_launch(n->seg_launches[seg]);
  return 0;
}
int kfp16_net_segment_grads(const kfp16_net* n, int seg, size_t* first_elem, size_t* count) {
  if (!n || seg < 0 || seg >= (int)n->seg_off.size() || !first_elem || !count) { set_error("kfp16_net_segment_grads: bad argument"); return -1; }
  *first_elem = n->seg_off[seg];
  *count = n->seg_cnt[seg];
  return 0;
}

int kfp16_net_launch(kfp16_net* n, int phases) {
  if (!n || !n->graph.count(phases) || !n->graph[phases]) { set_error("kfp16_net_launch: phases %d not captured", phases); return -1; }
  if (!check_cuda(cudaGraphLaunch(n->graph[phases], n->ctx->stream), "cudaGraphLaunch")) return -1;
  count_launch(n->graph_launches[phases]);
  return 0;
}
int kfp16_net_launches_per_step(const kfp16_net* n, int phases) {
  if (!n) return 0;
  auto it = n->graph_launches.find(phases);
  return it == n->graph_launches.end() ? 0 : it->second;
}

}  // extern "C"
