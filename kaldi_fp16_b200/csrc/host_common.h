// Host-side plumbing shared by the C-ABI translation units: the thread-local error buffer
// (same convention as the reference: int 0 / -1 + message read by *_last_error(),
// /root/reference/cpp/cuda/ops.cu:12-20,329-330), the launch counter and the context.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

struct kfp16_ctx {
  int device = 0;
  int num_sms = 148;
  int max_ctas = 0;
  cudaStream_t stream = nullptr;
  // split-K workspace owned by the context (used when the caller passes none)
  float* ws = nullptr;
  size_t ws_bytes = 0;
  // profile mode: a CUDA event pair around every GEMM launch (bench.py roofline leg)
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;     // start/stop pairs
  std::vector<double> prof_flops;       // per pair
  std::vector<std::string> prof_desc;   // per pair: shape / tiling of the launch
};

namespace kfp16 {

void set_error(const char* fmt, ...);
const char* get_error();   // nullptr when clear
void clear_error();

extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

cudaStream_t default_stream();
void count_gemm_kind(int ek);   // per-EpiKind launch counter (kfp16_gemm_kind_launches)

// returns false (and sets the error) when a launch / runtime call failed
bool check_launch(const char* what);
bool check_cuda(cudaError_t e, const char* what);

// Programmatic dependent launch (KFP16_PDL=0 disables): kernels that call griddep_wait() before their first
// global access are launched with the stream-serialization attribute, so consecutive launches overlap their
// launch latency / prologue with the predecessor's tail (also inside captured graphs: programmatic edges).
bool pdl_enabled();
// once per kernel: ask for the maximum shared-memory carve-out, the configuration the GEMM kernels need, so that the
// SMs are not re-partitioned (L1 <-> shared) every time a small helper kernel runs between two GEMMs
void prefer_max_smem_carveout(const void* kern);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  prefer_max_smem_carveout(reinterpret_cast<const void*>(kern));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// stream-explicit forms of a few reference ops, used by the network executor (elementwise.cu)
int softmax_on_stream(cudaStream_t stream, void* data, int rows, int cols, bool log);
int ops_concat_cols_on(cudaStream_t stream, void* dst, int T, int dst_cols, const void* src, int src_cols, int dst_col_offset);
int ops_slice_cols_on(cudaStream_t stream, const void* src, int T, int src_cols, void* dst, int dst_cols, int src_col_offset);
int ops_slice_add_on(cudaStream_t stream, const void* src, int T, int src_cols, void* dst, int dst_cols, int src_col_offset);
int ops_combine_feature_maps_on(cudaStream_t stream, void* data, int T, int total_dim, int height, int nf1, int nf2, int inverse);

}  // namespace kfp16
