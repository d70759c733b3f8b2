// Gradient exchange of the data-parallel step over NVLink peer memory: ONE kernel per rank that reduces and
// redistributes the FP16 gradient bucket through the peers' memory, instead of a library collective between the step
// graph and the SGD graph.
//
// The reference has no multi-GPU path (cpp/cuda/bridge.cu:38-47 pins device 0); the exchange sits where
// internal/nnet/train_step.go:212-221 goes from Backward to the optimizer updates, and carries what the reference's
// gradients are: FP16 tensors (internal/gpu/backward_ops.go:195-225).
//
// Every rank maps the other ranks' buckets and flag blocks (CUDA IPC handles, exchanged by the host).  The kernel, launched by every rank on its own stream after its gradients are in its bucket:
//   A  tell every peer "my bucket is ready" (flag store into the peer's flag block), wait for every peer's flag;
//   1  for the slice of the bucket this rank owns: load the slice of EVERY rank (16-byte loads over NVLink), add in FP32
//      in rank order, round once to FP16 and store the result into EVERY rank's bucket (16-byte stores over NVLink);
//   B  when the last CTA of the grid has finished: tell every peer "my slice is in your bucket", wait for every peer.
// Each element is reduced by exactly one rank and the same rounded value lands everywhere: the buckets -- hence the
// weights after the update -- stay bit-identical across the ranks.  2 x (N-1)/N of the bucket crosses NVLink in each
// direction per rank, the two directions at the same time.
// Flags carry a step counter kept in device memory, so the launch can sit in a captured graph and nothing is ever reset.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/kaldi_fp16_nnet.h"
#include "host_common.h"

using namespace kfp16;

namespace {

constexpr int kMaxPeers = 8;
// the flag block of a rank (32-bit words, in that rank's memory)
constexpr int kFlagA = 0;       // [kFlagA + p]: step number up to which rank p's bucket has been ready
constexpr int kFlagB = 16;      // [kFlagB + p]: step number up to which rank p has delivered its slice here
constexpr int kEpoch = 32;      // exchanges completed by this rank
constexpr int kDone = 33;       // CTAs of the running exchange that have stored their part
constexpr int kError = 34;      // set when a wait ran into the time limit (a peer never arrived)
constexpr int kFlagWords = 64;
constexpr int kChannels = 4;

struct PeerArgs {
  __half* buf[kMaxPeers];
  unsigned* flag[kMaxPeers];
  int rank, world;
  unsigned long long first, count;  // the FP16 elements [first, first + count) of the bucket are exchanged
  unsigned long long timeout_ns;
  int flag_off;                     // channel * kFlagWords: exchanges on different channels may be in flight together
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// system-scope accesses: never served from this SM's L1 (the lines are written by other GPUs while the kernel runs).
// Weak loads / stores (legal behind the flag acquire and in front of the fence + flag release) and L1::no_allocate
// measured the same: 69.5-70.9 us at 2 ranks, 124-126 us at 8 (profiles/r02_peer_allreduce.txt) -- NVLink bound.
__device__ __forceinline__ uint4 ld_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned short ld_sys_u16(const void* p) {
  unsigned short v;
  asm volatile("ld.relaxed.sys.global.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_u16(void* p, unsigned short v) {
  asm volatile("st.relaxed.sys.global.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// threads < world each wait for one peer's flag to reach `epoch`; the block continues together.  Relaxed polls and ONE
// acquire fence at the end: an acquire load per poll would invalidate the SM's L1 every time, and the exchange may share
// its SMs with the kernels of the backward pass.
__device__ __forceinline__ void wait_peers(unsigned* lf, int base, int world, unsigned epoch, unsigned long long timeout_ns) {
  if ((int)threadIdx.x < world) {
    const unsigned* f = lf + base + threadIdx.x;
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while ((int)(ld_relaxed_sys(f) - epoch) < 0) {
      if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) {
        atomicExch(lf + kError, 1u);     // (lf already points at the channel's block)
        break;
      }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  __syncthreads();
}

__device__ __forceinline__ void add8(float (&acc)[8], uint4 x) {
  const __half2* h = reinterpret_cast<const __half2*>(&x);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    acc[2 * j] += f.x;
    acc[2 * j + 1] += f.y;
  }
}

// WORLD: compile-time bound of the rank loop (a.world <= WORLD), U vectors of 8 elements per thread and pass
template <int WORLD, int U>
__global__ void __launch_bounds__(512) peer_allreduce_f16_kernel(const PeerArgs a) {
  unsigned* lf = a.flag[a.rank] + a.flag_off;
  // every CTA reads the counter before the last one to finish advances it
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(lf + kEpoch) + 1u;
  if (blockIdx.x == 0 && (int)threadIdx.x < a.world) st_release_sys(a.flag[threadIdx.x] + a.flag_off + kFlagA + a.rank, epoch);
  wait_peers(lf, kFlagA, a.world, epoch, a.timeout_ns);

  // whole 16-byte vectors of the range are split over the ranks; the up to 7 + 7 elements in front of / behind them are
  // the last rank's
  const unsigned long long end = a.first + a.count;
  const unsigned long long vf = (a.first + 7) >> 3, vl = end >> 3;         // vectors [vf, vl)
  const unsigned long long nvec = vl > vf ? vl - vf : 0;
  const unsigned long long per = (nvec + a.world - 1) / a.world;
  const unsigned long long v0 = vf + per * a.rank;
  const unsigned long long v1 = v0 + per < vf + nvec ? v0 + per : vf + nvec;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long v = v0 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += stride * U) {
    uint4 x[U][WORLD];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int p = 0; p < WORLD; ++p)
        if (p < a.world && v + u * stride < v1) x[u][p] = ld_sys_v4(a.buf[p] + ((v + u * stride) << 3));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (v + u * stride >= v1) break;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int p = 0; p < WORLD; ++p)        // rank order: the sum does not depend on which rank owns the slice
        if (p < a.world) add8(acc, x[u][p]);
      uint4 r;
      __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
#pragma unroll
      for (int p = 0; p < WORLD; ++p)
        if (p < a.world) st_sys_v4(a.buf[p] + ((v + u * stride) << 3), r);
    }
  }
  if (a.rank == a.world - 1 && blockIdx.x == 0 && threadIdx.x < 16) {
    // thread i < 8: element first + i while in front of the first whole vector; thread 8 + i: element vl*8 + i
    unsigned long long e;
    bool mine;
    if (nvec == 0) {                       // no whole vector: at most 14 elements, one thread each
      e = a.first + threadIdx.x;
      mine = e < end;
    } else if (threadIdx.x < 8) {
      e = a.first + threadIdx.x;
      mine = e < (vf << 3);
    } else {
      e = (vl << 3) + (threadIdx.x - 8);
      mine = e < end;
    }
    if (mine) {
      float acc = 0.f;
      for (int p = 0; p < a.world; ++p) {
        __half_raw raw;
        raw.x = ld_sys_u16(a.buf[p] + e);
        acc += __half2float(__half(raw));
      }
      const __half_raw out = static_cast<__half_raw>(__float2half_rn(acc));
      for (int p = 0; p < a.world; ++p) st_sys_u16(a.buf[p] + e, out.x);
    }
  }

  // B: the peers may read their buckets once EVERY CTA of this grid has stored its part
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(lf + kDone, 1u);
    s_last = prev == gridDim.x - 1;
    if (s_last) {
      lf[kDone] = 0u;
      lf[kEpoch] = epoch;
      __threadfence_system();
    }
  }
  __syncthreads();
  // only that CTA stays: the kernel -- and with it the stream -- must not complete before every peer has delivered, but
  // CTAs that merely wait would keep the SMs from CTAs that have not run yet (a grid larger than one wave)
  if (!s_last) return;
  if ((int)threadIdx.x < a.world) st_release_sys(a.flag[threadIdx.x] + a.flag_off + kFlagB + a.rank, epoch);
  wait_peers(lf, kFlagB, a.world, epoch, a.timeout_ns);
}

}  // namespace

struct kfp16_peer_comm {
  kfp16_ctx* ctx = nullptr;
  PeerArgs args{};
  unsigned* flags = nullptr;            // this rank's flag block
  void* opened[2 * kMaxPeers] = {};     // IPC mappings to close
  int n_opened = 0;
  bool connected = false;
  size_t count = 0;                     // FP16 elements in the bucket
};

extern "C" {

kfp16_peer_comm* kfp16_peer_comm_create(kfp16_ctx* ctx, int rank, int world, void* bucket_f16, size_t count) {
  if (!ctx || !bucket_f16) { set_error("kfp16_peer_comm_create: null argument"); return nullptr; }
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) {
    set_error("kfp16_peer_comm_create: rank %d / world %d (1..%d ranks of one node)", rank, world, kMaxPeers);
    return nullptr;
  }
  if (reinterpret_cast<uintptr_t>(bucket_f16) & 15) { set_error("kfp16_peer_comm_create: the bucket must be 16-byte aligned"); return nullptr; }
  if (!check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice")) return nullptr;
  kfp16_peer_comm* c = new kfp16_peer_comm();
  c->ctx = ctx;
  if (!check_cuda(cudaMalloc(&c->flags, kChannels * kFlagWords * sizeof(unsigned)), "cudaMalloc (peer flags)") ||
      !check_cuda(cudaMemset(c->flags, 0, kChannels * kFlagWords * sizeof(unsigned)), "cudaMemset (peer flags)")) {
    if (c->flags) cudaFree(c->flags);
    delete c;
    return nullptr;
  }
  c->args.rank = rank;
  c->args.world = world;
  c->args.first = 0;
  c->args.count = count;
  c->count = count;
  c->args.timeout_ns = 20ull * 1000 * 1000 * 1000;
  c->args.buf[rank] = static_cast<__half*>(bucket_f16);
  c->args.flag[rank] = c->flags;
  c->connected = world == 1;
  return c;
}

int kfp16_peer_comm_handle(kfp16_peer_comm* c, void* out) {
  if (!c || !out) { set_error("kfp16_peer_comm_handle: null argument"); return -1; }
  static_assert(2 * sizeof(cudaIpcMemHandle_t) == KFP16_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h[2];
  if (!check_cuda(cudaIpcGetMemHandle(&h[0], c->args.buf[c->args.rank]), "cudaIpcGetMemHandle (bucket)") ||
      !check_cuda(cudaIpcGetMemHandle(&h[1], c->flags), "cudaIpcGetMemHandle (flags)")) return -1;
  memcpy(out, h, sizeof(h));
  return 0;
}

int kfp16_peer_comm_connect(kfp16_peer_comm* c, const void* handles) {
  if (!c || !handles) { set_error("kfp16_peer_comm_connect: null argument"); return -1; }
  if (c->connected) { set_error("kfp16_peer_comm_connect: already connected"); return -1; }
  if (!check_cuda(cudaSetDevice(c->ctx->device), "cudaSetDevice")) return -1;
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
  for (int p = 0; p < c->args.world; ++p) {
    if (p == c->args.rank) continue;
    void *buf = nullptr, *flag = nullptr;
    cudaIpcMemHandle_t hb, hf;      // (the caller's buffer need not be aligned for the handle type)
    memcpy(&hb, reinterpret_cast<const char*>(h) + (size_t)p * KFP16_PEER_HANDLE_BYTES, sizeof(hb));
    memcpy(&hf, reinterpret_cast<const char*>(h) + (size_t)p * KFP16_PEER_HANDLE_BYTES + sizeof(hb), sizeof(hf));
    if (!check_cuda(cudaIpcOpenMemHandle(&buf, hb, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer bucket)")) return -1;
    c->opened[c->n_opened++] = buf;
    if (!check_cuda(cudaIpcOpenMemHandle(&flag, hf, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer flags)")) return -1;
    c->opened[c->n_opened++] = flag;
    c->args.buf[p] = static_cast<__half*>(buf);
    c->args.flag[p] = static_cast<unsigned*>(flag);
  }
  c->connected = true;
  return 0;
}

int kfp16_peer_comm_set_timeout(kfp16_peer_comm* c, double seconds) {
  if (!c || !(seconds > 0)) { set_error("kfp16_peer_comm_set_timeout: bad argument"); return -1; }
  c->args.timeout_ns = (unsigned long long)(seconds * 1e9);
  return 0;
}

int kfp16_peer_allreduce_f16_range(kfp16_peer_comm* c, size_t first, size_t count, int channel, int threads, int max_ctas, void* stream) {
  if (!c) { set_error("kfp16_peer_allreduce_f16: null communicator"); return -1; }
  if (!c->connected) { set_error("kfp16_peer_allreduce_f16: peers are not connected (kfp16_peer_comm_connect)"); return -1; }
  if (first > c->count || count > c->count - first) { set_error("kfp16_peer_allreduce_f16_range: [%zu, +%zu) is outside the bucket of %zu elements", first, count, c->count); return -1; }
  if (channel < 0 || channel >= kChannels) { set_error("kfp16_peer_allreduce_f16_range: channel %d (0..%d)", channel, kChannels - 1); return -1; }
  if (threads == 0) threads = 512;
  if (threads < 32 || threads > 512 || (threads & 31)) { set_error("kfp16_peer_allreduce_f16_range: %d threads per CTA (32..512, whole warps)", threads); return -1; }
  const int world = c->args.world;
  if (world == 1 || count == 0) return 0;
  PeerArgs a = c->args;
  a.first = first;
  a.count = count;
  a.flag_off = channel * kFlagWords;
  const int wt = world <= 2 ? 2 : (world <= 4 ? 4 : 8), unroll = 8 / wt;
  const unsigned long long nvec = count >> 3, per = (nvec + world - 1) / world;
  unsigned long long ctas = (per + (unsigned long long)threads * unroll - 1) / ((unsigned long long)threads * unroll);
  const unsigned long long cap = max_ctas > 0 ? (unsigned long long)max_ctas : (unsigned long long)c->ctx->num_sms;
  if (ctas < 1) ctas = 1;
  if (ctas > cap) ctas = cap;
  const dim3 grid((unsigned)ctas), block((unsigned)threads);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->ctx->stream;
  if (wt == 2) peer_allreduce_f16_kernel<2, 4><<<grid, block, 0, st>>>(a);
  else if (wt == 4) peer_allreduce_f16_kernel<4, 2><<<grid, block, 0, st>>>(a);
  else peer_allreduce_f16_kernel<8, 1><<<grid, block, 0, st>>>(a);
  count_launch();
  return check_launch("kfp16_peer_allreduce_f16") ? 0 : -1;
}

int kfp16_peer_allreduce_f16(kfp16_peer_comm* c) {
  if (!c) { set_error("kfp16_peer_allreduce_f16: null communicator"); return -1; }
  return kfp16_peer_allreduce_f16_range(c, 0, c->count, 0, 512, 0, nullptr);
}

int kfp16_peer_comm_status(kfp16_peer_comm* c) {
  if (!c) { set_error("kfp16_peer_comm_status: null communicator"); return -1; }
  // exchanges may have been queued on other streams (kfp16_peer_allreduce_f16_range): wait for the device
  if (!check_cuda(cudaSetDevice(c->ctx->device), "cudaSetDevice") || !check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return -1;
  unsigned words[kChannels * kFlagWords];
  if (!check_cuda(cudaMemcpy(words, c->flags, sizeof(words), cudaMemcpyDeviceToHost), "cudaMemcpy (peer status)")) return -1;
  for (int ch = 0; ch < kChannels; ++ch)
    if (words[ch * kFlagWords + kError]) {
      set_error("kfp16_peer_allreduce_f16: a peer did not arrive within the time limit (channel %d); the buckets are not reduced", ch);
      return -1;
    }
  return 0;
}

void kfp16_peer_comm_destroy(kfp16_peer_comm* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  for (int i = 0; i < c->n_opened; ++i) cudaIpcCloseMemHandle(c->opened[i]);
  if (c->flags) cudaFree(c->flags);
  delete c;
}

}  // extern "C"
