// Kernel instantiation + launch for one tile width BN; included by gemm_bn*.cu so the ~30
// instantiations per width compile in parallel translation units.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "gemm_sm100.cuh"
#include "host_common.h"

namespace kfp16 {

struct GemmLaunch {
  int bn;            // 64 / 128 / 160 / 256
  bool a_mn, b_mn;   // operand majors
  int ek;            // EpiKind
  int cg;            // 1 = one CTA per tile, 2 = CTA pair (cta_group::2, 256-row tiles)
  int share;         // kernel MODE: 1 both splice slabs read one A tile; 3 merged groups (spliced weight gradients, BN 160);
                     // 4 convolution weight gradient; 5 convolution forward / input gradient with a shared input box
  int grid;          // CTAs (even when cg == 2)
};

template <int BN, bool A_MN, bool B_MN, int EK, int CG, int SHARE>
static bool launch_cfg(kfp16_ctx* ctx, const GemmParams& p, int grid) {
  using Cfg = GemmCfg<BN, A_MN, B_MN, EK, CG, SHARE>;
  auto kern = gemm_f16_sm100<BN, A_MN, B_MN, EK, CG, SHARE>;
  static bool attr_done = false;   // per instantiation
  if (!attr_done) {
    if (!check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes),
                    "cudaFuncSetAttribute(gemm smem)"))
      return false;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // the kernel calls griddep_wait() after its prologue
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  count_launch();
  count_gemm_kind(EK);            // the epilogue body that actually ran (template argument, not the requested flags)
  return check_cuda(e, "gemm_f16_sm100 launch");
}

// Which specialisations exist (everything else runs the run-time-flag EK_GENERIC body, one CTA per tile):
//   A K-major,  B MN-major (forward NN)        : all kinds, CG 1/2; spliced-tile variants for plain / affine
//   A K-major,  B K-major  (input gradients NT): plain, residual, bn+gradmask, CG 1/2; spliced-tile plain / residual
//   A MN-major, B MN-major (weight gradients)  : split-K (CG 1/2), plain
//   A MN-major, B K-major  (kaldi_gemm TT)     : generic, split-K
template <int BN>
bool launch_gemm_bn(kfp16_ctx* ctx, const GemmParams& p, const GemmLaunch& L) {
  const int g = L.grid;
#define KFP16_CASE(A, B, EK, CG, SH) return launch_cfg<BN, A, B, EK, CG, SH>(ctx, p, g)
  if constexpr (BN == 160) {
    if (L.share == 3) {
      if (L.a_mn && L.b_mn && L.ek == EK_SPLITK && L.cg == 2) KFP16_CASE(true, true, EK_SPLITK, 2, 3);
      set_error("internal: merged-group kernel needs MN-major operands, split-K and CTA pairs");
      return false;
    }
  }
  if (L.share == 5) {     // convolution forward / input gradient reading one shared input box per channel chunk
    if constexpr (BN != 160) {
      if (!L.a_mn && L.b_mn) {
        if (L.ek == EK_AFFINE) { if (L.cg == 2) KFP16_CASE(false, true, EK_AFFINE, 2, 5); KFP16_CASE(false, true, EK_AFFINE, 1, 5); }
        if (L.ek == EK_GENERIC && L.cg == 1) KFP16_CASE(false, true, EK_GENERIC, 1, 5);
      } else if (!L.a_mn && !L.b_mn && L.ek == EK_PLAIN) {
        if (L.cg == 2) KFP16_CASE(false, false, EK_PLAIN, 2, 5);
        KFP16_CASE(false, false, EK_PLAIN, 1, 5);
      } else if (!L.a_mn && !L.b_mn && L.ek == EK_BN_GRADMASK) {     // input gradient + the producer's batch-norm / ReLU backward
        if (L.cg == 2) KFP16_CASE(false, false, EK_BN_GRADMASK, 2, 5);
        KFP16_CASE(false, false, EK_BN_GRADMASK, 1, 5);
      }
    }
    set_error("internal: no shared-box convolution kernel for this operand / epilogue combination");
    return false;
  }
  if (L.share == 4) {     // convolution weight gradient (80-row k-blocks of (time, height), 4-D boxes for A)
    if constexpr (BN != 160) {
      if (L.a_mn && L.b_mn && L.ek == EK_SPLITK) {
        if (L.cg == 2) KFP16_CASE(true, true, EK_SPLITK, 2, 4);
        KFP16_CASE(true, true, EK_SPLITK, 1, 4);
      }
    }
    set_error("internal: the convolution weight-gradient kernel needs MN-major operands, split-K and a 64 / 128 / 256 tile");
    return false;
  }
  if (!L.a_mn && L.b_mn) {
    if (L.share) {
      switch (L.ek) {
        case EK_PLAIN: KFP16_CASE(false, true, EK_PLAIN, 2, true);
        case EK_AFFINE: KFP16_CASE(false, true, EK_AFFINE, 2, true);
        case EK_AFFINE_RES: KFP16_CASE(false, true, EK_AFFINE_RES, 2, true);
        case EK_AFFINE_DROP: KFP16_CASE(false, true, EK_AFFINE_DROP, 2, true);
        case EK_AFFINE_RES_DROP: KFP16_CASE(false, true, EK_AFFINE_RES_DROP, 2, true);
        default: break;
      }
      set_error("internal: no shared-tile kernel for epilogue kind %d", L.ek);
      return false;
    }
    if (L.cg == 2) {
      switch (L.ek) {
        case EK_PLAIN: KFP16_CASE(false, true, EK_PLAIN, 2, false);
        case EK_AFFINE: KFP16_CASE(false, true, EK_AFFINE, 2, false);
        case EK_AFFINE_RES: KFP16_CASE(false, true, EK_AFFINE_RES, 2, false);
        case EK_RESID: KFP16_CASE(false, true, EK_RESID, 2, false);
        case EK_BN_GRADMASK: KFP16_CASE(false, true, EK_BN_GRADMASK, 2, false);
        case EK_BN: KFP16_CASE(false, true, EK_BN, 2, false);
        case EK_BIAS: KFP16_CASE(false, true, EK_BIAS, 2, false);
        case EK_AFFINE_DROP: KFP16_CASE(false, true, EK_AFFINE_DROP, 2, false);
        case EK_AFFINE_RES_DROP: KFP16_CASE(false, true, EK_AFFINE_RES_DROP, 2, false);
        default: break;
      }
      set_error("internal: no CTA-pair kernel for epilogue kind %d", L.ek);
      return false;
    }
    switch (L.ek) {
      case EK_PLAIN: KFP16_CASE(false, true, EK_PLAIN, 1, false);
      case EK_AFFINE: KFP16_CASE(false, true, EK_AFFINE, 1, false);
      case EK_AFFINE_RES: KFP16_CASE(false, true, EK_AFFINE_RES, 1, false);
      case EK_RESID: KFP16_CASE(false, true, EK_RESID, 1, false);
      case EK_BN_GRADMASK: KFP16_CASE(false, true, EK_BN_GRADMASK, 1, false);
      case EK_BN: KFP16_CASE(false, true, EK_BN, 1, false);
      case EK_BIAS: KFP16_CASE(false, true, EK_BIAS, 1, false);
      case EK_SPLITK: KFP16_CASE(false, true, EK_SPLITK, 1, false);
      case EK_AFFINE_DROP: KFP16_CASE(false, true, EK_AFFINE_DROP, 1, false);
      case EK_AFFINE_RES_DROP: KFP16_CASE(false, true, EK_AFFINE_RES_DROP, 1, false);
      default: KFP16_CASE(false, true, EK_GENERIC, 1, false);
    }
  }
  if (!L.a_mn && !L.b_mn) {
    if (L.share) {
      switch (L.ek) {
        case EK_PLAIN: KFP16_CASE(false, false, EK_PLAIN, 2, true);
        case EK_RESID: KFP16_CASE(false, false, EK_RESID, 2, true);
        default: break;
      }
      set_error("internal: no shared-tile kernel for epilogue kind %d", L.ek);
      return false;
    }
    if (L.cg == 2) {
      switch (L.ek) {
        case EK_PLAIN: KFP16_CASE(false, false, EK_PLAIN, 2, false);
        case EK_RESID: KFP16_CASE(false, false, EK_RESID, 2, false);
        case EK_BN_GRADMASK: KFP16_CASE(false, false, EK_BN_GRADMASK, 2, false);
        default: break;
      }
      set_error("internal: no CTA-pair kernel for epilogue kind %d", L.ek);
      return false;
    }
    switch (L.ek) {
      case EK_PLAIN: KFP16_CASE(false, false, EK_PLAIN, 1, false);
      case EK_RESID: KFP16_CASE(false, false, EK_RESID, 1, false);
      case EK_BN_GRADMASK: KFP16_CASE(false, false, EK_BN_GRADMASK, 1, false);
      case EK_SPLITK: KFP16_CASE(false, false, EK_SPLITK, 1, false);
      default: KFP16_CASE(false, false, EK_GENERIC, 1, false);
    }
  }
  if (L.a_mn && L.b_mn) {
    if (L.cg == 2) {
      if (L.ek == EK_SPLITK) KFP16_CASE(true, true, EK_SPLITK, 2, false);
      set_error("internal: no CTA-pair kernel for epilogue kind %d with MN-major operands", L.ek);
      return false;
    }
    switch (L.ek) {
      case EK_PLAIN: KFP16_CASE(true, true, EK_PLAIN, 1, false);
      case EK_SPLITK: KFP16_CASE(true, true, EK_SPLITK, 1, false);
      default: KFP16_CASE(true, true, EK_GENERIC, 1, false);
    }
  }
  if (L.ek == EK_SPLITK) KFP16_CASE(true, false, EK_SPLITK, 1, false);
  KFP16_CASE(true, false, EK_GENERIC, 1, false);
#undef KFP16_CASE
}

// true when launch_gemm_bn has a kernel for this combination
inline bool gemm_variant_exists(bool a_mn, bool b_mn, int ek, int cg, int share) {
  if (share) {
    if (a_mn || cg != 2) return false;   // the shared splice tile exists for CTA pairs only
    if (b_mn) return ek == EK_PLAIN || ek == EK_AFFINE || ek == EK_AFFINE_RES || ek == EK_AFFINE_DROP || ek == EK_AFFINE_RES_DROP;
    return ek == EK_PLAIN || ek == EK_RESID;
  }
  if (cg == 2) {
    if (!a_mn && b_mn) return ek != EK_GENERIC && ek != EK_SPLITK;
    if (!a_mn && !b_mn) return ek == EK_PLAIN || ek == EK_RESID || ek == EK_BN_GRADMASK;
    if (a_mn && b_mn) return ek == EK_SPLITK;
    return false;
  }
  return true;
}

}  // namespace kfp16
