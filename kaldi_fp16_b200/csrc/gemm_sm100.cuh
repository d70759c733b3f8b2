// Hand-written sm_100a FP16 GEMM (FP32 accumulate in TMEM) with fused epilogue.
//
// Replaces the reference's only matrix multiply, cublasGemmEx behind ops_gemm
// (/root/reference/cpp/cuda/ops.cu:366-400), and the transposes + extra GEMMs its Go
// layer builds around it (internal/gpu/backward_ops.go:162-253, ops.go:335-351).
//
// One persistent CTA per SM, 12 warps, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      MMA issuer     (one elected thread, tcgen05.mma kind::f16, M=128 x N=BN x K=16)
//   warp 2      TMEM allocator, then output store warp (TMA stores of finished 64-column chunks)
//   warp 3      residual / C-in prefetcher (TMA loads of the R tile, two 64-column chunks ahead)
//   warps 4..11 epilogue       (tcgen05.ld -> fused epilogue -> swizzled smem -> TMA store,
//                               or fp32 red.add for split-K).  Warp w owns TMEM lanes
//                               32*(w%4)..+31 and the 32-column half (w-4)/4 of each 64-column chunk.
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// The epilogue is specialised at compile time (EpiKind) for the flag combinations the network
// executor issues, so the hot variants are short straight-line code; every other combination
// (reference rounding mode, dropout, cuBLAS beta, ...) runs the EK_GENERIC body with run-time flags.
// Per-column vectors (bias, batch-norm scale/shift) are staged in shared memory once per tile.
//
// Operands may be K-major or MN-major (no transpose kernels: the UMMA descriptors take
// X^T / W^T directly) and the K dimension may be made of up to 2 "slabs" whose TMA
// coordinates are offset independently -- that is how time-splicing [X(t-s) | X(t)] is fed
// to the tensor cores without materialising the spliced matrix
// (reference: internal/nnet/forward.go:699-790 builds it with 3 copies + 2 concat kernels).
#pragma once
#include "sm100_ptx.cuh"

namespace kfp16 {

constexpr int kBM = 128;          // UMMA M (cta_group::1)
constexpr int kBK = 64;           // one 128-byte swizzle row of K per stage
constexpr int kEpiThreads = 256;  // 8 epilogue warps
constexpr int kGemmThreads = 128 + kEpiThreads;
constexpr int kChunkBytes = kBM * 64 * 2;   // [128 x 64] fp16 staging chunk, SW128
constexpr int kVecBytes = 256 * (4 + 4 + 4);   // per-column bias, bn scale, bn shift as fp32 [256] each
constexpr int kMaxGroups = 4;    // 2 per problem, up to two problems per launch (split-K only)
constexpr int kMaxSlabs = 2;
constexpr int kMaxTaps = 16;      // implicit-GEMM convolution: taps addressed by one launch

enum EpiFlags : uint32_t {
  EPI_BIAS = 1u << 0,       // + bias[n] (fp16 bias row, as gpu.AddBias)
  EPI_RELU = 1u << 1,       // max(x,0) keeping NaN like kernel_relu
  EPI_BN = 1u << 2,         // x*bn_scale[n] + bn_shift[n]   (inference batch-norm)
  EPI_RESID = 1u << 3,      // out = res_scale*R + x         (TDNN-F bypass, ops_add_scaled)
  EPI_BETA = 1u << 4,       // acc = alpha*acc + beta*R      (cuBLAS beta path)
  EPI_REF_ROUND = 1u << 5,  // round to fp16 between fused stages exactly where the reference stores fp16
  EPI_MASK = 1u << 6,       // emit relu bit-mask (x > 0), 1 bit / element
  EPI_SPLITK = 1u << 7,     // fp32 red.add of the partial tile into ws (no fp16 store)
  EPI_DROPOUT = 1u << 8,    // inverted dropout after the batch-norm: keep if u(seed,row,col) > p, scale 1/(1-p) (go/gotorch/layers.go:365-399)
  EPI_GRADMASK = 1u << 9,   // x = mask_in bit ? x : 0      (relu backward fused in producer)
};

// compile-time epilogue variants
enum EpiKind : int {
  EK_GENERIC = 0,      // run-time flags
  EK_PLAIN,            // D = h(alpha*acc)
  EK_AFFINE,           // bias, relu, bn, mask            (prefinal affine, tdnnf affine without bypass)
  EK_AFFINE_RES,       // bias, relu, bn, mask, residual  (tdnnf affine with bypass)
  EK_RESID,            // residual only                   (tdnnf input gradient + bypass gradient)
  EK_BN_GRADMASK,      // bn scale, relu-backward mask    (prefinal backward)
  EK_BN,               // bn only                         (prefinal linear)
  EK_BIAS,             // bias only                       (output layer)
  EK_SPLITK,           // fp32 accumulate into the workspace (weight gradients)
  EK_AFFINE_DROP,      // bias, relu, bn, dropout, mask            (tdnnf affine in training with dropout-proportion > 0)
  EK_AFFINE_RES_DROP,  // bias, relu, bn, dropout, mask, residual
  EK_COUNT
};

__host__ __device__ constexpr uint32_t epi_kind_flags(int k) {
  return k == EK_PLAIN ? 0u
       : k == EK_AFFINE ? (EPI_BIAS | EPI_RELU | EPI_BN | EPI_MASK)
       : k == EK_AFFINE_RES ? (EPI_BIAS | EPI_RELU | EPI_BN | EPI_MASK | EPI_RESID)
       : k == EK_RESID ? (uint32_t)EPI_RESID
       : k == EK_BN_GRADMASK ? (EPI_BN | EPI_GRADMASK)
       : k == EK_BN ? (uint32_t)EPI_BN
       : k == EK_BIAS ? (uint32_t)EPI_BIAS
       : k == EK_SPLITK ? (uint32_t)EPI_SPLITK
       : k == EK_AFFINE_DROP ? (EPI_BIAS | EPI_RELU | EPI_BN | EPI_MASK | EPI_DROPOUT)
       : k == EK_AFFINE_RES_DROP ? (EPI_BIAS | EPI_RELU | EPI_BN | EPI_MASK | EPI_DROPOUT | EPI_RESID)
       : 0u;
}

// One problem of a grouped split-K launch (weight gradients of many layers in one persistent kernel): the table lives
// in device memory (a launch can carry more problems than kernel parameters could hold tensor maps for).
struct alignas(64) GroupProb {
  CUtensorMap tmA, tmB;
  int a_row_off[2], b_row_off[2];   // per row-shifted group of the problem
  float* ws[2];                     // fp32 accumulation targets
  int ws_ld, ws_transposed;
  // merged-group kernels (MODE 3): one tile per problem; tmA / tmB boxes are 64 + span k-rows starting at
  // a_row_off[0] / b_row_off[0], group g reads them from k-row a_shift[g] / b_shift[g]; merge_tx = bytes per k-block per CTA
  int a_shift[2], b_shift[2];
  int merge_tx;
};

struct GemmParams {
  const GroupProb* probs;       // grouped launch: tile group g belongs to problem g / 2 (nullptr = single problem)
  CUtensorMap tmA, tmB;
  CUtensorMap tmA2, tmB2;       // second split-K problem of the launch: tile groups >= groups2_from read these
  int groups2_from;             // first group index of the second problem (0 = none)
  int ws_ld2, ws_transposed2;   // its accumulation target layout
  CUtensorMap tmD[kMaxGroups];  // per-group output view  (TMA store clips to the view)
  CUtensorMap tmR[kMaxGroups];  // per-group residual / C-in view (TMA load)
  int M, N, K;                  // per-group problem size; K = kslabs * kslab_len
  int groups, kslabs, kslab_len;
  int split_k;                  // >1 => EPI_SPLITK
  int a_row_off[kMaxGroups][kMaxSlabs], a_col_off[kMaxGroups][kMaxSlabs];
  int b_row_off[kMaxGroups][kMaxSlabs], b_col_off[kMaxGroups][kMaxSlabs];
  uint32_t flags;
  float alpha, beta, res_scale;
  const __half* bias;           // [N]   (group g uses bias + g*vec_gstride)
  const float* bn_scale;        // [N]
  const float* bn_shift;        // [N]
  int vec_gstride;              // per-group offset into bias/bn vectors
  uint32_t* mask_out;           // [M x mask_ld] words
  const uint32_t* mask_in;
  int mask_ld;
  float* ws[kMaxGroups];        // split-K fp32 accumulation target [M x ws_ld]
  int ws_ld;
  int ws_transposed;            // split-K target is stored [N x M] (ws[n*ws_ld + m]): warp lanes = consecutive m
  float drop_p; uint32_t drop_seed;
  const uint32_t* drop_seed_dev;   // optional device word XOR-ed into drop_seed when the kernel runs (a per-step counter: graph replays draw new masks)
  // shared splice tile (SHARE kernels): slab s reads the A tile from row a_shift[s] (0..8) on
  int a_shift[kMaxSlabs];
  int a_box_bytes;              // bytes of the A box (128 + span rows) x 128 B
  // merged groups (MODE 3): group g reads the A / B tile from k-row a_shift[g] / b_shift[g] (0..8) on; the tiles are
  // loaded from row offsets a_row_off[0][0] / b_row_off[0][0]; merge_tx = bytes one CTA's loads deliver per k-block
  int b_shift[kMaxSlabs];
  int merge_tx;
  int tile_rows;                // rows of M one CTA tile advances by (128; a convolution tile is tbox x height <= 128 rows)
  // ---- implicit-GEMM convolution (conv != 0): the A operand is a 4-D tensor [time][height][parity][channel] and K (conv 1)
  // or M (conv 2) runs over (tap, channel): tap s reads the box shifted by (dt, hq) in (time, height) at parity `par`;
  // out-of-bounds box elements are zero-filled by the TMA unit = the convolution's zero padding.  No patch matrix.
  //   conv 1: A K-major (forward / input gradient).  M row = t * conv_h + h; k-block kb -> tap kb / (conv_c/64)
  //   conv 2: A MN-major (weight gradient).  M row = tap * conv_c + channel; k-block = conv_tbox time steps x conv_h heights
  int conv;
  int conv_h, conv_c, conv_taps, conv_tbox, conv_k16;   // heights per time step, channels per tap, taps, box time steps, UMMA K steps per k-block (conv 2)
  int conv_tx;                  // bytes one CTA's loads deliver per k-block
  int8_t conv_dt[kMaxTaps], conv_hq[kMaxTaps], conv_par[kMaxTaps];
  int conv_brow[kMaxTaps];      // conv 1: B row offset of tap s (MN-major B: added to the k row; K-major B: to the n row)
  // conv 1, shared box (MODE 5): per 64-channel chunk ONE box of (tbox + time span) x conv_sl rows per parity plane is loaded
  // and every tap reads it through a row-shifted descriptor; the tile's accumulator rows are (frame, slot) with conv_sl slots
  // per frame, of which the first conv_h are real outputs
  int conv_sl;                  // slots per frame in the accumulator-row index (conv_h + height span of the taps)
  int conv_box_t0, conv_box_h0; // box origin relative to the tile's first frame / height 0 (= min dt, min hq)
  int conv_box_bytes;           // bytes of one plane's box
  int conv_plane_bytes;         // shared-memory bytes reserved per plane (box + read-ahead slack, multiple of 1024)
  int conv_planes;              // parity planes loaded (1 or 2)
  int conv_plane_par[2];        // their parity coordinates
  int conv_nstages;             // B ring depth (2..8)
  uint32_t conv_aoff[kMaxTaps]; // descriptor offset (16-byte units) of tap s inside an A buffer
  // rows whose index modulo zero_period lies outside [zero_lo, zero_lo + zero_len) are written as zeros with a zero mask
  // (halo rows of the padded minibatch layout: the zero padding the next convolution reads); zero_period == 0: off
  uint32_t zero_period, zero_lo, zero_len;
  long long* dbg;               // profiling: per-CTA role timestamps [grid][3 roles][8 tiles][16] (nullptr = off)
};

// (dropout_uniform, the counter-based generator of the dropout masks, lives in sm100_ptx.cuh: elementwise.cu draws the
//  SpecAugment masks from it too)

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
// max(x, 0) that lets NaN through, like the reference's `x < 0 ? 0 : x` (ops.cu:26-37)
__device__ __forceinline__ float relu_nan(float x) {
  float r;
  asm("max.NaN.f32 %0, %1, 0f00000000;\n" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }
// packed fp32 arithmetic (sm_100: one FFMA2 / FMUL2 per two values)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}

// MODE: 0 plain; 1 shared splice tile (one A tile of 128+8 rows serves both K slabs);
//       3 merged groups (weight gradients of a spliced layer: both row-shifted groups read ONE A and ONE B tile of 64+8
//         k-rows through shifted descriptors and accumulate into the two TMEM accumulator stages)
//       4 convolution weight gradient: MN-major operands in k-blocks of up to 80 (time, height) rows, the A chunks loaded
//         as 4-D boxes of the layer input shifted per tap
//       5 convolution forward / input gradient with a SHARED input box: per 64-channel chunk one box of the layer input
//         (all frames and heights any tap of the tile touches) is loaded once and the taps read it through row-shifted
//         descriptors, while their weight tiles stream through the ring -- a 3x3 layer moves 27 KB of input per chunk
//         instead of 9 x 15 KB
template <int BN, bool A_MN, bool B_MN, int EK, int CG, int MODE>
struct GemmCfg {
  static constexpr bool SHARE = MODE == 1;
  static constexpr bool kConvW = MODE == 4;
  static constexpr bool kConvS = MODE == 5;
  static constexpr bool kMerge = MODE == 3;
  static_assert(CG == 1 || CG == 2, "cta_group 1 or 2");
  static_assert(!SHARE || !A_MN, "the shared splice tile is implemented for a K-major A");
  static_assert(MODE == 0 || MODE == 1 || MODE == 3 || MODE == 4 || MODE == 5, "kernel mode");
  static_assert(!kConvS || (!A_MN && EK != EK_SPLITK), "shared convolution box: K-major A, no split-K");
  static_assert(!kConvW || (A_MN && B_MN && EK == EK_SPLITK), "convolution weight gradient: MN-major operands, split-K");
  static_assert(!kMerge || (A_MN && B_MN && EK == EK_SPLITK && CG == 2), "merged groups: MN-major operands, split-K, CTA pairs");
  static constexpr bool kSplitK = EK == EK_SPLITK;
  static constexpr uint32_t kFlags = epi_kind_flags(EK);
  static constexpr bool kMayUseR = EK == EK_GENERIC || (kFlags & (EPI_RESID | EPI_BETA)) != 0;
  static constexpr bool kUsesVec = EK == EK_GENERIC || (kFlags & (EPI_BIAS | EPI_BN)) != 0;
  // staging ring: 4 chunks; 6 when a residual tile is prefetched into it (in-place epilogue) and the tile is narrow enough
  // to afford it (the residual prefetch then runs 4 chunks ahead)
  static constexpr int kRing = kSplitK ? 0 : (kMayUseR ? (BN <= 160 ? 6 : 4) : 4);
  // TMA stores left in flight when a chunk is handed over (a store's smem-read latency is ~1000 cycles:
  // with none in flight every 64-column chunk paid it in full)
  static constexpr int kStoreWait = kRing >= 5 ? kRing - 3 : (kMayUseR ? 1 : 2);   // < kRing
  static constexpr int kUmmaM = kBM * CG;                      // 256 rows over a CTA pair
  static constexpr int kBNLocal = BN / CG;                     // B rows / columns staged by this CTA
  static constexpr int kBChunks = B_MN ? (kBNLocal + 63) / 64 : 1;   // 64-wide N chunks (MN-major)
  // SHARE: one A tile of 128+8 rows serves both splice slabs (row-shifted UMMA descriptors)
  // MN-major chunk = [k rows][64 M/N elements]: 64 k-rows, or 64+8 when the groups are merged (row-shifted reads)
  static constexpr int kMnRows = kConvW ? 80 : (kMerge ? kBK + 8 : kBK);
  static constexpr int kMnChunkBytes = kMnRows * 64 * 2;
  static constexpr int kABytes = SHARE ? (kBM + 8) * kBK * 2 : (A_MN ? 2 * kMnChunkBytes : kBM * kBK * 2);
  static constexpr int kBTileBytes = B_MN ? kBChunks * kMnChunkBytes : kBNLocal * kBK * 2;
  static constexpr int kNumBTiles = SHARE ? 2 : 1;
  static constexpr int kStageBytes = kConvS ? kBTileBytes : kABytes + kNumBTiles * kBTileBytes;   // MODE 5: the ring holds weight tiles only
  static constexpr int kEpiBytes = kRing * kChunkBytes + (kUsesVec ? 2 * kVecBytes : 0);   // vectors double-buffered
  static constexpr int kSmemBudget = 227 * 1024 - 1024 /*align*/ - 512 /*barriers*/;
  static constexpr int kStagesRaw = (kSmemBudget - kEpiBytes) / kStageBytes;
  static constexpr int kStages = kConvS ? 8 : (kStagesRaw > 8 ? 8 : kStagesRaw);   // MODE 5: barrier slots; the depth is a launch parameter
  static constexpr int kAccCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;   // power of two
  static constexpr int kSmemBytes = kConvS ? 227 * 1024 : kStages * kStageBytes + kEpiBytes + 1024 + 512;
  static_assert(kStages >= 2, "tile too large for shared memory");
  static_assert(kStageBytes % 1024 == 0, "stage must keep the 1024-byte swizzle alignment");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "UMMA N; epilogue works on 32-column halves");
};

// Tile bookkeeping shared by all warp roles.  A "unit" is a CTA (CG = 1) or a CTA pair (CG = 2);
// a tile is kBM*CG rows x BN columns and this CTA owns rows m_row0 .. m_row0+127 of it.
template <int BN, int CG>
struct TileIter {
  int m_tiles, n_tiles, tiles_per_split, total_tiles, unit, nunits, rank, tile_rows;
  int first, last, step;   // this unit's tiles: first, first + step, ... < last
  __device__ __forceinline__ TileIter(const GemmParams& p) {
    tile_rows = p.tile_rows;
    m_tiles = (p.M + tile_rows * CG - 1) / (tile_rows * CG);
    n_tiles = (p.N + BN - 1) / BN;
    tiles_per_split = m_tiles * n_tiles * p.groups;
    total_tiles = tiles_per_split * p.split_k;
    unit = blockIdx.x / CG;
    nunits = gridDim.x / CG;
    rank = CG == 2 ? (int)cluster_ctarank() : 0;
    p_groups = p.groups;
    first = unit; last = total_tiles; step = nunits;     // round-robin
  }
  __device__ __forceinline__ void decode(int tile, int& n_blk, int& m_row0, int& g, int& ks) const {
    int id = tile;
    n_blk = id % n_tiles; id /= n_tiles;
    m_row0 = (id % m_tiles) * (tile_rows * CG) + rank * tile_rows; id /= m_tiles;
    g = id % p_groups; ks = id / p_groups;
  }
  int p_groups;
};

// profiling hook: role 0 = TMA producer, 1 = MMA issuer, 2 = epilogue (warp 4); slot < 16, tile index < 8
__device__ __forceinline__ void dbg_stamp(const GemmParams& p, int role, int tile_i, int slot) {
  if (p.dbg != nullptr && tile_i < 8 && (threadIdx.x & 31) == 0)
    p.dbg[(((size_t)blockIdx.x * 3 + role) * 8 + tile_i) * 16 + slot] = clock64();
}

template <int BN, bool A_MN, bool B_MN, int EK, int CG, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_sm100(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, A_MN, B_MN, EK, CG, MODE>;
  constexpr bool SHARE = Cfg::SHARE;
  constexpr bool kMerge = Cfg::kMerge;
  constexpr int kStages = Cfg::kStages;
  constexpr int kRing = Cfg::kRing;
  constexpr bool kGeneric = EK == EK_GENERIC;
  constexpr bool kSplitK = Cfg::kSplitK;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int a_region = Cfg::kConvS ? 2 * p.conv_planes * p.conv_plane_bytes : 0;   // MODE 5: two A buffers in front of the ring
  const int nstages = Cfg::kConvS ? p.conv_nstages : kStages;
  uint8_t* smem_ring = smem + a_region;            // operand stages
  uint8_t* smem_epi = smem_ring + nstages * Cfg::kStageBytes;
  uint8_t* smem_vec = smem_epi + kRing * kChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;                  // [kStages]  TMA -> MMA          (CG=2: the leader's is used)
  uint64_t* empty_bar = bars + kStages;       // [kStages]  MMA -> TMA          (CG=2: commit multicast to both CTAs)
  uint64_t* tfull_bar = bars + 2 * kStages;   // [2]        MMA -> epilogue     (CG=2: commit multicast)
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]        epilogue -> MMA     (CG=2: the leader's, 16 arrivals)
  uint64_t* rfull_bar = tempty_bar + 2;       // [8]        residual TMA -> epilogue
  uint64_t* rempty_bar = rfull_bar + 8;       // [8]        output store drained -> residual TMA / next writer
  uint64_t* cfull_bar = rempty_bar + 8;       // [8]        epilogue warps wrote a chunk -> store warp
  uint64_t* afull_bar = cfull_bar + 8;        // [2]        MODE 5: input box landed (CG=2: the leader's is used)
  uint64_t* aempty_bar = afull_bar + 2;       // [2]        MODE 5: MMAs on the box done (commit multicast)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 2) dbg_stamp(p, 0, 7, 0);      // kernel entry (profiling row: role 0, tile slot 7)
  const uint32_t flags = kGeneric ? p.flags : Cfg::kFlags;
  const bool use_r = Cfg::kMayUseR && (flags & (EPI_RESID | EPI_BETA)) != 0;

  const TileIter<BN, CG> ti(p);
  const int total_tiles = ti.total_tiles;
  const int rank = ti.rank;
  const int kb_per_slab = (p.kslab_len + kBK - 1) / kBK;
  // SHARE: the slabs are consumed inside every k-block; otherwise they are laid end to end along K
  // (implicit-GEMM convolution: conv 1 has one slab of kslab_len = channels per tap; conv 2 walks the time axis in
  //  k-blocks of conv_tbox frames, K = frames x heights)
  const int kb_total = Cfg::kConvW ? (p.K / p.conv_h + p.conv_tbox - 1) / p.conv_tbox : (SHARE ? kb_per_slab : p.kslabs * kb_per_slab);
  const int kb_per_split = (kb_total + p.split_k - 1) / p.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (kSplitK && p.groups2_from > 0) { tma_prefetch_desc(&p.tmA2); tma_prefetch_desc(&p.tmB2); }
    if (!kSplitK) tma_prefetch_desc(&p.tmD[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], CG); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CG * (kEpiThreads / 32));   // one arrive per epilogue warp (of both CTAs)
    }
    for (int i = 0; i < 8; ++i) { mbar_init(&rfull_bar[i], 1); mbar_init(&rempty_bar[i], 1); mbar_init(&cfull_bar[i], kEpiThreads / 32); }
    for (int i = 0; i < 2; ++i) { mbar_init(&afull_bar[i], CG); mbar_init(&aempty_bar[i], 1); }
    mbar_fence_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  // programmatic dependent launch: the next kernel in the stream may be scheduled from here on (its CTAs start as
  // ours exit); everything above touched only parameters / shared memory / TMEM, everything below may read what the
  // previous kernel wrote, so wait for it here
  griddep_launch();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 2) dbg_stamp(p, 0, 7, 1);      // prologue done (barriers, TMEM, cluster sync)
  griddep_wait();
  if (warp == 2) dbg_stamp(p, 0, 7, 2);      // predecessor grid complete

  // The three single-issuer roles below run with the whole warp converged and elect ONE lane only around
  // the TMA / tcgen05 instructions: every address and descriptor is then warp-uniform, so ptxas keeps them
  // in uniform registers.  (Measured: with an `if (lane == 0)` body the issue loop cost ~100 cycles per
  // tcgen05.mma -- R2UR moves on a divergent path -- and bounded every GEMM regardless of tile shape.)
  if (warp == 0) {
    // ===================================================== TMA producer
    int stage = 0; uint32_t phase = 0;
    const uint32_t stage_tx = kMerge ? (uint32_t)p.merge_tx
                              : p.conv ? (uint32_t)p.conv_tx
                              : SHARE ? (uint32_t)(p.a_box_bytes + p.kslabs * Cfg::kBTileBytes)
                                      : (uint32_t)(Cfg::kABytes + Cfg::kBTileBytes);
    int tile_i = 0;
    uint32_t a_cnt = 0;            // MODE 5: input boxes loaded so far (buffer = a_cnt & 1)
    for (int tile = ti.first; tile < ti.last; tile += ti.step, ++tile_i) {
      int n_blk, m_row0, g, ks;
      ti.decode(tile, n_blk, m_row0, g, ks);
      const int kb0 = ks * kb_per_split;
      const int kb1 = min(kb0 + kb_per_split, kb_total);
      const int n_loc = n_blk * BN + rank * Cfg::kBNLocal;     // first B row / column staged by this CTA
      const bool second = kSplitK && p.groups2_from > 0 && g >= p.groups2_from;
      const CUtensorMap* mapA = second ? &p.tmA2 : &p.tmA;
      const CUtensorMap* mapB = second ? &p.tmB2 : &p.tmB;
      int grp_a_row = 0, grp_b_row = 0;
      const bool grouped = kSplitK && p.probs != nullptr;
      const int gi = grouped ? 0 : g;     // index into the per-group parameter arrays (unused, zero, in a grouped launch)
      uint32_t tile_tx = stage_tx;
      if (grouped) {
        const GroupProb& pr = p.probs[kMerge ? g : g >> 1];
        mapA = &pr.tmA; mapB = &pr.tmB;
        grp_a_row = pr.a_row_off[kMerge ? 0 : g & 1]; grp_b_row = pr.b_row_off[kMerge ? 0 : g & 1];
        if (kMerge) tile_tx = (uint32_t)pr.merge_tx;
        if (elect_one()) { tma_prefetch_desc(mapA); tma_prefetch_desc(mapB); }
        __syncwarp();
      }
      dbg_stamp(p, 0, tile_i, 0);
      // convolution: the tile's first time step (conv 1) / the two 64-wide M chunks' (tap, first channel) (conv 2)
      const int conv_t0 = p.conv == 1 ? m_row0 / p.conv_h : 0;
      int cw_tap[2] = {0, 0}, cw_c0[2] = {0, 0};
      if (Cfg::kConvW) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int m = m_row0 + c * 64;
          cw_tap[c] = m / p.conv_c; cw_c0[c] = m - cw_tap[c] * p.conv_c;
        }
      }
      if constexpr (Cfg::kConvS) {
        // shared input box: per 64-channel chunk ONE box per parity plane (every frame / height a tap of this tile reads),
        // then the taps' weight tiles through the ring
        const int f0 = m_row0 / p.conv_h;                 // this CTA's first output frame
        const int nchunks = p.conv_c / kBK;
        for (int c = 0; c < nchunks; ++c, ++a_cnt) {
          const uint32_t ab = a_cnt & 1u;
          mbar_wait(&aempty_bar[ab], ((a_cnt >> 1) & 1u) ^ 1u);
          if (elect_one()) {
            const uint32_t fa = CG == 2 ? mapa_u32(smem_u32(&afull_bar[ab]), 0) : 0;
            const uint32_t tx = (uint32_t)(p.conv_planes * p.conv_box_bytes);
            if (CG == 2) mbar_arrive_expect_tx_cluster(fa, tx); else mbar_arrive_expect_tx(&afull_bar[ab], tx);
            for (int pl = 0; pl < p.conv_planes; ++pl) {
              uint8_t* dst = smem + (ab * p.conv_planes + pl) * p.conv_plane_bytes;
              if (CG == 2) tma_load_4d_pair(dst, mapA, fa, c * kBK, p.conv_plane_par[pl], p.conv_box_h0, f0 + p.conv_box_t0);
              else tma_load_4d(dst, mapA, &afull_bar[ab], c * kBK, p.conv_plane_par[pl], p.conv_box_h0, f0 + p.conv_box_t0);
            }
          }
          __syncwarp();
          for (int tap = 0; tap < p.conv_taps; ++tap) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              uint8_t* sb = smem_ring + stage * Cfg::kStageBytes;
              const uint32_t fb = CG == 2 ? mapa_u32(smem_u32(&full_bar[stage]), 0) : 0;
              if (CG == 2) mbar_arrive_expect_tx_cluster(fb, (uint32_t)Cfg::kBTileBytes);
              else mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)Cfg::kBTileBytes);
              auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
                if (CG == 2) tma_load_2d_pair(dst, m, fb, c0, c1);
                else tma_load_2d(dst, m, &full_bar[stage], c0, c1);
              };
              if (!B_MN) load(sb, mapB, c * kBK, n_loc + p.conv_brow[tap]);
              else {
#pragma unroll
                for (int cc = 0; cc < Cfg::kBChunks; ++cc)
                  load(sb + cc * Cfg::kMnChunkBytes, mapB, n_loc + cc * 64, c * kBK + p.conv_brow[tap]);
              }
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
      } else
      for (int kb = kb0; kb < kb1; ++kb) {
        const int slab = SHARE ? 0 : kb / kb_per_slab;
        const int k_in = (kb - slab * kb_per_slab) * kBK;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem_ring + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          // completion is signalled on this CTA's full barrier (CG=1) or the pair leader's (CG=2)
          const uint32_t fb = CG == 2 ? mapa_u32(smem_u32(&full_bar[stage]), 0) : 0;
          if (CG == 2) mbar_arrive_expect_tx_cluster(fb, tile_tx);
          else mbar_arrive_expect_tx(&full_bar[stage], tile_tx);
          auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
            if (CG == 2) tma_load_2d_pair(dst, m, fb, c0, c1);
            else tma_load_2d(dst, m, &full_bar[stage], c0, c1);
          };
          auto load4 = [&](void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3) {
            if (CG == 2) tma_load_4d_pair(dst, m, fb, c0, c1, c2, c3);
            else tma_load_4d(dst, m, &full_bar[stage], c0, c1, c2, c3);
          };
          if (Cfg::kConvW) {
            // weight gradient: k-block kb = frames [kb*tbox, +tbox) x all heights; chunk c = 64 channels of tap cw_tap[c]
            // read from the layer input shifted by that tap (a tap index past the last one addresses nothing: zeros)
            const int t_k = kb * p.conv_tbox;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int tp = cw_tap[c];
              const bool ok = tp < p.conv_taps;
              load4(sa + c * Cfg::kMnChunkBytes, mapA, cw_c0[c], ok ? p.conv_par[tp] : 0, ok ? p.conv_hq[tp] : 0,
                    ok ? t_k + p.conv_dt[tp] : 0x3FFFFFFF);
            }
#pragma unroll
            for (int c = 0; c < Cfg::kBChunks; ++c)
              load(sb + c * Cfg::kMnChunkBytes, mapB, n_loc + c * 64, kb * p.conv_tbox * p.conv_h);
          } else if (!A_MN && p.conv == 1) {
            // forward / input gradient: slab = tap; the A box [tbox frames][conv_h heights][64 channels] starts at the
            // tile's first frame shifted by the tap; B rows (tap, channel) are contiguous
            load4(sa, mapA, k_in, p.conv_par[slab], p.conv_hq[slab], conv_t0 + p.conv_dt[slab]);
            if (!B_MN) load(sb, mapB, k_in, n_loc + p.conv_brow[slab]);
            else {
#pragma unroll
              for (int c = 0; c < Cfg::kBChunks; ++c)
                load(sb + c * Cfg::kMnChunkBytes, mapB, n_loc + c * 64, k_in + p.conv_brow[slab]);
            }
          } else {
            if (!A_MN) {
              load(sa, mapA, k_in + p.a_col_off[g][slab], m_row0 + p.a_row_off[g][slab]);
            } else {
#pragma unroll
              for (int c = 0; c < 2; ++c)
                load(sa + c * Cfg::kMnChunkBytes, mapA, m_row0 + c * 64 + p.a_col_off[gi][slab],
                     k_in + (grouped ? grp_a_row : p.a_row_off[gi][slab]));
            }
            const int nb = SHARE ? p.kslabs : 1;
            for (int t = 0; t < nb; ++t) {
              const int bs = SHARE ? t : slab;
              uint8_t* sbt = sb + t * Cfg::kBTileBytes;
              if (!B_MN) {
                load(sbt, mapB, k_in + p.b_col_off[g][bs], n_loc + p.b_row_off[g][bs]);
              } else {
#pragma unroll
                for (int c = 0; c < Cfg::kBChunks; ++c)
                  load(sbt + c * Cfg::kMnChunkBytes, mapB, n_loc + c * 64 + p.b_col_off[gi][bs],
                       k_in + (grouped ? grp_b_row : p.b_row_off[gi][bs]));
              }
            }
          }
        }
        __syncwarp();
        if (kb == kb0) dbg_stamp(p, 0, tile_i, 1);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
      dbg_stamp(p, 0, tile_i, 2);
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer (CG=2: the pair leader only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(Cfg::kUmmaM, BN, A_MN, B_MN);
      // K-major: 8-row groups 1024 B apart (SBO); MN-major: 64-wide chunks 8 KB apart (LBO),
      // 8-k groups 1024 B apart (SBO).
      constexpr uint32_t a_lbo = A_MN ? Cfg::kMnChunkBytes : 16, a_sbo = 1024;
      constexpr uint32_t b_lbo = B_MN ? Cfg::kMnChunkBytes : 16, b_sbo = 1024;
      constexpr uint32_t a_kstep = (A_MN ? 2048 : 32) >> 4;   // descriptor units (16 B) per UMMA K=16
      constexpr uint32_t b_kstep = (B_MN ? 2048 : 32) >> 4;
      // (MODE 5: A descriptors address the input-box buffers in front of the ring, the ring holds weight tiles only)
      const uint64_t adesc0 = make_smem_desc(smem_u32(Cfg::kConvS ? smem : smem_ring), a_lbo, a_sbo, kLayoutSW128);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_ring) + (Cfg::kConvS ? 0 : Cfg::kABytes), b_lbo, b_sbo, kLayoutSW128);
      uint32_t a_cnt = 0;
      auto mma = [&](uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t accum) {
        if (CG == 2) umma_f16_pair(d_tmem, ad, bd, idesc, accum);
        else umma_f16(d_tmem, ad, bd, idesc, accum);
      };
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int tile_i = 0;
      for (int tile = ti.first; tile < ti.last; tile += ti.step, ++tile_i) {
        const int ks = tile / ti.tiles_per_split;
        const int kb0 = ks * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_total);
        dbg_stamp(p, 1, tile_i, 0);
        int sh_a[2] = {p.a_shift[0], p.a_shift[1]}, sh_b[2] = {p.b_shift[0], p.b_shift[1]};
        if (kMerge && p.probs != nullptr) {      // grouped launch: this tile's problem carries its own shifts
          const GroupProb& pr = p.probs[(tile % ti.tiles_per_split) / (ti.m_tiles * ti.n_tiles)];
          sh_a[0] = pr.a_shift[0]; sh_a[1] = pr.a_shift[1]; sh_b[0] = pr.b_shift[0]; sh_b[1] = pr.b_shift[1];
        }
        if (kMerge) {      // a merged tile owns BOTH accumulator stages (one per group)
          mbar_wait(&tempty_bar[0], acc_phase ^ 1);
          mbar_wait(&tempty_bar[1], acc_phase ^ 1);
        } else {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        }
        tc_fence_after();
        dbg_stamp(p, 1, tile_i, 1);
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
        if constexpr (Cfg::kConvS) {
          const int nchunks = p.conv_c / kBK;
          for (int c = 0; c < nchunks; ++c, ++a_cnt) {
            const uint32_t ab = a_cnt & 1u;
            mbar_wait(&afull_bar[ab], (a_cnt >> 1) & 1u);
            tc_fence_after();
            const uint64_t abuf = adesc0 + (uint64_t)((uint32_t)(ab * p.conv_planes * p.conv_plane_bytes) >> 4);
            for (int tap = 0; tap < p.conv_taps; ++tap) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              if (elect_one()) {
                // the tap reads the box from its (frame, height) offset on: a descriptor start moved by whole 128-byte rows
                const uint64_t ad = abuf + (uint64_t)p.conv_aoff[tap];
                const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)(stage * Cfg::kStageBytes) >> 4);
                mma(d_tmem, ad, bd, (c > 0 || tap > 0) ? 1u : 0u);
                mma(d_tmem, ad + a_kstep, bd + b_kstep, 1u);
                mma(d_tmem, ad + 2 * a_kstep, bd + 2 * b_kstep, 1u);
                mma(d_tmem, ad + 3 * a_kstep, bd + 3 * b_kstep, 1u);
                if (CG == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
              }
              __syncwarp();
              if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) { if (CG == 2) umma_commit_pair(&aempty_bar[ab]); else umma_commit(&aempty_bar[ab]); }
            __syncwarp();
          }
        } else
        for (int kb = kb0; kb < kb1; ++kb) {
          const int slab = SHARE ? 0 : kb / kb_per_slab;
          const int k_in = (kb - slab * kb_per_slab) * kBK;
          const int k16s = Cfg::kConvW ? p.conv_k16 : (min(kBK, p.kslab_len - k_in + 15) >> 4);   // partial last block (K % 64)
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == kb0) dbg_stamp(p, 1, tile_i, 2);
          if (kMerge) {
            if (elect_one()) {
              const uint64_t so = (uint64_t)((uint32_t)(stage * Cfg::kStageBytes) >> 4);
#pragma unroll
              for (int gq = 0; gq < 2; ++gq) {
                // group gq: both operands start a_shift / b_shift whole 128-byte k-rows into their tiles
                const uint64_t ad = adesc0 + so + (uint64_t)((uint32_t)sh_a[gq] * 8u);
                const uint64_t bd = bdesc0 + so + (uint64_t)((uint32_t)sh_b[gq] * 8u);
                const uint32_t dt = tmem_base + gq * Cfg::kAccCols;
                for (int k = 0; k < k16s; ++k) mma(dt, ad + k * a_kstep, bd + k * b_kstep, (kb > kb0 || k > 0) ? 1u : 0u);
              }
              umma_commit_pair(&empty_bar[stage]);
            }
          } else if (elect_one()) {
            const uint64_t so = (uint64_t)((uint32_t)(stage * Cfg::kStageBytes) >> 4);
            const int nb = SHARE ? p.kslabs : 1;
            for (int t = 0; t < nb; ++t) {
              // SHARE: slab t reads the A tile from row a_shift[t] on: the descriptor start moves by whole
              // 128-byte rows.  Measured on B200: the 128B swizzle is applied on the absolute shared-memory
              // address bits, so the descriptor's base-offset field stays 0 (setting it to the row phase,
              // as the CUTLASS comment on that field suggests for unaligned starts, gives wrong products).
              const uint64_t ad = adesc0 + so + (SHARE ? (uint64_t)((uint32_t)p.a_shift[t] * 8u) : 0ull);
              const uint64_t bd = bdesc0 + so + (uint64_t)((uint32_t)(t * Cfg::kBTileBytes) >> 4);
              const uint32_t first = (kb > kb0 || t > 0) ? 1u : 0u;
              if (k16s == 4) {
                mma(d_tmem, ad, bd, first);
                mma(d_tmem, ad + a_kstep, bd + b_kstep, 1u);
                mma(d_tmem, ad + 2 * a_kstep, bd + 2 * b_kstep, 1u);
                mma(d_tmem, ad + 3 * a_kstep, bd + 3 * b_kstep, 1u);
              } else {
                for (int k = 0; k < k16s; ++k)
                  mma(d_tmem, ad + k * a_kstep, bd + k * b_kstep, (first || k > 0) ? 1u : 0u);
              }
            }
            if (CG == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (kMerge) {
          if (elect_one()) { umma_commit_pair(&tfull_bar[0]); umma_commit_pair(&tfull_bar[1]); }
          __syncwarp();
          acc = 1;      // the bookkeeping below then flips the phase once per merged tile
        } else if (elect_one()) {
          if (CG == 2) umma_commit_pair(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        dbg_stamp(p, 1, tile_i, 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // =========================================== residual / C-in prefetch
    if (Cfg::kMayUseR && use_r) {
      int buf = 0; uint32_t round = 0;     // ring position of the running chunk counter
      for (int tile = ti.first; tile < ti.last; tile += ti.step) {
        int n_blk, m_row0, g, ks;
        ti.decode(tile, n_blk, m_row0, g, ks);
        for (int c64 = 0; c64 < BN; c64 += 64) {
          if (round > 0) mbar_wait(&rempty_bar[buf], (round + 1) & 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&rfull_bar[buf], kChunkBytes);
            tma_load_2d(smem_epi + buf * kChunkBytes, &p.tmR[g], &rfull_bar[buf], n_blk * BN + c64, m_row0);
          }
          __syncwarp();
          if (++buf == kRing) { buf = 0; ++round; }
        }
      }
    }
  } else if (warp == 2) {
    // ================================================== output store warp
    // Waits until the 8 epilogue warps have written a 64-column chunk, issues its TMA store and recycles
    // staging buffers as earlier stores drain -- so the math warps never wait on a store or on each other
    // (measured before: a 256-thread barrier + the store issue on the critical thread cost as much as the math).
    if (!kSplitK) {
      int buf = 0; uint32_t round = 0; int fbuf = 0; int k = 0;
      for (int tile = ti.first; tile < ti.last; tile += ti.step) {
        int n_blk, m_row0, g, ks;
        ti.decode(tile, n_blk, m_row0, g, ks);
        for (int c64 = 0; c64 < BN; c64 += 64, ++k) {
          mbar_wait(&cfull_bar[buf], round & 1);
          if (elect_one()) {
            tma_store_2d(&p.tmD[g], smem_epi + buf * kChunkBytes, n_blk * BN + c64, m_row0);
            tma_store_commit();
            tma_store_wait_read<Cfg::kStoreWait>();          // stores up to chunk k - kStoreWait have drained
            if (k >= Cfg::kStoreWait) mbar_arrive(&rempty_bar[fbuf]);
          }
          __syncwarp();
          if (++buf == kRing) { buf = 0; ++round; }
          if (k >= Cfg::kStoreWait && ++fbuf == kRing) fbuf = 0;
        }
      }
      dbg_stamp(p, 0, 7, 3);                 // last store issued
      // the staging buffers only have to be READ before the CTA exits; the writes complete asynchronously and are
      // ordered before the next kernel by the grid boundary (waiting for full completion cost ~1 us per launch)
      if (elect_one()) tma_store_wait_read<0>();
      __syncwarp();
      dbg_stamp(p, 0, 7, 4);                 // stores drained
    }
  } else if (warp >= 4) {
    // ========================================================= epilogue
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int hsel = (warp - 4) >> 2;       // which 32-column half of a 64-column chunk
    const int row_in_tile = q * 32 + lane;
    const int epi_tid = threadIdx.x - 128;
    const bool ref_round = kGeneric && (flags & EPI_REF_ROUND) != 0;
    uint32_t drop_seed = 0; float drop_inv = 1.0f;
    if ((kGeneric || (Cfg::kFlags & EPI_DROPOUT)) && (flags & EPI_DROPOUT)) {
      drop_seed = p.drop_seed ^ (p.drop_seed_dev ? __ldg(p.drop_seed_dev) : 0u);
      drop_inv = 1.0f / (1.0f - p.drop_p);
    }
    int acc = 0; uint32_t acc_phase = 0;
    int k = 0;   // running 64-column chunk counter
    int buf = 0; uint32_t round = 0;   // its position in the staging ring
    // per-column vectors (bias, bn scale/shift): double-buffered in smem; the loads for the NEXT tile are
    // issued at the start of the current one so their latency hides behind its chunks
    const bool use_vec = Cfg::kUsesVec && (flags & (EPI_BIAS | EPI_BN)) != 0;
    int vsel = 0;
    __half nbias = __float2half(0.f); float nscale = 0.f, nshift = 0.f;
    auto vec_fetch = [&](int tile) {
      if (epi_tid < BN && tile < ti.last) {
        int n_blk, m_row0, g, ks;
        ti.decode(tile, n_blk, m_row0, g, ks);
        const int col = n_blk * BN + epi_tid;
        const bool ok = col < p.N;
        const int vi = g * p.vec_gstride + col;
        if (flags & EPI_BIAS) nbias = ok ? p.bias[vi] : __float2half(0.f);
        if (flags & EPI_BN) {
          nscale = ok ? __ldg(p.bn_scale + vi) : 0.f;
          nshift = ok ? __ldg(p.bn_shift + vi) : 0.f;
        }
      }
    };
    auto vec_store = [&](int sel) {
      if (epi_tid < BN) {
        uint8_t* vb = smem_vec + sel * kVecBytes;
        if (flags & EPI_BIAS) reinterpret_cast<float*>(vb)[epi_tid] = __half2float(nbias);   // converted once per tile
        if (flags & EPI_BN) {
          reinterpret_cast<float*>(vb + 1024)[epi_tid] = nscale;
          reinterpret_cast<float*>(vb + 2048)[epi_tid] = nshift;
        }
      }
    };
    if (use_vec) { vec_fetch(ti.first); vec_store(0); }
    const uint32_t tempty_leader[2] = {CG == 2 ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u,
                                       CG == 2 ? mapa_u32(smem_u32(&tempty_bar[1]), 0) : 0u};
    int tile_i = 0;
    // merged groups: a tile is visited twice, group g's partial product sits in accumulator stage g
    for (int tile = ti.first; tile < ti.last; tile += ti.step, ++tile_i)
    for (int gg = 0; gg < (kMerge ? 2 : 1); ++gg) {
      int n_blk, m_row0, g, ks;
      ti.decode(tile, n_blk, m_row0, g, ks);
      const int prob = kMerge ? g : g >> 1;      // grouped launches: index into the problem table
      if (kMerge) g = gg;
      // accumulator row -> row of the output tile.  MODE 5: accumulator rows are (frame, slot) with conv_sl slots per frame,
      // the first conv_h of which are real outputs; they are compacted to (frame, height) in the staging chunk
      int out_idx = row_in_tile;
      bool row_ok = row_in_tile < ti.tile_rows;
      if (Cfg::kConvS) {
        const int fl = row_in_tile / p.conv_sl, sl = row_in_tile - fl * p.conv_sl;
        out_idx = fl * p.conv_h + sl;
        row_ok = sl < p.conv_h && out_idx < ti.tile_rows;
      }
      const int row = m_row0 + out_idx;
      const int n0 = n_blk * BN;
      const bool zero_row = p.zero_period != 0 && ((uint32_t)row % p.zero_period - p.zero_lo) >= p.zero_len;
      if (warp == 4) dbg_stamp(p, 2, tile_i, 0);

      const uint32_t s_bias = smem_u32(smem_vec + vsel * kVecBytes);            // fp32 [BN] each
      const uint32_t s_scale = s_bias + 1024, s_shift = s_bias + 2048;
      if (use_vec) {
        named_bar_sync(2, kEpiThreads);      // this tile's vectors (stored at the end of the previous tile) are visible
        vec_fetch(tile + ti.step);           // next tile's: in flight while this tile is processed
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (warp == 4) dbg_stamp(p, 2, tile_i, 1);
      const uint32_t t_acc = tmem_base + acc * Cfg::kAccCols + ((uint32_t)(q * 32) << 16);

      if constexpr (kSplitK) {
        const bool second = p.groups2_from > 0 && g >= p.groups2_from;
        int ws_ld = second ? p.ws_ld2 : p.ws_ld;
        int ws_tr = second ? p.ws_transposed2 : p.ws_transposed;
        float* ws_base = p.probs ? nullptr : p.ws[g < kMaxGroups ? g : 0];
        if (p.probs) {
          const GroupProb& pr = p.probs[prob];
          ws_base = pr.ws[g & 1]; ws_ld = pr.ws_ld; ws_tr = pr.ws_transposed;
        }
        float* ws_row = ws_base + (size_t)row * ws_ld;
        // The split_k work items of one output tile finish together and accumulate into the SAME addresses: each
        // starts its walk over the 32-column groups at a different group (rotated by its split index) so that they
        // hit different L2 lines at any one time.
        const int my_groups = hsel == 0 ? (BN / 32 + 1) / 2 : (BN / 32) / 2;
        const int rot = ks + (kMerge ? g : 0);
#pragma unroll 1
        for (int ci = 0; ci < my_groups; ++ci) {
          const int c = (hsel + 2 * ((ci + rot) % my_groups)) * 32;
          uint32_t v[32];
          tmem_ld_32x32(t_acc + c, v);
          tmem_ld_wait();
          if (ws_tr) {
            // D^T accumulation: element (row, col) -> ws[col*ws_ld + row]; the 32 lanes of a warp hold 32
            // consecutive rows, so each red is one coalesced 128-byte line
            if (row < p.M) {
              float* wcol = ws_base + (size_t)(n0 + c) * ws_ld + row;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + c + j < p.N)
                  asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(wcol + (size_t)j * ws_ld), "f"(__uint_as_float(v[j]) * p.alpha) : "memory");
            }
          } else if (row < p.M) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int col = n0 + c + j;
              if (col + 4 <= p.N) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(ws_row + col),
                             "f"(__uint_as_float(v[j]) * p.alpha),
                             "f"(__uint_as_float(v[j + 1]) * p.alpha),
                             "f"(__uint_as_float(v[j + 2]) * p.alpha),
                             "f"(__uint_as_float(v[j + 3]) * p.alpha)
                             : "memory");
              }
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c64 = 0; c64 < BN; c64 += 64, ++k) {
          uint8_t* sbuf = smem_epi + buf * kChunkBytes;
          const uint32_t srow = smem_u32(sbuf) + out_idx * 128;
          const int c = c64 + hsel * 32;          // first tile column of this thread's 32
          if (use_r) mbar_wait(&rfull_bar[buf], round & 1);                       // residual tile landed (buffer was free)
          else if (round > 0) mbar_wait(&rempty_bar[buf], (round + 1) & 1);      // the store that last used it drained
          if (warp == 4 && c64 < 256) dbg_stamp(p, 2, tile_i, 2 + 3 * (c64 >> 6));
          if (c < BN) {
            uint32_t v32[32];
            // (the mask word of this row -- one 4-byte load per lane, rows mask_ld words apart -- is requested before the
            //  accumulator read so that both latencies overlap)
            uint32_t maskword = 0, gm_word = 0xFFFFFFFFu;
            if (flags & EPI_GRADMASK)
              gm_word = (row < p.M && row_ok && n0 + c < p.N) ? __ldg(p.mask_in + (size_t)row * p.mask_ld + ((n0 + c) >> 5)) : 0u;
            if (!kGeneric) {
              tmem_ld_32x32(t_acc + c, v32);
              tmem_ld_wait();
            }
#pragma unroll(kGeneric ? 1 : 4)
            for (int j = 0; j < 4; ++j) {          // 8 columns = one 16-byte smem unit
              const int ct = c + j * 8;            // column inside the tile
              const uint32_t sptr = srow + (((hsel * 4 + j) ^ (out_idx & 7)) << 4);
              uint32_t v8[8];
              if (kGeneric) {
                tmem_ld_32x32_x8(t_acc + ct, v8);
                tmem_ld_wait();
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v8[e] = v32[j * 8 + e];
              }
              float r[8], bia[8], bsc[8], bsh[8];
              if (use_r) {
                const uint4 rv = lds128(sptr);
                float2 f;
                f = unpack_f16x2(rv.x); r[0] = f.x; r[1] = f.y;
                f = unpack_f16x2(rv.y); r[2] = f.x; r[3] = f.y;
                f = unpack_f16x2(rv.z); r[4] = f.x; r[5] = f.y;
                f = unpack_f16x2(rv.w); r[6] = f.x; r[7] = f.y;
              }
              if (flags & EPI_BIAS) {
#pragma unroll
                for (int e = 0; e < 8; e += 4) {
                  const float4 b4 = lds128f(s_bias + (ct + e) * 4);
                  bia[e] = b4.x; bia[e + 1] = b4.y; bia[e + 2] = b4.z; bia[e + 3] = b4.w;
                }
              }
              if (flags & EPI_BN) {
#pragma unroll
                for (int e = 0; e < 8; e += 4) {
                  const float4 s4 = lds128f(s_scale + (ct + e) * 4);
                  const float4 h4 = lds128f(s_shift + (ct + e) * 4);
                  bsc[e] = s4.x; bsc[e + 1] = s4.y; bsc[e + 2] = s4.z; bsc[e + 3] = s4.w;
                  bsh[e] = h4.x; bsh[e + 1] = h4.y; bsh[e + 2] = h4.z; bsh[e + 3] = h4.w;
                }
              }
              float xo[8];
              if constexpr (!kGeneric) {
                // specialised kinds: two columns per instruction with the packed fp32 FMA of sm_100 (FFMA2); each half
                // is the same fused multiply-add the scalar form compiles to, so the results are bit-identical
                const float2 alpha2 = make_float2(p.alpha, p.alpha), res2 = make_float2(p.res_scale, p.res_scale);
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  const int bit = j * 8 + e;
                  float2 x = make_float2(__uint_as_float(v8[e]), __uint_as_float(v8[e + 1]));
                  if (flags & EPI_BIAS) x = ffma2(x, alpha2, make_float2(bia[e], bia[e + 1]));
                  else x = fmul2(x, alpha2);
                  bool keep0 = true, keep1 = true;
                  if (flags & EPI_DROPOUT) {
                    keep0 = dropout_uniform(drop_seed, (uint32_t)row, (uint32_t)(n0 + ct + e)) > p.drop_p;
                    keep1 = dropout_uniform(drop_seed, (uint32_t)row, (uint32_t)(n0 + ct + e + 1)) > p.drop_p;
                  }
                  if (flags & EPI_RELU) {
                    x.x = relu_nan(x.x); x.y = relu_nan(x.y);
                    if (x.x > 0.0f && keep0) maskword |= (1u << bit);      // gradient gate: ReLU active AND kept by dropout
                    if (x.y > 0.0f && keep1) maskword |= (2u << bit);
                  }
                  if (flags & EPI_BN) x = ffma2(x, make_float2(bsc[e], bsc[e + 1]), make_float2(bsh[e], bsh[e + 1]));
                  if (flags & EPI_DROPOUT) {
                    x.x = keep0 ? x.x * drop_inv : 0.0f;
                    x.y = keep1 ? x.y * drop_inv : 0.0f;
                  }
                  if (flags & EPI_RESID) x = ffma2(res2, make_float2(r[e], r[e + 1]), x);
                  if (flags & EPI_GRADMASK) {
                    x.x = ((gm_word >> bit) & 1u) ? x.x : 0.0f;
                    x.y = ((gm_word >> bit) & 2u) ? x.y : 0.0f;
                  }
                  xo[e] = x.x; xo[e + 1] = x.y;
                }
              } else
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int bit = j * 8 + e;           // bit inside the 32-column mask word
                float x = __uint_as_float(v8[e]) * p.alpha;
                if (flags & EPI_BETA) x = fmaf(p.beta, r[e], round_h(x));   // cuBLAS rounds alpha*acc to fp16 first (measured)
                if (ref_round) x = round_h(x);
                if (flags & EPI_BIAS) {
                  x += bia[e];
                  if (ref_round) x = round_h(x);
                }
                const bool keep = !(flags & EPI_DROPOUT) || dropout_uniform(drop_seed, (uint32_t)row, (uint32_t)(n0 + ct + e)) > p.drop_p;
                if (flags & EPI_RELU) {
                  x = relu_nan(x);
                  if (x > 0.0f && keep) maskword |= (1u << bit);
                }
                if (flags & EPI_BN) {
                  x = fmaf(x, bsc[e], bsh[e]);
                  if (ref_round) x = round_h(x);
                }
                if (flags & EPI_DROPOUT) {
                  x = keep ? x * drop_inv : 0.0f;
                  if (ref_round) x = round_h(x);
                }
                if (flags & EPI_RESID) x = fmaf(p.res_scale, r[e], x);
                if (flags & EPI_GRADMASK) x = ((gm_word >> bit) & 1u) ? x : 0.0f;
                xo[e] = x;
              }
              uint4 ov;
              ov.x = pack_f16x2(xo[0], xo[1]); ov.y = pack_f16x2(xo[2], xo[3]);
              ov.z = pack_f16x2(xo[4], xo[5]); ov.w = pack_f16x2(xo[6], xo[7]);
              if (zero_row) ov = make_uint4(0u, 0u, 0u, 0u);
              if (!Cfg::kConvS || row_ok) sts128(sptr, ov);
            }
            if (zero_row) maskword = 0u;
            if ((flags & EPI_MASK) && row < p.M && row_ok && n0 + c < p.N)
              p.mask_out[(size_t)row * p.mask_ld + ((n0 + c) >> 5)] = maskword;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&cfull_bar[buf]);      // this warp's part of the chunk is in smem
          if (warp == 4 && c64 < 256) dbg_stamp(p, 2, tile_i, 3 + 3 * (c64 >> 6));
          if (++buf == kRing) { buf = 0; ++round; }
        }
        if (use_vec) {            // every warp passed this tile's vector barrier: the other buffer is free
          vsel ^= 1;
          vec_store(vsel);
        }
      }
      // all tcgen05.ld of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_leader[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
      if (warp == 4) dbg_stamp(p, 2, tile_i, 15);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer may still arrive on the leader's barriers
  if (warp == 2) {
    dbg_stamp(p, 0, 7, 5);                   // all roles done
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    dbg_stamp(p, 0, 7, 6);                   // TMEM released
  }
}

}  // namespace kfp16
