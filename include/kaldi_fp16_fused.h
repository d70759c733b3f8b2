/* kaldi_fp16_fused.h -- B200-native fused operator ABI (extension of the reference surface).
 *
 * The reference's Go layer (internal/gpu, internal/nnet) builds every layer out of
 * ops_gemm + 5..7 one-element-per-thread kernels + explicit transposes
 * (/root/reference/internal/gpu/ops.go:55-79,335-351, backward_ops.go:162-253,
 *  internal/nnet/forward.go:589-790).  These entry points expose the same math as ONE
 * tcgen05 GEMM launch with the surrounding ops fused into its epilogue and the
 * splice / transpose folded into TMA coordinates and UMMA operand majors.
 * Plain C: pointers are device pointers unless noted, sizes are ints, no C++/torch types.
 */
#ifndef KALDI_FP16_FUSED_H
#define KALDI_FP16_FUSED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- context: replaces the implicit (device 0, default stream, cuBLAS handle) state of
 * ops_cublas_create (cpp/cuda/ops.cu:336-348).  The handle returned by ops_cublas_create IS a
 * kfp16 context, so old callers keep working; multi-GPU callers create one per device. */
typedef struct kfp16_ctx kfp16_ctx;
kfp16_ctx *kfp16_ctx_create(int device_id);
void kfp16_ctx_destroy(kfp16_ctx *ctx);
/* cudaStream_t to launch on (NULL = legacy default stream, what the reference uses) */
int kfp16_ctx_set_stream(kfp16_ctx *ctx, void *cuda_stream);
void *kfp16_ctx_get_stream(kfp16_ctx *ctx);
int kfp16_ctx_num_sms(kfp16_ctx *ctx);
/* cap the persistent grid (0 = all SMs); used by tests to exercise multi-tile-per-CTA paths */
int kfp16_ctx_set_max_ctas(kfp16_ctx *ctx, int max_ctas);
/* profile mode: a CUDA-event pair is recorded around every GEMM launch made through this context;
 * profile_read waits for the stream, returns launches / summed kernel ms / summed 2*M*N*K and resets */
int kfp16_ctx_set_profile(kfp16_ctx *ctx, int on);
int kfp16_ctx_profile_read(kfp16_ctx *ctx, int *launches, double *total_ms, double *total_flops);
/* stream used by the context-free reference entry points (ops_relu, ...); NULL = default */
void kfp16_set_default_stream(void *cuda_stream);
/* number of kernels this library has launched in the calling process (bench "gpu_launches") */
unsigned long long kfp16_launch_count(void);
/* GEMM launches so far whose kernel was compiled for epilogue kind `kind` (0 generic run-time flags, 1 plain, 2 bias+relu+bn+mask,
 * 3 same + residual, 4 residual, 5 bn+relu-backward mask, 6 bn, 7 bias, 8 split-K): lets tests assert WHICH epilogue body ran */
unsigned long long kfp16_gemm_kind_launches(int kind);
const char *kfp16_last_error(void);

enum {
    KFP16_EPI_BIAS = 1 << 0,      /* + bias[n]            gpu.AddBias              ops.go:335 */
    KFP16_EPI_RELU = 1 << 1,      /* max(x,0)             ops_relu                 ops.cu:26 */
    KFP16_EPI_BN = 1 << 2,        /* x*scale[n]+shift[n]  ops_batchnorm_forward    ops.cu:171 */
    KFP16_EPI_RESID = 1 << 3,     /* res_scale*R + x      ops_add_scaled (bypass)  ops.cu:207 */
    KFP16_EPI_BETA = 1 << 4,      /* alpha*acc + beta*R   cublasGemmEx beta        ops.cu:381 */
    KFP16_EPI_REF_ROUND = 1 << 5, /* fp16 rounding between stages where the reference stores fp16 */
    KFP16_EPI_MASK = 1 << 6,      /* emit 1-bit relu mask (x>0) for the backward pass */
    KFP16_EPI_SPLITK = 1 << 7,    /* internal */
    KFP16_EPI_DROPOUT = 1 << 8,   /* inverted dropout (go/gotorch/layers.go:348-399 semantics) */
    KFP16_EPI_GRADMASK = 1 << 9   /* x = mask_in bit ? x : 0   (ops_relu_backward fused) */
};

/* operand storage */
enum { KFP16_K_MAJOR = 0, KFP16_MN_MAJOR = 1 };

typedef struct {
    const void *ptr; /* fp16, row-major as stored, points at row 0 */
    int rows, cols;  /* stored shape */
    int ld;          /* elements between rows (multiple of 8) */
    int halo;        /* rows readable before row 0 and after row rows-1 */
} kfp16_mat;

/* Implicit-GEMM convolution addressing (the conv-relu-batchnorm layer, internal/nnet/forward.go:418-524, without the
 * patch matrix): one operand of the GEMM is the 4-D activation tensor x[T][H][P][C] (frames, heights, parity planes,
 * channels; C innermost, C % 64 == 0) and the GEMM's K (mode 1) or M (mode 2) index runs over (tap, channel): tap s reads
 * x[t + dt[s]][h + hq[s]][par[s]][:], zeros outside [0,T) x [0,H) -- TMA out-of-bounds fill is the zero padding.
 *   mode 1 (forward, input gradient): A rows m = t*rows_h + h, K = ntaps*C; B as usual with brow[s] = where tap s's
 *           weights start (MN-major B: first k row; K-major B: first n row).  M = T*rows_h.
 *   mode 2 (weight gradient): dW[(s, c), n] = sum over (t, h) of x[t+dt[s]][h+hq[s]][par[s]][c] * B[(t*rows_h + h), n];
 *           A and B MN-major, fp32 accumulation into ws[0] (split_k >= 1).  M = ntaps*C, K = T*rows_h.
 * Height subsampling by 2 is a parity plane: x viewed as [T][H/2][2][C], tap dh -> par = dh & 1, hq = (dh - par) / 2. */
typedef struct {
    int mode;
    const void *x;
    int T, H, P, C;
    int rows_h;   /* heights per frame in the GEMM index (<= H; <= 128 in mode 1) */
    int ntaps;    /* 1..16 */
    int dt[16], hq[16], par[16], brow[16];
} kfp16_conv_addr;

typedef struct {
    int M, N, K; /* per-group problem; K = kslabs*kslab_len */
    /* A: K-major = stored [M x K]; MN-major = stored [K x M] (i.e. A^T, used by weight gradients)
     * B: K-major = stored [N x K] (i.e. B^T);  MN-major = stored [K x N] (what ops_gemm takes) */
    int a_major, b_major;
    kfp16_mat A, B;
    int groups;    /* 1..2 independent outputs sharing the launch */
    int kslabs;    /* 1..2 K-slabs (time splice) */
    int kslab_len; /* K per slab */
    /* storage-coordinate offsets added per (group, slab): rows / cols of the STORED matrix */
    int a_row_off[2][2], a_col_off[2][2];
    int b_row_off[2][2], b_col_off[2][2];
    void *D[2];       /* fp16 outputs [M x N], ld = ldd */
    int ldd;
    const void *R[2]; /* residual or C-in, fp16 [M x N], ld = ldr */
    int ldr;
    uint32_t flags;
    float alpha, beta, res_scale;
    const void *bias;      /* fp16 [N] */
    const float *bn_scale; /* fp32 [N]  gamma/sqrt(var+eps) */
    const float *bn_shift; /* fp32 [N]  beta - mean*scale   */
    int vec_gstride;       /* group g reads bias/bn at + g*vec_gstride */
    uint32_t *mask_out;
    const uint32_t *mask_in;
    int mask_ld; /* words per row */
    /* split-K: >1 accumulates alpha*partial into fp32 ws[g] (must be zeroed / hold the running sum) */
    int split_k;
    float *ws[2];
    int ws_ld;
    int ws_transposed; /* 1: the split-K target is D^T, fp32 [N x M] with ld ws_ld >= M (weight gradients computed as dY^T * X) */
    float drop_p;
    uint32_t drop_seed;
    int force_bn; /* 0 = heuristic tile width; else 64/128/160/256 */
    int force_generic; /* 1 = run the run-time-flag epilogue even when a specialised one exists (tests) */
    int force_cg;      /* 0 = heuristic; 1 = one CTA per 128-row tile; 2 = CTA pairs (cta_group::2, 256-row tiles) */
    int no_share;      /* 1 = load each splice slab's A tile separately even when one shifted tile could serve both;
                          4 = merge the two row-shifted groups of a split-K weight gradient into one tile pass (what the
                          grouped launch does; for a single problem only on request);
                          convolutions (conv.mode 1): 1 = every tap loads its own input box, 5 = one shared box per channel
                          chunk whatever the tile width (default: shared for N <= 128, where it measured faster) */
    void *debug_clock_buf; /* profiling: device int64 [grid][3 roles][8 tiles][16] clock64 stamps per warp role; NULL = off */
    /* A second split-K problem of the same shape (M, N, K, majors, groups) sharing the launch -- the two weight
     * gradients of one TDNN-F layer: tile groups [groups, 2*groups) compute A2^T * B2 into ws2[].  One launch and half
     * the split count (= half the fp32 reduction traffic) of two separate calls.  A2.ptr == NULL: none. */
    kfp16_mat A2, B2;
    int a2_row_off[2], b2_row_off[2]; /* row offsets per group of the second problem */
    float *ws2[2];
    int ws2_ld, ws2_transposed;
    kfp16_conv_addr conv; /* conv.mode != 0: the A operand is addressed as a convolution input (A.ptr unused) */
    /* KFP16_EPI_DROPOUT: element (row, col) is kept iff u(seed, row, col) > drop_p, kept values scaled by 1/(1-drop_p)
     * (inverted dropout, go/gotorch/layers.go:365-399), applied after the batch-norm and before the bypass; the emitted
     * mask bit is (ReLU active AND kept).  u is a counter-based hash (kfp16_dropout_uniform); seed = drop_seed XOR
     * *drop_seed_dev when that device word is given (a per-step counter, so CUDA-graph replays draw new masks). */
    const uint32_t *drop_seed_dev;
    /* zero_row_period > 0: output rows r with (r mod period) outside [lo, hi) are stored as zeros and get a zero mask --
     * the halo rows of the padded minibatch layout, which the next convolution reads as its zero padding in time */
    int zero_row_period, zero_row_lo, zero_row_hi;
} kfp16_gemm_desc;
/* the dropout hash, for callers that need the mask on the host: uniform in [0,1) */
float kfp16_dropout_uniform(uint32_t seed, uint32_t row, uint32_t col);

/* ---- grouped weight gradients: `count` split-K problems of the SAME shape (dW = A^T B with MN-major operands, two
 * row-shifted groups each -- the spliced weight gradients of many TDNN-F layers) in ONE persistent launch.  The
 * problem table (tensor maps, offsets, targets) is built once in device memory; the launch itself takes no per-call
 * host work, so it can be captured in a CUDA graph.  Against one launch per layer: no per-launch fixed cost, 2-3x
 * fewer partial-tile reductions (split_k is chosen for the whole set). */
typedef struct {
    kfp16_mat A, B;                   /* stored [K x M] and [K x N] */
    int a_row_off[2], b_row_off[2];   /* per group */
    float *ws[2];                     /* fp32 targets (accumulated into) */
    int ws_ld, ws_transposed;
} kfp16_wgrad_prob;
typedef struct kfp16_wgrad_group kfp16_wgrad_group;
kfp16_wgrad_group *kfp16_wgrad_group_create(kfp16_ctx *ctx, int M, int N, int K, const kfp16_wgrad_prob *probs, int count);
int kfp16_wgrad_group_launch(kfp16_ctx *ctx, kfp16_wgrad_group *grp);
void kfp16_wgrad_group_destroy(kfp16_wgrad_group *grp);

/* returns 0 on success, -1 on error (message via kfp16_last_error / ops_last_error) */
int kfp16_gemm_ex(kfp16_ctx *ctx, const kfp16_gemm_desc *d);

/* plain GEMM with operand majors: C[MxN] = alpha*op(A)*op(B) + beta*C, dense leading dims.
 * transA: A stored [K x M]; transB: B stored [N x K].  (kaldi_gemm semantics, cgo_interface.cu:206) */
int kfp16_gemm(kfp16_ctx *ctx, int M, int N, int K, float alpha, const void *A, int transA,
               const void *B, int transB, float beta, void *C);

/* ---- fused elementwise helpers used by the layer executor ---- */
/* bn_scale = gamma/sqrt(var+eps), bn_shift = beta - mean*bn_scale   (ops.cu:171-187) */
int kfp16_bn_fold(kfp16_ctx *ctx, const float *mean, const float *var, const float *gamma,
                  const float *beta, float eps, float target_rms, int D, float *scale, float *shift);
/* dZ = mask ? h(dY*scale[n]) : 0   (ops_batchnorm_backward + ops_relu_backward in one pass) */
int kfp16_bn_relu_backward(kfp16_ctx *ctx, const void *dY, int ldy, const float *scale,
                           const uint32_t *mask, int mask_ld, void *dZ, int ldz, int rows, int cols);
/* x[t,:] = h(x[t,:] + bias)  (gpu.AddBias, ops.go:335-351; the reference runs a K=1 GEMM) */
int kfp16_add_bias(kfp16_ctx *ctx, void *x, int ld, const void *bias, int rows, int cols);
/* out[n] = sum_t X[t,n]  (AffineBackwardBias, backward_ops.go:228-253), fp32 accumulate */
int kfp16_colsum(kfp16_ctx *ctx, const void *X, int ld, int rows, int cols, float *out_f32,
                 void *out_f16);
/* dst[i] = src[i] * c  (fp32 vectors: per-layer copies of batch-norm scales with a dropout factor folded in) */
int kfp16_scale_f32(kfp16_ctx *ctx, const float *src, float *dst, int n, float c);
/* *counter += 1 on the stream (the per-step dropout seed word) */
int kfp16_bump_counter(kfp16_ctx *ctx, uint32_t *counter_dev);
/* fp32 -> fp16 (round to nearest even), n elements */
int kfp16_f32_to_f16(kfp16_ctx *ctx, const float *src, void *dst, size_t n);
/* Padded activation layout: n_seq blocks of (seq_len + 2*halo) rows; X points at the very first
 * row (the first halo row of sequence 0).  pad_edges replicates each sequence's first / last real
 * row into its halo rows (the per-sequence form of the clamp in forward.go:714-722,760-770);
 * fold_edges is its adjoint: edge row += sum of its halo rows, halo rows = 0. */
int kfp16_pad_edges(kfp16_ctx *ctx, void *X, int ld, int n_seq, int seq_len, int cols, int halo);
int kfp16_fold_edges(kfp16_ctx *ctx, void *G, int ld, int n_seq, int seq_len, int cols, int halo);
/* multi-tensor SGD over one flat parameter bucket (ops_sgd_update semantics, backward_wrappers.cu:129-142)
 *   g = grad_is_f32 ? (round_grad ? float(half(g32)) : g32) : float(g16) */
int kfp16_sgd_update_flat(kfp16_ctx *ctx, float *w32, void *w16, const void *grad, int grad_is_f32,
                          int round_grad, float grad_scale, float *velocity, float lr,
                          float momentum, size_t n);

/* same update with {lr, momentum, grad_scale} read from a 3-float DEVICE block when the kernel runs: a captured CUDA
 * graph of the update follows SGDOptimizer.SetLR (internal/gpu/optimize.go:123) without re-capture */
int kfp16_sgd_update_flat_hp(kfp16_ctx *ctx, float *w32, void *w16, const void *grad, int grad_is_f32,
                             int round_grad, float *velocity, const float *hyper_dev, size_t n);
/* dst_f16[i] = half(src[i] * *scale_dev)  (scale_dev NULL = 1): FP32 gradient bucket -> the FP16 gradient tensors of
 * the reference (AffineBackwardWeights output, backward_ops.go:195-225) */
int kfp16_scale_f32_to_f16(kfp16_ctx *ctx, const float *src, void *dst_f16, size_t n, const float *scale_dev);

/* ---- padded minibatch layout: dense [n_seq*seq_len x cols] <-> padded [n_seq*(seq_len+2*halo) x ld] */
/* mode 0: halo rows = 0 (conv zero padding, forward.go:449)   1: replicate the edge row (splice clamp) */
int kfp16_pack_rows(kfp16_ctx *ctx, const void *src, void *dst, int ld, int n_seq, int seq_len, int halo,
                    int cols, int mode);
/* same from dense FP32 rows, converting FP32 -> FP16 round-to-nearest-even on the device (what
 * fp16.ConvertFloat32ToFloat16 does on the CPU for the features: internal/gpu/bridge.go:141, internal/fp16/fp16.go:13-70) */
int kfp16_pack_rows_f32(kfp16_ctx *ctx, const float *src, void *dst, int ld, int n_seq, int seq_len, int halo,
                        int cols, int mode);
int kfp16_unpack_rows(kfp16_ctx *ctx, const void *src, int ld, int col0, void *dst, int n_seq, int seq_len,
                      int halo, int cols);
/* dst[r, col0 + c] = src[r / blk, c]: per-sequence vector (ivector) appended to every frame of its block */
int kfp16_bcast_rows(kfp16_ctx *ctx, const void *src, int cols, void *dst, int ld, int col0, int rows, int blk);
/* adjoint of the broadcast: out[s, c] = sum over real rows of block s of G[r, col0 + c] */
int kfp16_seq_sum(kfp16_ctx *ctx, const void *G, int ld, int col0, void *out, int cols, int n_seq, int seq_len,
                  int halo);
int kfp16_zero_halo(kfp16_ctx *ctx, void *X, int ld, int n_seq, int seq_len, int cols, int halo);
/* y = h(x*scale[c] + shift[c]); shift may be NULL (batch-norm forward / backward with folded params) */
int kfp16_scale_shift(kfp16_ctx *ctx, const void *x, void *y, int rows, int cols, const float *scale,
                      const float *shift);
/* ---- train-mode batch-norm (batch statistics): cpp/cuda/cnn_kernels.cu:236-320 (training branch), go/gotorch/layers.go:257-300.
 * Row filter of all three calls: period > 0 restricts to rows r with lo <= r % period < lo + len (the real frames of the
 * padded layout: period = seq_len + 2*halo, lo = halo, len = seq_len; times the height count for conv activations).
 *  stats[0..cols) = sum, stats[cols..2*cols) = sum of squares (fp32; zeroed by the call).  A data-parallel job sums `stats`
 *  over the ranks before kfp16_bn_finalize and passes the global row count. */
int kfp16_bn_batch_stats(kfp16_ctx *ctx, const void *X, int ld, int rows, int cols, float *stats, int period, int lo, int len);
/* mean = S1/n, var = S2/n - mean^2 (biased); running = (1-momentum)*running + momentum*batch; scale = gamma/sqrt(var+eps)
 * (gamma NULL: target_rms), shift = beta - mean*scale, scale_bwd = scale*bwd_mul (may be NULL) */
int kfp16_bn_finalize(kfp16_ctx *ctx, const float *stats, double n_rows, int D, float *run_mean, float *run_var,
                      const float *gamma, const float *beta, float eps, float target_rms, float momentum, float bwd_mul,
                      float *scale, float *shift, float *scale_bwd);
/* Z = h(Z*scale[c % col_mod] + shift[c % col_mod] (+ res_scale*R)) in place on the filtered rows */
int kfp16_bn_apply(kfp16_ctx *ctx, void *Z, int ld, const float *scale, const float *shift, const void *R, int ldr,
                   float res_scale, int rows, int cols, int col_mod, int period, int lo, int len);
/* ---- restricted self-attention (attention-relu-batchnorm-layer, internal/nnet/forward.go:795-909: a CPU loop nest between a D2H and
 * an H2D copy in the reference).  proj: [rows x ldp] projection, per head [key (K) | value (V) | query key (K) | query context (C)],
 * C = 1 + n_left + n_right <= 32; output frame t attends to frames t + (o - n_left)*stride of its own sequence (zeros outside);
 * rows = the padded layout (n_seq blocks of seq_len + 2*halo, halo rows -> zeros).  z receives ReLU(out) (fp16, kept for the
 * backward pass: mask + attention weights), y = z*scale + shift (folded batch-norm), both [rows x ldy], heads*(V + C) columns. */
int kfp16_attention_forward(kfp16_ctx *ctx, const void *proj, int ldp, void *z, void *y, int ldy, const float *scale,
                            const float *shift, int n_seq, int seq_len, int halo, int heads, int key_dim, int value_dim,
                            int n_left, int n_right, int stride, float key_scale);
/* exact transpose: dproj [rows x ldp] from dy [rows x ldy]; db_scratch: fp32 [rows x heads x 32] */
int kfp16_attention_backward(kfp16_ctx *ctx, const void *proj, int ldp, const void *z, const void *dy, int ldy,
                             const float *scale, float *db_scratch, void *dproj, int n_seq, int seq_len, int halo, int heads,
                             int key_dim, int value_dim, int n_left, int n_right, int stride, float key_scale);
/* SpecAugment (go/gotorch/cnn_tdnn.go:612-668) on the padded layout: per sequence nfreq frequency masks of width <= fmax and
 * ntime time masks of width <= tmax are zeroed (halo rows copied through; x == y allowed).  The masks are drawn from the
 * counter-based generator kfp16_dropout_uniform(seed ^ *seed_dev, sequence, draw) -- see csrc/elementwise.cu for the draw
 * order -- so applying the same call to a gradient is the layer's backward pass. */
int kfp16_spec_augment(kfp16_ctx *ctx, const void *x, void *y, int ld, int n_seq, int seq_len, int halo, int dim, int fmax,
                       int nfreq, int tmax, int ntime, uint32_t seed, const uint32_t *seed_dev);
/* the same on row-strided views: row r of x / y starts ldx / ldy elements after row r-1 (cols, ldx, ldy multiples of 8) */
int kfp16_scale_shift_ld(kfp16_ctx *ctx, const void *x, long long ldx, void *y, long long ldy, int rows, int cols,
                         const float *scale, const float *shift);
/* X[r, :] = 0 for every row r that is not row0 + k*step (k >= 0): turns a gradient that was only written on the
 * objective's output frames (frame subsampling, ops_subsample_rows ops.cu:290-304) into a dense one */
int kfp16_zero_rows_except(kfp16_ctx *ctx, void *X, int ld, int rows, int cols, int row0, int step);
/* dY = Y on real rows (0 on halo rows); *loss_dev += 0.5*sum(Y^2)   (cmd/sgdtest/main.go:258-267) */
int kfp16_half_sq_loss(kfp16_ctx *ctx, const void *Y, void *dY, int n_seq, int seq_len, int halo, int cols,
                       float *loss_dev);
/* kfp16_bn_relu_backward + bias gradient in the same pass: db_accum[c] += sum_r dZ[r,c] (fp32, may be NULL) */
int kfp16_bn_relu_backward_bias(kfp16_ctx *ctx, const void *dY, int ldy, const float *scale,
                                const uint32_t *mask, int mask_ld, void *dZ, int ldz, int rows, int cols,
                                float *db_accum);
/* same on a padded minibatch [n_seq x (seq_len + 2*halo)] rows, with the adjoint of kfp16_pad_edges applied to dY
 * first and in place (edge row += its halo rows, halo rows = 0): one pass instead of fold + backward */
int kfp16_bn_relu_backward_bias_fold(kfp16_ctx *ctx, void *dY, int ldy, const float *scale, const uint32_t *mask,
                                     int mask_ld, void *dZ, int ldz, int n_seq, int seq_len, int halo, int cols,
                                     float *db_accum);
/* out_f32[n] += sum_t X[t,n]  (no memset: accumulates into the flat gradient bucket) */
int kfp16_colsum_accum(kfp16_ctx *ctx, const void *X, int ld, int rows, int cols, float *out_f32);

/* ---- CNN front-end: patch gather / scatter for conv-relu-batchnorm-layer
 * (the reference builds the patch matrix in Go on the CPU between a D2H and an H2D copy,
 *  internal/nnet/forward.go:429-455).  x: padded rows [n_seq*(seq_len+2*halo) x hin*fin], height-major;
 * P: [rows*hout x Kp] with P[(r*hout+ho), tap*fin+f] = x[r+dt[tap], (ho*sub+dh[tap])*fin+f], zero outside the
 * sequence's real frames / [0,hin) and in the K padding columns [ntaps*fin, Kp). */
int kfp16_im2col(kfp16_ctx *ctx, const void *x, void *P, int Kp, int n_seq, int seq_len, int halo, int hin,
                 int hout, int sub, int fin, int ntaps, const int *dt, const int *dh);
/* adjoint of kfp16_im2col: dx[r, h*fin+f] = sum of the dP entries that read it (fp32 sum, one fp16 rounding) */
int kfp16_col2im(kfp16_ctx *ctx, const void *dP, int Kp, void *dx, int n_seq, int seq_len, int halo, int hin,
                 int hout, int sub, int fin, int ntaps, const int *dt, const int *dh);

/* ---- egs feature decode on the device: Kaldi compressed matrices -> FP16 rows
 * (the reference decodes on the CPU while parsing: internal/parser/matrix.go:11-165, then converts to FP16 on the CPU,
 *  internal/gpu/bridge.go:141).  The payload is the matrix body as it sits in the archive, after the global header
 *  (format token, global min, global range, rows, cols -- parsed on the host, internal/parser/parser.go:300-365):
 *    KFP16_CM   "CM"  cols x 4 uint16 percentiles, then rows*cols bytes COLUMN-major      (ReadCompressedMatrix)
 *    KFP16_CM2  "CM2" rows*cols uint16 row-major                                            (ReadCompressedMatrix2)
 *    KFP16_CM3  "CM3" rows*cols uint8  row-major                                            (ReadCompressedMatrix3)
 *    KFP16_FM   "FM"  rows*cols float32 row-major                                           (ReadFullMatrix)
 * Same float32 arithmetic operation for operation, then round-to-nearest-even FP16: bit-identical to decode + convert
 * on the CPU.  One launch decodes up to 64 matrices (the sequences of a minibatch), each into dst rows [dst_row, +rows). */
enum { KFP16_CM = 1, KFP16_CM2 = 2, KFP16_CM3 = 3, KFP16_FM = 4 };
typedef struct {
    int format;
    int rows, cols;
    float global_min, global_range;
    size_t payload_offset; /* byte offset of this matrix's body inside the payload buffer (even for CM / CM2) */
    int dst_row;           /* first destination row */
} kfp16_cm_desc;
size_t kfp16_cm_payload_bytes(const kfp16_cm_desc *d);
int kfp16_decode_matrices(kfp16_ctx *ctx, const void *payload_dev, size_t payload_size, const kfp16_cm_desc *descs, int count,
                          void *dst_f16, int ld, int dst_rows);

#ifdef __cplusplus
}
#endif
#endif /* KALDI_FP16_FUSED_H */
