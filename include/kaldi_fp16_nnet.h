/* kaldi_fp16_nnet.h -- network executor: the B200-native form of the reference's Go package
 * internal/nnet (Model / Network.Forward / Network.Backward / TrainStep) behind a C ABI.
 *
 * The reference walks the layer list in Go and issues 10-20 cgo calls per layer
 * (/root/reference/internal/nnet/forward.go:148-1001, network_backward.go:94-656,
 * train_step.go:142-283).  Here the same xconfig text is compiled ONCE into a fixed launch
 * plan over preallocated buffers (no per-op cudaMalloc), each layer is 1-2 fused tcgen05 GEMM
 * launches per direction, parameters / FP32 masters / velocities / gradients live in flat
 * buckets (one SGD launch, one all-reduce), and the whole step can be captured in a CUDA graph.
 *
 * Row layout: a minibatch is n_seq sequences of seq_len frames; every per-frame activation is
 * stored "padded": n_seq blocks of (seq_len + 2*halo) rows, so that time splicing is a TMA row
 * offset and clamps per sequence.  n_seq = 1 reproduces the reference's
 * whole-minibatch clamping exactly.
 */
#ifndef KALDI_FP16_NNET_H
#define KALDI_FP16_NNET_H

#include <stddef.h>
#include <stdint.h>

#include "kaldi_fp16_chain.h"
#include "kaldi_fp16_fused.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kfp16_net kfp16_net;

typedef struct {
    int n_seq;      /* sequences per minibatch (per GPU) */
    int seq_len;    /* frames per sequence */
    int ref_round;  /* 1: keep the reference's FP16 store between fused stages (bit-closer, same speed) */
    int train;      /* 1: allocate backward buffers, FP32 masters, velocities, gradient bucket */
    float lr;       /* SGDOptimizer.LR       (internal/gpu/optimize.go:30-38) */
    float momentum; /* SGDOptimizer.Momentum */
    int conv_cartesian; /* 1: time x height offsets as Cartesian product (Kaldi); 0: paired (quirk Q4) */
    float grad_scale;   /* captured SGD phase: g *= grad_scale (0 = 1.0), e.g. 1/frames or 1/(N*frames) */
    int round_grad;     /* captured SGD phase: round g to fp16 first (the reference's FP16 gradient tensors) */
} kfp16_net_opts;

/* xconfig: the reference's model description (internal/nnet/xconfig.go:143); returns NULL on error */
kfp16_net *kfp16_net_create(kfp16_ctx *ctx, const char *xconfig, const kfp16_net_opts *opts);
void kfp16_net_destroy(kfp16_net *net);

/* ---- introspection */
int kfp16_net_num_layers(const kfp16_net *net);
const char *kfp16_net_layer_name(const kfp16_net *net, int i);
const char *kfp16_net_layer_type(const kfp16_net *net, int i);
int kfp16_net_layer_dim(const kfp16_net *net, int i); /* output dim */
int kfp16_net_padded_rows(const kfp16_net *net);      /* n_seq*(seq_len+2*halo) */
int kfp16_net_halo(const kfp16_net *net);
double kfp16_net_flops_forward(const kfp16_net *net); /* 2*M*N*K over the GEMMs, real rows only */
/* same for the backward pass as executed: weight-gradient GEMMs of every layer on the gradient path plus the
 * input-gradient GEMMs that feed a layer with parameters (the xent branch gets no gradient, network_backward.go:104-107) */
double kfp16_net_flops_backward(const kfp16_net *net);
/* of those, the flops the last training step (or backward pass) did NOT execute because the objective only reads / writes
 * its subsampled output frames (kfp16_net_set_sparse_output_grad): executed = forward + backward - skipped */
double kfp16_net_flops_skipped(const kfp16_net *net);

/* ---- parameters: flat buckets.  name = "<layer>.<param>" as SGDOptimizer.RegisterParam keys
 * (optimize.go:52): W, LinearW, AffineW, AffineBias, BigW, BigBias, SmallW, Bias */
int kfp16_net_num_params(const kfp16_net *net);
const char *kfp16_net_param_name(const kfp16_net *net, int i);
int kfp16_net_param_shape(const kfp16_net *net, int i, int *rows, int *cols);
size_t kfp16_net_param_offset(const kfp16_net *net, int i); /* element offset into the buckets */
size_t kfp16_net_bucket_size(const kfp16_net *net);         /* elements */
void *kfp16_net_params_f16(kfp16_net *net);   /* device, fp16 [bucket]  weights used by the GEMMs */
float *kfp16_net_params_f32(kfp16_net *net);  /* device, fp32 master weights */
float *kfp16_net_velocity(kfp16_net *net);    /* device, fp32 */
float *kfp16_net_grads_f32(kfp16_net *net);   /* device, fp32 gradient bucket (all-reduce this) */
/* host fp32 [rows x cols] -> truncating fp16 (gpu.TensorFromFP32, tensor.go:67-91) -> device; the
 * FP32 master is set to float(fp16) as RegisterParam does (optimize.go:52-93) */
int kfp16_net_set_param(kfp16_net *net, const char *name, const float *host, int rows, int cols);
/* idct-layer: load the matrix instead of computing it (Kaldi's `idct` component, weight_loader.go:766-776); fp32 [dim x dim],
 * [in x out], through the truncating converter */
int kfp16_net_set_idct(kfp16_net *net, const char *layer, const float *host_f32, int dim);
int kfp16_net_get_param(kfp16_net *net, const char *name, uint16_t *host_f16, int rows, int cols);
/* which: for "batchnorm-component" layers "", for tdnnf "AffBN", prefinal "PfBN" / "BN", conv "BN" */
int kfp16_net_set_bn(kfp16_net *net, const char *layer, const char *which, const float *mean,
                     const float *var, const float *gamma, const float *beta, float eps, int dim);
/* randTensor rule N(0,1)*sqrt(2/(rows+cols)), zero biases, identity BN (forward.go:1161-1187) */
int kfp16_net_init_random(kfp16_net *net, uint64_t seed);

/* ---- data: dense host fp16 [n_seq*seq_len x dim] (per-frame) or [n_seq x dim] (per-sequence
 * inputs consumed through ReplaceIndex(x, t, 0)); uploads and lays the rows out padded */
int kfp16_net_set_input(kfp16_net *net, const char *input_name, const uint16_t *host_f16, int rows, int cols);
/* same from a device buffer already holding the dense fp16 rows (inputs resident in HBM) */
int kfp16_net_set_input_device(kfp16_net *net, const char *input_name, const void *dev_f16, int rows, int cols);
/* double-buffered asynchronous form: prefetch copies the NEXT minibatch's dense rows from PINNED host memory on a
 * copy stream (overlapping the current step), commit scatters them into the padded layout on the step's stream.
 * (the reference's TransferBatchPinned is a synchronous cudaMemcpy: internal/gpu/bridge.go:273-366, bridge.cu:257-267) */
int kfp16_net_prefetch_input(kfp16_net *net, const char *input_name, const uint16_t *host_pinned_f16, int rows, int cols);
int kfp16_net_commit_input(kfp16_net *net, const char *input_name);
/* FP32 forms: the features arrive as FP32 (as the egs hold them) and are converted FP32 -> FP16 round-to-nearest-even
 * ON THE DEVICE while they are scattered into the padded layout -- the conversion internal/gpu/bridge.go:141 does on
 * the CPU through fp16.ConvertFloat32ToFloat16 (internal/fp16/fp16.go:13-70).  commit_input handles either kind. */
int kfp16_net_set_input_f32(kfp16_net *net, const char *input_name, const float *host_f32, int rows, int cols);
int kfp16_net_prefetch_input_f32(kfp16_net *net, const char *input_name, const float *host_pinned_f32, int rows, int cols);
/* egs payloads straight from the archive: one compressed / full matrix per sequence (descs[i].rows = seq_len frames of
 * sequence i; dst_row is ignored -- the executor knows where sequence i lives in its padded layout).  The payload bytes
 * are uploaded (pinned or pageable host memory) and decoded + converted to FP16 on the device (kfp16_decode_matrices):
 * the on-device form of parser.ReadCompressedMatrix* + gpu.TransferBatch (internal/gpu/bridge.go:123-221). */
int kfp16_net_set_input_compressed(kfp16_net *net, const char *input_name, const void *payload_host, size_t payload_size,
                                   const kfp16_cm_desc *descs, int count);
int kfp16_net_forward(kfp16_net *net);
/* dense real rows of a layer's output -> host fp16 [n_seq*seq_len x dim] */
int kfp16_net_get_output(kfp16_net *net, const char *layer, uint16_t *host_f16, int rows, int cols);

/* ReLU mask (x > 0) saved by a tdnnf / prefinal layer's fused epilogue, one byte per element of the
 * dense real rows (debug / tests) */
int kfp16_net_get_mask(kfp16_net *net, const char *layer, uint8_t *host, int rows, int cols);

/* ---- training step pieces (train = 1) */
int kfp16_net_zero_grads(kfp16_net *net);
/* loss = 0.5*||out||^2 over real rows of `layer` ("" = the layer named "output"), dOut = out */
int kfp16_net_loss_half_sq(kfp16_net *net, const char *layer);
int kfp16_net_set_output_grad(kfp16_net *net, const char *layer, const uint16_t *host_f16, int rows, int cols);
/* chain LF-MMI objective on `layer` ("" = "output"): ComputeChainLossBatch (internal/nnet/chain_loss.go:221-294) in one
 * launch.  Output frame t of every sequence is the layer's frame left_context + t*subsampling; the gradient is written on
 * those rows of the layer's gradient buffer (zero elsewhere) and the accumulated loss (kfp16_net_read_loss) grows by the
 * sum over sequences of -(num_logprob - den_logprob).  The chain object must have been created for this network's
 * n_seq and hold the minibatch's numerator FSTs. */
int kfp16_net_loss_chain(kfp16_net *net, const char *layer, kfp16_chain *chain, int subsampling, int left_context,
                         float supervision_weight);
/* make the chain objective the one the captured step graph (phase 1) computes instead of 0.5*||out||^2; chain = NULL
 * switches back.  Call before kfp16_net_capture. */
int kfp16_net_set_chain(kfp16_net *net, kfp16_chain *chain, int subsampling, int left_context, float supervision_weight);
/* Frame subsampling (ops_subsample_rows before the loss, ops.cu:290-304 / chain_loss.go:221-294) makes the output
 * gradient zero on every row that is not an output frame.  on = 1 (default): the chain objective writes only the output
 * frames' rows and the row-wise layers behind the output (output, prefinal, linear, batch-norm) back-propagate exactly
 * those rows -- a third of their GEMM work at factor 3 -- while the first layer that mixes rows (time splice, convolution)
 * gets the dense gradient.  Same results as on = 0 (clear everything, dense backward); kfp16_net_get_grad returns the
 * dense form either way.  The training step (kfp16_net_launch / kfp16_net_capture, phase 1) with a chain objective also
 * restricts the FORWARD pass of the row-wise layers that feed nothing but the objective (prefinal-chain, output) to those
 * rows -- Kaldi evaluates only the requested output frames the same way; kfp16_net_forward itself always computes every row. */
int kfp16_net_set_sparse_output_grad(kfp16_net *net, int on);
/* on = 1 (default): a conv-relu-batchnorm layer with a single consumer gets dZ = mask ? h(dY * bn_scale) : 0 straight from
 * that consumer's input-gradient GEMM epilogue (ops_batchnorm_backward + ops_relu_backward, backward_wrappers.cu:41-115,
 * folded into the producing kernel) and only sums its bias gradient; on = 0: one elementwise pass per conv layer. */
/* Train-mode batch-norm (cpp/cuda/cnn_kernels.cu:236-320 training branch, go/gotorch/layers.go:257-330): on = 1 makes every
 * batch-norm of a training network normalise with the statistics of the current minibatch (all real frames; per filter over
 * frames x heights in conv layers; biased variance) and update the stored running statistics with `momentum`
 * (running = (1-m)*running + m*batch); the backward pass scales by gamma/sqrt(batch var + eps) and, like the reference,
 * does not differentiate through the statistics.  The producing GEMM then runs with an identity batch-norm and three small
 * passes follow it (kfp16_bn_batch_stats / _finalize / _apply).  on = 0 (default): running statistics, fused epilogue. */
int kfp16_net_set_train_batchnorm(kfp16_net *net, int on, float momentum);
/* Data-parallel training: `hook(user, stats_dev, count, stream)` is called between the statistics pass and the finalize
 * pass of every batch-norm with the fp32 [sum | sum of squares] vector on the device; it must sum it over the `world` ranks
 * (stream-ordered on `stream`) and return 0.  The step cannot be captured into a graph while a hook is set. */
typedef int (*kfp16_bn_stats_hook)(void *user, float *stats_dev, int count, void *stream);
int kfp16_net_set_bn_stats_hook(kfp16_net *net, kfp16_bn_stats_hook hook, void *user, int world);
/* running mean / variance of a layer's batch-norm (fp32 [dim]); `which` as in kfp16_net_set_bn */
int kfp16_net_get_bn(kfp16_net *net, const char *layer, const char *which, float *mean, float *var, int dim);
/* spec-augment-layer: the reference's GPU executor passes the features through (internal/nnet/forward.go:377-383, a TODO).
 * on = 1 (training networks): apply kfp16_spec_augment with one frequency mask of width <= freq-max-proportion*dim and
 * round(time-zeroed-proportion*seq_len / (time-mask-max-frames/2)) time masks of width <= time-mask-max-frames per
 * sequence, new masks every step (the per-step seed word of the dropout masks); the backward pass masks the gradient. */
int kfp16_net_set_spec_augment(kfp16_net *net, int on);
int kfp16_net_set_fuse_conv_backward(kfp16_net *net, int on);
/* on = 1 (default): in the training step the chain objective is queued on a second stream as soon as the output layer is
 * done, and the layers it does not depend on (the xent branch, which the reference's Forward computes as well) run beside
 * it on the SMs the objective's one-CTA-per-sequence kernel leaves free; both join before the backward pass.  Needs a
 * non-default context stream; on = 0: everything in order on one stream. */
int kfp16_net_set_overlap_loss(kfp16_net *net, int on);
int kfp16_net_backward(kfp16_net *net);
/* gradient wrt a layer's output, dense real rows (tests) */
int kfp16_net_get_grad(kfp16_net *net, const char *layer, uint16_t *host_f16, int rows, int cols);
/* v = m*v + g*grad_scale; w32 -= lr*v; w16 = half(w32) over the whole bucket (one launch);
 * round_grad = 1 rounds g to fp16 first, as the reference's FP16 gradient tensors do */
int kfp16_net_sgd_step(kfp16_net *net, float grad_scale, int round_grad);
/* SGDOptimizer.SetLR (internal/gpu/optimize.go:123).  lr / momentum / grad_scale live in a small device block the
 * update kernel reads when it runs, so this also takes effect for already captured graphs (stream-ordered). */
int kfp16_net_set_lr(kfp16_net *net, float lr);
int kfp16_net_set_momentum(kfp16_net *net, float momentum);
float kfp16_net_get_lr(const kfp16_net *net);
/* Dropout (tdnnf-layer dropout-proportion=p, training networks only; the reference has it on the CPU only,
 * go/gotorch/layers.go:348-399): inverted dropout after the layer's batch-norm, fused into the GEMM epilogue.  The mask
 * of layer index i at (padded row, column) is kfp16_dropout_uniform(seed ^ i*0x9E3779B9, row, col) > p; the seed word
 * lives on the device and is incremented at the start of every captured / phase-1 step. */
int kfp16_net_set_dropout_seed(kfp16_net *net, uint32_t seed);
int kfp16_net_get_dropout_seed(kfp16_net *net, uint32_t *seed);
/* FP16 gradient bucket: g16 = half(g32 * grad_scale), the FP16 gradient tensors of the reference
 * (backward_ops.go:195-225) as one flat buffer.  A data-parallel loop all-reduces THIS buffer (half the bytes of the
 * FP32 bucket) and applies kfp16_net_sgd_step_f16 (ops_sgd_update arithmetic on FP16 gradients, scale 1). */
int kfp16_net_grads_to_f16(kfp16_net *net);
void *kfp16_net_grads_f16(kfp16_net *net);
int kfp16_net_sgd_step_f16(kfp16_net *net);
/* accumulated loss since the last call (device->host sync) */
int kfp16_net_read_loss(kfp16_net *net, float *loss);
/* pipelined form: queue the download (+ reset) of the loss accumulated so far into pinned slot 0/1 behind the work
 * already in the stream; kfp16_net_wait_loss blocks on that copy only (the host can queue the next minibatch first) */
int kfp16_net_read_loss_async(kfp16_net *net, int slot);
int kfp16_net_wait_loss(kfp16_net *net, int slot, float *loss);

/* ---- CUDA graph of the step: phases bitmask 1 = zero_grads+forward+loss+backward, 2 = SGD (FP32 bucket),
 * 4 = kfp16_net_grads_to_f16, 8 = kfp16_net_sgd_step_f16; any combination is one graph.
 * Capture once (buffers are fixed), then launch per minibatch after kfp16_net_set_input*.
 * Capturing runs the phases once eagerly first (kernel attributes, grouped weight-gradient tables); weights,
 * velocities, both gradient buckets and the loss accumulator are saved before and restored after that pass, so a
 * capture does NOT take an optimiser step or change the training state (activations are recomputed scratch). */
int kfp16_net_capture(kfp16_net *net, int phases);
int kfp16_net_launch(kfp16_net *net, int phases);
/* The step graph (phases = 1) cut into up to nseg graphs along the backward pass: segment 0 = zero grads + forward +
 * objective + backward of the top layers, segment k = backward of the next layer group.  After segment k the gradient
 * bucket elements [first_elem, first_elem + count) of kfp16_net_segment_grads(k) are final, so a data-parallel loop
 * all-reduces them while segment k+1 computes.  Returns the number of segments captured, -1 on error. */
int kfp16_net_capture_segments(kfp16_net *net, int nseg);
/* extended form: cut_layers = NULL (equal parameter shares) or a comma-separated list of layer names, top of the network
 * first -- segment k back-propagates down to and including the k-th named layer, one more segment takes the rest;
 * export_f16 = 1: every segment ends by exporting its gradient range to the FP16 bucket (kfp16_net_grads_f16);
 * tail_max_ctas > 0: segments 1.. size their persistent grids for that many CTAs (SMs left to the collective that
 * reduces the previous segment's gradients beside them). */
int kfp16_net_capture_segments_ex(kfp16_net *net, int nseg, const char *cut_layers, int export_f16, int tail_max_ctas);
int kfp16_net_launch_segment(kfp16_net *net, int seg);
int kfp16_net_segment_grads(const kfp16_net *net, int seg, size_t *first_elem, size_t *count);
/* kernels launched by one forward+loss+backward(+sgd) pass of this network */
int kfp16_net_launches_per_step(const kfp16_net *net, int phases);

/* ---- Gradient exchange of the data-parallel step over NVLink peer memory (the reference is single-GPU,
 * cpp/cuda/bridge.cu:38-47 -- the exchange sits between Backward and the optimizer updates of
 * internal/nnet/train_step.go:212-221).  One process per GPU, up to 8 GPUs of one node.  Every rank creates a
 * communicator over its FP16 gradient bucket (kfp16_net_grads_f16; any 16-byte aligned cudaMalloc'ed buffer of `count`
 * FP16 elements, the same count on every rank), publishes its handle, and connects with the handles of all ranks in
 * rank order (the host exchanges the bytes: torch.distributed / MPI / a file).  kfp16_peer_allreduce_f16 then queues ONE
 * kernel on the context's stream that replaces the buckets of all ranks by their sum: each rank adds the slice it owns
 * over all ranks (FP32 accumulation in rank order, one rounding to FP16) through loads from, and stores to, the peers'
 * memory; flags in peer memory order the ranks, so every rank must queue the call once per step -- it can be captured
 * into a CUDA graph.  The result is bit-identical on all ranks.  A peer that does not arrive within the time limit
 * (default 20 s) makes the kernel give up; kfp16_peer_comm_status (synchronises the device) then returns -1. */
#define KFP16_PEER_HANDLE_BYTES 128
typedef struct kfp16_peer_comm kfp16_peer_comm;
kfp16_peer_comm *kfp16_peer_comm_create(kfp16_ctx *ctx, int rank, int world, void *bucket_f16, size_t count);
int kfp16_peer_comm_handle(kfp16_peer_comm *comm, void *handle_out /* KFP16_PEER_HANDLE_BYTES */);
int kfp16_peer_comm_connect(kfp16_peer_comm *comm, const void *handles /* world x KFP16_PEER_HANDLE_BYTES */);
int kfp16_peer_comm_set_timeout(kfp16_peer_comm *comm, double seconds);
int kfp16_peer_allreduce_f16(kfp16_peer_comm *comm);
/* The same exchange for the elements [first, first + count) of the bucket only (any alignment), on `stream` (NULL: the
 * context's stream), with CTAs of `threads` threads (0: 512) and at most `max_ctas` CTAs (0: one per SM).  Exchanges that may
 * be in flight at the same time -- a finished part of the gradient exchanged on a second stream beside the rest of the
 * backward pass (kfp16_net_capture_segments_ex) -- must use different channels (0..3); all ranks use the same channel for the
 * same range.  Small CTAs on every SM (threads = 64) keep the exchange thin beside the compute kernels. */
int kfp16_peer_allreduce_f16_range(kfp16_peer_comm *comm, size_t first, size_t count, int channel, int threads, int max_ctas,
                                   void *stream);
int kfp16_peer_comm_status(kfp16_peer_comm *comm);
/* every rank must have finished its last exchange before any rank destroys its communicator */
void kfp16_peer_comm_destroy(kfp16_peer_comm *comm);

#ifdef __cplusplus
}
#endif
#endif /* KALDI_FP16_NNET_H */
