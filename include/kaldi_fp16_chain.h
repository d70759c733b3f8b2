/* kaldi_fp16_chain.h -- chain LF-MMI objective, batched over the sequences of a minibatch.
 *
 * B200-native form of the reference's chain objective path
 *   cpp/cuda/chain.cu:80-352 (forward / backward / posterior kernels), :475 chain_compute_loss,
 *   internal/nnet/chain_loss.go:221-294 ComputeChainLossBatch (per-sequence loop, ops_subsample_rows, FST upload per
 *   sequence, cudaMalloc + cudaMemcpy + host sync per call)
 * with the same arithmetic -- log-semiring forward-backward over a per-sequence numerator FST and a shared denominator
 * FST, loss = -(num_logprob - den_logprob), grad = clamp((den_post - num_post) * weight, +-30) stored as FP16 -- as ONE
 * kernel launch for the whole minibatch and no host involvement per frame or per sequence.
 *
 * The reference's core chain entry points (cpp/include/chain.h:47-160: chain_forward_backward, chain_compute_posteriors,
 * chain_compute_loss, chain_workspace_bytes, chain_last_error / chain_clear_error) are exported too, with the reference's
 * signatures, on top of these kernels (csrc/chain_compat.cu) -- internal/nnet/chain_loss.go links unchanged.  The den_* (leaky-HMM
 * denominator on host buffers), chain_num_* / chain_*_det and chain_backward_api.h symbols are NOT redefined: a Go build that
 * still wants them links those reference objects beside this library.
 */
#ifndef KALDI_FP16_CHAIN_H
#define KALDI_FP16_CHAIN_H

#include <stddef.h>
#include <stdint.h>

#include "kaldi_fp16_fused.h"

#ifdef __cplusplus
extern "C" {
#endif

/* FST in CSR form on the HOST (the layout of sparse.CSR / ChainFstGPU, cpp/include/chain.h:24-36):
 * arcs of state s are [row_ptr[s], row_ptr[s+1]); labels are pdf-ids, 1-indexed, 0 = epsilon (skipped, chain.cu:118);
 * weights and final weights are log-weights. */
typedef struct {
    const int32_t *row_ptr;      /* [num_states + 1] */
    const int32_t *col_idx;      /* [num_arcs] destination states */
    const int32_t *labels;       /* [num_arcs] */
    const float *weights;        /* [num_arcs] */
    const int32_t *final_states; /* [num_final] */
    const float *final_weights;  /* [num_final] */
    int num_states, num_arcs, num_final, start_state;
} kfp16_chain_fst;

typedef struct kfp16_chain kfp16_chain;

/* den: the denominator graph, shared by every sequence; frames_per_seq: output frames per sequence (after subsampling) */
kfp16_chain *kfp16_chain_create(kfp16_ctx *ctx, int num_pdfs, int n_seq, int frames_per_seq, const kfp16_chain_fst *den);
void kfp16_chain_destroy(kfp16_chain *chain);
/* numerator FSTs of the current minibatch, one per sequence (batch.PerSeqCSRs, train_step.go:183-193) */
int kfp16_chain_set_numerators(kfp16_chain *chain, const kfp16_chain_fst *nums, int n_seq);
/* nnet_out: FP16 network output (NOT log-softmax).  Output frame t of sequence s is the row
 *     s * seq_rows + row0 + t * row_step          (ld elements between rows)
 * i.e. row_step = the frame-subsampling factor, row0 = halo + left context in the executor's padded layout (the
 * reference materialises these rows with ops_subsample_rows, ops.cu:290-304).  grad_out (same addressing, may be NULL)
 * receives the FP16 gradient on exactly those rows.  loss_accum_dev (may be NULL) += sum over sequences of the loss. */
int kfp16_chain_loss(kfp16_chain *chain, const void *nnet_out, void *grad_out, int ld, int seq_rows, int row0, int row_step,
                     float supervision_weight, float *loss_accum_dev);
/* per sequence {num_logprob, den_logprob, loss, 0} of the last kfp16_chain_loss (ChainLossResult, chain.h:39-45) */
int kfp16_chain_read_results(kfp16_chain *chain, float *host, int n_seq);
/* two kernels compute the same thing: one keeps both graphs and the running alpha / beta vectors in shared memory (taken
 * whenever they fit), one works from global memory (any size).  on = 1 forces the second (tests). */
int kfp16_chain_force_general(kfp16_chain *chain, int on);
/* profiling: device int64[8] receiving clock64 stamps of the first sequence's phases (start, forward loop start / end,
 * backward loop start / end); NULL = off */
int kfp16_chain_set_debug(kfp16_chain *chain, void *dev_i64x8);
int kfp16_chain_num_sequences(const kfp16_chain *chain);
int kfp16_chain_frames(const kfp16_chain *chain);

/* ---- the reference's own interface (cpp/include/chain.h:24-160), same names / structs / semantics ------------------------- */
/* FST in CSR form with DEVICE pointers (chain.h:24-36) */
typedef struct {
    int32_t *row_ptr;      /* [num_states + 1] */
    int32_t *col_idx;      /* [num_arcs] destination states */
    int32_t *labels;       /* [num_arcs] pdf-ids, 1-indexed, 0 = epsilon */
    float *weights;        /* [num_arcs] log-weights */
    int32_t *final_states; /* [num_final] */
    float *final_weights;  /* [num_final] */
    int num_states, num_arcs, num_final, start_state;
} ChainFstGPU;
typedef struct { float num_logprob, den_logprob, loss; } ChainLossResult; /* chain.h:39-45 */

/* alpha / beta: caller-allocated fp32 [(T+1) x num_states] on the device (chain_workspace_bytes covers both);
 * *total_logprob (host) = logsum over final states of alpha[T][s] + final weight.  chain.h:47-66 */
int chain_forward_backward(const void *nnet_output, const ChainFstGPU *fst, int T, int num_pdfs, float *alpha, float *beta,
                           float *total_logprob);
/* posteriors: fp32 [T x num_pdfs] on the device, zeroed by the call.  chain.h:68-84 */
int chain_compute_posteriors(const void *nnet_output, const ChainFstGPU *fst, int T, int num_pdfs, const float *alpha,
                             const float *beta, float total_logprob, float *posteriors);
/* loss = -(num_logprob - den_logprob); grad_output (FP16 [T x num_pdfs], may be NULL) = clamp(den_post - num_post, +-30).
 * chain.h:86-108; here one launch of the batched objective kernel for a single sequence */
int chain_compute_loss(const void *nnet_output, const ChainFstGPU *num_fst, const ChainFstGPU *den_fst, int T, int num_pdfs,
                       void *grad_output, ChainLossResult *result);
size_t chain_workspace_bytes(int T, int num_states);
const char *chain_last_error(void);
void chain_clear_error(void);

#ifdef __cplusplus
}
#endif
#endif /* KALDI_FP16_CHAIN_H */
