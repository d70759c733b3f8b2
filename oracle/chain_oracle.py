"""CPU restatement of the reference's chain LF-MMI objective (TEST INFRASTRUCTURE ONLY -- nothing in
kaldi_fp16_b200/ imports it).

Follows /root/reference/cpp/cuda/chain.cu:
  * kernel_chain_forward  (80-138):  alpha[t+1][dst] (+)= alpha[t][src] + nnet[t][pdf-1] + w   in the log semiring,
        arcs with pdf <= 0 (epsilon) or pdf > P are skipped, alpha[0][start] = 0
  * kernel_chain_backward (140-183): beta[t][src] (+)= beta[t+1][dst] + nnet[t][pdf-1] + w,  beta[T][final] = final weight
  * kernel_total_logprob  (220-250): total = logsum over finals of alpha[T][f] + final weight
  * kernel_chain_posteriors (252-300): post[t][pdf-1] += exp(min(0, alpha[t][src] + nnet + w + beta[t+1][dst] - total))
  * kernel_chain_gradient (302-318) + chain_compute_loss (475-600): loss = -(num - den),
        grad = clamp((den_post - num_post) * weight, -30, 30) stored as FP16
and internal/nnet/chain_loss.go:221-294 (ComputeChainLossBatch): per sequence, output frame t = input frame
left_context + t * subsampling (gpu.SubsampleRows / ops_subsample_rows).

float64 accumulation (the reference accumulates in float32 with atomics in arbitrary order: agreement is to float32
rounding).  PINNED on the GPU box against the reference's own chain.cu compiled into oracle/_ref
(tests/test_chain_gpu.py); on the CPU its gradient is checked against finite differences of its own loss
(tests/test_chain_oracle_cpu.py), the check internal/nnet/backward_test.go:24-140 performs."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

LOG_ZERO = -1.0e30


@dataclass
class Fst:
    """host CSR, the layout of sparse.CSR / ChainFstGPU (cpp/include/chain.h:24-36)"""
    row_ptr: np.ndarray      # int32 [S+1]
    col_idx: np.ndarray      # int32 [A]
    labels: np.ndarray       # int32 [A]  pdf-id, 1-indexed, 0 = epsilon
    weights: np.ndarray      # float32 [A] log-weights
    final_states: np.ndarray
    final_weights: np.ndarray
    start_state: int = 0

    @property
    def num_states(self) -> int:
        return len(self.row_ptr) - 1

    @property
    def num_arcs(self) -> int:
        return len(self.col_idx)

    def src_of_arcs(self) -> np.ndarray:
        return np.repeat(np.arange(self.num_states, dtype=np.int32), np.diff(self.row_ptr))


def linear_chain_fst(T: int, num_pdfs: int, offset: int = 0) -> Fst:
    """the numerator fixture of internal/nnet/backward_test.go:42-58: state t -> t+1 on pdf (t + offset) % P, weight 0.
    Labels are 1-indexed here (the C structs' convention, chain.h:29)."""
    return Fst(row_ptr=np.concatenate([np.arange(T, dtype=np.int32), np.array([T, T], np.int32)]),
               col_idx=np.arange(1, T + 1, dtype=np.int32),
               labels=((np.arange(T) + offset) % num_pdfs + 1).astype(np.int32),
               weights=np.zeros(T, np.float32),
               final_states=np.array([T], np.int32), final_weights=np.zeros(1, np.float32), start_state=0)


def random_ergodic_fst(rng: np.random.Generator, num_states: int, arcs_per_state: int, num_pdfs: int) -> Fst:
    """small random ergodic HMM standing in for the denominator graph: every state has arcs_per_state
    outgoing arcs with random destinations / pdfs and normalised log-weights; every state is final with weight 0"""
    S, K = num_states, arcs_per_state
    col = rng.integers(0, S, size=S * K).astype(np.int32)
    lab = (rng.integers(0, num_pdfs, size=S * K) + 1).astype(np.int32)
    w = rng.random((S, K)) + 0.1
    w = np.log(w / w.sum(1, keepdims=True)).astype(np.float32).reshape(-1)
    return Fst(row_ptr=(np.arange(S + 1) * K).astype(np.int32), col_idx=col, labels=lab, weights=w,
               final_states=np.arange(S, dtype=np.int32), final_weights=np.zeros(S, np.float32), start_state=0)


def _logaddexp_at(dst: np.ndarray, idx: np.ndarray, vals: np.ndarray) -> None:
    """dst[idx] = log(exp(dst[idx]) + sum exp(vals)) per index, stable"""
    mx = np.full(dst.shape, -np.inf)
    np.maximum.at(mx, idx, vals)
    mx = np.maximum(mx, np.where(dst > LOG_ZERO, dst, -np.inf))
    live = np.isfinite(mx)
    acc = np.zeros(dst.shape)
    np.add.at(acc, idx, np.exp(vals - np.where(live, mx, 0.0)[idx]))
    acc += np.where((dst > LOG_ZERO) & live, np.exp(dst - np.where(live, mx, 0.0)), 0.0)
    dst[live] = mx[live] + np.log(acc[live])


def forward_backward(nnet: np.ndarray, fst: Fst):
    """nnet: [T x P] float (the FP16 network output as float).  Returns alpha [T+1 x S], beta [T+1 x S], total"""
    T, P = nnet.shape
    S = fst.num_states
    src, dst, pdf, w = fst.src_of_arcs(), fst.col_idx, fst.labels, fst.weights.astype(np.float64)
    ok = (pdf > 0) & (pdf <= P)
    src, dst, pdf, w = src[ok], dst[ok], pdf[ok] - 1, w[ok]
    nn = nnet.astype(np.float64)
    alpha = np.full((T + 1, S), LOG_ZERO)
    alpha[0, fst.start_state] = 0.0
    for t in range(T):
        a = alpha[t, src]
        live = a > LOG_ZERO
        if live.any():
            _logaddexp_at(alpha[t + 1], dst[live], a[live] + nn[t, pdf[live]] + w[live])
    fin = alpha[T, fst.final_states] + fst.final_weights.astype(np.float64)
    fin = fin[fin > LOG_ZERO]
    total = float(np.logaddexp.reduce(fin)) if fin.size else LOG_ZERO
    beta = np.full((T + 1, S), LOG_ZERO)
    beta[T, fst.final_states] = fst.final_weights.astype(np.float64)
    for t in range(T - 1, -1, -1):
        b = beta[t + 1, dst]
        live = b > LOG_ZERO
        if live.any():
            _logaddexp_at(beta[t], src[live], b[live] + nn[t, pdf[live]] + w[live])
    return alpha, beta, total


def posteriors(nnet: np.ndarray, fst: Fst, alpha, beta, total) -> np.ndarray:
    T, P = nnet.shape
    src, dst, pdf, w = fst.src_of_arcs(), fst.col_idx, fst.labels, fst.weights.astype(np.float64)
    ok = (pdf > 0) & (pdf <= P)
    src, dst, pdf, w = src[ok], dst[ok], pdf[ok] - 1, w[ok]
    nn = nnet.astype(np.float64)
    post = np.zeros((T, P))
    for t in range(T):
        a, b = alpha[t, src], beta[t + 1, dst]
        live = (a > LOG_ZERO) & (b > LOG_ZERO)
        lp = np.minimum(a[live] + nn[t, pdf[live]] + w[live] + b[live] - total, 0.0)
        np.add.at(post[t], pdf[live], np.exp(lp))
    return post


def chain_loss(nnet: np.ndarray, num: Fst, den: Fst, weight: float = 1.0):
    """chain_compute_loss: returns (num_logprob, den_logprob, loss, grad [T x P] as FP16-rounded float32)"""
    an, bn, tn = forward_backward(nnet, num)
    ad, bd, td = forward_backward(nnet, den)
    g = (posteriors(nnet, den, ad, bd, td) - posteriors(nnet, num, an, bn, tn)) * weight
    g = np.clip(g, -30.0, 30.0).astype(np.float32).astype(np.float16).astype(np.float32)
    return tn, td, -(tn - td), g


def chain_loss_batch(out: np.ndarray, nums: list, den: Fst, n_seq: int, seq_len: int, frames: int, subsampling: int,
                     left_context: int, weight: float = 1.0):
    """ComputeChainLossBatch on dense rows [n_seq*seq_len x P]: output frame t of sequence s = row
    s*seq_len + left_context + t*subsampling.  Returns (per-sequence [num, den, loss], gradient on the same dense rows)"""
    P = out.shape[1]
    grad = np.zeros_like(out, dtype=np.float32)
    res = np.zeros((n_seq, 3))
    for s in range(n_seq):
        rows = s * seq_len + left_context + np.arange(frames) * subsampling
        tn, td, loss, g = chain_loss(out[rows], nums[s], den, weight)
        res[s] = (tn, td, loss)
        grad[rows] = g
    return res, grad
