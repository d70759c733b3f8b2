"""Host-side mirror of the reference's chain-objective wrappers (internal/nnet/chain_loss.go) over the C ABI
(include/kaldi_fp16_chain.h): ``ChainFst`` is NewChainFstGPU's input (a sparse.CSR), ``ChainObjective`` owns the
denominator graph and the batched workspace, ``ComputeChainLossBatch`` is chain_loss.go:221-294 -- here one kernel
launch for the whole minibatch instead of a per-sequence loop.  All compute happens in libkaldi_fp16.so."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib, gpu
from ._lib import ChainFst as _CFst


class ChainError(gpu.GPUError):
    pass


@dataclass
class ChainLossResult:
    """chain_loss.go ChainLossResult: batch means, as ComputeChainLossBatch reports them (289-293)"""
    NumLogprob: float
    DenLogprob: float
    Loss: float
    PerSeq: np.ndarray     # [n_seq x 3] num_logprob, den_logprob, loss


class ChainFst:
    """sparse.CSR as NewChainFstGPU takes it (chain_loss.go:60-140): host arrays; labels 1-indexed pdf-ids"""

    def __init__(self, row_ptr, col_idx, labels, weights, final_states, final_weights, start_state: int = 0):
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int32)
        self.col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.labels = np.ascontiguousarray(labels, np.int32)
        self.weights = np.ascontiguousarray(weights, np.float32)
        self.final_states = np.ascontiguousarray(final_states, np.int32)
        self.final_weights = np.ascontiguousarray(final_weights, np.float32)
        self.start_state = int(start_state)

    def c_struct(self) -> _CFst:
        f = _CFst()
        f.row_ptr, f.col_idx, f.labels = self.row_ptr.ctypes.data, self.col_idx.ctypes.data, self.labels.ctypes.data
        f.weights, f.final_states = self.weights.ctypes.data, self.final_states.ctypes.data
        f.final_weights = self.final_weights.ctypes.data
        f.num_states, f.num_arcs = len(self.row_ptr) - 1, len(self.col_idx)
        f.num_final, f.start_state = len(self.final_states), self.start_state
        return f


class ChainObjective:
    def __init__(self, handle: gpu.Handle, num_pdfs: int, n_seq: int, frames_per_seq: int, den: ChainFst):
        self.lib = _lib.load()
        self.n_seq, self.frames, self.num_pdfs = n_seq, frames_per_seq, num_pdfs
        self._den = den
        d = den.c_struct()
        self.ptr = self.lib.kfp16_chain_create(handle.ptr, num_pdfs, n_seq, frames_per_seq, C.byref(d))
        if not self.ptr:
            raise ChainError(f"NewChainObjective: {_lib.last_error()}")

    def SetNumerators(self, nums: Sequence[ChainFst]) -> None:
        arr = (_CFst * len(nums))(*[f.c_struct() for f in nums])
        if self.lib.kfp16_chain_set_numerators(self.ptr, arr, len(nums)) != 0:
            raise ChainError(f"SetNumerators: {_lib.last_error()}")

    def Results(self) -> ChainLossResult:
        r = np.empty((self.n_seq, 4), np.float32)
        if self.lib.kfp16_chain_read_results(self.ptr, r.ctypes.data, self.n_seq) != 0:
            raise ChainError(f"chain results: {_lib.last_error()}")
        return ChainLossResult(float(r[:, 0].mean()), float(r[:, 1].mean()), float(r[:, 2].mean()), r[:, :3].copy())

    def Free(self) -> None:
        if self.ptr:
            self.lib.kfp16_chain_destroy(self.ptr)
        self.ptr = None


def ComputeChainLossBatch(net, chain: ChainObjective, subsampling: int = 3, left_context: int = 0,
                          supervision_weight: float = 1.0, layer: str = "") -> ChainLossResult:
    """chain_loss.go:221-294 on the network's current output: writes the gradient into the output layer's gradient
    buffer (rows of the subsampling grid, zero elsewhere) and returns the batch-mean log-probabilities / loss"""
    if net.lib.kfp16_net_loss_chain(net.ptr, layer.encode(), chain.ptr, subsampling, left_context, supervision_weight) != 0:
        raise ChainError(f"ComputeChainLossBatch: {_lib.last_error()}")
    return chain.Results()
