"""Data-parallel plumbing for the training step: one process per GPU, sequences of the
egs minibatch sharded across ranks, ONE exchange step per minibatch -- the sum of the flat gradient
bucket over the ranks -- then the identical SGD update on every rank.

The reference has no multi-GPU path (single process, device 0: cpp/cuda/bridge.cu:38-47); this is
where nnet.TrainStep (internal/nnet/train_step.go:142-283) is split: shard before TransferBatch,
exchange between Backward (212) and the optimizer updates (221).

PeerGradAllReducer: the library's own exchange kernel over NVLink peer memory (FP16 bucket, what
bench.py --gpus N runs).  GradAllReducer: torch.distributed's all-reduce on the same bucket (NCCL
on the GPU box, gloo in the CPU tests).  torch.distributed is plumbing only; the buffer that is
summed is the library's own gradient bucket (kfp16_net_grads_f16 / kfp16_net_grads_f32), in place.
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    first_seq: int   # first sequence of the global minibatch owned by this rank
    n_seq: int       # sequences owned by this rank

    def rows(self, seq_len: int) -> slice:
        """rows of the dense [n_seq_global*seq_len x dim] minibatch matrix owned by this rank
        (sequences are contiguous row blocks: internal/batch merges examples back to back)"""
        return slice(self.first_seq * seq_len, (self.first_seq + self.n_seq) * seq_len)

    def seqs(self) -> slice:
        return slice(self.first_seq, self.first_seq + self.n_seq)


def shard_sequences(n_seq_global: int, world: int, rank: int) -> Shard:
    """Sequences [r*B/N, (r+1)*B/N) go to rank r; a remainder is spread over the first ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    if n_seq_global < world:
        raise ValueError(f"cannot shard {n_seq_global} sequences over {world} ranks")
    base, rem = divmod(n_seq_global, world)
    first = rank * base + min(rank, rem)
    return Shard(rank, world, first, base + (1 if rank < rem else 0))


def env_rank_world() -> tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (1 process per GPU)"""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


class GradAllReducer:
    """Sum all-reduce of the flat gradient bucket (and of the scalar objective).

    `bucket` is a torch tensor aliasing the bucket memory (CUDA tensor over kfp16_net_grads_f32 on
    the GPU box, a CPU tensor under gloo).  Sum semantics: the N-rank step equals the 1-rank step on
    the concatenated minibatch up to FP32 summation order (the reference does not scale gradients
    by 1/B: internal/nnet/chain_loss.go:289-293 divides only the reported objective)."""

    def __init__(self, bucket, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.bucket = bucket
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def all_reduce(self, async_op: bool = False):
        if self.world == 1:
            return None
        return self.dist.all_reduce(self.bucket, op=self.dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def all_reduce_range(self, first: int, count: int, async_op: bool = True):
        """all-reduce bucket[first : first + count] only (the part of the gradient a finished segment of the backward
        pass has completed: kfp16_net_segment_grads), by default asynchronously so that it overlaps the next segment"""
        if self.world == 1 or count <= 0:
            return None
        return self.dist.all_reduce(self.bucket[first:first + count], op=self.dist.ReduceOp.SUM, group=self.group,
                                    async_op=async_op)

    def all_reduce_scalars(self, t):
        """objective / frame-count exchange (3 floats in the chain objective, 1 for 0.5*||out||^2)"""
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t


class BNStatsAllReducer:
    """Cross-rank sum of the batch-norm statistics (train-mode batch-norm): plugged into the executor with
    Network.SetBNStatsHook(reducer, reducer.world).  The executor calls it between the statistics pass and the finalize
    pass of every batch-norm with the device address of the fp32 [sum | sum of squares] vector; `as_tensor(ptr, count)`
    must return a torch tensor aliasing that memory (CUDA: torch.as_tensor over __cuda_array_interface__)."""

    def __init__(self, as_tensor, group=None):
        import torch.distributed as dist

        self.dist, self.group, self.as_tensor = dist, group, as_tensor
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.calls = 0

    def __call__(self, ptr: int, count: int, stream: int) -> None:
        self.calls += 1
        if self.world > 1:
            self.dist.all_reduce(self.as_tensor(ptr, count), op=self.dist.ReduceOp.SUM, group=self.group)


class PeerGradAllReducer:
    """Sum all-reduce of the FP16 gradient bucket by the library's own kernel over NVLink peer memory
    (kfp16_peer_allreduce_f16, include/kaldi_fp16_nnet.h): every rank reduces the slice it owns through loads from the
    peers' buckets and stores the rounded sums into all of them -- one launch on the step's stream between the step graph
    and the SGD graph, no library collective.  torch.distributed only carries the 128-byte memory handles at set-up.

    `bucket_ptr` / `count`: kfp16_net_grads_f16 and kfp16_net_bucket_size (any cudaMalloc'ed FP16 buffer).  Raises when
    the peers cannot be mapped (GPUs of different nodes, no peer access): the caller then keeps GradAllReducer."""

    def __init__(self, lib, handle_ptr, bucket_ptr, count, group=None, timeout_s: float = 20.0):
        import ctypes as C

        import torch.distributed as dist

        from . import _lib

        self.lib, self.dist, self.group = lib, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.comm = lib.kfp16_peer_comm_create(handle_ptr, self.rank, self.world, bucket_ptr, count)
        ok = bool(self.comm)
        err = "" if ok else _lib.last_error()
        mine = b""
        if ok:
            buf = C.create_string_buffer(128)
            ok = lib.kfp16_peer_comm_handle(self.comm, buf) == 0
            err = "" if ok else _lib.last_error()
            mine = bytes(buf.raw)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (ok, mine, int(count)), group=group)
        if ok and all(g[0] for g in gathered):
            if len({g[2] for g in gathered}) != 1:
                ok, err = False, f"bucket sizes differ between the ranks: {[g[2] for g in gathered]}"
            else:
                ok = lib.kfp16_peer_comm_connect(self.comm, b"".join(g[1] for g in gathered)) == 0 and \
                    lib.kfp16_peer_comm_set_timeout(self.comm, float(timeout_s)) == 0
                err = "" if ok else _lib.last_error()
        elif ok:
            ok, err = False, "a peer could not export its bucket"
        # all ranks agree on the outcome before anyone launches (a rank that waits for a peer that gave up would time out)
        agreed = [None] * self.world
        dist.all_gather_object(agreed, ok, group=group)
        if not all(agreed):
            if self.comm:
                lib.kfp16_peer_comm_destroy(self.comm)
            self.comm = None
            raise RuntimeError(f"peer-memory gradient exchange unavailable on rank {self.rank}: {err or 'a peer failed'}")

    def all_reduce(self) -> None:
        """queue the exchange on the context's stream (every rank, once per step)"""
        from . import _lib

        if self.lib.kfp16_peer_allreduce_f16(self.comm) != 0:
            raise RuntimeError(_lib.last_error())

    def all_reduce_range(self, first: int, count: int, channel: int = 0, stream_ptr: int = 0, threads: int = 0,
                         max_ctas: int = 0) -> None:
        """the exchange for bucket[first : first + count] only, on `stream_ptr` (0: the context's stream); ranges that may
        be in flight together (one beside the rest of the backward pass) use different channels"""
        from . import _lib

        if self.lib.kfp16_peer_allreduce_f16_range(self.comm, first, count, channel, threads, max_ctas, stream_ptr or None) != 0:
            raise RuntimeError(_lib.last_error())

    def check(self) -> None:
        """synchronise the device and raise if a peer failed to arrive in an exchange"""
        from . import _lib

        if self.lib.kfp16_peer_comm_status(self.comm) != 0:
            raise RuntimeError(_lib.last_error())

    def close(self) -> None:
        if self.comm:
            self.lib.kfp16_peer_comm_status(self.comm)
            self.dist.barrier(self.group)        # nobody unmaps memory a peer's kernel may still touch
            self.lib.kfp16_peer_comm_destroy(self.comm)
            self.comm = None
