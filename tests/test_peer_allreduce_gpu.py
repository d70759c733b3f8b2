"""Gradient exchange over NVLink peer memory (kfp16_peer_allreduce_f16) ON THE GPUs: one process per GPU, the
buckets mapped into each other through CUDA IPC handles, one kernel per rank and exchange.

  * the exchanged bucket is, on EVERY rank, bit for bit half(sum over the ranks in FP32, in rank order) -- the numpy
    restatement below; repeated exchanges (the flags carry a step counter, nothing is reset) and a bucket whose size is not
    a multiple of the 16-byte vectors are covered, and so are two ranges of the bucket exchanged at the same time on two
    streams / flag channels with an unaligned cut, and a range shorter than one vector;
  * the time of one exchange of the benchmark's 36 MB bucket is written to gpurun_out/ (not asserted).

Skipped when fewer than 2 GPUs are visible (run with `gpurun --gpus 2`)."""
import socket
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COUNT = 1_000_003            # 125 000 vectors of 8 + 3 elements
EPOCHS = 4
BENCH_COUNT = 18_000_000     # the CNN-TDNN's bucket: 36 MB of FP16


def n_gpus() -> int:
    try:
        import ctypes
        n = ctypes.c_int(0)
        return n.value if ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0 else 0
    except OSError:
        return 0


def rank_data(rank: int, epoch: int) -> np.ndarray:
    rng = np.random.default_rng(1000 * epoch + rank)
    return rng.standard_normal(COUNT).astype(np.float16)


def expected(world: int, epoch: int) -> np.ndarray:
    acc = np.zeros(COUNT, np.float32)
    for r in range(world):                      # rank order, FP32 accumulation, one rounding: the kernel's arithmetic
        acc = acc + rank_data(r, epoch).astype(np.float32)
    return acc.astype(np.float16)


def run_rank(rank, world, port, q):
    import torch.distributed as dist

    from kaldi_fp16_b200 import _lib, cudart, dp, gpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = _lib.load()
    gpu.Init(rank)
    h = gpu.NewHandle()
    st = cudart.Stream()
    lib.kfp16_ctx_set_stream(h.ptr, st.ptr)
    t = gpu.TensorFromFP16(rank_data(rank, 0).reshape(1, -1))
    red = dp.PeerGradAllReducer(lib, h.ptr, t.Ptr, COUNT)
    lib.kfp16_peer_comm_set_timeout(red.comm, 10.0)
    bad = 0
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    for epoch in range(EPOCHS):
        if epoch:
            bits = rank_data(rank, epoch).view(np.uint16)       # (kept alive across the call)
            assert lib.bridge_transfer_fp16(t.Ptr, bits.ctypes.data, COUNT) == 0
        cudart.synchronize()
        red.all_reduce()
        red.check()
        got = t.ToBits().reshape(-1)
        want = expected(world, epoch).view(np.uint16)
        if not np.array_equal(got, want):
            idx = np.nonzero(got != want)[0]
            with open("gpurun_out/peer_allreduce_mismatch.txt", "a") as f:
                f.write(f"rank {rank} epoch {epoch}: {idx.size} differ, first at {idx[:12].tolist()}: got {got[idx[:4]].tolist()} "
                        f"want {want[idx[:4]].tolist()}\n")
            bad += int(idx.size)
    # the same sum as two ranges in flight together (an unaligned cut, small CTAs on a second stream and channel: what the
    # bucketed exchange beside the backward pass does), then a range without a whole 16-byte vector
    cut = 333_331
    bits = rank_data(rank, EPOCHS).view(np.uint16)
    assert lib.bridge_transfer_fp16(t.Ptr, bits.ctypes.data, COUNT) == 0
    cudart.synchronize()
    st2 = cudart.Stream()
    red.all_reduce_range(0, cut, channel=1, stream_ptr=st2.ptr, threads=64)
    red.all_reduce_range(cut, COUNT - cut, channel=0)
    red.check()
    want = expected(world, EPOCHS)
    bad += int(np.sum(t.ToBits().reshape(-1) != want.view(np.uint16)))
    red.all_reduce_range(5, 9, channel=2, threads=32)
    red.check()
    acc = np.zeros(9, np.float32)
    for _ in range(world):
        acc = acc + want[5:14].astype(np.float32)
    want[5:14] = acc.astype(np.float16)
    bad += int(np.sum(t.ToBits().reshape(-1) != want.view(np.uint16)))
    st2.destroy()
    red.close()
    t.Free()
    # one exchange of the benchmark's bucket, timed on the device between host barriers (zeros: sums stay finite)
    big = gpu.ZeroTensor(1, BENCH_COUNT)
    red = dp.PeerGradAllReducer(lib, h.ptr, big.Ptr, BENCH_COUNT)
    lib.kfp16_peer_comm_set_timeout(red.comm, 10.0)
    for _ in range(3):
        red.all_reduce()
    red.check()
    dist.barrier()
    e0, e1 = cudart.Event(), cudart.Event()
    e0.record(st.ptr)
    for _ in range(20):
        red.all_reduce()
    e1.record(st.ptr)
    red.check()
    us = round(e0.elapsed_ms(e1) * 1e3 / 20, 1)
    red.close()
    big.Free()
    q.put((rank, bad, us))
    dist.destroy_process_group()


def launch(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [ctx.Process(target=run_rank, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    deadline = time.time() + 300
    while len(res) < world:
        try:
            r, bad, us = q.get(timeout=5)
            res[r] = (bad, us)
        except Exception:
            if time.time() > deadline or any(p.exitcode not in (None, 0) for p in procs):
                for p in procs:
                    if p.is_alive():
                        p.kill()
                raise AssertionError(f"a rank failed: exit codes {[p.exitcode for p in procs]}")
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_allreduce_is_the_rank_ordered_fp32_sum_on_every_rank(world):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    res = launch(world)
    for r in range(world):
        assert res[r][0] == 0, f"rank {r}: {res[r][0]} of {EPOCHS * COUNT} elements differ from the rank-ordered FP32 sum"
    try:
        import os
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/peer_allreduce_n{world}.txt", "w") as f:
            f.write(f"kfp16_peer_allreduce_f16, {world} ranks, {BENCH_COUNT * 2 / 1e6:.0f} MB FP16 bucket: " +
                    ", ".join(f"rank {r} {res[r][1]} us" for r in range(world)) + " per exchange\n")
    except OSError:
        pass
