"""Data-parallel host logic on CPU: world_size-2 gloo processes, each running the numpy network
oracle on its shard of the minibatch's sequences; the all-reduced (summed) gradient bucket and the
SGD'd weights must equal the single-process step on the whole minibatch."""
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kaldi_fp16_b200 import dp
from oracle import kaldi_oracle as O
from oracle.nnet_oracle import OracleNet

XCONFIG = """
input name=input dim=32
linear-component name=lin0 dim=48
tdnnf-layer name=tdnnf1 dim=48 bottleneck-dim=16 time-stride=3 bypass-scale=0.66
tdnnf-layer name=tdnnf2 dim=48 bottleneck-dim=16 time-stride=3 bypass-scale=0.66
prefinal-layer name=prefinal small-dim=16 big-dim=48
output-layer name=output dim=24 include-log-softmax=false
"""
N_SEQ, L = 6, 20


def make_inputs():
    rng = np.random.default_rng(7)
    net = OracleNet(XCONFIG, N_SEQ, L)
    net.init_random(rng)
    x = O.to_f16_rne(rng.standard_normal((N_SEQ * L, 32)).astype(np.float32))
    return net.params, x


def flat(wg: dict, names) -> np.ndarray:
    return np.concatenate([np.asarray(wg[k], np.float32).reshape(-1) for k in names])


def step(params, x, n_seq):
    net = OracleNet(XCONFIG, n_seq, L)
    net.params = {k: v.copy() for k, v in params.items()}
    out = net.forward({"input": x})["output"]
    wg, _ = net.backward("output", out)         # 0.5*||out||^2: dY = Y (cmd/sgdtest/main.go:258-267)
    return wg, float(0.5 * np.sum(out.astype(np.float64) ** 2))


def worker(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        params, x = make_inputs()
        sh = dp.shard_sequences(N_SEQ, world, rank)
        wg, loss = step(params, x[sh.rows(L)], sh.n_seq)
        names = sorted(wg)
        bucket = torch.from_numpy(flat(wg, names).copy())
        red = dp.GradAllReducer(bucket)
        # bucketed exchange as the overlapped loop issues it: ranges from the top of the bucket down, asynchronously
        whole = bucket.clone()
        n = bucket.numel()
        cuts = [n, (2 * n) // 3, n // 4, 0]
        works = [red.all_reduce_range(lo, hi - lo, async_op=True) for hi, lo in zip(cuts, cuts[1:])]
        for w in works:
            w.wait()
        dist.all_reduce(whole, op=dist.ReduceOp.SUM)
        assert torch.equal(whole, bucket), "range-wise all-reduce differs from the whole-bucket all-reduce"
        obj = red.all_reduce_scalars(torch.tensor([loss], dtype=torch.float64))
        if rank == 0:
            q.put((bucket.numpy().copy(), float(obj[0])))
    finally:
        dist.destroy_process_group()


def test_shard_sequences_partitions_the_minibatch():
    for n, w in [(64, 1), (64, 2), (64, 8), (7, 3), (5, 5)]:
        shards = [dp.shard_sequences(n, w, r) for r in range(w)]
        assert shards[0].first_seq == 0 and sum(s.n_seq for s in shards) == n
        for a, b in zip(shards, shards[1:]):
            assert a.first_seq + a.n_seq == b.first_seq
        assert max(s.n_seq for s in shards) - min(s.n_seq for s in shards) <= 1
    assert dp.shard_sequences(64, 8, 3).rows(150) == slice(3 * 8 * 150, 4 * 8 * 150)
    with pytest.raises(ValueError):
        dp.shard_sequences(2, 4, 0)
    with pytest.raises(ValueError):
        dp.shard_sequences(8, 2, 2)


def test_two_rank_step_equals_single_process_step():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_bucket, got_obj = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params, x = make_inputs()
    wg, loss = step(params, x, N_SEQ)
    names = sorted(wg)
    want = flat(wg, names)
    # per-sequence clamping makes sequences independent units: the only difference is that each
    # rank rounds its partial weight gradient to FP16 before the sum (the reference stores FP16 grads)
    assert abs(got_obj - loss) <= 1e-6 * abs(loss)
    assert O.max_err_vs_scale(got_bucket, want) < 2e-3
    # identical SGD on every rank -> identical weights
    w32 = np.concatenate([params[k].reshape(-1) for k in names]).astype(np.float32)
    a, _, _ = O.sgd_update(w32, got_bucket, np.zeros_like(w32), 1e-3, 0.9)
    b, _, _ = O.sgd_update(w32, want, np.zeros_like(w32), 1e-3, 0.9)
    assert np.max(np.abs(a - b)) < 1e-4


# ------------------------------------------------------------------ train-mode batch-norm across ranks
def step_train_bn(params, x, n_seq, reducer=None, world=1):
    allreduce = None
    if reducer is not None:
        def allreduce(stats):
            t = torch.from_numpy(np.ascontiguousarray(stats, np.float32).copy())
            reducer(t, t.numel(), 0)
            return t.numpy()
    net = OracleNet(XCONFIG, n_seq, L, train=True, train_bn=True, bn_momentum=0.1, stats_allreduce=allreduce, world=world)
    net.params = {k: v.copy() for k, v in params.items()}
    out = net.forward({"input": x})["output"]
    wg, _ = net.backward("output", out)
    running = {k: (v["mean"].copy(), v["var"].copy()) for k, v in net.bn.items()}
    return wg, out, running


def worker_train_bn(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        params, x = make_inputs()
        sh = dp.shard_sequences(N_SEQ, world, rank)
        # the reducer the executor hook uses on the GPU box (there: a CUDA tensor over the device vector; here the tensor itself)
        red = dp.BNStatsAllReducer(lambda t, count: t)
        wg, out, running = step_train_bn(params, x[sh.rows(L)], sh.n_seq, red, world)
        assert red.calls == 4      # tdnnf1, tdnnf2, prefinal x 2
        names = sorted(wg)
        bucket = torch.from_numpy(flat(wg, names).copy())
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
        q.put((rank, bucket.numpy().copy(), out, {f"{k[0]}/{k[1]}": v for k, v in running.items()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_train_batchnorm_uses_global_statistics():
    """train-mode batch-norm under data parallelism: the [sum | sum of squares] vectors are summed over the ranks before
    the statistics are finalised, so both ranks normalise with the statistics of the WHOLE minibatch -- outputs, running
    statistics and the all-reduced gradients equal the single-process step"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker_train_bn, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120), q.get(timeout=120)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params, x = make_inputs()
    wg, out, running = step_train_bn(params, x, N_SEQ)
    want = flat(wg, sorted(wg))
    got_out = np.concatenate([res[0][2], res[1][2]], axis=0)
    assert O.max_err_vs_scale(got_out, out) < 2e-3
    for r in res:
        assert O.max_err_vs_scale(r[1], want) < 3e-3
        for k, (m, v) in running.items():
            gm, gv = r[3][f"{k[0]}/{k[1]}"]
            assert np.allclose(gm, m, rtol=1e-4, atol=1e-5) and np.allclose(gv, v, rtol=1e-4, atol=1e-5), k
    assert np.array_equal(res[0][1], res[1][1])
    # ... and differ from per-rank statistics (what a missing exchange would give)
    wg_local, out_local, _ = step_train_bn(params, x[:N_SEQ // 2 * L], N_SEQ // 2)
    assert O.max_err_vs_scale(out_local, out[:N_SEQ // 2 * L]) > 5e-3
