#!/usr/bin/env python
"""Restatement of the reference's cmd/fwdtest (cmd/fwdtest/main.go:14-204) on synthetic egs-shaped input:
build the CNN-TDNN model from xconfig, random-init weights, one minibatch of 64 sequences x 150 frames
(40-dim MFCC-like features + 100-dim ivector), time Network.Forward, print frames/sec
(BASELINE.json configs[0]; the reference main needs Go and /opt/kaldi egs files).

    python scripts/fwdtest.py [--check]      --check compares a small instance against the CPU oracle
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from kaldi_fp16_b200 import _lib, cudart, gpu, nnet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-seq", type=int, default=64)
    ap.add_argument("--seq-len", type=int, default=150)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    gpu.Init(0)
    free, total = gpu.MemoryInfo()
    print(f"GPU memory: {free / 2**30:.1f} / {total / 2**30:.1f} GiB free")
    h = gpu.NewHandle()
    st = cudart.Stream()
    lib.kfp16_ctx_set_stream(h.ptr, st.ptr)
    model = nnet.BuildModelFromString(bench.cnn_tdnn_xconfig())
    net = nnet.NewNetwork(model, h, args.n_seq, args.seq_len, train=False)
    print(f"model: {len(net.layers)} layers, {sum(r * c for r, c, _ in net.params.values()) / 1e6:.1f} M parameters")
    T = args.n_seq * args.seq_len
    rng = np.random.default_rng(1234)
    feats = rng.standard_normal((T, 40)).astype(np.float32) * (10.0 * 0.9 ** np.arange(40, dtype=np.float32))
    feats[:, 0] = np.clip(60 + 20 * rng.standard_normal(T), -20, 105)
    ivec = np.clip(rng.standard_normal((args.n_seq, 100)), -3, 3).astype(np.float32)
    out = net.Forward(feats, ivec)            # includes the H2D transfer and the D2H of the output, like the reference main
    print(f"output: {out.shape}, finite: {bool(np.isfinite(out).all())}, absmax {np.abs(out).max():.3f}")
    net.SetInput("input", feats)
    net.SetInput("ivector", ivec)
    for _ in range(3):
        assert lib.kfp16_net_forward(net.ptr) == 0
    cudart.synchronize()
    e0, e1 = cudart.Event(), cudart.Event()
    t0 = time.perf_counter()
    e0.record(st.ptr)
    for _ in range(args.iters):
        assert lib.kfp16_net_forward(net.ptr) == 0, _lib.last_error()
    e1.record(st.ptr)
    e1.synchronize()
    ms = e0.elapsed_ms(e1) / args.iters
    flops = lib.kfp16_net_flops_forward(net.ptr)
    print(f"forward: {ms:.3f} ms / minibatch  ->  {T / ms * 1e3:,.0f} frames/sec   ({flops / ms / 1e9:.0f} TFLOP/s over {flops / 1e9:.0f} GFLOP; "
          f"wall {1e3 * (time.perf_counter() - t0) / args.iters:.3f} ms)")
    if args.check:
        from oracle import kaldi_oracle as O
        from oracle.nnet_oracle import OracleNet
        n_seq, L = 2, 12
        on = OracleNet(bench.cnn_tdnn_xconfig(pdfs=96), n_seq, L)
        on.init_random(np.random.default_rng(3))
        small = nnet.NewNetwork(nnet.BuildModelFromString(bench.cnn_tdnn_xconfig(pdfs=96)), h, n_seq, L, train=False, ref_round=True)
        for k, w in on.params.items():
            small.SetParam(k, w)
        x = O.to_f16_rne(feats[: n_seq * L])
        iv = O.to_f16_rne(ivec[:n_seq])
        want = on.forward({"input": x, "ivector": iv})["output"]
        got = small.Forward(x, iv)
        err = O.max_err_vs_scale(got, want)
        print(f"oracle check (2 x 12 frames, 96 pdfs): max err / scale = {err:.2e}")
        assert err < 2e-3
        small.Free()
    net.Free()


if __name__ == "__main__":
    main()
