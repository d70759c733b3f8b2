"""CPU tests of the network-level oracle (oracle/nnet_oracle.py) and the CPU-baseline port
(oracle/gotorch_port.c): the reference's acceptance mains restated on the CPU
(cmd/backtest numerical-gradient check, cmd/sgdtest / cmd/traintest loss-decrease checks)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import kaldi_oracle as O
from oracle.nnet_oracle import OracleNet, idct_matrix, parse_xconfig

ROOT = Path(__file__).resolve().parent.parent

SGDTEST = """
input name=input dim=40
linear-component name=linear1 dim=128
batchnorm-component name=bn1
prefinal-layer name=prefinal small-dim=64 big-dim=128
output-layer name=output dim=40 include-log-softmax=false
"""
TRAINTEST = SGDTEST.replace("prefinal-layer", "tdnnf-layer name=tdnnf1 dim=128 bottleneck-dim=64 time-stride=0 bypass-scale=0.66\nprefinal-layer")
NUMGRAD = """
input name=input dim=40
idct-layer name=idct input=input dim=40
linear-component name=linear1 dim=128
batchnorm-component name=bn1
tdnnf-layer name=tdnnf1 dim=128 bottleneck-dim=64 time-stride=0 bypass-scale=0.66
tdnnf-layer name=tdnnf2 dim=128 bottleneck-dim=64 time-stride=0 bypass-scale=0.66
prefinal-layer name=prefinal input=tdnnf2 small-dim=64 big-dim=128
output-layer name=output dim=40 include-log-softmax=false
"""


def loss_of(net, x):
    out = net.forward({"input": x})["output"]
    return 0.5 * float(np.sum(out.astype(np.float64) ** 2)), out


def test_xconfig_parser_and_dims():
    layers = parse_xconfig(NUMGRAD)
    assert [l.name for l in layers] == ["input", "idct", "linear1", "bn1", "tdnnf1", "tdnnf2", "prefinal", "output"]
    net = OracleNet(NUMGRAD, 1, 4)
    assert net.params["tdnnf1.LinearW"].shape == (128, 64) and net.params["tdnnf1.AffineW"].shape == (64, 128)
    assert net.params["prefinal.BigW"].shape == (128, 128) and net.params["prefinal.SmallW"].shape == (128, 64)
    assert net.params["output.W"].shape == (64, 40)
    spliced = OracleNet("input name=input dim=32\ntdnnf-layer name=t dim=32 bottleneck-dim=16 time-stride=3\n", 2, 10)
    assert spliced.params["t.LinearW"].shape == (64, 16) and spliced.params["t.AffineW"].shape == (32, 32)   # true spliced shapes (Q1)


def test_idct_matrix_formula():
    """forward.go:1190-1210: cos(pi*j*(i+0.5)/D)*sqrt((j?2:1)/D), liftered by 1+(L/2)sin(pi*j/L)"""
    D, L = 40, 22.0
    M = idct_matrix(D, L)
    i, j = 3, 5
    want = np.cos(np.pi * j * (i + 0.5) / D) * np.sqrt(2.0 / D) * (1 + (L / 2) * np.sin(np.pi * j / L))
    assert abs(M[i, j] - want) <= 2.0 ** -10 * abs(want) + 1e-6      # stored through the truncating fp16 converter
    assert abs(M[7, 0] - np.sqrt(1.0 / D)) < 2e-4


def test_numerical_gradient_of_output_weights():
    """cmd/backtest/main.go:299-427: T=4 rows, eps=0.1 on 20 sampled output-layer weights, loss 0.5*||out||^2;
    a sample fails only if rel > 0.2 AND abs > 0.1 (the reference's acceptance rule)."""
    rng = np.random.default_rng(11)
    net = OracleNet(NUMGRAD, 1, 4)
    net.init_random(rng)
    x = O.to_f16_trunc(rng.uniform(-1, 1, (4, 40)).astype(np.float32))
    _, out = loss_of(net, x)
    wg, _ = net.backward("output", out)
    W = net.params["output.W"]
    eps, bad = 0.1, 0
    for _ in range(20):
        i, j = rng.integers(W.shape[0]), rng.integers(W.shape[1])
        w0 = W[i, j]
        W[i, j] = float(O.to_f16_trunc(np.array([w0 + eps], np.float32))[0])
        lp, _ = loss_of(net, x)
        W[i, j] = float(O.to_f16_trunc(np.array([w0 - eps], np.float32))[0])
        lm, _ = loss_of(net, x)
        W[i, j] = w0
        num = (lp - lm) / float(O.to_f16_trunc(np.array([w0 + eps], np.float32))[0] - O.to_f16_trunc(np.array([w0 - eps], np.float32))[0])
        ana = float(wg["output.W"][i, j])
        if abs(num - ana) / max(abs(num), 1e-6) > 0.2 and abs(num - ana) > 0.1:
            bad += 1
    assert bad == 0


def test_backward_matches_float64_autograd(monkeypatch):
    """The oracle's backward FORMULAS (transpose of the forward, spliced tdnnf + bypass + prefinal included --
    which the reference's own backward gets wrong, quirks Q2/Q8) against an independent float64 autograd of the
    same network.  The FP16 stores are switched off (h = identity) so only the formulas are compared."""
    import torch

    import oracle.nnet_oracle as NO

    ident = lambda v: np.asarray(v, dtype=np.float32)   # noqa: E731
    monkeypatch.setattr(O, "h", ident)
    monkeypatch.setattr(NO, "h", ident)
    xc = ("input name=input dim=32\nlinear-component name=lin dim=64\n"
          "tdnnf-layer name=t1 dim=64 bottleneck-dim=32 time-stride=3 bypass-scale=0.66\n"
          "prefinal-layer name=pf small-dim=24 big-dim=48\n"
          "output-layer name=output dim=16 include-log-softmax=false\n")
    n_seq, L, s = 2, 24, 3
    rng = np.random.default_rng(3)
    net = OracleNet(xc, n_seq, L)
    net.init_random(rng)
    x = rng.standard_normal((n_seq * L, 32)).astype(np.float32)
    out = net.forward({"input": x})["output"]
    wg, _ = net.backward("output", out)

    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in net.params.items()}

    def shift(a, d):
        idx = torch.clamp(torch.arange(L) + d, 0, L - 1)
        return a.reshape(n_seq, L, -1)[:, idx, :].reshape(n_seq * L, -1)

    bn = 1.0 / np.sqrt(1.0 + 1e-3)                      # identity batch-norm, eps = 1e-3
    h1 = torch.tensor(x, dtype=torch.float64) @ P["lin.W"]
    b = torch.cat([shift(h1, -s), h1], 1) @ P["t1.LinearW"]
    z = torch.relu(torch.cat([b, shift(b, s)], 1) @ P["t1.AffineW"] + P["t1.AffineBias"]) * bn
    y = z + 0.66 * h1
    g = torch.relu(y @ P["pf.BigW"] + P["pf.BigBias"]) * bn
    o = (g @ P["pf.SmallW"]) * bn @ P["output.W"] + P["output.Bias"]
    (0.5 * (o ** 2).sum()).backward()
    assert np.abs(o.detach().numpy() - out).max() < 1e-4
    for k, gk in wg.items():
        want = P[k].grad.numpy()
        assert np.abs(want - gk).max() <= 1e-5 * np.abs(want).max(), k


@pytest.mark.parametrize("xconfig,steps", [(SGDTEST, 20), (TRAINTEST, 10)])
def test_acceptance_loss_decreases(xconfig, steps):
    """cmd/sgdtest/main.go:196-321 (20 steps) and cmd/traintest/main.go:34-162 (10 steps): T=32 rows of U(-1,1),
    lr 1e-3, momentum 0.9, loss 0.5*||out||^2 with dY = Y; pass criterion loss[last] < loss[0]"""
    rng = np.random.default_rng(42)
    net = OracleNet(xconfig, 1, 32)
    net.init_random(rng)
    x = O.to_f16_trunc(rng.uniform(-1, 1, (32, 40)).astype(np.float32))
    losses, state = [], {}
    for _ in range(steps):
        l, out = loss_of(net, x)
        losses.append(l)
        wg, _ = net.backward("output", out)
        state = net.sgd(state, wg, 1e-3, 0.9)
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    assert all(np.abs(w).max() > 0 for k, w in net.params.items() if not k.endswith("ias"))   # traintest: weights non-zero


def test_sequences_are_independent_units():
    """per-sequence clamping (quirk Q3 fixed): a sequence's outputs do not depend on its neighbours in the
    minibatch -- the property data-parallel sharding relies on"""
    xc = ("input name=input dim=16\ntdnnf-layer name=t1 dim=16 bottleneck-dim=8 time-stride=3\n"
          "tdnnf-layer name=t2 dim=16 bottleneck-dim=8 time-stride=3\n")
    rng = np.random.default_rng(0)
    net = OracleNet(xc, 3, 12)
    net.init_random(rng)
    x = O.to_f16_rne(rng.standard_normal((36, 16)).astype(np.float32))
    full = net.forward({"input": x})["t2"]
    one = OracleNet(xc, 1, 12)
    one.params = net.params
    for s in range(3):
        assert np.array_equal(one.forward({"input": x[s * 12:(s + 1) * 12]})["t2"], full[s * 12:(s + 1) * 12])


# ------------------------------------------------------------------ CPU-baseline port of go/gotorch
@pytest.fixture(scope="module")
def port():
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "port"], check=True, capture_output=True)
    lib = C.CDLL(str(ROOT / "oracle" / "_build" / "libgotorch_port.so"))
    dp = C.POINTER(C.c_double)
    lib.gt_matmul.argtypes = [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.gt_tdnn_forward.argtypes = [dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    lib.gt_tdnn_backward.argtypes = [dp, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    lib.gt_bench_tdnnf_stack.restype = C.c_double
    lib.gt_bench_tdnnf_stack.argtypes = [C.c_int] * 7 + [dp]
    return lib


def P(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_gotorch_port_matmul_and_tdnn(port):
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((70, 33)), rng.standard_normal((33, 21))
    c = np.zeros((70, 21))
    port.gt_matmul(P(a), P(b), P(c), 70, 33, 21, 4)          # goroutine row split (go/gotorch/ops.go:49-81)
    assert np.allclose(c, a @ b, rtol=1e-12, atol=1e-12)
    batch, T, din, dout = 2, 9, 5, 4
    ctx = np.array([-3, 0], np.int32)
    x, W, bias = rng.standard_normal((batch, T, din)), rng.standard_normal((2 * din, dout)), rng.standard_normal(dout)
    y = np.zeros((batch, T, dout))
    port.gt_tdnn_forward(P(x), P(W), P(bias), P(y), batch, T, din, dout, ctx.ctypes.data_as(C.POINTER(C.c_int)), 2)
    idx = np.clip(np.arange(T)[:, None] + ctx[None, :], 0, T - 1)            # layers.go:444-481 clamped context
    want = np.concatenate([x[:, idx[:, 0]], x[:, idx[:, 1]]], axis=2) @ W + bias
    assert np.allclose(y, want, rtol=1e-12, atol=1e-12)
    gy = rng.standard_normal(y.shape)
    gW, gb, gx = np.zeros_like(W), np.zeros(dout), np.zeros_like(x)
    port.gt_tdnn_backward(P(x), P(W), P(gy), P(gW), P(gb), P(gx), batch, T, din, dout, ctx.ctypes.data_as(C.POINTER(C.c_int)), 2)
    S = np.concatenate([x[:, idx[:, 0]], x[:, idx[:, 1]]], axis=2).reshape(-1, 2 * din)
    assert np.allclose(gW, S.T @ gy.reshape(-1, dout)) and np.allclose(gb, gy.sum((0, 1)))
    cs = C.c_double()
    assert port.gt_bench_tdnnf_stack(2, 32, 8, 3, 1, 6, 1, C.byref(cs)) > 0 and np.isfinite(cs.value)


def test_spec_augment_masks_geometry():
    """oracle SpecAugment (go/gotorch/cnn_tdnn.go:612-668): one frequency band of width <= freq-max-proportion*dim over every
    frame and time bands of width <= time-mask-max-frames over every bin, per sequence; deterministic in the seed; the
    backward pass masks the gradient with the same pattern"""
    from oracle.nnet_oracle import OracleNet
    xc = ("input name=input dim=40\n"
          "linear-component name=pre dim=40\n"
          "spec-augment-layer name=sa freq-max-proportion=0.25 time-zeroed-proportion=0.3 time-mask-max-frames=8\n"
          "output-layer name=output include-log-softmax=false dim=16\n")
    n_seq, L = 5, 60
    rng = np.random.default_rng(0)
    on = OracleNet(xc, n_seq, L, train=True, dropout_seed=77, spec_augment=True)
    on.init_random(rng)
    x = (rng.standard_normal((n_seq * L, 40)) + 4).astype(np.float16).astype(np.float32)
    acts = on.forward({"input": x})
    keep = on.saved["sa"]["keep"].reshape(n_seq, L, 40)
    for s in range(n_seq):
        rows_all = ~keep[s].any(axis=1)                   # fully masked frames = time masks
        cols_all = ~keep[s].any(axis=0)                   # fully masked bins = the frequency mask
        assert cols_all.sum() <= 10
        assert np.array_equal(keep[s], ~(rows_all[:, None] | cols_all[None, :]))      # nothing but whole rows / columns
        if rows_all.any():                                # each band at most 8 frames long
            runs = np.diff(np.flatnonzero(np.diff(np.r_[0, rows_all.astype(int), 0])))[::2]
            assert runs.max() <= 8 * 8
    again = OracleNet(xc, n_seq, L, train=True, dropout_seed=77, spec_augment=True)
    again.params = on.params
    again.forward({"input": x})
    assert np.array_equal(again.saved["sa"]["keep"], on.saved["sa"]["keep"])
    other = OracleNet(xc, n_seq, L, train=True, dropout_seed=78, spec_augment=True)
    other.params = on.params
    other.forward({"input": x})
    assert not np.array_equal(other.saved["sa"]["keep"], on.saved["sa"]["keep"])
    # gradient: zero exactly where the features were masked
    wg, dact = on.backward("output", acts["output"], {})
    g = dact["pre"].reshape(n_seq, L, 40)
    assert not g[~keep].any() and np.abs(g[keep]).max() > 0
    off = OracleNet(xc, n_seq, L, train=True, dropout_seed=77)          # default: the reference executor's pass-through
    off.params = on.params
    assert np.array_equal(off.forward({"input": x})["sa"], off.acts["pre"])


def test_attention_backward_matches_float64_autograd(monkeypatch):
    """restricted self-attention (internal/nnet/forward.go:795-909): the oracle's forward against a direct restatement of
    the reference's loop nest (padded rows, per-head softmax over the context positions) and its backward -- the exact
    transpose, which the reference replaces by a plain affine backward (quirk Q2) -- against a float64 autograd"""
    import torch

    import oracle.nnet_oracle as NO

    ident = lambda v: np.asarray(v, dtype=np.float32)   # noqa: E731
    monkeypatch.setattr(O, "h", ident)
    monkeypatch.setattr(NO, "h", ident)
    H, V, K, nl, nr, s = 2, 8, 6, 2, 1, 2
    C = 1 + nl + nr
    per = 2 * K + V + C
    xc = ("input name=input dim=24\nlinear-component name=lin dim=32\n"
          f"attention-relu-batchnorm-layer name=att num-heads={H} value-dim={V} key-dim={K} num-left-inputs={nl} num-right-inputs={nr} time-stride={s}\n"
          "output-layer name=output dim=16 include-log-softmax=false\n")
    n_seq, L = 2, 11
    rng = np.random.default_rng(5)
    net = OracleNet(xc, n_seq, L)
    net.init_random(rng)
    assert net.params["att.W"].shape == (32, H * per) and net.by_name["att"].out_dim == H * (V + C)
    net.params["att.Bias"] = (rng.standard_normal((1, H * per)) * 0.1).astype(np.float32)
    x = rng.standard_normal((n_seq * L, 24)).astype(np.float32)
    acts = net.forward({"input": x})
    out = acts["output"]
    wg, _ = net.backward("output", out)
    ks = 1.0 / np.sqrt(K)
    bn = 1.0 / np.sqrt(1.0 + 1e-3)
    # the reference's loop nest on one sequence (forward.go:833-895), float64
    proj = (acts["lin"].astype(np.float64) @ net.params["att.W"] + net.params["att.Bias"])[:L]
    padded = np.zeros((L + (nl + nr) * s, H * per))
    padded[nl * s:nl * s + L] = proj
    want = np.zeros((L, H * (V + C)))
    for hd in range(H):
        for t in range(L):
            q = padded[t + nl * s, hd * per:(hd + 1) * per]
            b = np.array([q[2 * K + V + o] + ks * np.dot(q[K + V:2 * K + V], padded[t + o * s, hd * per:hd * per + K]) for o in range(C)])
            w = np.exp(b - b.max())
            w /= w.sum()
            for o in range(C):
                want[t, hd * (V + C):hd * (V + C) + V] += w[o] * padded[t + o * s, hd * per + K:hd * per + K + V]
            want[t, hd * (V + C) + V:(hd + 1) * (V + C)] = w
    assert np.abs(acts["att"][:L] - np.maximum(want, 0) * bn).max() < 1e-5
    # float64 autograd of the whole net
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in net.params.items()}
    h1 = torch.tensor(x, dtype=torch.float64) @ P["lin.W"]
    pj = (h1 @ P["att.W"] + P["att.Bias"]).reshape(n_seq, L, H, per)

    def ctx(a, d):
        o = torch.zeros_like(a)
        lo, hi = max(0, -d), min(L, L - d)
        if hi > lo:
            o[:, lo:hi] = a[:, lo + d:hi + d]
        return o

    key, val, qk, qc = pj[..., :K], pj[..., K:K + V], pj[..., K + V:2 * K + V], pj[..., 2 * K + V:]
    bb = torch.stack([qc[..., o] + ks * (qk * ctx(key, (o - nl) * s)).sum(-1) for o in range(C)], -1)
    ww = torch.softmax(bb, -1)
    u = sum(ww[..., o:o + 1] * ctx(val, (o - nl) * s) for o in range(C))
    a = torch.relu(torch.cat([u, ww], -1).reshape(n_seq * L, H * (V + C))) * bn
    o_ = a @ P["output.W"] + P["output.Bias"]
    (0.5 * (o_ ** 2).sum()).backward()
    assert np.abs(o_.detach().numpy() - out).max() < 1e-4
    for k, gk in wg.items():
        want_g = P[k].grad.numpy()
        assert np.abs(want_g - gk).max() <= 2e-5 * np.abs(want_g).max(), k
