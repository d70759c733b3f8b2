"""A whole network forward through the REFERENCE'S OWN operator library (oracle/_ref/libkaldi_fp16_ref.so, compiled
unmodified from /root/reference/cpp), driven in the order its Go executor issues the operators:

    forwardLinear    forward.go:333-346   ops_gemm
    forwardTDNNF     forward.go:589-695   splice copies + concat, ops_gemm, splice, ops_gemm, AddBias (K = 1 GEMM, beta = 1:
                                          internal/gpu/ops.go:335-351), ops_relu, ops_batchnorm_forward, ops_add_scaled
    forwardPrefinal  forward.go:912-968   ops_gemm, AddBias, ops_relu, ops_batchnorm_forward, ops_gemm, ops_batchnorm_forward
    forwardOutput    forward.go:971-1001  ops_gemm, AddBias, ops_log_softmax

This pins the NETWORK-LEVEL oracle (oracle/nnet_oracle.py, a numpy composition of the operator oracle in the same order)
to outputs of the reference itself, layer by layer, and compares the product executor with the same outputs directly.
One sequence (n_seq = 1), where the reference's whole-matrix splice clamp and the per-sequence clamp coincide (SURVEY quirk
Q3).  Tolerance as SURVEY 8c: <= 2e-3 of the tensor scale per activation (cuBLAS and numpy order the FP32 sums differently;
every elementwise operator is bit-identical on its own: tests/test_ops_gpu.py)."""
import numpy as np
import pytest

from oracle import kaldi_oracle as O
from tests.refbind import load_ref
from tests.test_layer_vs_reference_gpu import RefOps
from tests.test_nnet_gpu import SPLICED, make_pair, rel_to_scale

pytestmark = pytest.mark.gpu

T = 120


def ref_forward(R, on, x):
    """every layer of `on` (an OracleNet: topology, weights, batch-norm statistics) on the reference library"""
    lib = R.lib
    ones = R.up(np.ones((T, 1), np.float32))
    acts = {}

    def W(name):
        return R.up(on.params[name])

    def bn(buf, rows, dim, stats):
        assert lib.ops_batchnorm_forward(buf.ptr, rows, dim, R.f32(stats["mean"]).ptr, R.f32(stats["var"]).ptr,
                                         R.f32(stats["gamma"]).ptr, R.f32(stats["beta"]).ptr, stats["eps"]) == 0

    def add_bias(buf, dim, bias):
        R.gemm(T, dim, 1, ones, bias, buf, beta=1.0)

    for l in on.layers:
        if l.type == "input":
            acts[l.name] = R.up(x)
            continue
        assert len(l.inputs) == 1, l.name
        src = acts[l.inputs[0]]
        din, dout = l.in_dim, l.out_dim
        if l.type == "linear-component":
            y = R.alloc(T, dout)
            R.gemm(T, dout, din, src, W(f"{l.name}.W"), y)
        elif l.type == "tdnnf-layer":
            s, b = int(l.kv.get("time-stride", 3)), int(l.kv["bottleneck-dim"])
            s1 = R.splice(src, T, din, -s)
            bt = R.alloc(T, b)
            R.gemm(T, b, 2 * din, s1, W(f"{l.name}.LinearW"), bt)
            s2 = R.splice(bt, T, b, +s)
            y = R.alloc(T, dout)
            R.gemm(T, dout, 2 * b, s2, W(f"{l.name}.AffineW"), y)
            add_bias(y, dout, W(f"{l.name}.AffineBias"))
            assert lib.ops_relu(y.ptr, T * dout) == 0
            bn(y, T, dout, on.bn[(l.name, "AffBN")])
            assert lib.ops_add_scaled(y.ptr, src.ptr, T * dout, float(l.kv.get("bypass-scale", 0.66)), 1.0) == 0
        elif l.type == "prefinal-layer":
            big, small = int(l.kv["big-dim"]), int(l.kv["small-dim"])
            g = R.alloc(T, big)
            R.gemm(T, big, din, src, W(f"{l.name}.BigW"), g)
            add_bias(g, big, W(f"{l.name}.BigBias"))
            assert lib.ops_relu(g.ptr, T * big) == 0
            bn(g, T, big, on.bn[(l.name, "PfBN")])
            y = R.alloc(T, small)
            R.gemm(T, small, big, g, W(f"{l.name}.SmallW"), y)
            bn(y, T, small, on.bn[(l.name, "BN")])
        elif l.type == "output-layer":
            y = R.alloc(T, dout)
            R.gemm(T, dout, din, src, W(f"{l.name}.W"), y)
            add_bias(y, dout, W(f"{l.name}.Bias"))
            if l.kv.get("include-log-softmax", "true").lower() in ("true", "1", "yes"):
                assert lib.ops_log_softmax(y.ptr, T, dout) == 0
        else:
            raise AssertionError(l.type)
        acts[l.name] = y
    R.sync()
    return {k: v.f32() for k, v in acts.items()}


def test_network_forward_against_the_reference_operator_library(handle, lib):
    ref = load_ref()
    if ref is None:
        pytest.skip("oracle/_ref/libkaldi_fp16_ref.so is not built")
    on, net, rng = make_pair(handle, SPLICED, 1, T, seed=21)
    x = O.to_f16_rne(rng.standard_normal((T, 64)).astype(np.float32))
    R = RefOps(ref)
    try:
        want = ref_forward(R, on, x)
    finally:
        R.close()
    oracle_acts = on.forward({"input": x})
    net.SetInput("input", x)
    assert lib.kfp16_net_forward(net.ptr) == 0
    for l in on.layers:
        if l.type == "input":
            continue
        e_oracle = rel_to_scale(oracle_acts[l.name], want[l.name])
        e_net = rel_to_scale(net.Output(l.name), want[l.name])
        assert e_oracle <= 2e-3, f"{l.name}: network oracle vs the reference's operators: {e_oracle:.2e}"
        assert e_net <= 2e-3, f"{l.name}: executor vs the reference's operators: {e_net:.2e}"
    net.Free()
