"""Kaldi nnet3 text import (kaldi_fp16_b200/weight_loader.py) against the reference's own parser test
(internal/nnet/weight_loader_test.go: the fixture `testComponents` is committed as tests/golden/nnet3_text_fixture.txt by
scripts/gen_nnet3_fixture.py; the assertions below restate TestParseNnet3Text / TestParseRealBatchNormLine /
TestComponentCount / TestParseInlineVector / TestBatchNormComputation)."""
from pathlib import Path

import numpy as np

from kaldi_fp16_b200 import weight_loader as WL

FIXTURE = (Path(__file__).parent / "golden" / "nnet3_text_fixture.txt").read_text()


def close(a, b, tol):
    assert abs(float(a) - float(b)) <= tol, (a, b)


def test_parse_nnet3_text_reference_fixture():
    comps = WL.ParseNnet3Text(FIXTURE)
    c = comps["idct"]
    assert c.Type == "FixedAffineComponent" and (c.LinearRows, c.LinearCols) == (2, 4) and c.LinearParams.size == 8
    close(c.LinearParams.reshape(-1)[0], 0.1581139, 1e-5)
    close(c.LinearParams.reshape(-1)[4], 0.1581139, 1e-5)
    assert c.BiasParams.size == 4
    c = comps["ivector-linear"]
    assert c.Type == "LinearComponent" and (c.LinearRows, c.LinearCols) == (2, 3)
    assert c.LearningRate == 0.0001 and c.L2Regularize == 0.03 and c.MaxChange == 0.75
    c = comps["ivector-batchnorm"]
    assert c.Type == "BatchNormComponent" and c.Epsilon == 0.001 and c.TargetRms == 0.025 and c.Count == 176000
    assert c.StatsMean.size == 4 and c.StatsVar.size == 4
    close(c.StatsMean[0], -0.005183299, 1e-6)
    close(c.StatsVar[0], 0.1, 1e-6)
    c = comps["cnn1.conv"]
    assert c.Type == "TimeHeightConvolutionComponent"
    assert (c.NumFiltersIn, c.NumFiltersOut, c.HeightIn, c.HeightOut) == (6, 48, 40, 40)
    assert (c.LinearRows, c.LinearCols) == (2, 3) and c.BiasParams.size == 3
    close(c.BiasParams[0], 0.05598261, 1e-6)
    assert c.Offsets == [(dt, dh) for dt in (-1, 0, 1) for dh in (-1, 0, 1)]       # time-major: the executor's tap order
    c = comps["tdnnf7.linear"]
    assert c.Type == "TdnnComponent" and (c.LinearRows, c.LinearCols) == (2, 2)
    assert c.BiasParams is None                                                     # "<BiasParams>  [ ]"
    close(c.LinearParams.reshape(-1)[0], 3.699428e-43, 1e-45)                       # near-zero SVD init: FP32 subnormals survive
    assert c.TimeOffsets == [0]
    c = comps["tdnnf7.affine"]
    assert (c.LinearRows, c.LinearCols) == (2, 3) and c.BiasParams.size == 3
    close(c.BiasParams[0], -1.943402e-05, 1e-8)
    c = comps["prefinal-chain.affine"]
    assert c.Type == "NaturalGradientAffineComponent" and (c.LinearRows, c.LinearCols) == (2, 2) and c.BiasParams.size == 2
    c = comps["output.affine"]
    assert (c.LinearRows, c.LinearCols) == (3, 3)
    close(c.LinearParams.reshape(-1)[8], 0.9, 1e-6)
    assert comps["noop1"].Type == "NoOpComponent" and comps["noop1"].LinearParams is None
    assert comps["output-xent.log-softmax"].Type == "LogSoftmaxComponent"
    for name in ("idct", "ivector-linear", "ivector-batchnorm", "cnn1.conv", "cnn1.relu", "cnn1.batchnorm", "tdnnf7.linear",
                 "tdnnf7.affine", "tdnnf7.batchnorm", "prefinal-chain.affine", "output.affine", "noop1", "output-xent.log-softmax"):
        assert name in comps, name


def test_parse_real_batchnorm_line_and_inline_vectors():
    line = ("<ComponentName> prefinal-chain.batchnorm2 <BatchNormComponent> <Dim> 192 <BlockDim> 192 <Epsilon> 0.001 <TargetRms> 1 "
            "<TestMode> F <Count> 41344 <StatsMean>  [ 4.844032e-10 -4.039575e-09 -7.640916e-11 ]\n<StatsVar>  [ 0.001 0.002 0.003 ]")
    c = WL.ParseNnet3Text(line)["prefinal-chain.batchnorm2"]
    assert c.Epsilon == 0.001 and c.TargetRms == 1.0 and c.Count == 41344 and c.StatsMean.size == 3
    close(c.StatsMean[0], 4.844032e-10, 1e-15)
    text = ("<ComponentName> test <BatchNormComponent> <Dim> 3 <Epsilon> 0.001 <TargetRms> 1 <Count> 100 <StatsMean>  [ 0.1 0.2 0.3 ]\n"
            "<StatsVar>  [ 0.4 0.5 0.6 ]")
    c = WL.ParseNnet3Text(text)["test"]
    assert np.allclose(c.StatsMean, [0.1, 0.2, 0.3]) and np.allclose(c.StatsVar, [0.4, 0.5, 0.6])


def test_write_parse_round_trip():
    comps = WL.ParseNnet3Text(FIXTURE)
    again = WL.ParseNnet3Text(WL.WriteNnet3Text(comps))
    assert set(again) == set(comps)
    for k, c in comps.items():
        d = again[k]
        assert d.Type == c.Type
        for attr in ("LinearParams", "BiasParams", "StatsMean", "StatsVar"):
            a, b = getattr(c, attr), getattr(d, attr)
            assert (a is None) == (b is None), (k, attr)
            if a is not None:
                assert a.shape == b.shape and np.array_equal(a, b), (k, attr)      # repr(float32) round-trips exactly
        assert (d.Epsilon, d.TargetRms, d.Count, d.Offsets, d.TimeOffsets) == (c.Epsilon, c.TargetRms, c.Count, c.Offsets, c.TimeOffsets)


class FakeNet:
    """the slice of nnet.Network LoadWeights uses"""

    def __init__(self, layers, params):
        self.layers, self.params = layers, {k: (r, c, 0) for k, (r, c) in params.items()}
        self.set, self.bn, self.idct = {}, {}, {}

    def SetParam(self, name, w):
        r, c, _ = self.params[name]
        assert w.shape == (r, c) and w.dtype == np.float32
        self.set[name] = w

    def SetBN(self, layer, which, mean, var, gamma, beta, eps):
        self.bn[(layer, which)] = (np.asarray(mean), np.asarray(var), np.asarray(gamma), np.asarray(beta), eps)

    def SetIDCT(self, layer, m):
        self.idct[layer] = m


def test_load_weights_maps_components_to_parameters():
    """LoadWeights (weight_loader.go:754-946): Kaldi [out x in] -> [in x out]; batch-norm as makeBN (gamma = target-rms, beta = 0)"""
    rng = np.random.default_rng(0)

    def comp(name, typ, out, inn, bias=True):
        return WL.KaldiComponent(Name=name, Type=typ, LinearParams=rng.standard_normal((out, inn)).astype(np.float32),
                                 BiasParams=rng.standard_normal(out).astype(np.float32) if bias else None)

    def bn(name, dim, rms=1.0):
        return WL.KaldiComponent(Name=name, Type="BatchNormComponent", StatsMean=rng.standard_normal(dim).astype(np.float32),
                                 StatsVar=(rng.random(dim) + 0.5).astype(np.float32), TargetRms=rms, Epsilon=0.001)

    comps = {c.Name: c for c in [
        comp("idct", "FixedAffineComponent", 8, 8), comp("lin", "LinearComponent", 16, 8, bias=False), bn("norm", 16, 0.025),
        comp("cnn1.conv", "TimeHeightConvolutionComponent", 64, 27), bn("cnn1.batchnorm", 64),
        comp("tdnnf2.linear", "TdnnComponent", 32, 2 * 128, bias=False), comp("tdnnf2.affine", "TdnnComponent", 128, 2 * 32),
        bn("tdnnf2.batchnorm", 128), comp("prefinal-chain.affine", "NaturalGradientAffineComponent", 256, 128),
        comp("prefinal-chain.linear", "LinearComponent", 64, 256, bias=False), bn("prefinal-chain.batchnorm1", 256),
        bn("prefinal-chain.batchnorm2", 64), comp("output.affine", "NaturalGradientAffineComponent", 48, 64)]}
    layers = [("input", "input", 8), ("idct", "idct-layer", 8), ("lin", "linear-component", 16), ("norm", "batchnorm-component", 16),
              ("cnn1", "conv-relu-batchnorm-layer", 1024), ("tdnnf2", "tdnnf-layer", 128), ("prefinal-chain", "prefinal-layer", 64),
              ("output", "output-layer", 48), ("aug", "spec-augment-layer", 8)]
    params = {"lin.W": (8, 16), "cnn1.W": (27, 64), "cnn1.Bias": (1, 64), "tdnnf2.LinearW": (256, 32), "tdnnf2.AffineW": (64, 128),
              "tdnnf2.AffineBias": (1, 128), "prefinal-chain.BigW": (128, 256), "prefinal-chain.BigBias": (1, 256),
              "prefinal-chain.SmallW": (256, 64), "output.W": (64, 48), "output.Bias": (1, 48)}
    net = FakeNet(layers, params)
    rep = WL.LoadWeights(net, comps)
    assert rep["loaded"] == 7 and rep["skipped"] == ["aug"]
    assert rep["params"] == sum(c.LinearParams.size + (0 if c.BiasParams is None else c.BiasParams.size) for c in comps.values()
                                if c.LinearParams is not None) - 8 + 4 * (16 + 64 + 128 + 256 + 64)     # (idct bias is not loaded)
    assert np.array_equal(net.set["tdnnf2.LinearW"], comps["tdnnf2.linear"].LinearParams.T)
    assert np.array_equal(net.set["cnn1.W"], comps["cnn1.conv"].LinearParams.T)
    assert np.array_equal(net.set["output.Bias"], comps["output.affine"].BiasParams.reshape(1, -1))
    assert np.array_equal(net.idct["idct"], comps["idct"].LinearParams.T)
    mean, var, gamma, beta, eps = net.bn[("norm", "")]
    assert np.array_equal(mean, comps["norm"].StatsMean) and np.all(gamma == np.float32(0.025)) and not beta.any() and eps == 0.001
    assert set(net.bn) == {("norm", ""), ("cnn1", "BN"), ("tdnnf2", "AffBN"), ("prefinal-chain", "PfBN"), ("prefinal-chain", "BN")}
    # the folded factor TestBatchNormComputation checks: gamma / sqrt(var + eps) with StatsVar taken as the variance
    m0, v0 = np.float32(-0.005183299), np.float32(0.1)
    close(np.float32(0.025) / np.sqrt(v0 + np.float32(0.001)), 0.025 / np.sqrt(0.1 + 0.001), 1e-6)
    # a missing component is an error (strict) or a skipped layer
    del comps["tdnnf2.batchnorm"]
    import pytest
    with pytest.raises(WL.WeightLoadError):
        WL.LoadWeights(FakeNet(layers, params), comps)
    assert "tdnnf2" in WL.LoadWeights(FakeNet(layers, params), comps, strict=False)["skipped"]
    # a shape that does not fit is reported with both shapes
    comps["output.affine"].LinearParams = comps["output.affine"].LinearParams[:, :10]
    with pytest.raises(WL.WeightLoadError, match="does not fit"):
        WL.LoadWeights(FakeNet(layers, params), comps, strict=False) if False else WL._matrix(FakeNet(layers, params), "output.W", comps["output.affine"])
