"""Network-level parity on the GPU: the native executor (kaldi_fp16_b200.nnet over the C ABI) against
the numpy oracle (oracle/nnet_oracle.py) on the reference's own acceptance nets
(cmd/sgdtest/main.go:199-203, cmd/traintest/main.go:37-42, cmd/backtest/main.go:219-227)
and on a spliced TDNN-F stack.  Tolerances: activations max-rel-err-vs-scale <= 2e-3,
gradients <= 5e-3 of the tensor's max-abs, SGD'd weights <= 1e-3."""
import numpy as np
import pytest

from kaldi_fp16_b200 import gpu, nnet
from oracle import kaldi_oracle as O
from oracle.nnet_oracle import OracleNet

pytestmark = pytest.mark.gpu

SGDTEST = """
input name=input dim=40
linear-component name=linear1 dim=128
batchnorm-component name=bn1
prefinal-layer name=prefinal small-dim=64 big-dim=128
output-layer name=output dim=40 include-log-softmax=false
"""
TRAINTEST = """
input name=input dim=40
linear-component name=linear1 dim=128
batchnorm-component name=bn1
tdnnf-layer name=tdnnf1 dim=128 bottleneck-dim=64 time-stride=0 bypass-scale=0.66
prefinal-layer name=prefinal small-dim=64 big-dim=128
output-layer name=output dim=40 include-log-softmax=false
"""
BACKTEST = """
input name=input dim=40
input name=ivector dim=32
idct-layer name=idct input=input dim=40
linear-component name=linear1 input=Append(idct, ivector) dim=128
batchnorm-component name=bn1
tdnnf-layer name=tdnnf1 dim=128 bottleneck-dim=64 time-stride=0 bypass-scale=0.66
tdnnf-layer name=tdnnf2 dim=128 bottleneck-dim=64 time-stride=0 bypass-scale=0.66
prefinal-layer name=prefinal input=tdnnf2 small-dim=64 big-dim=128
output-layer name=output dim=40 include-log-softmax=false
"""
SPLICED = """
input name=input dim=64
linear-component name=lin0 dim=256
tdnnf-layer name=tdnnf1 dim=256 bottleneck-dim=64 time-stride=3 bypass-scale=0.66
tdnnf-layer name=tdnnf2 dim=256 bottleneck-dim=64 time-stride=3 bypass-scale=0.66
tdnnf-layer name=tdnnf3 dim=256 bottleneck-dim=64 time-stride=1 bypass-scale=0.66
linear-component name=prefinal-l dim=64
prefinal-layer name=prefinal-chain input=prefinal-l big-dim=256 small-dim=64
output-layer name=output include-log-softmax=false dim=104
prefinal-layer name=prefinal-xent input=prefinal-l big-dim=256 small-dim=64
output-layer name=output-xent dim=104
"""
IVECTOR = """
input dim=16 name=ivector
input dim=40 name=input
idct-layer name=idct input=input dim=40 cepstral-lifter=22
linear-component name=ivector-linear dim=24 input=ReplaceIndex(ivector, t, 0)
batchnorm-component name=ivector-batchnorm target-rms=0.025
batchnorm-component name=idct-batchnorm input=idct
combine-feature-maps-layer name=combine_inputs input=Append(idct-batchnorm, ivector-batchnorm) num-filters1=5 num-filters2=3 height=8
linear-component name=lin1 dim=128
tdnnf-layer name=tdnnf1 dim=128 bottleneck-dim=32 time-stride=2 bypass-scale=0.66
output-layer name=output include-log-softmax=true dim=48
"""


def make_pair(handle, xconfig, n_seq, L, seed, randomize_bn=True, **kw):
    # ref_round: keep the reference's FP16 store between fused epilogue stages, so that the ReLU
    # masks (whose flips change a gradient element by its full value) agree with the op-by-op oracle
    kw.setdefault("ref_round", True)
    rng = np.random.default_rng(seed)
    on = OracleNet(xconfig, n_seq, L)
    on.init_random(rng)
    for k in on.params:
        if k.endswith("Bias"):
            on.params[k] = O.to_f16_trunc((rng.standard_normal(on.params[k].shape) * 0.1).astype(np.float32))
    net = nnet.NewNetwork(nnet.BuildModelFromString(xconfig), handle, n_seq, L, **kw)
    assert set(net.params) == set(on.params), (sorted(net.params), sorted(on.params))
    for k, w in on.params.items():
        net.SetParam(k, w)
    if randomize_bn:
        for (layer, which), bn in on.bn.items():
            d = bn["mean"].size
            bn["mean"] = (rng.standard_normal(d) * 0.1).astype(np.float32)
            bn["var"] = (rng.random(d) + 0.5).astype(np.float32)
            if not (which == "" and float(on.by_name[layer].kv.get("target-rms", 1.0)) != 1.0):
                bn["gamma"] = (rng.random(d) + 0.5).astype(np.float32)
                bn["beta"] = (rng.standard_normal(d) * 0.1).astype(np.float32)
            net.SetBN(layer, which, bn["mean"], bn["var"], bn["gamma"], bn["beta"], bn["eps"])
    return on, net, rng


def rel_to_scale(got, want):
    return O.max_err_vs_scale(got, want)


def check_forward_backward(on, net, inputs, per_seq=(), act_tol=2e-3, grad_tol=5e-3, check_layers=None):
    acts = on.forward(inputs)
    net.MarkPerSequence(*per_seq)
    for name, x in inputs.items():
        net.SetInput(name, x)
    assert net.lib.kfp16_net_forward(net.ptr) == 0
    for l in on.layers:
        if l.type == "input" or (check_layers and l.name not in check_layers):
            continue
        got = net.Output(l.name)
        err = rel_to_scale(got, acts[l.name])
        assert err <= act_tol, f"forward {l.name}: err {err:.2e} > {act_tol}"
    # backward with dOut = out
    out = acts["output"]
    # ReLU masks: must agree with the oracle's except where a pre-activation is within rounding of
    # zero (< 0.5 % of elements); the gradient comparison then uses the kernel's masks
    masks = {}
    for l in on.layers:
        if l.type in ("tdnnf-layer", "prefinal-layer") and l.name in on.saved and "mask" in on.saved[l.name]:
            want = on.saved[l.name]["mask"]
            got = net.Mask(l.name, want.shape[1])
            assert np.mean(got != want) < 5e-3, f"relu mask {l.name}: {np.mean(got != want):.2%} differ"
            masks[l.name] = got
    wg, dact = on.backward("output", out, masks)
    net.ZeroGrads()
    net.Backward(None)
    loss = net.ReadLoss()
    want_loss = 0.5 * float((out.astype(np.float64) ** 2).sum())
    assert abs(loss - want_loss) <= 2e-3 * max(want_loss, 1e-6), (loss, want_loss)
    got_wg = net.WeightGrads()
    for k, g in wg.items():
        err = rel_to_scale(got_wg[k], g)
        # bias gradients are short, cancellation-heavy column sums: the reference's own gate (1e-2,
        # cmd/sgdtest/main.go:71) is the bound there
        tol = 1e-2 if k.endswith("Bias") else grad_tol
        assert err <= tol, f"weight grad {k}: err {err:.2e} > {tol}"
    # parameters without a gradient path (xent branch) stay zero (network_backward.go:104-107)
    for k in got_wg:
        if k not in wg:
            assert not got_wg[k].any(), f"{k} should receive no gradient"
    return acts, wg, dact


# ref_round = True keeps the reference's FP16 store between fused stages (generic run-time-flag epilogue);
# ref_round = False is what bench.py runs: the compile-time specialised epilogues (EK_AFFINE, EK_AFFINE_RES, EK_BN,
# EK_BN_GRADMASK, EK_BIAS, EK_RESID, EK_PLAIN), which skip those intermediate roundings.  Both must meet the same bounds.
REF_ROUND = pytest.mark.parametrize("ref_round", [True, False], ids=["ref_round", "specialised"])
SEEDS = {"sgdtest": 11, "traintest": 23}


@REF_ROUND
@pytest.mark.parametrize("xconfig,name", [(SGDTEST, "sgdtest"), (TRAINTEST, "traintest")])
def test_acceptance_nets_forward_backward(handle, xconfig, name, ref_round):
    on, net, rng = make_pair(handle, xconfig, 1, 32, seed=SEEDS[name], ref_round=ref_round)
    x = O.to_f16_rne((rng.random((32, 40)) * 2 - 1).astype(np.float32))     # rand.Float32()*2-1
    check_forward_backward(on, net, {"input": x})
    net.Free()


@REF_ROUND
def test_backtest_net_with_append(handle, ref_round):
    """cmd/backtest/main.go:217-293: Append(idct, ivector) with a [T x 32] ivector; every layer gets gradients"""
    on, net, rng = make_pair(handle, BACKTEST, 1, 32, seed=5, ref_round=ref_round)
    x = O.to_f16_rne((rng.random((32, 40)) * 2 - 1).astype(np.float32))
    iv = O.to_f16_rne((rng.random((32, 32)) * 2 - 1).astype(np.float32))
    acts, wg, dact = check_forward_backward(on, net, {"input": x, "ivector": iv})
    for k in ("linear1.W", "tdnnf1.LinearW", "tdnnf2.AffineW", "prefinal.BigW", "output.W"):
        assert np.abs(net.WeightGrads()[k]).max() > 0
    net.Free()


@REF_ROUND
@pytest.mark.parametrize("n_seq,L", [(1, 96), (4, 50), (3, 41)])
def test_spliced_tdnnf_stack(handle, lib, n_seq, L, ref_round):
    """time-stride > 0: splice as TMA row offsets over the padded layout, per-sequence clamp
    (n_seq=1 is the reference's whole-minibatch clamp, forward.go:714-722,760-770)"""
    kinds0 = [lib.kfp16_gemm_kind_launches(k) for k in range(9)]
    on, net, rng = make_pair(handle, SPLICED, n_seq, L, seed=n_seq * 100 + L, ref_round=ref_round)
    x = O.to_f16_rne(rng.standard_normal((n_seq * L, 64)).astype(np.float32))
    acts, wg, dact = check_forward_backward(on, net, {"input": x})
    # activation gradients too
    for lname in ("tdnnf2", "tdnnf1", "lin0"):
        err = rel_to_scale(net.Grad(lname), dact[lname])
        assert err <= 5e-3, f"dact {lname}: {err:.2e}"
    # which epilogue bodies ran: kind 0 = generic run-time flags, 2 affine, 3 affine+bypass, 4 residual, 5 bn+gradmask,
    # 6 bn, 7 bias (include/kaldi_fp16_fused.h kfp16_gemm_kind_launches)
    ran = [lib.kfp16_gemm_kind_launches(k) - kinds0[k] for k in range(9)]
    if ref_round:
        assert ran[0] > 0 and ran[2] == 0 and ran[3] == 0
    else:
        assert ran[0] == 0, f"specialised run used the generic epilogue: {ran}"
        assert all(ran[k] > 0 for k in (1, 2, 3, 4, 5, 6, 7, 8)), f"expected every specialised kind to run: {ran}"
    net.Free()


@REF_ROUND
def test_ivector_branch_and_combine(handle, ref_round):
    """ReplaceIndex(ivector,t,0) -> per-sequence branch broadcast into Append + combine-feature-maps"""
    n_seq, L = 3, 20
    on, net, rng = make_pair(handle, IVECTOR, n_seq, L, seed=77, ref_round=ref_round)
    x = O.to_f16_rne((rng.standard_normal((n_seq * L, 40)) * 3).astype(np.float32))
    iv = O.to_f16_rne(rng.standard_normal((n_seq, 16)).astype(np.float32))
    check_forward_backward(on, net, {"input": x, "ivector": iv}, per_seq=("ivector", "ivector-linear", "ivector-batchnorm"))
    net.Free()


def test_ref_round_mode_is_bit_closer(handle):
    """EPI_REF_ROUND keeps the reference's FP16 store between fused stages: with it the fused TDNN-F
    output equals the op-by-op oracle to <= 1 fp16 ulp almost everywhere"""
    on, net, rng = make_pair(handle, TRAINTEST, 1, 64, seed=9, ref_round=True)
    x = O.to_f16_rne((rng.random((64, 40)) * 2 - 1).astype(np.float32))
    acts = on.forward({"input": x})
    net.SetInput("input", x)
    assert net.lib.kfp16_net_forward(net.ptr) == 0
    got, want = net.Output("tdnnf1"), acts["tdnnf1"]
    ulp = np.abs(got - want) / np.maximum(np.abs(want) * 2.0 ** -10, 2.0 ** -24)
    assert np.mean(ulp <= 1.01) > 0.995                       # bypass add can cancel: bound the rest by scale
    assert O.max_err_vs_scale(got, want) <= 1e-3
    net.Free()


def test_sgd_step_matches_oracle_and_loss_decreases(handle):
    """cmd/sgdtest/main.go:196-321 + cmd/traintest/main.go:34-162: lr 1e-3, momentum 0.9, loss
    0.5*||out||^2, dY = Y; first-step weights match the oracle; loss[last] < loss[0]"""
    on, net, rng = make_pair(handle, TRAINTEST, 1, 32, seed=12, randomize_bn=False, lr=1e-3, momentum=0.9)
    x = O.to_f16_rne((rng.random((32, 40)) * 2 - 1).astype(np.float32))
    trainer = nnet.Trainer(net)
    state = {}
    losses, olosses = [], []
    for step in range(20):
        acts = on.forward({"input": x})
        out = acts["output"]
        olosses.append(0.5 * float((out.astype(np.float64) ** 2).sum()))
        wg, _ = on.backward("output", out)
        on.sgd(state, wg, 1e-3, 0.9)
        losses.append(trainer.Step(x))
        if step == 0:
            for k in on.params:
                err = rel_to_scale(net.GetParam(k), on.params[k])
                assert err <= 1e-3, f"after 1 step {k}: {err:.2e}"
    assert losses[-1] < losses[0], losses
    assert abs(losses[0] - olosses[0]) <= 2e-3 * olosses[0]
    assert abs(losses[-1] - olosses[-1]) <= 0.05 * olosses[0]
    # Trainer.SetLR + weights non-zero (traintest/main.go:145-160)
    trainer.SetLR(5e-4)
    assert all(np.abs(w).max() > 0 for k, w in net.MasterWeights().items() if not k.endswith("Bias"))
    net.Free()


def test_graph_replay_equals_eager(handle, lib):
    """the captured CUDA graph of zero_grads+forward+loss+backward+SGD reproduces the eager step"""
    from kaldi_fp16_b200 import cudart
    st = cudart.Stream()
    lib.kfp16_ctx_set_stream(handle.ptr, st.ptr)
    try:
        on, net, rng = make_pair(handle, SPLICED, 2, 40, seed=4, lr=1e-3, momentum=0.9)
        x = O.to_f16_rne(rng.standard_normal((80, 64)).astype(np.float32))
        net.SetInput("input", x)
        w0 = {k: net.GetParam(k) for k in net.params}
        trainer = nnet.Trainer(net)
        l_eager = trainer.Step(x)
        w_eager = {k: net.GetParam(k) for k in net.params}
        for k, w in w0.items():
            net.SetParam(k, w)
        m0 = net.MasterWeights()
        net.Capture(3)           # its eager warm-up pass is undone: capturing takes no optimiser step
        m1 = net.MasterWeights()
        for k in m0:
            assert np.array_equal(m0[k], m1[k]), f"capture changed master weights of {k}"
        assert not np.any(net._bucket_f32(lib.kfp16_net_velocity)), "capture left a velocity behind"
        net.ReadLoss()
        net.Launch(3)
        st.synchronize()
        l_graph = net.ReadLoss()
        assert abs(l_graph - l_eager) <= 1e-4 * abs(l_eager)
        for k in w0:
            assert rel_to_scale(net.GetParam(k), w_eager[k]) <= 1e-3, k
        assert lib.kfp16_net_launches_per_step(net.ptr, 3) > 10
        net.Free()
    finally:
        lib.kfp16_ctx_set_stream(handle.ptr, None)
        st.destroy()


def test_xconfig_errors(handle, lib):
    """unsupported / malformed models fail loudly with a message (no silent fallback)"""
    for bad, frag in [("input name=input dim=40\nattention-relu-batchnorm-layer name=a num-heads=2", b"value-dim, key-dim must be positive"),
                      ("input name=input dim=40\nlstm-layer name=l cell-dim=64", b"unsupported layer type"),
                      ("input name=input dim=40\nlinear-component name=l", b"missing dim"),
                      ("input name=input dim=40\nlinear-component name=l dim=64 input=nope", b"not found"),
                      ("input name=input dim=40\ntdnnf-layer name=t dim=100 bottleneck-dim=20 time-stride=3", b"multiple")]:
        with pytest.raises(nnet.NNetError) as e:
            nnet.NewNetwork(nnet.BuildModelFromString(bad), handle, 1, 8)
        assert frag.decode() in str(e.value), str(e.value)


# ------------------------------------------------------------------ CNN front-end (conv-relu-batchnorm-layer)
CNN_SMALL = """
input dim=24 name=ivector
input dim=16 name=input
idct-layer name=idct input=input dim=16 cepstral-lifter=22
linear-component name=ivector-linear dim=32 input=ReplaceIndex(ivector, t, 0)
batchnorm-component name=ivector-batchnorm target-rms=0.025
batchnorm-component name=idct-batchnorm input=idct
combine-feature-maps-layer name=combine_inputs input=Append(idct-batchnorm, ivector-batchnorm) num-filters1=1 num-filters2=2 height=16
conv-relu-batchnorm-layer name=cnn1 height-in=16 height-out=16 time-offsets=-1,0,1 height-offsets=-1,0,1 num-filters-out=F1
conv-relu-batchnorm-layer name=cnn2 height-in=16 height-out=8 height-subsample-out=2 time-offsets=-1,0,1 height-offsets=-1,0,1 num-filters-out=64
conv-relu-batchnorm-layer name=cnn3 height-in=8 height-out=8 time-offsets=-1,0,1 height-offsets=-1,0,1 num-filters-out=64
tdnnf-layer name=tdnnf4 dim=256 bottleneck-dim=64 time-stride=0
tdnnf-layer name=tdnnf5 dim=256 bottleneck-dim=64 time-stride=3
output-layer name=output include-log-softmax=false dim=72
"""
# F1 = 64: cnn2 (height subsampling 2) and cnn3 run as implicit GEMMs (64-channel inputs), cnn1 (3 input filters) through a
# patch matrix; F1 = 32: cnn2 has a 32-channel input and takes the patch-matrix path too


def check_conv_net(handle, n_seq, L, seed, ref_round=True, f1=64, fuse=True):
    on, net, rng = make_pair(handle, CNN_SMALL.replace("F1", str(f1)), n_seq, L, seed=seed, ref_round=ref_round)
    net.SetFuseConvBackward(fuse)
    x = O.to_f16_rne((rng.standard_normal((n_seq * L, 16)) * 2).astype(np.float32))
    iv = O.to_f16_rne(np.clip(rng.standard_normal((n_seq, 24)), -3, 3).astype(np.float32))
    inputs = {"input": x, "ivector": iv}
    acts = on.forward(inputs)
    net.MarkPerSequence("ivector", "ivector-linear", "ivector-batchnorm")
    net.SetInput("input", x)
    net.SetInput("ivector", iv)
    assert net.lib.kfp16_net_forward(net.ptr) == 0
    for l in on.layers:
        if l.type == "input":
            continue
        err = rel_to_scale(net.Output(l.name), acts[l.name])
        assert err <= 2e-3, f"forward {l.name}: err {err:.2e}"
    masks = {}
    for l in on.layers:
        if l.type in ("tdnnf-layer", "conv-relu-batchnorm-layer"):
            want = on.saved[l.name]["mask"]
            got = net.Mask(l.name, l.out_dim).reshape(want.shape)
            assert np.mean(got != want) < 5e-3, f"relu mask {l.name}: {np.mean(got != want):.2%} differ"
            masks[l.name] = got
    wg, dact = on.backward("output", acts["output"], masks)
    net.ZeroGrads()
    net.Backward(None)
    got_wg = net.WeightGrads()
    for k, g in wg.items():
        err = rel_to_scale(got_wg[k], g)
        tol = 1e-2 if k.endswith("Bias") else 5e-3
        assert err <= tol, f"weight grad {k}: err {err:.2e} > {tol}"
    from oracle.nnet_oracle import h
    for name in ("cnn3", "cnn1", "combine_inputs"):
        want = dact[name]
        if fuse and name.startswith("cnn") and (name != "cnn1" or f1 == 64):     # (f1 = 32: cnn2 takes the patch-matrix path, which does not fuse)
            # the consumer's input-gradient epilogue already applied this layer's batch-norm scale and ReLU mask: the
            # buffer holds dZ = mask ? h(dY * scale) : 0 (what backwardConvReluBN computes first)
            fout = on.saved[name]["mask"].shape[1]
            sc = on._bn_scale(on.bn[(name, "BN")])
            want = np.where(masks[name].reshape(-1, fout), h(want.reshape(-1, fout) * sc), np.float32(0)).reshape(want.shape)
        err = rel_to_scale(net.Grad(name), want)
        assert err <= 5e-3, f"activation grad {name}: err {err:.2e}"
    net.Free()


@REF_ROUND
@pytest.mark.parametrize("n_seq,L", [(2, 30), (3, 17), (5, 64)])
def test_cnn_front_end_forward_backward(handle, n_seq, L, ref_round):
    """conv-relu-batchnorm layers (3x3 Cartesian taps, height subsampling, 3-filter input with K = 27) lowered to
    tcgen05 GEMMs (implicit GEMM over 4-D TMA boxes; the 3-filter first layer through a patch matrix), against the
    numpy oracle's explicit patch matrices"""
    check_conv_net(handle, n_seq, L, seed=5 + n_seq, ref_round=ref_round)


@REF_ROUND
def test_cnn_front_end_unfused_backward(handle, ref_round):
    """the same with one elementwise batch-norm / ReLU backward pass per conv layer (kfp16_net_set_fuse_conv_backward 0)"""
    check_conv_net(handle, 3, 17, seed=8, ref_round=ref_round, fuse=False)


def test_cnn_front_end_patch_matrix_path(handle):
    """a 32-channel conv input cannot be addressed by 64-wide TMA boxes: im2col + GEMM + col2im, same results"""
    check_conv_net(handle, 2, 30, seed=3, ref_round=False, f1=32)


def test_set_lr_reaches_a_captured_sgd_graph(handle, lib):
    """SGDOptimizer.SetLR (optimize.go:123) on the graph path: lr / momentum live in a device block the captured update
    kernel reads at run time, so kfp16_net_set_lr changes the step size of an already captured graph"""
    from kaldi_fp16_b200 import cudart
    st = cudart.Stream()
    lib.kfp16_ctx_set_stream(handle.ptr, st.ptr)
    try:
        on, net, rng = make_pair(handle, TRAINTEST, 1, 32, seed=3, randomize_bn=False, lr=1e-3, momentum=0.0)
        x = O.to_f16_rne((rng.random((32, 40)) * 2 - 1).astype(np.float32))
        net.SetInput("input", x)
        net.Capture(1)
        net.Capture(2)
        w0 = net.MasterWeights()

        def delta():
            before = net.MasterWeights()
            net.Launch(1)
            net.Launch(2)
            st.synchronize()
            after = net.MasterWeights()
            return {k: after[k] - before[k] for k in before}

        d1 = delta()
        for k, w in w0.items():          # same weights again -> same gradient
            net.SetParam(k, w)
        net.SetLR(2.5e-4)
        assert abs(lib.kfp16_net_get_lr(net.ptr) - 2.5e-4) < 1e-9
        d2 = delta()
        for k in d1:
            if k.endswith("Bias") or not np.abs(d1[k]).max() > 0:
                continue
            ratio = np.abs(d2[k]).sum() / np.abs(d1[k]).sum()
            assert abs(ratio - 0.25) < 2e-3, f"{k}: update ratio {ratio} after lr 1e-3 -> 2.5e-4"
        net.Free()
    finally:
        lib.kfp16_ctx_set_stream(handle.ptr, None)
        st.destroy()


def test_fp32_input_is_converted_on_the_device(handle, lib):
    """kfp16_net_set_input_f32 / prefetch_input_f32: FP32 rows -> RNE FP16 on the device == the host-side conversion
    (internal/gpu/bridge.go:141, fp16.ConvertFloat32ToFloat16), including overflow to Inf and subnormals"""
    import ctypes as C
    n_seq, L = 3, 25
    on, net, rng = make_pair(handle, SPLICED, n_seq, L, seed=21)
    x = (rng.standard_normal((n_seq * L, 64)) * 3).astype(np.float32)
    x[0, :6] = [65504.0, 65519.9, 65520.0, 1e-7, 6.1e-5, -2.0 ** -25]
    net.SetInput("input", x)                  # host RNE conversion
    assert lib.kfp16_net_forward(net.ptr) == 0
    want_in, want = net.Output("input"), net.Output("lin0")
    net.SetInputF32("input", x)               # device conversion
    assert lib.kfp16_net_forward(net.ptr) == 0
    assert np.array_equal(net.Output("input"), want_in, equal_nan=True)
    assert np.array_equal(net.Output("lin0"), want, equal_nan=True)
    pinned = lib.bridge_host_alloc(x.nbytes)
    C.memmove(pinned, x.ctypes.data, x.nbytes)
    net.SetInput("input", np.zeros_like(x))
    net.PrefetchInputF32("input", pinned, n_seq * L, 64)
    net.CommitInput("input")
    assert lib.kfp16_net_forward(net.ptr) == 0
    assert np.array_equal(net.Output("input"), want_in, equal_nan=True)
    lib.bridge_host_free(pinned)
    net.Free()


def test_fp16_gradient_bucket_and_update(handle, lib):
    """kfp16_net_grads_to_f16 + kfp16_net_sgd_step_f16 (the data-parallel path: FP16 gradient bucket, as the reference's
    FP16 gradient tensors) == kfp16_net_sgd_step with round_grad = 1 on the FP32 bucket, bit for bit"""
    on, net, rng = make_pair(handle, SPLICED, 2, 40, seed=8, lr=1e-3, momentum=0.9, grad_scale=1.0 / 64)
    x = O.to_f16_rne(rng.standard_normal((80, 64)).astype(np.float32))
    w0 = {k: net.GetParam(k) for k in net.params}
    net.SetInput("input", x)
    net.ZeroGrads()
    assert lib.kfp16_net_forward(net.ptr) == 0
    net.Backward(None)
    net.SGDStep(1.0 / 64, True)
    a = net.MasterWeights()
    for k, w in w0.items():
        net.SetParam(k, w)
    net.GradsToF16()
    g16 = np.empty(lib.kfp16_net_bucket_size(net.ptr), np.uint16)
    gpu.Sync()
    assert lib.bridge_read_fp16(g16.ctypes.data, lib.kfp16_net_grads_f16(net.ptr), g16.size) == 0
    g32 = net._bucket_f32(lib.kfp16_net_grads_f32)
    assert np.array_equal(g16.view(np.float16), (g32 * np.float32(1.0 / 64)).astype(np.float16))
    net.SGDStepF16()
    b = net.MasterWeights()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    net.Free()


def test_prefetched_input_equals_synchronous_input(handle, lib):
    """kfp16_net_prefetch_input / commit_input (async H2D from pinned memory, double buffered) == kfp16_net_set_input"""
    import ctypes as C

    n_seq, L = 3, 25
    on, net, rng = make_pair(handle, SPLICED, n_seq, L, seed=9)
    outs = []
    pinned = lib.bridge_host_alloc(n_seq * L * 64 * 2)
    xs = [O.to_f16_rne(rng.standard_normal((n_seq * L, 64)).astype(np.float32)) for _ in range(3)]
    for x in xs:                      # synchronous path
        outs.append(net.Forward(x))
    bits0 = nnet.rne_fp16_bits(xs[0])
    C.memmove(pinned, bits0.ctypes.data, bits0.nbytes)
    net.PrefetchInput("input", pinned, n_seq * L, 64)
    for i, x in enumerate(xs):        # prefetch i+1 while i runs
        net.CommitInput("input")
        gpu.Sync()                    # (the test reuses ONE pinned buffer, so wait before overwriting it)
        if i + 1 < len(xs):
            nb = nnet.rne_fp16_bits(xs[i + 1])
            C.memmove(pinned, nb.ctypes.data, nb.nbytes)
            net.PrefetchInput("input", pinned, n_seq * L, 64)
        assert lib.kfp16_net_forward(net.ptr) == 0
        assert np.array_equal(net.Output(""), outs[i])
    lib.bridge_host_free(pinned)
    net.Free()


def test_pipelined_loss_read_equals_synchronous_read(handle, lib):
    """kfp16_net_read_loss_async / kfp16_net_wait_loss (download queued behind the step, collected one step later)
    return the same per-step objective as the synchronous kfp16_net_read_loss"""
    n_seq, L = 3, 25
    on, net, rng = make_pair(handle, SPLICED, n_seq, L, seed=11)
    xs = [O.to_f16_rne(rng.standard_normal((n_seq * L, 64)).astype(np.float32)) for _ in range(4)]
    sync_losses = []
    net.ReadLoss()
    for x in xs:
        net.Forward(x)
        net.Backward(None)
        sync_losses.append(net.ReadLoss())
    got = []
    for i, x in enumerate(xs):
        net.Forward(x)
        net.Backward(None)
        net.ReadLossAsync(i & 1)
        if i > 0:
            got.append(net.WaitLoss((i - 1) & 1))
    got.append(net.WaitLoss((len(xs) - 1) & 1))
    # the objective is an atomic fp32 sum: equal up to summation order
    assert all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(got, sync_losses)) and all(v > 0 for v in got)
    assert lib.kfp16_net_read_loss_async(net.ptr, 2) == -1
    net.Free()


@pytest.mark.parametrize("nseg", [1, 3, 8])
def test_segmented_step_graphs_equal_the_single_graph(handle, lib, nseg):
    """kfp16_net_capture_segments cuts the step along the backward pass for a bucketed gradient all-reduce: the
    segments' gradient ranges tile the bucket in reverse layer order, and launching them in turn gives the same
    objective and gradients as the eager step"""
    from kaldi_fp16_b200 import cudart
    st = cudart.Stream()
    lib.kfp16_ctx_set_stream(handle.ptr, st.ptr)
    try:
        on, net, rng = make_pair(handle, SPLICED, 2, 40, seed=6)
        x = O.to_f16_rne(rng.standard_normal((80, 64)).astype(np.float32))
        net.SetInput("input", x)
        net.ZeroGrads()
        assert lib.kfp16_net_forward(net.ptr) == 0
        net.Backward(None)
        st.synchronize()
        l_eager, g_eager = net.ReadLoss(), net.WeightGrads()
        k = net.CaptureSegments(nseg)
        assert 1 <= k <= nseg
        ranges = [net.SegmentGrads(i) for i in range(k)]
        bucket = lib.kfp16_net_bucket_size(net.ptr)
        assert ranges[0][0] + ranges[0][1] == bucket and ranges[-1][0] == 0
        for (a0, c0), (a1, c1) in zip(ranges, ranges[1:]):
            assert a1 + c1 == a0 and c1 > 0          # contiguous, walking down the bucket
        net.ReadLoss()
        for i in range(k):
            net.LaunchSegment(i)
        st.synchronize()
        assert abs(net.ReadLoss() - l_eager) <= 1e-5 * abs(l_eager)
        g = net.WeightGrads()
        for name in g_eager:
            assert rel_to_scale(g[name], g_eager[name]) <= 1e-5, name     # fp32 reduction order only
        assert lib.kfp16_net_launch_segment(net.ptr, k) == -1
        net.Free()
    finally:
        lib.kfp16_ctx_set_stream(handle.ptr, None)
        st.destroy()


DROPNET = """
input name=input dim=64
linear-component name=lin0 dim=256
tdnnf-layer name=tdnnf1 dim=256 bottleneck-dim=64 time-stride=3 bypass-scale=0.66 dropout-proportion=0.2
tdnnf-layer name=tdnnf2 dim=256 bottleneck-dim=64 time-stride=0 bypass-scale=0.66 dropout-proportion=0.35
tdnnf-layer name=tdnnf3 dim=384 bottleneck-dim=64 time-stride=1 dropout-proportion=0.1
output-layer name=output include-log-softmax=false dim=104
"""


@REF_ROUND
def test_dropout_in_the_training_step(handle, lib, ref_round):
    """tdnnf-layer dropout-proportion (training networks): the fused dropout epilogue + its backward (mask bit = ReLU active
    AND kept, 1/(1-p) folded into the batch-norm factor) against the oracle's op-by-op dropout with the same mask; a
    phase-1 step draws a new mask, an inference network ignores the option"""
    n_seq, L, seed = 3, 41, 0xC0FFEE
    rng = np.random.default_rng(17)
    on = OracleNet(DROPNET, n_seq, L, train=True, dropout_seed=seed)
    on.init_random(rng)
    for k in on.params:
        if k.endswith("Bias"):
            on.params[k] = O.to_f16_trunc((rng.standard_normal(on.params[k].shape) * 0.1).astype(np.float32))
    net = nnet.NewNetwork(nnet.BuildModelFromString(DROPNET), handle, n_seq, L, ref_round=ref_round)
    for k, w in on.params.items():
        net.SetParam(k, w)
    assert lib.kfp16_net_set_dropout_seed(net.ptr, seed) == 0
    x = O.to_f16_rne(rng.standard_normal((n_seq * L, 64)).astype(np.float32))
    acts = on.forward({"input": x})
    net.SetInput("input", x)
    assert lib.kfp16_net_forward(net.ptr) == 0
    for name in ("tdnnf1", "tdnnf2", "tdnnf3", "output"):
        err = rel_to_scale(net.Output(name), acts[name])
        assert err <= 2e-3, f"forward {name} with dropout: {err:.2e}"
    dropped = np.mean(net.Output("tdnnf3") == 0)
    assert 0.05 < dropped < 0.6          # tdnnf3 has no bypass: dropped elements are exact zeros
    masks = {}
    for name in ("tdnnf1", "tdnnf2", "tdnnf3"):
        want = on.saved[name]["mask"]
        got = net.Mask(name, want.shape[1])
        assert np.mean(got != want) < 5e-3, f"gradient gate {name}: {np.mean(got != want):.2%} differ"
        masks[name] = got
    wg, dact = on.backward("output", acts["output"], masks)
    net.ZeroGrads()
    net.Backward(None)
    got_wg = net.WeightGrads()
    for k, g in wg.items():
        err = rel_to_scale(got_wg[k], g)
        assert err <= (1e-2 if k.endswith("Bias") else 5e-3), f"weight grad {k} with dropout: {err:.2e}"
    # a phase-1 step (what the captured graph replays) bumps the device seed word -> a different mask
    import ctypes as C
    s0, s1 = C.c_uint32(0), C.c_uint32(0)
    assert lib.kfp16_net_get_dropout_seed(net.ptr, C.byref(s0)) == 0 and s0.value == seed
    y0 = net.Output("tdnnf3")
    from kaldi_fp16_b200 import cudart
    st = cudart.Stream()
    lib.kfp16_ctx_set_stream(handle.ptr, st.ptr)
    try:
        net.Capture(1)
        assert lib.kfp16_net_get_dropout_seed(net.ptr, C.byref(s1)) == 0
        base = s1.value
        net.Launch(1)
        st.synchronize()
        assert lib.kfp16_net_get_dropout_seed(net.ptr, C.byref(s1)) == 0 and s1.value == base + 1
        y1 = net.Output("tdnnf3")
        assert np.mean((y0 == 0) != (y1 == 0)) > 0.05
    finally:
        lib.kfp16_ctx_set_stream(handle.ptr, None)
        st.destroy()
    net.Free()
    # inference network: dropout off
    inf = nnet.NewNetwork(nnet.BuildModelFromString(DROPNET), handle, n_seq, L, train=False)
    for k, w in on.params.items():
        inf.SetParam(k, w)
    on_inf = OracleNet(DROPNET, n_seq, L, train=False)
    on_inf.params = on.params
    want = on_inf.forward({"input": x})["output"]
    assert rel_to_scale(inf.Forward(x), want) <= 2e-3
    inf.Free()


SPECNET = """
input name=input dim=40
spec-augment-layer name=spec-augment freq-max-proportion=0.5 time-zeroed-proportion=0.2 time-mask-max-frames=10
linear-component name=lin0 dim=128
tdnnf-layer name=tdnnf1 dim=128 bottleneck-dim=32 time-stride=3 bypass-scale=0.66
output-layer name=output include-log-softmax=false dim=48
"""


def test_spec_augment_layer_masks(handle, lib):
    """spec-augment-layer in training (kfp16_net_set_spec_augment): per-sequence frequency / time masks
    (go/gotorch/cnn_tdnn.go:612-668) from the counter-based generator, bit-identical to the oracle's masks; the backward pass
    masks the gradient the same way; switched off (the default, = the reference executor's pass-through) features pass unchanged"""
    n_seq, L, seed = 6, 50, 0xBEEF
    rng = np.random.default_rng(23)
    on = OracleNet(SPECNET, n_seq, L, train=True, dropout_seed=seed, spec_augment=True)
    on.init_random(rng)
    net = nnet.NewNetwork(nnet.BuildModelFromString(SPECNET), handle, n_seq, L, ref_round=True)
    for k, w in on.params.items():
        net.SetParam(k, w)
    x = O.to_f16_rne((rng.standard_normal((n_seq * L, 40)) + 3.0).astype(np.float32))      # no exact zeros in the input
    net.SetInput("input", x)
    assert lib.kfp16_net_forward(net.ptr) == 0
    assert np.array_equal(net.Output("spec-augment"), x)                    # default: pass-through
    net.SetSpecAugment(True)
    assert lib.kfp16_net_set_dropout_seed(net.ptr, seed) == 0
    acts = on.forward({"input": x})
    assert lib.kfp16_net_forward(net.ptr) == 0
    got = net.Output("spec-augment")
    assert np.array_equal(got, acts["spec-augment"]), "masked features differ from the oracle's"
    keep = on.saved["spec-augment"]["keep"]
    zeroed = 1.0 - keep.mean()
    assert 0.05 < zeroed < 0.8, zeroed
    per_seq = keep.reshape(n_seq, L, 40)
    assert len({per_seq[s].tobytes() for s in range(n_seq)}) > 1           # masks differ between sequences
    for name in ("lin0", "tdnnf1", "output"):
        assert rel_to_scale(net.Output(name), acts[name]) <= 2e-3, name
    masks = {"tdnnf1": net.Mask("tdnnf1", 128)}
    wg, dact = on.backward("output", acts["output"], masks)
    net.ZeroGrads()
    net.Backward(None)
    got_wg = net.WeightGrads()
    for k, g in wg.items():
        err = rel_to_scale(got_wg[k], g)
        assert err <= (1e-2 if k.endswith("Bias") else 5e-3), f"weight grad {k} under SpecAugment: {err:.2e}"
    net.Free()
    inf = nnet.NewNetwork(nnet.BuildModelFromString(SPECNET), handle, n_seq, L, train=False)
    with pytest.raises(nnet.NNetError):
        inf.SetSpecAugment(True)
    inf.Free()


TRAINBN_NET = CNN_SMALL.replace("F1", "64").replace(
    "output-layer name=output include-log-softmax=false dim=72",
    "prefinal-layer name=prefinal small-dim=64 big-dim=256\noutput-layer name=output include-log-softmax=false dim=72")


@REF_ROUND
def test_train_mode_batchnorm(handle, lib, ref_round):
    """kfp16_net_set_train_batchnorm: every batch-norm of the training network (batchnorm-component, conv per filter, TDNN-F,
    both prefinal ones) normalises with the minibatch's statistics (cpp/cuda/cnn_kernels.cu:236-320 training branch),
    updates the running statistics with the momentum, and the backward pass scales by gamma / sqrt(batch var + eps)
    (go/gotorch/layers.go:302-330) -- against the oracle's train-mode batch-norm; switching it off restores the running
    statistics path"""
    n_seq, L, mom = 4, 33, 0.25
    on, net, rng = make_pair(handle, TRAINBN_NET, n_seq, L, seed=31, ref_round=ref_round)
    on.train, on.train_bn, on.bn_momentum = True, True, np.float32(mom)
    x = O.to_f16_rne((rng.standard_normal((n_seq * L, 16)) * 2).astype(np.float32))
    iv = O.to_f16_rne(np.clip(rng.standard_normal((n_seq, 24)), -3, 3).astype(np.float32))
    net.MarkPerSequence("ivector", "ivector-linear", "ivector-batchnorm")
    net.SetInput("input", x)
    net.SetInput("ivector", iv)
    net.SetTrainBatchNorm(True, mom)
    acts = on.forward({"input": x, "ivector": iv})
    assert lib.kfp16_net_forward(net.ptr) == 0
    for l in on.layers:
        if l.type == "input":
            continue
        err = rel_to_scale(net.Output(l.name), acts[l.name])
        # (the batch statistics are fp32 sums accumulated with atomics -- their order, and so the last bits of mean / variance,
        #  change from run to run -- and every layer normalises with its own: 5e-3 at the end of seven normalised layers)
        assert err <= 5e-3, f"train-BN forward {l.name}: err {err:.2e}"
    # running statistics moved towards the batch statistics
    for (layer, which), bn in on.bn.items():
        if layer == "ivector-batchnorm":
            continue                                   # per-sequence rows: running statistics in both implementations
        m, v = net.GetBN(layer, which, bn["mean"].size)
        assert np.allclose(m, bn["mean"], rtol=2e-3, atol=2e-3), (layer, which)
        assert np.allclose(v, bn["var"], rtol=5e-3, atol=2e-3), (layer, which)
    masks = {}
    for l in on.layers:
        if l.type in ("tdnnf-layer", "conv-relu-batchnorm-layer", "prefinal-layer"):
            want = on.saved[l.name]["mask"]
            got = net.Mask(l.name, l.out_dim if l.type == "conv-relu-batchnorm-layer" else want.shape[-1]).reshape(want.shape)
            assert np.mean(got != want) < 5e-3, f"relu mask {l.name}"
            masks[l.name] = got
    wg, _ = on.backward("output", acts["output"], masks)
    net.ZeroGrads()
    net.Backward(None)
    got_wg = net.WeightGrads()
    for k, g in wg.items():
        err = rel_to_scale(got_wg[k], g)
        assert err <= (1.5e-2 if k.endswith("Bias") else 8e-3), f"train-BN weight grad {k}: err {err:.2e}"
    # off again: running statistics (now the updated ones) through the fused epilogues
    net.SetTrainBatchNorm(False)
    on.train_bn = False
    acts2 = on.forward({"input": x, "ivector": iv})
    assert lib.kfp16_net_forward(net.ptr) == 0
    assert rel_to_scale(net.Output("output"), acts2["output"]) <= 5e-3
    assert rel_to_scale(acts2["output"], acts["output"]) > 1e-2       # ... which is a different function
    net.Free()


def test_kaldi_weight_import_export(handle, lib):
    """weight_loader: a network's parameters written as Kaldi nnet3 text ([out x in] matrices, batch-norm statistics), parsed
    back and loaded into a fresh network (LoadWeights, weight_loader.go:754-946) give the same forward pass as the oracle with
    those weights; ComponentsFromNetwork exports what was loaded"""
    from kaldi_fp16_b200 import weight_loader as WL
    n_seq, L = 2, 30
    xc = CNN_SMALL.replace("F1", "64").replace(
        "output-layer name=output include-log-softmax=false dim=72",
        "prefinal-layer name=prefinal-chain small-dim=64 big-dim=256\noutput-layer name=output include-log-softmax=false dim=72")
    rng = np.random.default_rng(77)
    on = OracleNet(xc, n_seq, L)
    on.init_random(rng)
    comps = {}

    def add(name, typ, w=None, b=None, bn=None):
        c = WL.KaldiComponent(Name=name, Type=typ)
        if w is not None:
            c.LinearParams = np.ascontiguousarray(w.T, np.float32)
        if b is not None:
            c.BiasParams = np.asarray(b, np.float32).reshape(-1)
        if bn is not None:
            c.StatsMean, c.StatsVar, c.TargetRms, c.Epsilon = bn
        comps[name] = c

    def rand_bn(key, rms=1.0):
        d = on.bn[key]["mean"].size
        mean, var = (rng.standard_normal(d) * 0.1).astype(np.float32), (rng.random(d) + 0.5).astype(np.float32)
        on.bn[key].update(mean=mean, var=var, gamma=np.full(d, rms, np.float32), beta=np.zeros(d, np.float32), eps=1e-3)
        return mean, var, rms, 1e-3

    P = on.params
    for k in P:
        if k.endswith("Bias"):
            P[k] = O.to_f16_trunc((rng.standard_normal(P[k].shape) * 0.1).astype(np.float32))
    from oracle.nnet_oracle import idct_matrix
    add("idct", "FixedAffineComponent", w=idct_matrix(16, 22.0))          # [in x out]; Kaldi's component stores the transpose
    add("ivector-linear", "LinearComponent", w=P["ivector-linear.W"])
    add("ivector-batchnorm", "BatchNormComponent", bn=rand_bn(("ivector-batchnorm", ""), 0.025))
    add("idct-batchnorm", "BatchNormComponent", bn=rand_bn(("idct-batchnorm", "")))
    for cn in ("cnn1", "cnn2", "cnn3"):
        add(f"{cn}.conv", "TimeHeightConvolutionComponent", w=P[f"{cn}.W"], b=P[f"{cn}.Bias"])
        add(f"{cn}.batchnorm", "BatchNormComponent", bn=rand_bn((cn, "BN")))
    for tn in ("tdnnf4", "tdnnf5"):
        add(f"{tn}.linear", "TdnnComponent", w=P[f"{tn}.LinearW"])
        add(f"{tn}.affine", "TdnnComponent", w=P[f"{tn}.AffineW"], b=P[f"{tn}.AffineBias"])
        add(f"{tn}.batchnorm", "BatchNormComponent", bn=rand_bn((tn, "AffBN")))
    add("prefinal-chain.affine", "NaturalGradientAffineComponent", w=P["prefinal-chain.BigW"], b=P["prefinal-chain.BigBias"])
    add("prefinal-chain.linear", "LinearComponent", w=P["prefinal-chain.SmallW"])
    add("prefinal-chain.batchnorm1", "BatchNormComponent", bn=rand_bn(("prefinal-chain", "PfBN")))
    add("prefinal-chain.batchnorm2", "BatchNormComponent", bn=rand_bn(("prefinal-chain", "BN")))
    add("output.affine", "NaturalGradientAffineComponent", w=P["output.W"], b=P["output.Bias"])
    text = WL.WriteNnet3Text(comps)
    parsed = WL.ParseNnet3Text(text)
    net = nnet.NewNetwork(nnet.BuildModelFromString(xc), handle, n_seq, L, train=False, seed=5)      # different random init
    rep = WL.LoadWeights(net, parsed)
    assert rep["loaded"] == 11 and rep["skipped"] == ["combine_inputs"], rep
    x = O.to_f16_rne((rng.standard_normal((n_seq * L, 16)) * 2).astype(np.float32))
    iv = O.to_f16_rne(np.clip(rng.standard_normal((n_seq, 24)), -3, 3).astype(np.float32))
    net.MarkPerSequence("ivector", "ivector-linear", "ivector-batchnorm")
    net.SetInput("input", x)
    net.SetInput("ivector", iv)
    assert lib.kfp16_net_forward(net.ptr) == 0
    acts = on.forward({"input": x, "ivector": iv})
    for name in ("cnn1", "cnn3", "tdnnf5", "prefinal-chain", "output"):
        assert rel_to_scale(net.Output(name), acts[name]) <= 2e-3, name
    # export: what was loaded comes back bit for bit (FP16-representable values), statistics included
    out = WL.ComponentsFromNetwork(net)
    for name in ("cnn2.conv", "tdnnf4.linear", "tdnnf5.affine", "prefinal-chain.linear", "output.affine"):
        assert np.array_equal(out[name].LinearParams, O.to_f16_trunc(parsed[name].LinearParams)), name
    assert np.array_equal(out["tdnnf4.batchnorm"].StatsVar, parsed["tdnnf4.batchnorm"].StatsVar)
    assert np.array_equal(out["cnn3.batchnorm"].StatsMean, parsed["cnn3.batchnorm"].StatsMean)
    net.Free()


ATTNET = """
input name=input dim=40
linear-component name=lin0 dim=64
tdnnf-layer name=tdnnf1 dim=64 bottleneck-dim=32 time-stride=1 bypass-scale=0.66
attention-relu-batchnorm-layer name=attention1 num-heads=4 value-dim=10 key-dim=8 num-left-inputs=3 num-right-inputs=2 time-stride=3
tdnnf-layer name=tdnnf2 dim=64 bottleneck-dim=32 time-stride=3 bypass-scale=0.66
output-layer name=output include-log-softmax=false dim=48
"""


@REF_ROUND
@pytest.mark.parametrize("n_seq,L", [(3, 29), (1, 7)])
def test_restricted_self_attention_layer(handle, lib, n_seq, L, ref_round):
    """attention-relu-batchnorm-layer on the device (the reference runs it on the CPU between a D2H and an H2D copy,
    internal/nnet/forward.go:795-909): projection GEMM + one attention / ReLU / batch-norm kernel, per-sequence zero padding of
    the context; backward = the exact transpose in two gather passes -- against the oracle (itself checked against the
    reference's loop nest and a float64 autograd in tests/test_nnet_oracle_cpu.py).  4 heads x (10 + 6) = 64 outputs,
    4 x (8 + 10 + 8 + 6) = 128 projection columns; L = 7 is shorter than the context span."""
    on, net, rng = make_pair(handle, ATTNET, n_seq, L, seed=41 + L, ref_round=ref_round)
    x = O.to_f16_rne(rng.standard_normal((n_seq * L, 40)).astype(np.float32))
    acts = on.forward({"input": x})
    net.SetInput("input", x)
    assert lib.kfp16_net_forward(net.ptr) == 0
    for name in ("tdnnf1", "attention1", "tdnnf2", "output"):
        err = rel_to_scale(net.Output(name), acts[name])
        assert err <= 2e-3, f"forward {name}: err {err:.2e}"
    att = net.Output("attention1").reshape(n_seq * L, 4, 16)
    masks = {name: net.Mask(name, 64) for name in ("tdnnf1", "tdnnf2")}
    for name, m in masks.items():
        assert np.mean(m != on.saved[name]["mask"]) < 5e-3
    wg, dact = on.backward("output", acts["output"], masks)
    net.ZeroGrads()
    net.Backward(None)
    got = net.WeightGrads()
    for k, g in wg.items():
        err = rel_to_scale(got[k], g)
        assert err <= (1e-2 if k.endswith("Bias") else 5e-3), f"weight grad {k}: err {err:.2e}"
    assert rel_to_scale(net.Grad("tdnnf1"), dact["tdnnf1"]) <= 5e-3
    # identity-free check of the softmax: before the batch-norm the last 6 columns of every head are weights that sum to 1;
    # with the randomised batch-norm undone they still do
    bn = on.bn[("attention1", "BN")]
    sc = (bn["gamma"] / np.sqrt(bn["var"] + np.float32(bn["eps"]))).reshape(4, 16)
    sh = (bn["beta"] - bn["mean"] * sc.reshape(-1)).reshape(4, 16)
    wsum = ((att - sh) / sc)[..., 10:].sum(-1)
    assert np.abs(wsum - 1.0).max() < 2e-2
    net.Free()
