"""Chain LF-MMI objective on the GPU (kfp16_chain_*, one launch per minibatch) against
  (a) the numpy oracle (oracle/chain_oracle.py, restating cpp/cuda/chain.cu), and
  (b) the REFERENCE'S OWN chain.cu compiled unmodified into oracle/_ref, called per sequence on materialised
      subsampled rows the way internal/nnet/chain_loss.go:221-294 does.
Tolerances: objective per frame |delta| <= 1e-3; gradient (FP16 posteriors in [-1, 1]) |delta| <= 2e-3."""
import ctypes as C

import numpy as np
import pytest

from kaldi_fp16_b200 import chain as KC
from kaldi_fp16_b200 import gpu, nnet
from oracle import chain_oracle as CO
from oracle import kaldi_oracle as O
from oracle.nnet_oracle import OracleNet
from tests.refbind import RefBuf, RefChainResult, read_f32, ref_chain_fst

pytestmark = pytest.mark.gpu


def to_kc(f: CO.Fst) -> KC.ChainFst:
    return KC.ChainFst(f.row_ptr, f.col_idx, f.labels, f.weights, f.final_states, f.final_weights, f.start_state)


def make_case(rng, n_seq, frames, P, den_states, arcs, branching_num=False):
    den = CO.random_ergodic_fst(rng, den_states, arcs, P)
    nums = []
    for s in range(n_seq):
        f = CO.linear_chain_fst(frames, P, offset=int(rng.integers(P)))
        if branching_num:      # alternative pronunciation: a second arc on some frames (two paths through the numerator)
            extra_src = np.arange(0, frames, 3, dtype=np.int32)
            order = np.argsort(np.concatenate([np.arange(frames), extra_src]), kind="stable")
            src = np.concatenate([np.arange(frames), extra_src])[order]
            dst = np.concatenate([np.arange(1, frames + 1), extra_src + 1])[order].astype(np.int32)
            lab = np.concatenate([f.labels, ((extra_src * 7 + s) % P + 1)])[order].astype(np.int32)
            w = np.concatenate([np.full(frames, -0.3), np.full(len(extra_src), -1.2)])[order].astype(np.float32)
            row_ptr = np.searchsorted(src, np.arange(frames + 2)).astype(np.int32)
            f = CO.Fst(row_ptr, dst, lab, w, f.final_states, f.final_weights, 0)
        nums.append(f)
    return den, nums


@pytest.mark.parametrize("general", [0, 1], ids=["smem_kernel", "global_kernel"])
@pytest.mark.parametrize("n_seq,seq_rows,frames,sub,left,P,S,arcs,branch", [
    (4, 40, 12, 3, 1, 48, 16, 3, False),
    (6, 31, 10, 3, 0, 104, 64, 4, True),
    (3, 20, 20, 1, 0, 3080, 40, 5, True),        # ragged pdf count, no subsampling
    (64, 156, 50, 3, 3, 6016, 256, 4, False),    # the benchmark's shape: 64 sequences x 50 output frames, 6016 pdfs
    (8, 156, 50, 3, 0, 1024, 512, 8, True),      # 4096 denominator arcs: too many to gather all frames' values up front
])
def test_batched_chain_objective_matches_oracle_and_reference(handle, lib, reflib, general, n_seq, seq_rows, frames, sub, left, P, S, arcs, branch):
    rng = np.random.default_rng(n_seq * 1000 + P)
    den, nums = make_case(rng, n_seq, frames, P, S, arcs, branch)
    out = O.to_f16_rne((rng.standard_normal((n_seq * seq_rows, P)) * 0.7).astype(np.float32))
    t_out = gpu.TensorFromFP16(out)
    t_grad = gpu.TensorFromFP16(np.full_like(out, 3.0))
    obj = KC.ChainObjective(handle, P, n_seq, frames, to_kc(den))
    obj.SetNumerators([to_kc(f) for f in nums])
    assert lib.kfp16_chain_force_general(obj.ptr, general) == 0
    weight = 0.75
    loss_acc = gpu.DeviceF32(n=4)
    assert lib.kfp16_chain_loss(obj.ptr, t_out.Ptr, t_grad.Ptr, P, seq_rows, left, sub, weight, loss_acc.Ptr) == 0, lib.kfp16_last_error()
    gpu.Sync()
    res = obj.Results()
    got_grad = t_grad.ToFP32()
    # ---- (a) numpy oracle
    want_res, want_grad = CO.chain_loss_batch(out, nums, den, n_seq, seq_rows, frames, sub, left, weight)
    assert np.abs(res.PerSeq[:, 0] - want_res[:, 0]).max() / frames <= 1e-3
    assert np.abs(res.PerSeq[:, 1] - want_res[:, 1]).max() / frames <= 1e-3
    assert np.abs(res.PerSeq[:, 2] - want_res[:, 2]).max() / frames <= 1e-3
    assert abs(loss_acc.ToHost()[0] - want_res[:, 2].sum()) <= 1e-3 * frames * n_seq
    used = np.zeros(n_seq * seq_rows, bool)
    for s in range(n_seq):
        used[s * seq_rows + left + np.arange(frames) * sub] = True
    assert np.abs(got_grad[used] - want_grad[used]).max() <= 2e-3
    assert np.all(got_grad[~used] == 3.0)            # rows off the subsampling grid are not touched
    # ---- (b) the reference's chain_compute_loss, per sequence on materialised subsampled rows (a few sequences)
    den_ref, den_bufs = ref_chain_fst(reflib, den)
    for s in list(range(n_seq))[:4]:
        rows = s * seq_rows + left + np.arange(frames) * sub
        sub_out = RefBuf(reflib, np.ascontiguousarray(out[rows]).astype(np.float16).view(np.uint16))
        g = RefBuf(reflib, np.zeros((frames, P), np.uint16))
        num_ref, num_bufs = ref_chain_fst(reflib, nums[s])
        r = RefChainResult()
        assert reflib.chain_compute_loss(sub_out.ptr, C.byref(num_ref), C.byref(den_ref), frames, P, g.ptr, C.byref(r)) == 0, reflib.chain_last_error()
        assert abs(res.PerSeq[s, 0] - r.num_logprob) / frames <= 1e-3, (s, res.PerSeq[s], r.num_logprob)
        assert abs(res.PerSeq[s, 1] - r.den_logprob) / frames <= 1e-3, (s, res.PerSeq[s], r.den_logprob)
        # the reference's gradient is unweighted here? no: chain_compute_loss applies supervision weight 1.0
        ref_grad = g.f32() * np.float32(weight)
        assert np.abs(got_grad[rows] - ref_grad).max() <= 2e-3 + 2.0 ** -10, s
        for p in num_bufs:
            reflib.bridge_gpu_free(p)
        sub_out.free()
        g.free()
    for p in den_bufs:
        reflib.bridge_gpu_free(p)
    obj.Free()
    t_out.Free()
    t_grad.Free()


CHAIN_NET = """
input name=input dim=64
linear-component name=lin0 dim=256
tdnnf-layer name=tdnnf1 dim=256 bottleneck-dim=64 time-stride=3 bypass-scale=0.66
tdnnf-layer name=tdnnf2 dim=256 bottleneck-dim=64 time-stride=3 bypass-scale=0.66
linear-component name=prefinal-l dim=64
prefinal-layer name=prefinal-chain input=prefinal-l big-dim=256 small-dim=64
output-layer name=output include-log-softmax=false dim=104
"""


def test_chain_training_step(handle, lib):
    """forward -> ComputeChainLossBatch -> Backward on a TDNN-F net: the parameter gradients equal the oracle's backward of
    the oracle's chain gradient; a captured step graph with the chain objective reproduces it; SGD on it lowers the loss"""
    n_seq, L, frames, sub, left, P = 4, 45, 14, 3, 1, 104
    rng = np.random.default_rng(99)
    on = OracleNet(CHAIN_NET, n_seq, L)
    on.init_random(rng)
    net = nnet.NewNetwork(nnet.BuildModelFromString(CHAIN_NET), handle, n_seq, L, ref_round=True, lr=5e-4, momentum=0.0)
    for k, w in on.params.items():
        net.SetParam(k, w)
    den, nums = make_case(rng, n_seq, frames, P, 24, 4)
    obj = KC.ChainObjective(handle, P, n_seq, frames, to_kc(den))
    obj.SetNumerators([to_kc(f) for f in nums])
    x = O.to_f16_rne(rng.standard_normal((n_seq * L, 64)).astype(np.float32))
    acts = on.forward({"input": x})
    want_res, dy = CO.chain_loss_batch(acts["output"], nums, den, n_seq, L, frames, sub, left)
    net.SetInput("input", x)
    net.ZeroGrads()
    assert lib.kfp16_net_forward(net.ptr) == 0
    net.ReadLoss()
    r = KC.ComputeChainLossBatch(net, obj, sub, left)
    assert abs(r.Loss - want_res[:, 2].mean()) / frames <= 1e-3
    assert np.abs(net.Grad("output") - dy).max() <= 3e-3
    assert lib.kfp16_net_backward(net.ptr) == 0
    assert abs(net.ReadLoss() - want_res[:, 2].sum()) <= 1e-3 * frames * n_seq
    masks = {l.name: net.Mask(l.name, on.saved[l.name]["mask"].shape[1]) for l in on.layers if l.type in ("tdnnf-layer", "prefinal-layer")}
    wg, _ = on.backward("output", dy, masks)
    got = net.WeightGrads()
    for k, g in wg.items():
        err = O.max_err_vs_scale(got[k], g)
        assert err <= (2e-2 if k.endswith("Bias") else 1e-2), f"{k}: {err:.2e}"     # FP16 posterior gradients: coarser than dY = Y
    # captured step with the chain objective + SGD: the loss goes down
    from kaldi_fp16_b200 import cudart
    st = cudart.Stream()
    lib.kfp16_ctx_set_stream(handle.ptr, st.ptr)
    try:
        assert lib.kfp16_net_set_chain(net.ptr, obj.ptr, sub, left, 1.0) == 0
        net.Capture(3)
        losses = []
        net.ReadLoss()
        for _ in range(4):      # (a random denominator graph does not bound the objective: a few steps only)
            net.Launch(3)
            st.synchronize()
            losses.append(net.ReadLoss())
        assert abs(losses[0] - want_res[:, 2].sum()) <= 1e-3 * frames * n_seq
        assert np.all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    finally:
        lib.kfp16_net_set_chain(net.ptr, None, 3, 0, 1.0)
        lib.kfp16_ctx_set_stream(handle.ptr, None)
        st.destroy()
    obj.Free()
    net.Free()


@pytest.mark.parametrize("n_seq,L,frames,sub,left", [(4, 45, 14, 3, 1), (3, 45, 15, 3, 0), (5, 44, 14, 3, 2)])
def test_output_frame_rows_backward_equals_dense(handle, lib, n_seq, L, frames, sub, left):
    """frame subsampling: back-propagating only the output frames' rows through output / prefinal / prefinal-l (the rows
    row0 + 3k of the padded matrix; the first splicing layer gets the dense gradient) gives the gradients of the dense
    backward pass over a cleared output gradient.  (5 x 44: the block of 50 rows is not a multiple of 3 -> dense either way)"""
    P = 104
    rng = np.random.default_rng(5)
    on = OracleNet(CHAIN_NET, n_seq, L)
    on.init_random(rng)
    den, nums = make_case(rng, n_seq, frames, P, 24, 4)
    x = O.to_f16_rne(rng.standard_normal((n_seq * L, 64)).astype(np.float32))
    res = []
    for sparse in (True, False):
        net = nnet.NewNetwork(nnet.BuildModelFromString(CHAIN_NET), handle, n_seq, L, ref_round=False)
        for k, w in on.params.items():
            net.SetParam(k, w)
        net.SetSparseOutputGrad(sparse)
        obj = KC.ChainObjective(handle, P, n_seq, frames, to_kc(den))
        obj.SetNumerators([to_kc(f) for f in nums])
        net.SetInput("input", x)
        net.ZeroGrads()
        assert lib.kfp16_net_forward(net.ptr) == 0
        KC.ComputeChainLossBatch(net, obj, sub, left)
        assert lib.kfp16_net_backward(net.ptr) == 0
        res.append((net.WeightGrads(), net.Grad("tdnnf2"), net.Grad("output"), net.Grad("prefinal-l")))
        obj.Free()
        net.Free()
    (wa, ga, oa, pa), (wb, gb, ob, pb) = res
    assert np.array_equal(oa, ob)                       # the objective's gradient itself, dense read-back
    assert np.array_equal(ga, gb)                       # what the splicing layer receives: same GEMM rows, bit-identical
    assert np.array_equal(pa, pb)
    for k in wb:
        scale = max(np.abs(wb[k]).max(), 1e-12)
        assert np.abs(wa[k] - wb[k]).max() <= 2e-6 * scale + 1e-7, k      # fp32 sums over a third of the rows: order only


@pytest.mark.parametrize("T,P,S,arcs,branch", [(14, 104, 24, 4, False), (50, 6016, 256, 4, False), (9, 40, 7, 3, True)])
def test_reference_chain_entry_points(handle, lib, reflib, T, P, S, arcs, branch):
    """chain_forward_backward / chain_compute_posteriors / chain_compute_loss / chain_workspace_bytes exported with the
    reference's signatures (cpp/include/chain.h:47-160; internal/nnet/chain_loss.go:189,309,325 binds them): against the numpy
    oracle and against the same calls into the reference's own library (oracle/_ref), FSTs with device pointers as the Go side
    uploads them"""
    rng = np.random.default_rng(T * 7 + P)
    den, nums = make_case(rng, 1, T, P, S, arcs, branch)
    num = nums[0]
    out = O.to_f16_rne((rng.standard_normal((T, P)) * 0.7).astype(np.float32))
    # inputs live in the reference library's allocations (plain cudaMalloc): both libraries read the same buffers
    t_out = RefBuf(reflib, out.astype(np.float16).view(np.uint16))
    assert lib.chain_workspace_bytes(T, S) == reflib.chain_workspace_bytes(T, S) == 2 * (T + 1) * S * 4
    for fst in (num, den):
        f_ref, bufs = ref_chain_fst(reflib, fst)
        Sn = fst.num_states
        res = {}
        for name, L in (("ours", lib), ("ref", reflib)):
            alpha = RefBuf(reflib, np.zeros(((T + 1), Sn), np.float32))
            beta = RefBuf(reflib, np.zeros(((T + 1), Sn), np.float32))
            post = RefBuf(reflib, np.full((T, P), 5.0, np.float32))
            total = C.c_float(0)
            assert L.chain_forward_backward(t_out.ptr, C.byref(f_ref), T, P, alpha.ptr, beta.ptr, C.byref(total)) == 0, L.chain_last_error()
            assert L.chain_compute_posteriors(t_out.ptr, C.byref(f_ref), T, P, alpha.ptr, beta.ptr, total.value, post.ptr) == 0
            host = [read_f32(alpha.ptr, (T + 1, Sn)), read_f32(beta.ptr, (T + 1, Sn)), read_f32(post.ptr, (T, P))]
            res[name] = (total.value, host)
            for b_ in (alpha, beta, post):
                b_.free()
        a_o, b_o, tot_o = CO.forward_backward(out, fst)
        p_o = CO.posteriors(out, fst, a_o, b_o, tot_o)
        assert abs(res["ours"][0] - tot_o) <= 1e-3 * T and abs(res["ours"][0] - res["ref"][0]) <= 1e-3 * T
        ga, gb, gp = res["ours"][1]
        live = a_o > CO.LOG_ZERO / 2
        assert np.array_equal(ga > CO.LOG_ZERO / 2, live) and np.abs(ga[live] - a_o[live]).max() <= 1e-3 * T      # alpha, state by state
        live = b_o > CO.LOG_ZERO / 2
        assert np.array_equal(gb > CO.LOG_ZERO / 2, live) and np.abs(gb[live] - b_o[live]).max() <= 1e-3 * T      # beta
        assert np.abs(gp - p_o).max() <= 2e-3 and np.abs(gp - res["ref"][1][2]).max() <= 2e-3                      # posteriors
        assert np.abs(gp.sum(1) - 1.0).max() <= 5e-3          # every frame's posteriors sum to one
        for p in bufs:
            reflib.bridge_gpu_free(p)
    # chain_compute_loss: same struct, same buffers
    num_ref, nb = ref_chain_fst(reflib, num)
    den_ref, db = ref_chain_fst(reflib, den)
    want = CO.chain_loss(out, num, den)
    got = {}
    for name, L in (("ours", lib), ("ref", reflib)):
        g = RefBuf(reflib, np.zeros((T, P), np.uint16))
        r = RefChainResult()
        assert L.chain_compute_loss(t_out.ptr, C.byref(num_ref), C.byref(den_ref), T, P, g.ptr, C.byref(r)) == 0, L.chain_last_error()
        got[name] = (r.num_logprob, r.den_logprob, r.loss, g.f32())
        g.free()
    for k in range(3):
        assert abs(got["ours"][k] - want[k]) / T <= 1e-3, (k, got["ours"][k], want[k])
        assert abs(got["ours"][k] - got["ref"][k]) / T <= 1e-3
    assert np.abs(got["ours"][3] - want[3]).max() <= 2e-3
    assert np.abs(got["ours"][3] - got["ref"][3]).max() <= 2e-3 + 2.0 ** -10
    # no gradient requested
    r = RefChainResult()
    assert lib.chain_compute_loss(t_out.ptr, C.byref(num_ref), C.byref(den_ref), T, P, None, C.byref(r)) == 0
    assert abs(r.loss - want[2]) / T <= 1e-3
    # errors are reported through chain_last_error
    assert lib.chain_compute_loss(None, C.byref(num_ref), C.byref(den_ref), T, P, None, C.byref(r)) == -1
    assert b"bad argument" in lib.chain_last_error()
    lib.chain_clear_error()
    assert lib.chain_last_error() is None
    for p in nb + db:
        reflib.bridge_gpu_free(p)
    t_out.free()
