#!/usr/bin/env python
"""bench.py -- frames/sec of the FP16 fwd+bwd(+SGD) step on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload: the full CNN-TDNN (BASELINE configs[2], the config the metric is quoted on);
--workload tdnnf_stack runs BASELINE configs[1].  A "step" is one minibatch (64 sequences x 150
frames per GPU) through zero-grads, forward, 0.5*||out||^2 objective (dY = Y,
cmd/sgdtest/main.go:258-267), backward, gradient all-reduce (N > 1: the FP16 gradient bucket) and
the momentum-SGD update.  `value` is whole-job frames/s with the inputs resident in HBM; `e2e` is
the same step through the public host-buffer API (pinned host FP32 features -> H2D every step,
FP32 -> FP16 on the device, loss read back every step).  One JSON line is printed by rank 0.

--impl reference times the reference's own CPU engine (go/gotorch restated in C, oracle/
gotorch_port.c -- there is no Go toolchain here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_SEQ, SEQ_LEN = 64, 150
CHAIN_FRAMES = 50   # output frames per sequence: frame-subsampling factor 3 (internal/nnet/chain_loss.go:221-294)
LR = 1e-4   # with gradients scaled to the mean over frames x output dims (keeps the 0.5*||out||^2 objective stable)
LR_CHAIN = 1e-6   # the synthetic denominator graph does not bound the LF-MMI objective: a small step keeps 300 updates finite


def tdnnf_stack_xconfig(layers=16, dim=1536, bott=160, stride=3):
    lines = [f"input name=input dim={dim}"]
    for i in range(layers):
        lines.append(f"tdnnf-layer name=tdnnf{i + 1} dim={dim} bottleneck-dim={bott} time-stride={stride} bypass-scale=0.66")
    return "\n".join(lines) + "\n"


def cnn_tdnn_xconfig(pdfs=6016):
    """SURVEY Appendix D.2 (Kaldi cnn_tdnn_1a family in the reference's xconfig vocabulary)"""
    l = ["input dim=100 name=ivector", "input dim=40 name=input",
         "idct-layer name=idct input=input dim=40 cepstral-lifter=22",
         "linear-component name=ivector-linear dim=200 input=ReplaceIndex(ivector, t, 0)",
         "batchnorm-component name=ivector-batchnorm target-rms=0.025",
         "batchnorm-component name=idct-batchnorm input=idct",
         "combine-feature-maps-layer name=combine_inputs input=Append(idct-batchnorm, ivector-batchnorm) num-filters1=1 num-filters2=5 height=40"]
    conv = [("cnn1", 40, 40, 1, 64), ("cnn2", 40, 40, 1, 64), ("cnn3", 40, 20, 2, 128), ("cnn4", 20, 20, 1, 128),
            ("cnn5", 20, 10, 2, 256), ("cnn6", 10, 10, 1, 256)]
    for name, hin, hout, sub, f in conv:
        l.append(f"conv-relu-batchnorm-layer name={name} height-in={hin} height-out={hout} height-subsample-out={sub} "
                 f"time-offsets=-1,0,1 height-offsets=-1,0,1 num-filters-out={f}")
    l.append("tdnnf-layer name=tdnnf7 dim=1536 bottleneck-dim=256 time-stride=0")
    for i in range(8, 19):
        l.append(f"tdnnf-layer name=tdnnf{i} dim=1536 bottleneck-dim=160 time-stride=3")
    l += ["linear-component name=prefinal-l dim=256",
          "prefinal-layer name=prefinal-chain input=prefinal-l big-dim=1536 small-dim=256",
          f"output-layer name=output include-log-softmax=false dim={pdfs}",
          "prefinal-layer name=prefinal-xent input=prefinal-l big-dim=1536 small-dim=256",
          f"output-layer name=output-xent dim={pdfs}"]
    return "\n".join(l) + "\n"


WORKLOADS = {
    "tdnnf_stack": dict(xconfig=tdnnf_stack_xconfig, feat_dim=1536, ivec_dim=0,
                        desc="TDNN-F stack 16 x (1536 hidden, 160 bottleneck, stride 3, bypass 0.66) fwd+bwd+SGD, 64 seqs x 150 frames per GPU (BASELINE configs[1])"),
    "cnn_tdnn": dict(xconfig=cnn_tdnn_xconfig, feat_dim=40, ivec_dim=100, dp_cut="tdnnf7",
                     desc="full CNN-TDNN (6 conv + 12 TDNN-F + prefinal + 6016-pdf output) fwd+bwd+SGD, 64 seqs x 150 frames per GPU (BASELINE configs[2])"),
}


def build_synthetic_chain(handle, out_dim, rank=0):
    """synthetic supervision: per sequence a linear-chain numerator over the 50 output frames (pdf = (i + seq) mod P,
    weight 0: internal/nnet/backward_test.go:42-58), one small random ergodic HMM as the shared denominator graph"""
    from kaldi_fp16_b200 import chain as KC
    crng = np.random.default_rng(7)
    S, K = 256, 4
    w = crng.random((S, K)) + 0.1
    den = KC.ChainFst((np.arange(S + 1) * K), crng.integers(0, S, size=S * K), crng.integers(0, out_dim, size=S * K) + 1,
                      np.log(w / w.sum(1, keepdims=True)).reshape(-1), np.arange(S), np.zeros(S), 0)
    nums = [KC.ChainFst(np.concatenate([np.arange(CHAIN_FRAMES), [CHAIN_FRAMES, CHAIN_FRAMES]]), np.arange(1, CHAIN_FRAMES + 1),
                        (np.arange(CHAIN_FRAMES) + q + rank * N_SEQ) % out_dim + 1, np.zeros(CHAIN_FRAMES), [CHAIN_FRAMES], [0.0], 0)
            for q in range(N_SEQ)]
    obj = KC.ChainObjective(handle, out_dim, N_SEQ, CHAIN_FRAMES, den)
    obj.SetNumerators(nums)
    return obj


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the bench runs"""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.marks = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        self.marks.append(time.time())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = []
        for line in Path(self.f.name).read_text().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = time.mktime(time.strptime(parts[0].split(".")[0], "%Y/%m/%d %H:%M:%S")) + float("0." + parts[0].split(".")[1])
                rows.append((ts, float(parts[2]), float(parts[3]), float(parts[4]), parts[5:9]))
            except (ValueError, IndexError):
                continue
        os.unlink(self.f.name)
        sel = rows
        if len(self.marks) >= 2:
            inside = [r for r in rows if self.marks[0] - 0.05 <= r[0] <= self.marks[-1] + 0.05]
            if len(inside) >= 1:
                sel = inside
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in sel for i in range(4) if r[4][i].lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[1] for r in sel])), "sm_max_mhz": sel[0][2],
                "power_w_max": max(r[3] for r in sel), "samples": len(sel), "reasons": reasons}


# ------------------------------------------------------------------------------------------ CPU arm
def load_port():
    so = ROOT / "oracle" / "_build" / "libgotorch_port.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "port"], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.gt_bench_tdnnf_stack.restype = C.c_double
    lib.gt_bench_tdnnf_stack.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_double)]
    lib.gt_bench_cnn_tdnn.restype = C.c_double
    lib.gt_bench_cnn_tdnn.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_double)]
    return lib


def cpu_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))


def cpu_sample(lib, workload: str, frames: int) -> float:
    """one fwd+bwd of the workload's network on `frames` frames of one sequence; returns seconds"""
    cs = C.c_double()
    if workload == "cnn_tdnn":
        return lib.gt_bench_cnn_tdnn(frames, 6016, 1, cpu_threads(), C.byref(cs))
    return lib.gt_bench_tdnnf_stack(16, 1536, 160, 3, 1, frames, 1, C.byref(cs))


def cpu_sample_desc(workload: str, frames: int) -> str:
    if workload == "cnn_tdnn":
        return (f"full CNN-TDNN fwd+bwd on 1 sequence x {frames} frames, float64: gotorch Conv1DLayer loop nest "
                f"(go/gotorch/cnn_tdnn.go:85-172, extended over height with the 3x3 taps of forward.go:429-455), TDNNLayer "
                f"(layers.go:444-524) and AffineLayer (layers.go:57-110) restated in C; as in the reference only MatMul "
                f"(the affine layers) is multi-threaded ({cpu_threads()} threads), conv / TDNN / every Backward are "
                f"single-goroutine loops; host has {os.cpu_count()} cores")
    return (f"TDNN-F stack (16 layers) fwd+bwd on 1 sequence x {frames} frames, float64, gotorch TDNNLayer loops "
            f"(go/gotorch/layers.go:444-524, single goroutine in the reference; host has {os.cpu_count()} cores)")


def cpu_frames_for(lib, workload: str, budget_s: float) -> int:
    t_small = cpu_sample(lib, workload, 4)
    return int(max(4, min(SEQ_LEN, 4 * budget_s / max(t_small, 1e-3))))


def cpu_baseline(workload: str, budget_s: float = 12.0):
    lib = load_port()
    frames = cpu_frames_for(lib, workload, budget_s)
    t = cpu_sample(lib, workload, frames)
    return {"value": frames / t, "unit": "frames/s", "cores": cpu_threads() if workload == "cnn_tdnn" else 1, "kind": "port",
            "sample": cpu_sample_desc(workload, frames) + f", {t:.1f} s"}


def metric_name(workload: str) -> str:
    return {"cnn_tdnn": "CNN-TDNN fwd+bwd frames/sec", "tdnnf_stack": "TDNN-F stack fwd+bwd frames/sec"}[workload]


def run_reference(args, rank):
    if rank != 0:
        return
    lib = load_port()
    per_step = max(1.0, min(6.0, 150.0 / max(1, args.steps + args.warmup)))
    frames = cpu_frames_for(lib, args.workload, per_step)
    for _ in range(args.warmup):
        cpu_sample(lib, args.workload, frames)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample(lib, args.workload, frames)
    dt = time.perf_counter() - t0
    val = frames * args.steps / dt
    sample = cpu_sample_desc(args.workload, frames) + " per step; reference CPU engine restated in C (no Go toolchain)"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.workload), "value": val, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["desc"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cpu_threads() if args.workload == "cnn_tdnn" else 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("KFP16_WORKLOAD", "cnn_tdnn"), choices=sorted(WORKLOADS))
    ap.add_argument("--objective", default=None, choices=["chain", "half_sq"],
                    help="chain = LF-MMI on 50 output frames per sequence (default for cnn_tdnn: BASELINE configs[2] is a "
                         "chain-model SGD step); half_sq = 0.5*||out||^2, dY = Y (cmd/sgdtest/main.go:258-267)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 64 sequences per GPU (the contract's default); strong: 64 sequences in total, 64 / N per GPU (SURVEY 8d config 4)")
    ap.add_argument("--profile-steps", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    global N_SEQ
    global_seqs = N_SEQ * world
    if args.scaling == "strong":
        if N_SEQ % world:
            raise SystemExit(f"--scaling strong: {N_SEQ} sequences do not divide over {world} ranks")
        global_seqs = N_SEQ
        N_SEQ = N_SEQ // world        # every use below (network, inputs, supervision) is per rank

    from kaldi_fp16_b200 import _lib, cudart, gpu, nnet
    lib = _lib.load()          # raises if the CUDA library is missing: there is no fallback
    dist = torch = None
    # N > 1: the FP16 gradient bucket is all-reduced (the reference keeps FP16 gradient tensors: backward_ops.go:195-225):
    # 36 MB instead of the 72 MB FP32 bucket.  KFP16_DP_OVERLAP=1 additionally cuts the step graph once, where the backward
    # pass leaves the TDNN-F / output layers, and reduces their gradients (97 % of the bucket) on a second stream while the
    # convolutional front end back-propagates -- measured on 8 B200 it does NOT pay (3.64 against 3.61 ms per step: the
    # exposed all-reduce is ~0.2 ms and the compute kernels lose SMs to it), so it is off by default.
    overlap = world > 1 and os.environ.get("KFP16_DP_OVERLAP", "0") != "0"
    nccl_sms = int(os.environ.get("KFP16_NCCL_SMS", "16"))
    if world > 1:
        # NCCL's own banner / debug lines go to stderr so that stdout carries the one JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if overlap:
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(nccl_sms))
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    gpu.Init(local)
    handle = gpu.NewHandle()
    if world > 1:
        tstream = torch.cuda.Stream(device=local)
        stream_ptr = tstream.cuda_stream
    else:
        tstream = None
        st = cudart.Stream()
        stream_ptr = st.ptr
    lib.kfp16_ctx_set_stream(handle.ptr, stream_ptr)

    wl = WORKLOADS[args.workload]
    out_dim = {"tdnnf_stack": 1536, "cnn_tdnn": 6016}[args.workload]
    objective = args.objective or ("chain" if args.workload == "cnn_tdnn" else "half_sq")
    # chain: posterior gradients in [-1, 1] on the output frames, averaged over them; half_sq: mean over frames x dims
    # weak scaling keeps the per-rank scale (sum semantics over the ranks); strong scaling scales by the GLOBAL
    # minibatch, so that the N-GPU step equals the 1-GPU step on the same 64 sequences
    gs_seqs = global_seqs if args.scaling == "strong" else N_SEQ
    grad_scale = 1.0 / (gs_seqs * CHAIN_FRAMES) if objective == "chain" else 1.0 / (gs_seqs * SEQ_LEN * out_dim)
    lr = LR_CHAIN if objective == "chain" else LR
    net = nnet.NewNetwork(nnet.BuildModelFromString(wl["xconfig"]()), handle, N_SEQ, SEQ_LEN, train=True, lr=lr,
                          momentum=0.9, ref_round=False, seed=42, grad_scale=grad_scale)
    T, fd, ivd = N_SEQ * SEQ_LEN, wl["feat_dim"], wl["ivec_dim"]
    rng = np.random.default_rng(1234 + rank)
    if args.workload == "cnn_tdnn":   # MFCC-like columns
        feats = rng.standard_normal((T, fd)).astype(np.float32) * (10.0 * 0.9 ** np.arange(fd, dtype=np.float32))
        feats[:, 0] = np.clip(60 + 20 * rng.standard_normal(T), -20, 105)
    else:
        feats = rng.standard_normal((T, fd)).astype(np.float32)
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    ivecs = np.ascontiguousarray(np.clip(rng.standard_normal((N_SEQ, max(ivd, 1))), -3, 3), dtype=np.float32) if ivd else None
    feat_bits = nnet.rne_fp16_bits(feats)
    ivec_bits = nnet.rne_fp16_bits(ivecs) if ivd else None

    # inputs resident in HBM as FP16 (device-timed leg) and in pinned host memory as FP32 (e2e leg: the features are
    # FP32 in the egs; the reference converts them on the CPU, internal/gpu/bridge.go:141 -- here on the device)
    d_feat = gpu.TensorFromBits(feat_bits)
    d_ivec = gpu.TensorFromBits(ivec_bits) if ivd else None
    h_feat_ptr = lib.bridge_host_alloc(feats.nbytes)
    C.memmove(h_feat_ptr, feats.ctypes.data, feats.nbytes)
    h_ivec_ptr = None
    if ivd:
        h_ivec_ptr = lib.bridge_host_alloc(ivecs.nbytes)
        C.memmove(h_ivec_ptr, ivecs.ctypes.data, ivecs.nbytes)

    def set_inputs_device():
        assert lib.kfp16_net_set_input_device(net.ptr, b"input", d_feat.Ptr, T, fd) == 0, _lib.last_error()
        if ivd:
            assert lib.kfp16_net_set_input_device(net.ptr, b"ivector", d_ivec.Ptr, N_SEQ, ivd) == 0, _lib.last_error()

    chain_obj = None
    if objective == "chain":
        chain_obj = build_synthetic_chain(handle, out_dim, rank)
        assert lib.kfp16_net_set_chain(net.ptr, chain_obj.ptr, 3, 0, 1.0) == 0, _lib.last_error()
    set_inputs_device()
    # graphs: N = 1: [step] [SGD on the FP32 bucket, gradients rounded to FP16 as the reference's are]
    #         N > 1: [step + FP16 gradient export] all-reduce(g16) [SGD on the reduced FP16 bucket]
    n_seg, seg_ranges = 0, []
    reducer = None
    exchange = "nccl"
    if world > 1:
        from kaldi_fp16_b200 import dp
        # the exchange: the library's own kernel over NVLink peer memory (KFP16_DP_EXCHANGE=nccl: torch.distributed's
        # all-reduce; also what is left when the peers' buckets cannot be mapped -- the JSON line says which ran)
        if os.environ.get("KFP16_DP_EXCHANGE", "peer") == "peer":
            try:
                reducer = dp.PeerGradAllReducer(lib, handle.ptr, lib.kfp16_net_grads_f16(net.ptr), lib.kfp16_net_bucket_size(net.ptr))
                exchange = "peer"
            except RuntimeError as e:
                print(f"[bench] rank {rank}: {e}; using the NCCL all-reduce", file=sys.stderr)
        if overlap:
            # the peer kernel runs as small CTAs beside the compute kernels on every SM; NCCL gets SMs of its own
            tail_ctas = 0 if exchange == "peer" else ((148 - nccl_sms) if nccl_sms > 0 else 0)
            n_seg = net.CaptureSegments(2, cut_layers=wl.get("dp_cut"), export_f16=True, tail_max_ctas=tail_ctas)
            seg_ranges = [net.SegmentGrads(k) for k in range(n_seg)]
        else:
            net.Capture(1 | 4)
        net.Capture(8)
        if exchange != "peer":
            reducer = dp.GradAllReducer(torch.as_tensor(net.grads_as_cuda_array(f16=True), device=f"cuda:{local}"))
        comm_stream = torch.cuda.Stream(device=local, priority=-1)
        seg_ev = [torch.cuda.Event() for _ in range(max(n_seg, 1))]
        peer_threads = int(os.environ.get("KFP16_PEER_THREADS", "64"))
    else:
        net.Capture(1)
        net.Capture(2)

    def step_compute_and_reduce():
        """graph(s) of the step + the sum all-reduce of the gradient bucket: N-GPU step == 1-GPU step on the
        concatenated batch (up to FP16 rounding of the per-rank gradients)"""
        if world == 1:
            net.Launch(1)
            return
        if exchange == "peer" and not overlap:
            net.Launch(1 | 4)
            reducer.all_reduce()                           # one kernel on the step's stream
            return
        if exchange == "peer":
            for k in range(n_seg):
                net.LaunchSegment(k)                       # backward of one layer group (+ its FP16 export) ...
                if k < n_seg - 1:                          # ... its gradients are exchanged beside the next segment
                    seg_ev[k].record(tstream)
                    comm_stream.wait_event(seg_ev[k])
                    reducer.all_reduce_range(*seg_ranges[k], channel=1 + k, stream_ptr=comm_stream.cuda_stream, threads=peer_threads)
                else:
                    reducer.all_reduce_range(*seg_ranges[k], channel=0)
            seg_ev[n_seg - 1].record(comm_stream)
            tstream.wait_event(seg_ev[n_seg - 1])          # the SGD graph waits for both exchanges
            return
        if not overlap:
            net.Launch(1 | 4)
            with torch.cuda.stream(tstream):
                reducer.all_reduce()
            return
        with torch.cuda.stream(tstream):
            for k in range(n_seg):
                net.LaunchSegment(k)                       # backward of one layer group (+ its FP16 export) ...
                ev = torch.cuda.Event()
                ev.record(tstream)
                comm_stream.wait_event(ev)
                with torch.cuda.stream(comm_stream):       # ... its gradients reduce beside the next segment
                    reducer.all_reduce_range(*seg_ranges[k], async_op=False)
            ev = torch.cuda.Event()
            ev.record(comm_stream)
            tstream.wait_event(ev)                         # the SGD graph waits for the reduced bucket

    def step_device():
        set_inputs_device()
        step_compute_and_reduce()
        net.Launch(2 if world == 1 else 8)

    def sync_all():
        cudart.synchronize()
        if world > 1:
            dist.barrier()
            cudart.synchronize()

    net.ReadLoss()
    step_device()
    cudart.synchronize()
    first_loss = net.ReadLoss()
    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.kfp16_launch_count()
    e0, e1 = cudart.Event(), cudart.Event()
    if sampler:
        sampler.mark()
    e0.record(stream_ptr)
    for _ in range(args.steps):
        step_device()
    e1.record(stream_ptr)
    e1.synchronize()
    sync_all()
    if sampler:
        sampler.mark()
    ms = e0.elapsed_ms(e1)
    launches = lib.kfp16_launch_count() - launches0
    if world > 1:
        t_ms = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms = float(t_ms.item())
    frames_per_step = T * world
    value = frames_per_step * args.steps / (ms * 1e-3)

    # ---- end-to-end leg: every step's inputs come from pinned host memory as FP32 (H2D inside the timed region, the
    # FP32 -> FP16 conversion on the device) and the step's loss is read back to the host.  The copy of step i+1's
    # inputs is issued on the library's copy stream while step i computes (kfp16_net_prefetch_input_f32 / commit_input),
    # as a training loop feeding egs would.  Timed by CUDA events AND by the host clock; the larger one counts.
    e2e_steps = max(3, min(args.steps, 100))

    def prefetch_host():
        assert lib.kfp16_net_prefetch_input_f32(net.ptr, b"input", h_feat_ptr, T, fd) == 0, _lib.last_error()
        if ivd:
            assert lib.kfp16_net_prefetch_input_f32(net.ptr, b"ivector", h_ivec_ptr, N_SEQ, ivd) == 0, _lib.last_error()

    def commit_host():
        assert lib.kfp16_net_commit_input(net.ptr, b"input") == 0, _lib.last_error()
        if ivd:
            assert lib.kfp16_net_commit_input(net.ptr, b"ivector") == 0, _lib.last_error()

    def e2e_loop(nsteps):
        last = 0.0
        prefetch_host()
        for i in range(nsteps):
            commit_host()
            if i + 1 < nsteps:
                prefetch_host()
            step_compute_and_reduce()
            net.Launch(2 if world == 1 else 8)
            # every step's loss is read back to the host; the read of step i is queued behind step i and collected
            # after step i+1 has been queued, so the stream never drains between minibatches
            net.ReadLossAsync(i & 1)
            if i > 0:
                last = net.WaitLoss((i - 1) & 1)
        return net.WaitLoss((nsteps - 1) & 1)

    net.ReadLoss()
    e2e_loop(3)                 # warm-up of the host-fed path (staging buffers, copy streams)
    sync_all()
    t0 = time.perf_counter()
    e0.record(stream_ptr)
    last_loss = e2e_loop(e2e_steps)
    e1.record(stream_ptr)
    e1.synchronize()
    cudart.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sync_all()
    e2e_event_ms = e0.elapsed_ms(e1)
    e2e_ms = max(e2e_event_ms, wall_ms)
    if world > 1:
        t_ms = torch.tensor([e2e_ms], device=f"cuda:{local}")
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        e2e_ms = float(t_ms.item())
    e2e_value = frames_per_step * e2e_steps / (e2e_ms * 1e-3)
    h2d = feats.nbytes + (ivecs.nbytes if ivd else 0)
    clocks = sampler.stop() if sampler else None

    # ---- data-parallel invariant: every rank holds the same master weights after the timed loops
    ranks_identical = None
    if exchange == "peer":
        reducer.check()           # raises if a peer failed to arrive in any exchange
        reducer.close()
    if world > 1:
        w = torch.as_tensor(net.params_as_cuda_array(), device=f"cuda:{local}")
        wmax, wmin = w.clone(), w.clone()
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
        ranks_identical = bool(torch.equal(wmax, wmin))
        assert ranks_identical, "master weights differ between ranks after the timed loop"

    # ---- side table: CUDA-event pair around every GEMM launch in an eager replay of the same step (per-kernel rates)
    lib.kfp16_ctx_set_profile(handle.ptr, 1)
    ev0, ev1 = cudart.Event(), cudart.Event()
    ev0.record(stream_ptr)
    for _ in range(args.profile_steps):
        set_inputs_device()
        net.ZeroGrads()
        assert lib.kfp16_net_forward(net.ptr) == 0
        if chain_obj is not None:
            assert lib.kfp16_net_loss_chain(net.ptr, b"", chain_obj.ptr, 3, 0, 1.0) == 0, _lib.last_error()
            assert lib.kfp16_net_backward(net.ptr) == 0, _lib.last_error()
        else:
            net.Backward(None)
        net.SGDStep(grad_scale)
    ev1.record(stream_ptr)
    ev1.synchronize()
    profile_ms = ev0.elapsed_ms(ev1) / max(args.profile_steps, 1)
    n_l, g_ms, g_fl = C.c_int(), C.c_double(), C.c_double()
    lib.kfp16_ctx_profile_read(handle.ptr, C.byref(n_l), C.byref(g_ms), C.byref(g_fl))
    lib.kfp16_ctx_set_profile(handle.ptr, 0)
    burst, sustained, hbm, src = peaks()
    traffic = traffic_src = None
    tp = ROOT / "profiles" / f"r02_ncu_traffic_{args.workload}.json"
    if tp.exists():
        tj = json.loads(tp.read_text())
        traffic, traffic_src = tj.get("dram_bytes_per_step"), f"profiles/{tp.name}: {tj.get('how', '')}"
    # ALGORITHMIC work of the step as executed (real rows only): forward GEMMs of every layer (the xent
    # branch included) + weight- and input-gradient GEMMs of the layers on the gradient path
    flops_fwd = lib.kfp16_net_flops_forward(net.ptr)
    flops_bwd = lib.kfp16_net_flops_backward(net.ptr)
    # ... minus the rows the step does not compute: with the chain objective (frame subsampling 3) the row-wise layers behind
    # the objective (prefinal-chain, output, prefinal-l's backward) only run on the objective's output frames
    flops_skipped = lib.kfp16_net_flops_skipped(net.ptr)
    step_flops = flops_fwd + flops_bwd - flops_skipped
    ms_step = ms / args.steps
    achieved = step_flops / ms_step / 1e9        # TFLOP/s over the WHOLE timed step (GEMMs, epilogues, elementwise, SGD)
    gemm_ms = g_ms.value / max(args.profile_steps, 1)

    if rank == 0:
        out = {
            "impl": "ours", "metric": metric_name(args.workload), "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": wl["desc"] if args.scaling == "weak" else wl["desc"].replace("64 seqs x 150 frames per GPU", f"64 seqs x 150 frames in total, {N_SEQ} per GPU"),
                       "frames_per_gpu_step": T, "global_frames_per_step": frames_per_step,
                       "parallelism": f"dp{world}" + ((f", FP16 gradient bucket summed by the library's kernel over NVLink peer memory in {n_seg} parts, the first beside the conv front end's backward pass ({peer_threads}-thread CTAs on every SM)" if exchange == "peer" else f", FP16 gradient all-reduce in {n_seg} buckets, the first overlapped with the conv front end's backward pass ({nccl_sms} SMs left to NCCL there)") if overlap else ((", FP16 gradient bucket summed by one kernel per rank over NVLink peer memory (kfp16_peer_allreduce_f16)" if exchange == "peer" else ", FP16 gradient all-reduce (NCCL)") if world > 1 else "")),
                       "l2": "per-step working set (activations + gradients, > 1 GB) exceeds the 126 MB L2; no explicit flush",
                       "loss": ("chain LF-MMI (log-semiring numerator / denominator forward-backward, 50 output frames per sequence, one batched launch)"
                                if objective == "chain" else "0.5*||out||^2, dY=Y"),
                       "optimizer": f"momentum SGD on FP32 masters, lr {lr}, m 0.9, grads scaled by {grad_scale:.3g}",
                       "launch": "CUDA graphs (step, SGD) with programmatic dependent launch between kernels" + ("" if os.environ.get("KFP16_PDL", "1") != "0" else " OFF")},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "event_ms": e2e_event_ms, "wall_ms": wall_ms, "input": "FP32 features in pinned host memory, converted to FP16 on the device",
                    "first_loss": first_loss, "last_loss": last_loss},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "kfp16::gemm_f16_sm100 (all tile variants; whole step timed)",
                         "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                         "how": "algorithmic 2*M*N*K of the GEMMs the step EXECUTES (forward incl. xent branch + backward; rows outside the chain objective's output frames are not computed behind the last splicing layer and not counted) / ms_per_step of the timed region",
                         "flops_all_rows": flops_fwd + flops_bwd, "flops_skipped_rows": flops_skipped,
                         "frac_all_rows": (flops_fwd + flops_bwd) / ms_step / 1e9 / burst,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops burst ({src}); sustained {sustained}",
                         "frac_of_sustained": achieved / sustained, "frac_of_nominal_2250": achieved / 2250.0,
                         "flops_per_step": step_flops, "flops_forward": flops_fwd, "flops_backward": flops_bwd,
                         "traffic": traffic, "traffic_source": traffic_src,
                         # side table from an eager, event-bracketed replay (each GEMM launch timed alone; their sum can
                         # exceed the graph step, which overlaps consecutive launches through programmatic dependent launch)
                         "gemm_launches_per_step": n_l.value // max(args.profile_steps, 1),
                         "gemm_ms_per_step_eager": gemm_ms,
                         "gemm_tflops_eager": g_fl.value / max(g_ms.value, 1e-9) / 1e9,
                         "eager_replay_ms_per_step": profile_ms},
        }
        if ranks_identical is not None:
            out["ranks_identical_weights"] = ranks_identical
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(args.workload)
            except Exception as e:  # noqa: BLE001 - the GPU number must still be reported
                out["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(out), flush=True)
    net.Free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
