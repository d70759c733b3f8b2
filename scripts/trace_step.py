"""In-graph kernel timeline of the bench step (CUPTI through torch.profiler; no ncu, no serialisation).

    python scripts/trace_step.py [layers=16] [out=gpurun_out/trace_step.txt] [workload=tdnnf_stack|cnn_tdnn] [chain]

Replays the captured step graph a few times under the profiler and writes, for ONE step, every kernel in
launch order with its start offset, duration and the idle gap before it, plus totals per kernel name.
This is how the per-kernel shares quoted in DESIGN.md are obtained (ncu times are cold-cache / serialised).
"""
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from kaldi_fp16_b200 import _lib, cudart, gpu, nnet  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 16
out_path = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "trace_step.txt"
workload = sys.argv[3] if len(sys.argv) > 3 else "tdnnf_stack"

torch.cuda.init()
lib = _lib.load()
gpu.Init(0)
h = gpu.NewHandle()
st = cudart.Stream()
lib.kfp16_ctx_set_stream(h.ptr, st.ptr)
if workload == "tdnnf_stack":
    xc, fd, ivd, od = bench.tdnnf_stack_xconfig(layers=layers), 1536, 0, 1536
else:
    xc, fd, ivd, od = bench.cnn_tdnn_xconfig(), 40, 100, 6016
T = bench.N_SEQ * bench.SEQ_LEN
net = nnet.NewNetwork(nnet.BuildModelFromString(xc), h, bench.N_SEQ, bench.SEQ_LEN, train=True, lr=bench.LR,
                      grad_scale=1.0 / (T * od))
rng = np.random.default_rng(1234)
d_feat = gpu.TensorFromBits(nnet.rne_fp16_bits(rng.standard_normal((T, fd)).astype(np.float32)))
d_ivec = gpu.TensorFromBits(nnet.rne_fp16_bits(rng.standard_normal((bench.N_SEQ, 100)).astype(np.float32))) if ivd else None


def set_in():
    assert lib.kfp16_net_set_input_device(net.ptr, b"input", d_feat.Ptr, T, fd) == 0, _lib.last_error()
    if ivd:
        assert lib.kfp16_net_set_input_device(net.ptr, b"ivector", d_ivec.Ptr, bench.N_SEQ, ivd) == 0, _lib.last_error()


set_in()
if workload != "tdnnf_stack" and len(sys.argv) > 4 and sys.argv[4] == "chain":     # the bench's default objective for the CNN-TDNN
    chain_obj = bench.build_synthetic_chain(h, od)
    assert lib.kfp16_net_set_chain(net.ptr, chain_obj.ptr, 3, 0, 1.0) == 0, _lib.last_error()
net.Capture(1)
net.Capture(2)


def step():
    set_in()
    net.Launch(1)
    net.Launch(2)


for _ in range(5):
    step()
cudart.synchronize()
N = 4
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    cudart.synchronize()
    torch.cuda.synchronize()

ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
per = len(ev) // N
one = ev[per * (N - 2): per * (N - 1)]          # one warm step in the middle
t0 = one[0].time_range.start
lines, agg, prev_end = [], defaultdict(lambda: [0, 0.0]), None
busy = 0.0
for e in one:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = 0.0 if prev_end is None else e.time_range.start - prev_end
    prev_end = max(prev_end or 0, e.time_range.end)
    name = e.name
    if len(name) > 110:
        name = name[:110]
    lines.append(f"{s:10.1f} us  dur {d:8.2f}  gap {gap:7.2f}  {name}")
    a = agg[name]
    a[0] += 1
    a[1] += d
    busy += d
span = one[-1].time_range.end - t0
out_path.parent.mkdir(parents=True, exist_ok=True)
with open(out_path, "w") as f:
    f.write(f"# {workload} layers={layers}: one step, {len(one)} kernels, span {span:.1f} us, sum of kernel durations {busy:.1f} us\n")
    f.write("# totals per kernel (count, total us, avg us)\n")
    for name, (c, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"#  {c:4d}  {tot:9.1f}  {tot / c:8.2f}  {name}\n")
    f.write("\n".join(lines) + "\n")
print(f"step span {span:.1f} us, kernels {len(one)}, busy {busy:.1f} us -> {out_path}")
net.Free()
