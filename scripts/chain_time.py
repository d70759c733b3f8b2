"""Device time of the batched chain objective at the benchmark's shape (64 sequences x 50 output frames x 6016 pdfs,
256-state / 1024-arc denominator): python scripts/chain_time.py [general]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from kaldi_fp16_b200 import _lib, cudart, gpu  # noqa: E402

lib = _lib.load()
gpu.Init(0)
h = gpu.NewHandle()
st = cudart.Stream()
lib.kfp16_ctx_set_stream(h.ptr, st.ptr)
P, n_seq, rows = 6016, bench.N_SEQ, 156
obj = bench.build_synthetic_chain(h, P)
if len(sys.argv) > 1 and sys.argv[1] == "general":
    lib.kfp16_chain_force_general(obj.ptr, 1)
rng = np.random.default_rng(0)
out = gpu.TensorFromFP16((rng.standard_normal((n_seq * rows, P)) * 0.5).astype(np.float32))
grad = gpu.ZeroTensor(n_seq * rows, P)
acc = gpu.DeviceF32(n=4)
for _ in range(3):
    assert lib.kfp16_chain_loss(obj.ptr, out.Ptr, grad.Ptr, P, rows, 3, 3, 1.0, acc.Ptr) == 0, _lib.last_error()
e0, e1 = cudart.Event(), cudart.Event()
e0.record(st.ptr)
N = 20
for _ in range(N):
    lib.kfp16_chain_loss(obj.ptr, out.Ptr, grad.Ptr, P, rows, 3, 3, 1.0, acc.Ptr)
e1.record(st.ptr)
e1.synchronize()
dbg = gpu.DeviceF32(n=32)
lib.kfp16_chain_set_debug(obj.ptr, dbg.Ptr)
lib.kfp16_chain_loss(obj.ptr, out.Ptr, grad.Ptr, P, rows, 3, 3, 1.0, acc.Ptr)
cudart.synchronize()
stamps = dbg.ToHost().view(np.int64)[:5]
print("phase cycles (setup+gather, forward loop, totals+reload, backward loop):", np.diff(stamps))
print("backward frame 10 (arc pass, barrier, state pass, gradient row, alpha staging, barrier):", np.diff(dbg.ToHost().view(np.int64)[8:15]))
e0.record(st.ptr)
for _ in range(N):
    lib.kfp16_chain_loss(obj.ptr, out.Ptr, grad.Ptr, P, rows, 3, 3, 1.0, acc.Ptr)
e1.record(st.ptr)
e1.synchronize()
print(f"chain objective: {e0.elapsed_ms(e1) / N * 1e3:.1f} us per minibatch ({'global' if len(sys.argv) > 1 else 'shared-memory'} kernel)")
