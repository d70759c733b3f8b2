"""One eager fwd + chain loss + bwd step of the full CNN-TDNN (the bench workload), for ncu:

    ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -o gpurun_out/prof python scripts/profile_cnn_tdnn_step.py 1
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from kaldi_fp16_b200 import _lib, gpu, nnet  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = _lib.load()
gpu.Init(0)
h = gpu.NewHandle()
T = bench.N_SEQ * bench.SEQ_LEN
net = nnet.NewNetwork(nnet.BuildModelFromString(bench.cnn_tdnn_xconfig()), h, bench.N_SEQ, bench.SEQ_LEN, train=True,
                      lr=bench.LR_CHAIN, grad_scale=1.0 / (bench.N_SEQ * bench.CHAIN_FRAMES))
rng = np.random.default_rng(1234)
feats = rng.standard_normal((T, 40)).astype(np.float32) * (10.0 * 0.9 ** np.arange(40, dtype=np.float32))
ivecs = np.clip(rng.standard_normal((bench.N_SEQ, 100)), -3, 3).astype(np.float32)
net.SetInputF32("input", feats)
net.SetInputF32("ivector", ivecs)
chain_obj = bench.build_synthetic_chain(h, 6016)
for _ in range(iters):
    net.ZeroGrads()
    assert lib.kfp16_net_forward(net.ptr) == 0
    assert lib.kfp16_net_loss_chain(net.ptr, b"", chain_obj.ptr, 3, 0, 1.0) == 0, _lib.last_error()
    assert lib.kfp16_net_backward(net.ptr) == 0
    net.SGDStep(1.0 / (bench.N_SEQ * bench.CHAIN_FRAMES))
gpu.Sync()
print("launches", lib.kfp16_launch_count())
