"""One TDNN-F layer (1536 hidden, 160 bottleneck, stride 3, bypass) fwd+bwd, eager launches, for ncu:
6 GEMM launches per iteration (linear, affine, dB, dWaff, dX, dWlin) + the elementwise kernels.

    ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 12 -c 6 -o gpurun_out/prof \
        python scripts/profile_tdnnf_layer.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kaldi_fp16_b200 import _lib, gpu, nnet  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
xc = "input name=input dim=1536\n" + "".join(
    f"tdnnf-layer name=tdnnf{i + 1} dim=1536 bottleneck-dim=160 time-stride=3 bypass-scale=0.66\n" for i in range(layers))
gpu.Init(0)
h = gpu.NewHandle()
net = nnet.NewNetwork(nnet.BuildModelFromString(xc), h, 64, 150, train=True, lr=1e-4, grad_scale=1.0 / (9600 * 1536))
x = np.random.default_rng(0).standard_normal((9600, 1536)).astype(np.float32)
for _ in range(iters):
    net.ZeroGrads()
    net.Forward(x)
    net.Backward(None)
gpu.Sync()
print("launches", _lib.load().kfp16_launch_count())
