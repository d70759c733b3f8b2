import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from kaldi_fp16_b200 import gpu, nnet, _lib
import bench
gpu.Init(0); h = gpu.NewHandle()
net = nnet.NewNetwork(nnet.BuildModelFromString(bench.tdnnf_stack_xconfig()), h, 64, 150, lr=1e-4, grad_scale=1/9600)
rng = np.random.default_rng(0)
x = rng.standard_normal((9600, 1536)).astype(np.float32)
for step in range(6):
    net.ZeroGrads()
    out = net.Forward(x)
    print("step", step, "out absmax", np.abs(out).max(), "nan", np.isnan(out).sum(), "inf", np.isinf(out).sum())
    for name in ("tdnnf1", "tdnnf4", "tdnnf8", "tdnnf12"):
        o = net.Output(name); print("   ", name, np.abs(o).max(), np.isnan(o).sum())
    net.Backward(None)
    print("   loss", net.ReadLoss())
    wg = net.WeightGrads()
    print("   grad absmax", max(np.abs(g).max() for g in wg.values()), "nan", sum(np.isnan(g).sum() for g in wg.values()))
    for name in ("tdnnf15", "tdnnf8", "tdnnf1"):
        g = net.Grad(name); print("   d", name, np.abs(g).max(), np.isnan(g).sum(), np.isinf(g).sum())
    net.SGDStep(1/9600)
