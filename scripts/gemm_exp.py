"""GEMM micro-experiments on the TDNN-F shapes: time one launch configuration many times.
usage: python scripts/gemm_exp.py NAME [key=val ...]   keys: cg, share(0/1), bn, split, ctas, flush(0/1), iters, pf
shapes: F1 (bott = [X(t-3)|X(t)] Wlin), F2 (affine + epilogue), F2p (affine, plain), B2, B4, W3 (dWaff), W5 (dWlin)"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kaldi_fp16_b200 import _lib, cudart, gpu  # noqa: E402
from kaldi_fp16_b200._lib import EPI_BIAS, EPI_BN, EPI_MASK, EPI_RELU, EPI_RESID, GemmDesc, K_MAJOR, MN_MAJOR  # noqa: E402

name = sys.argv[1]
kv = dict(a.split("=") for a in sys.argv[2:])
cg, share, bn = int(kv.get("cg", 0)), int(kv.get("share", 1)), int(kv.get("bn", 0))
split, ctas, flush, iters = int(kv.get("split", 0)), int(kv.get("ctas", 0)), int(kv.get("flush", 0)), int(kv.get("iters", 30))
if "pf" in kv:
    os.environ["KFP16_PF"] = kv["pf"]
lib = _lib.load()
gpu.Init(0)
h = gpu.NewHandle()
rng = np.random.default_rng(0)
T, H, Bt, s, halo = 9984, 1536, 160, 3, 3


def dev(rows, cols, scale=0.05):
    t = gpu.NewTensor(rows, cols)
    noise = (rng.standard_normal((min(rows, 2048), cols)) * scale).astype(np.float16)
    for r0 in range(0, rows, 2048):
        n = min(2048, rows - r0)
        assert lib.bridge_transfer_fp16(t.Ptr + r0 * cols * 2, noise.ctypes.data, n * cols) == 0
    return t


X, Bo, Y, dZ = dev(T + 2 * halo, H), dev(T + 2 * halo, Bt), dev(T, H), dev(T + 2 * halo, H)
Wlin, Waff = dev(2 * H, Bt), dev(2 * Bt, H)
bias = dev(1, H)
sc, sh = gpu.DeviceF32(np.ones(H, np.float32)), gpu.DeviceF32(np.zeros(H, np.float32))
mask = gpu.DeviceF32(n=T * (H // 32))
ws = gpu.DeviceF32(n=2 * H * Bt * 2)
d = GemmDesc()
d.groups, d.kslabs, d.alpha = 1, 1, 1.0
d.force_bn, d.force_cg, d.no_share, d.split_k = bn, cg, 0 if share else 1, split


def setA(t, rows, cols, hl=0):
    d.A.ptr, d.A.rows, d.A.cols, d.A.ld, d.A.halo = t.Ptr + hl * cols * 2, rows, cols, cols, hl


def setB(t, rows, cols):
    d.B.ptr, d.B.rows, d.B.cols, d.B.ld, d.B.halo = t.Ptr, rows, cols, cols, 0


if name in ("F1", "B2"):
    K1 = H
    d.M, d.N, d.K, d.kslabs, d.kslab_len = T, Bt, 2 * K1, 2, K1
    setA(X if name == "F1" else dZ, T, H, halo)
    if name == "F1":
        d.a_major, d.b_major = K_MAJOR, MN_MAJOR
        setB(Wlin, 2 * H, Bt)
        d.a_row_off[0][0], d.a_row_off[0][1], d.b_row_off[0][0], d.b_row_off[0][1] = -s, 0, 0, H
    else:
        d.a_major, d.b_major = K_MAJOR, K_MAJOR
        setB(Waff, 2 * Bt, H)
        d.a_row_off[0][0], d.a_row_off[0][1], d.b_row_off[0][0], d.b_row_off[0][1] = 0, -s, 0, Bt
    d.D[0], d.ldd = Bo.Ptr + halo * Bt * 2, Bt
elif name in ("F2", "F2p", "B4"):
    d.M, d.N, d.K, d.kslabs, d.kslab_len = T, H, 2 * Bt, 2, Bt
    setA(Bo, T, Bt, halo)
    if name == "B4":
        d.a_major, d.b_major = K_MAJOR, K_MAJOR
        setB(Wlin, 2 * H, Bt)
        d.a_row_off[0][0], d.a_row_off[0][1], d.b_row_off[0][0], d.b_row_off[0][1] = s, 0, 0, H
        d.flags, d.res_scale, d.ldr = EPI_RESID, 0.66, H
        d.R[0] = dZ.Ptr + halo * H * 2
    else:
        d.a_major, d.b_major = K_MAJOR, MN_MAJOR
        setB(Waff, 2 * Bt, H)
        d.a_row_off[0][0], d.a_row_off[0][1], d.b_row_off[0][0], d.b_row_off[0][1] = 0, s, 0, Bt
        if name == "F2":
            d.flags = int(kv["flags"], 0) if "flags" in kv else (EPI_BIAS | EPI_RELU | EPI_BN | EPI_MASK | EPI_RESID)
            d.bias, d.bn_scale, d.bn_shift, d.mask_out, d.mask_ld = bias.Ptr, sc.Ptr, sh.Ptr, mask.Ptr, H // 32
            d.res_scale, d.ldr = 0.66, H
            d.R[0] = X.Ptr + halo * H * 2
    d.D[0], d.ldd = Y.Ptr, H
elif name in ("W3", "W5"):
    d.a_major, d.b_major = MN_MAJOR, MN_MAJOR
    d.groups = 2
    if name == "W3":       # dWaff[2*160 x 1536] = [B(t) | B(t+s)]^T dZ
        d.M, d.N, d.K, d.kslab_len = Bt, H, T, T
        setA(Bo, T, Bt, halo); setB(dZ, T, H)
        d.a_row_off[0][0], d.a_row_off[1][0] = 0, s
        d.ws[0], d.ws[1], d.ws_ld = ws.Ptr, ws.Ptr + Bt * H * 4, H
    else:                  # dWlin[2*1536 x 160] = [X(t-s) | X(t)]^T dB
        d.M, d.N, d.K, d.kslab_len = H, Bt, T, T
        setA(X, T, H, halo); setB(Bo, T, Bt)
        d.a_row_off[0][0], d.a_row_off[1][0] = -s, 0
        d.ws[0], d.ws[1], d.ws_ld = ws.Ptr, ws.Ptr + H * Bt * 4, Bt
    if not split:
        d.split_k = 13
elif name == "G":          # generic plain GEMM: m= n= k= bk=(0|1: B K-major)
    gm, gn, gk, bk = int(kv.get("m", 9984)), int(kv.get("n", 256)), int(kv.get("k", 3072)), int(kv.get("bk", 0))
    GA, GB, GD = dev(gm, gk), (dev(gn, gk) if bk else dev(gk, gn)), dev(gm, gn)
    d.M, d.N, d.K, d.kslab_len = gm, gn, gk, gk
    d.a_major, d.b_major = K_MAJOR, (K_MAJOR if bk else MN_MAJOR)
    setA(GA, gm, gk); setB(GB, gn if bk else gk, gk if bk else gn)
    d.D[0], d.ldd = GD.Ptr, gn
else:
    raise SystemExit("unknown shape " + name)

if ctas:
    lib.kfp16_ctx_set_max_ctas(h.ptr, ctas)
dbg = None
if int(kv.get("dbg", 0)):
    dbg = gpu.DeviceF32(n=148 * 3 * 8 * 16 * 2)
    cudart.memset(dbg.Ptr, 0, dbg.N * 4)
    d.debug_clock_buf = dbg.Ptr
fl = gpu.DeviceF32(n=64 * 1024 * 1024) if flush else None
e0, e1 = cudart.Event(), cudart.Event()
lib.kfp16_ctx_set_profile(h.ptr, 0)
if int(kv.get("prof", 0)):     # device-side kernel durations from CUPTI (no host launch overhead in the number)
    import torch
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.init()
    for _ in range(3):
        assert lib.kfp16_gemm_ex(h.ptr, C.byref(d)) == 0, _lib.last_error()
    cudart.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            assert lib.kfp16_gemm_ex(h.ptr, C.byref(d)) == 0, _lib.last_error()
        cudart.synchronize()
    durs = sorted(e.time_range.end - e.time_range.start for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA)
    print(f"{name:4s} {' '.join(sys.argv[2:]):40s} CUPTI kernel duration median {durs[len(durs) // 2]:.2f} us  min {durs[0]:.2f} us  (n={len(durs)})")
    raise SystemExit(0)
ts = []
for it in range(iters + 3):
    if fl is not None:
        cudart.memset(fl.Ptr, 0, fl.N * 4)
    e0.record()
    rc = lib.kfp16_gemm_ex(h.ptr, C.byref(d))
    assert rc == 0, _lib.last_error()
    e1.record()
    e1.synchronize()
    if it >= 3:
        ts.append(e0.elapsed_ms(e1) * 1e3)
if dbg is not None:
    st = dbg.ToHost().view(np.int64).reshape(148, 3, 8, 16)
    for cta in [int(x) for x in kv.get("ctas_dbg", "0,1,77").split(",")]:
        t0 = st[cta][st[cta] > 0].min() if (st[cta] > 0).any() else 0
        for role, rn in enumerate(["tma", "mma", "epi"]):
            for ti_ in range(8):
                row = st[cta, role, ti_]
                if (row > 0).any():
                    print(f"cta {cta:3d} {rn} tile {ti_}: " + " ".join(f"{(v - t0) if v > 0 else -1:6d}" for v in row))
flops = 2.0 * d.M * d.N * d.K * d.groups
med = float(np.median(ts))
print(f"{name:4s} {' '.join(sys.argv[2:]):40s} median {med:7.2f} us  min {min(ts):7.2f} us  {flops / med / 1e6:7.1f} TF")
