"""Per-role clock stamps of the implicit-GEMM convolution kernel on one layer shape (debug_clock_buf of kfp16_gemm_ex):
where a CTA's time goes per tile -- TMA producer, MMA issuer, epilogue warp 4.
usage: python scripts/conv_tile_profile.py [T=9984] [H=40] [C=64] [N=64] [no_share=0]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from kaldi_fp16_b200 import _lib, cudart, gpu  # noqa: E402
from tests.util import make_desc, run_desc  # noqa: E402
from kaldi_fp16_b200._lib import EPI_BIAS, EPI_BN, EPI_MASK, EPI_RELU  # noqa: E402

kv = dict(a.split("=") for a in sys.argv[1:])
T, H, Cc, N, ns = (int(kv.get(k, d)) for k, d in (("T", 9984), ("H", 40), ("C", 64), ("N", 64), ("no_share", 0)))
lib = _lib.load()
gpu.Init(0)
h = gpu.NewHandle()
rng = np.random.default_rng(0)
taps = [(dt, dh) for dt in (-1, 0, 1) for dh in (-1, 0, 1)]
x = gpu.TensorFromFP16((rng.standard_normal((T, H * Cc)) * 0.5).astype(np.float16))
W = gpu.TensorFromFP16((rng.standard_normal((len(taps) * Cc, N)) * 0.05).astype(np.float16))
b = gpu.TensorFromFP16(np.zeros((1, N), np.float16))
D = gpu.ZeroTensor(T * H, N)
sc, sh = gpu.DeviceF32(np.ones(N, np.float32)), gpu.DeviceF32(np.zeros(N, np.float32))
mask_ld = (N + 31) // 32
mask = gpu.DeviceF32(n=T * H * mask_ld)
d = make_desc(T * H, N, len(taps) * Cc, x, W, D, flags=EPI_BIAS | EPI_RELU | EPI_BN | EPI_MASK, no_share=ns)
d.A.ptr = None
d.bias, d.bn_scale, d.bn_shift = b.Ptr, sc.Ptr, sh.Ptr
d.mask_out, d.mask_ld = mask.Ptr, mask_ld
c = d.conv
c.mode, c.x, c.T, c.H, c.P, c.C, c.rows_h, c.ntaps = 1, x.Ptr, T, H, 1, Cc, H, len(taps)
for i, (dt, dh) in enumerate(taps):
    c.dt[i], c.hq[i], c.par[i], c.brow[i] = dt, dh, 0, i * Cc
for _ in range(3):
    run_desc(h, d)
gpu.Sync()
e0, e1 = cudart.Event(), cudart.Event()
e0.record()
for _ in range(20):
    run_desc(h, d)
e1.record()
e1.synchronize()
us = e0.elapsed_ms(e1) / 20 * 1e3
print(f"T={T} H={H} C={Cc} N={N} no_share={ns}: {us:.1f} us, {2.0 * T * H * N * len(taps) * Cc / us / 1e6:.0f} TF/s")
GRID = 148
dbg = gpu.DeviceF32(n=GRID * 3 * 8 * 16 * 2)      # int64 stamps
d.debug_clock_buf = dbg.Ptr
run_desc(h, d)
gpu.Sync()
st = dbg.ToHost().view(np.int64).reshape(GRID, 3, 8, 16)
for cta in (0, 2, 100):
    s = st[cta]
    t0 = s[0, 7, 0]
    print(f"-- CTA {cta}: entry 0, prologue done {s[0,7,1]-t0}, predecessor {s[0,7,2]-t0}, last store issued {s[0,7,3]-t0}, all roles done {s[0,7,5]-t0}")
    for i in range(1, 6):
        prod = s[0, i]; mma = s[1, i]; epi = s[2, i]
        if mma[0] == 0:
            continue
        nxt = s[1, i + 1, 0] if s[1, i + 1, 0] else 0
        print(f"   tile {i}: producer start {prod[0]-t0:7d} first-load {prod[1]-prod[0]:5d} done {prod[2]-prod[0]:5d} | "
              f"mma start {mma[0]-t0:7d} wait-acc {mma[1]-mma[0]:5d} wait-first-operands {mma[2]-mma[1]:5d} issue {mma[3]-mma[2]:5d} period {nxt-mma[0] if nxt else 0:5d} | "
              f"epi start {epi[0]-t0:7d} wait-acc-full {epi[1]-epi[0]:5d} chunk0 {epi[3]-epi[2]:5d} total {epi[15]-epi[0]:5d}")
