"""CPU oracle for the network level of the hot path -- TEST INFRASTRUCTURE ONLY (see kaldi_oracle.py).

numpy restatement of the reference's Go executor: xconfig parsing (internal/nnet/xconfig.go:143,
layers.go:126-355), per-layer forward op sequences (internal/nnet/forward.go:317-1001) and the
backward pass as the exact transpose of the forward (right column; the
reference's own backward ignores splicing / im2col -- quirk Q2 -- and agrees with this for
time-stride 0 nets).  Every intermediate the reference stores as FP16 is rounded here with h().

Deliberate, flagged deviations from reference quirks: true spliced weight shapes (Q1),
per-sequence clamping (Q3; n_seq=1 gives the reference's whole-minibatch clamp), Cartesian conv
taps (Q4; cartesian=False gives the paired taps), ivector broadcast per sequence (Q6), eps = 1e-3 in
both directions (Q5), prefinal backward = transpose of its forward (Q8), conv activations stay
height-major [T x H*F] with batch-norm per filter (the reference's filter-major re-layout is not
consumed consistently by its next layer, forward.go:450 vs 503-513).

Pinning.  The Go executor itself cannot be run here (no Go toolchain, and it links Kaldi).  FORWARD: pinned on
the GPU box against the reference's own compiled operator library driven in forward.go's operator order, layer by
layer over a whole network (tests/test_network_vs_reference_ops_gpu.py; full-size layers:
tests/test_layer_vs_reference_gpu.py, tests/test_cnn_tdnn_fullsize_gpu.py), and every operator it composes is
pinned bit for bit (tests/test_ops_gpu.py, tests/test_gemm_gpu.py, tests/golden/ref_ops.npz).  BACKWARD: the
orchestration is unpinned -- the reference's own Network.Backward is inconsistent with its forward (quirks
Q2/Q8), so there is nothing to pin it to; the formulas are checked against an independent float64 autograd
(tests/test_nnet_oracle_cpu.py) and, per layer at full size, against the reference's backward OPERATORS
(tests/test_layer_vs_reference_gpu.py).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np

from . import kaldi_oracle as O

f32 = np.float32
h = O.h


@dataclass
class OLayer:
    type: str
    name: str
    kv: dict
    inputs: list = field(default_factory=list)
    replace: bool = False
    in_dim: int = 0
    out_dim: int = 0
    per_seq: bool = False


def _split_top(s: str, sep: str) -> list:
    out, cur, depth = [], "", 0
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def parse_xconfig(text: str) -> list:
    """xconfig.go:143 ParseXConfig + layers.go:126 ResolveLayers (dims only)."""
    layers: list[OLayer] = []
    by_name = {}
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        toks = [t for t in _split_top(line.replace("\t", " "), " ") if t.strip()]
        kv = dict(t.split("=", 1) for t in toks[1:] if "=" in t)
        l = OLayer(toks[0], kv["name"], kv)
        spec = kv.get("input", "").strip()
        if spec.startswith("Append("):
            for part in _split_top(spec[7:-1], ","):
                part = part.strip()
                if part.startswith("ReplaceIndex("):
                    part = part[13:].split(",")[0].strip()
                l.inputs.append(part)
        elif spec.startswith("ReplaceIndex("):
            l.inputs.append(spec[13:].split(",")[0].strip())
            l.replace = True
        elif spec:
            l.inputs.append(spec)
        elif l.type != "input":
            l.inputs.append(layers[-1].name)
        layers.append(l)
        by_name[l.name] = l
    # per-sequence inputs: consumed only via ReplaceIndex
    for l in layers:
        if l.type == "input":
            cons = [c for c in layers if l.name in c.inputs]
            l.per_seq = bool(cons) and all(c.replace for c in cons)
    for l in layers:
        l.in_dim = sum(by_name[i].out_dim for i in l.inputs)
        if l.type != "input":
            l.per_seq = bool(l.inputs) and all(by_name[i].per_seq for i in l.inputs)
        t = l.type
        gi = lambda k, d: int(l.kv.get(k, d))
        if t == "input":
            l.out_dim = gi("dim", 0)
        elif t in ("idct-layer",):
            l.out_dim = gi("dim", l.in_dim)
        elif t in ("linear-component", "output-layer", "tdnnf-layer"):
            l.out_dim = gi("dim", 0)
        elif t in ("batchnorm-component", "spec-augment-layer", "combine-feature-maps-layer"):
            l.out_dim = l.in_dim
        elif t == "conv-relu-batchnorm-layer":
            hin = gi("height-in", 0)
            l.out_dim = gi("height-out", hin) * gi("num-filters-out", 0)
        elif t == "prefinal-layer":
            l.out_dim = gi("small-dim", 0)
        elif t == "attention-relu-batchnorm-layer":
            # internal/nnet/layers.go:298-318: output = num_heads * (value_dim + context_dim)
            l.out_dim = gi("num-heads", 1) * (gi("value-dim", 0) + 1 + gi("num-left-inputs", 0) + gi("num-right-inputs", 0))
        else:
            raise ValueError(f"unsupported layer type {t}")
    return layers


def idct_matrix(dim: int, lifter: float) -> np.ndarray:
    """makeIDCTMatrix (forward.go:1190-1210), through the truncating converter (TensorFromFP32)."""
    m = np.zeros((dim, dim), dtype=np.float64)
    for i in range(dim):
        for j in range(dim):
            v = math.cos(math.pi * j * (i + 0.5) / dim) * (math.sqrt(1.0 / dim) if j == 0 else math.sqrt(2.0 / dim))
            if lifter > 0 and j > 0:
                v *= 1.0 + (lifter / 2.0) * math.sin(math.pi * j / lifter)
            m[i, j] = v
    return O.to_f16_trunc(m.astype(f32))


def identity_bn(dim):
    """identityBN (forward.go:1170-1187)"""
    return dict(mean=np.zeros(dim, f32), var=np.ones(dim, f32), gamma=np.ones(dim, f32), beta=np.zeros(dim, f32),
                eps=1e-3)


class OracleNet:
    def __init__(self, xconfig: str, n_seq: int, seq_len: int, cartesian: bool = True, train: bool = False,
                 dropout_seed: int = 0, spec_augment: bool = False, train_bn: bool = False, bn_momentum: float = 0.1,
                 stats_allreduce=None, world: int = 1):
        # train = True enables the tdnnf-layer dropout-proportion (inverted dropout after the batch-norm,
        # go/gotorch/layers.go:348-399); the mask of layer index i is dropout_uniform(seed ^ i*0x9E3779B9, padded row, col) > p
        self.train, self.dropout_seed = train, dropout_seed
        # spec_augment = True (training): spec-augment-layer zeroes frequency / time masks per sequence
        # (go/gotorch/cnn_tdnn.go:612-668) drawn from the same counter-based generator; False = the reference executor's
        # pass-through (internal/nnet/forward.go:377-383)
        self.spec_augment = spec_augment
        # train_bn = True (training): batch statistics instead of the stored running statistics
        # (cpp/cuda/cnn_kernels.cu:236-320 training branch, go/gotorch/layers.go:257-330): mean / biased variance over the rows of
        # the minibatch, running = (1-m)*running + m*batch, backward = gamma/sqrt(batch var + eps) without differentiating through
        # the statistics.  stats_allreduce(vector) sums [S1 | S2] over `world` data-parallel ranks.
        self.train_bn, self.bn_momentum, self.stats_allreduce, self.world = train_bn, f32(bn_momentum), stats_allreduce, world
        self.layers = parse_xconfig(xconfig)
        self.by_name = {l.name: l for l in self.layers}
        self.n_seq, self.L = n_seq, seq_len
        self.cartesian = cartesian
        self.params: dict[str, np.ndarray] = {}
        self.bn: dict[tuple, dict] = {}
        for l in self.layers:
            t = l.type
            if t == "linear-component":
                self.params[f"{l.name}.W"] = np.zeros((l.in_dim, l.out_dim), f32)
            elif t == "tdnnf-layer":
                s = int(l.kv.get("time-stride", 3))
                sp = 2 if s > 0 else 1
                bn = int(l.kv["bottleneck-dim"])
                self.params[f"{l.name}.LinearW"] = np.zeros((sp * l.in_dim, bn), f32)
                self.params[f"{l.name}.AffineW"] = np.zeros((sp * bn, l.out_dim), f32)
                self.params[f"{l.name}.AffineBias"] = np.zeros((1, l.out_dim), f32)
                self.bn[(l.name, "AffBN")] = identity_bn(l.out_dim)
            elif t == "prefinal-layer":
                big, small = int(l.kv["big-dim"]), int(l.kv["small-dim"])
                self.params[f"{l.name}.BigW"] = np.zeros((l.in_dim, big), f32)
                self.params[f"{l.name}.BigBias"] = np.zeros((1, big), f32)
                self.params[f"{l.name}.SmallW"] = np.zeros((big, small), f32)
                self.bn[(l.name, "PfBN")] = identity_bn(big)
                self.bn[(l.name, "BN")] = identity_bn(small)
            elif t == "output-layer":
                self.params[f"{l.name}.W"] = np.zeros((l.in_dim, l.out_dim), f32)
                self.params[f"{l.name}.Bias"] = np.zeros((1, l.out_dim), f32)
            elif t == "batchnorm-component":
                self.bn[(l.name, "")] = identity_bn(l.in_dim)
            elif t == "attention-relu-batchnorm-layer":
                g = self.attention_geom(l)
                # true shape of the projection (forward.go:803-812); the reference's random init allocates [in x out] (quirk Q1)
                self.params[f"{l.name}.W"] = np.zeros((l.in_dim, g["affine"]), f32)
                self.params[f"{l.name}.Bias"] = np.zeros((1, g["affine"]), f32)
                self.bn[(l.name, "BN")] = identity_bn(l.out_dim)
            elif t == "conv-relu-batchnorm-layer":
                taps = self.conv_taps(l)
                fin = l.in_dim // int(l.kv["height-in"])
                fout = int(l.kv["num-filters-out"])
                self.params[f"{l.name}.W"] = np.zeros((len(taps) * fin, fout), f32)
                self.params[f"{l.name}.Bias"] = np.zeros((1, fout), f32)
                self.bn[(l.name, "BN")] = identity_bn(fout)

    def attention_geom(self, l: OLayer) -> dict:
        """restricted self-attention geometry (internal/nnet/forward.go:795-812): per head the projection row holds
        [key (key-dim) | value (value-dim) | query key part (key-dim) | query context part (context)]"""
        gi = lambda k, d: int(l.kv.get(k, d))
        H, V, K = gi("num-heads", 1), gi("value-dim", 0), gi("key-dim", 0)
        nl, nr = gi("num-left-inputs", 0), gi("num-right-inputs", 0)
        C = 1 + nl + nr
        per = K + V + K + C
        ks = float(l.kv["key-scale"]) if "key-scale" in l.kv else 1.0 / math.sqrt(K)      # weight_loader.go:267-271
        return dict(H=H, V=V, K=K, nl=nl, nr=nr, C=C, per=per, affine=H * per, stride=gi("time-stride", 1), ks=f32(ks))

    def _ctx_rows(self, a, shift):
        """a: [n_seq, L, ...]; result[t] = a[t + shift], zeros outside the sequence (the reference pads the whole minibatch
        matrix with zeros, forward.go:833-845: per sequence here, quirk Q3)"""
        out = np.zeros_like(a)
        L = a.shape[1]
        lo, hi = max(0, -shift), min(L, L - shift)
        if hi > lo:
            out[:, lo:hi] = a[:, lo + shift:hi + shift]
        return out

    def _attention_fwd(self, l, x, P, saved):
        """forward.go:795-909: affine projection -> per head and frame: b[o] = q_ctx[o] + key_scale * <q_key, key[t + (o-nl)*s]>,
        w = softmax(b), out = [sum_o w[o] * value[t + (o-nl)*s] | w] -> ReLU -> batch-norm (the attention itself in fp32)"""
        g = self.attention_geom(l)
        H, V, K, C, per = g["H"], g["V"], g["K"], g["C"], g["per"]
        proj = O.add_bias(O.gemm(x, P[f"{l.name}.W"]), P[f"{l.name}.Bias"])
        p4 = proj.reshape(self.n_seq, self.L, H, per).astype(f32)
        key, val, qk, qc = p4[..., :K], p4[..., K:K + V], p4[..., K + V:2 * K + V], p4[..., 2 * K + V:]
        keys = [self._ctx_rows(key, (o - g["nl"]) * g["stride"]) for o in range(C)]
        vals = [self._ctx_rows(val, (o - g["nl"]) * g["stride"]) for o in range(C)]
        b = np.stack([qc[..., o] + g["ks"] * np.sum(qk * keys[o], axis=-1, dtype=f32) for o in range(C)], axis=-1).astype(f32)
        e = np.exp(b - b.max(axis=-1, keepdims=True)).astype(f32)
        w = (e / e.sum(axis=-1, keepdims=True, dtype=f32)).astype(f32)
        u = np.zeros(p4.shape[:3] + (V,), f32)
        for o in range(C):
            u += w[..., o:o + 1] * vals[o]
        out = np.concatenate([u, w], axis=-1).reshape(self.n_seq * self.L, H * (V + C))
        z = h(np.maximum(out, f32(0)))
        saved[l.name].update(proj=proj, w=w, keys=keys, vals=vals, qk=qk, mask=z > 0, geom=g)
        return self._bn_fwd(z, self.bn[(l.name, "BN")])

    def _attention_bwd(self, l, dy, sv, P, wg):
        """exact transpose of _attention_fwd (the reference back-propagates the layer as a plain affine + ReLU + BN,
        network_backward.go:539-545: quirk Q2)"""
        g = sv["geom"]
        H, V, K, C, per = g["H"], g["V"], g["K"], g["C"], g["per"]
        dz = np.where(sv["mask"], h(dy * self._bn_scale(self.bn[(l.name, "BN")])), f32(0)).reshape(self.n_seq, self.L, H, V + C)
        dU, dWt = dz[..., :V], dz[..., V:]
        w = sv["w"]
        gw = np.stack([dWt[..., o] + np.sum(dU * sv["vals"][o], axis=-1, dtype=f32) for o in range(C)], axis=-1).astype(f32)
        db = (w * (gw - np.sum(w * gw, axis=-1, keepdims=True, dtype=f32))).astype(f32)
        dp = np.zeros((self.n_seq, self.L, H, per), f32)
        dp[..., 2 * K + V:] = db
        for o in range(C):
            shift = (o - g["nl"]) * g["stride"]
            dp[..., K + V:2 * K + V] += g["ks"] * db[..., o:o + 1] * sv["keys"][o]
            # rows t + shift received key / value gradients from query frame t: the adjoint of _ctx_rows is the opposite shift
            dp[..., :K] += self._ctx_rows(g["ks"] * db[..., o:o + 1] * sv["qk"], -shift)
            dp[..., K:K + V] += self._ctx_rows(w[..., o:o + 1] * dU, -shift)
        dproj = h(dp.reshape(self.n_seq * self.L, H * per))
        x = sv["x"]
        wg[f"{l.name}.W"] = x.T.astype(f32) @ dproj
        wg[f"{l.name}.Bias"] = dproj.sum(0, dtype=f32).reshape(1, -1)
        return O.gemm(dproj, P[f"{l.name}.W"], transB=True)

    def conv_taps(self, l: OLayer) -> list:
        to = [int(v) for v in l.kv.get("time-offsets", "0").split(",")]
        ho = [int(v) for v in l.kv.get("height-offsets", "0").split(",")]
        if self.cartesian:
            return [(dt, dh) for dt in to for dh in ho]
        return list(zip(to, ho))   # quirk Q4: paired (forward.go:426-446)

    def init_random(self, rng: np.random.Generator):
        """randTensor (forward.go:1161-1168): N(0,1)*sqrt(2/(rows+cols)) through the truncating
        converter; biases 0; identity BN."""
        for k, w in self.params.items():
            if w.shape[0] == 1 and k.endswith("Bias"):
                continue
            r, c = w.shape
            self.params[k] = O.to_f16_trunc((rng.standard_normal((r, c)) * math.sqrt(2.0 / (r + c))).astype(f32))

    # ------------------------------------------------------------------ helpers
    def halo(self) -> int:
        """halo rows per sequence side in the executor's padded layout: the largest time-stride / conv time offset"""
        hmax = 0
        for l in self.layers:
            if l.type == "tdnnf-layer":
                hmax = max(hmax, int(l.kv.get("time-stride", 3)))
            elif l.type == "conv-relu-batchnorm-layer":
                hmax = max([hmax] + [abs(int(v)) for v in l.kv.get("time-offsets", "0").split(",")])
        return hmax

    def dropout_keep(self, l, p: float) -> np.ndarray:
        """keep mask [T x out_dim] of layer l: rows are numbered as the executor stores them (padded layout)"""
        idx = self.layers.index(l)
        hl = self.halo()
        blk = self.L + 2 * hl
        rows = (np.arange(self.n_seq)[:, None] * blk + hl + np.arange(self.L)[None, :]).reshape(-1)
        seed = (self.dropout_seed ^ ((idx * 0x9E3779B9) & 0xFFFFFFFF)) & 0xFFFFFFFF
        return O.dropout_uniform(seed, rows, np.arange(l.out_dim)) > f32(p)

    def spec_augment_keep(self, l) -> np.ndarray:
        """keep mask [n_seq*L x dim] of a spec-augment-layer: the draw order and mask geometry of
        kaldi_fp16_b200/csrc/elementwise.cu::spec_augment_kernel / nnet.cu (L_SPECAUG)"""
        idx = self.layers.index(l)
        dim, L = l.out_dim, self.L
        fprop, tprop = float(l.kv.get("freq-max-proportion", 0.5)), float(l.kv.get("time-zeroed-proportion", 0.0))
        tmax = max(0, min(int(l.kv.get("time-mask-max-frames", 20)), L))
        fmax = max(0, min(dim, int(fprop * dim)))
        nfreq = 1 if fmax > 0 else 0
        ntime = min(8, max(1, int(tprop * L / (0.5 * tmax) + 0.5))) if (tprop > 0 and tmax > 0) else 0
        seed = (self.dropout_seed ^ ((idx * 0x9E3779B9) & 0xFFFFFFFF)) & 0xFFFFFFFF
        u = O.dropout_uniform(seed, np.arange(self.n_seq), np.arange(32))          # [n_seq x draws]
        keep = np.ones((self.n_seq, L, dim), bool)
        for m in range(8):
            f = (u[:, 4 * m] * f32(fmax + 1)).astype(np.int64) if m < nfreq else np.zeros(self.n_seq, np.int64)
            f0 = (u[:, 4 * m + 1] * (dim - f + 1).astype(f32)).astype(np.int64)
            t = (u[:, 4 * m + 2] * f32(tmax + 1)).astype(np.int64) if m < ntime else np.zeros(self.n_seq, np.int64)
            t0 = (u[:, 4 * m + 3] * (L - t + 1).astype(f32)).astype(np.int64)
            for s_ in range(self.n_seq):
                keep[s_, :, f0[s_]:f0[s_] + f[s_]] = False
                keep[s_, t0[s_]:t0[s_] + t[s_], :] = False
        return keep.reshape(self.n_seq * L, dim)

    def _shift(self, x, s):
        """rows t -> t+s inside each sequence, clamped at the sequence edges (forward.go:699-790)"""
        D = x.shape[1]
        x3 = x.reshape(self.n_seq, self.L, D)
        idx = np.clip(np.arange(self.L) + s, 0, self.L - 1)
        return x3[:, idx, :].reshape(-1, D)

    def _shift_T(self, g, s):
        """adjoint of _shift"""
        D = g.shape[1]
        g3 = g.reshape(self.n_seq, self.L, D)
        idx = np.clip(np.arange(self.L) + s, 0, self.L - 1)
        out = np.zeros_like(g3)
        np.add.at(out, (slice(None), idx, slice(None)), g3)
        return out.reshape(-1, D)

    def _bn_fwd(self, x, bn, rms=None):
        if self.train_bn and self.train and x.shape[0] == self.n_seq * self.L * max(1, x.shape[0] // (self.n_seq * self.L)):
            x32 = np.asarray(x, f32)
            stats = np.concatenate([x32.sum(0, dtype=f32), (x32 * x32).sum(0, dtype=f32)]).astype(f32)
            if self.stats_allreduce is not None:
                stats = np.asarray(self.stats_allreduce(stats), f32)
            n = f32(x.shape[0] * self.world)
            d = x.shape[1]
            mean = stats[:d] / n
            var = np.maximum(stats[d:] / n - mean * mean, f32(0))
            m = self.bn_momentum
            bn["mean"] = (bn["mean"] * (f32(1) - m) + mean * m).astype(f32)      # running statistics
            bn["var"] = (bn["var"] * (f32(1) - m) + var * m).astype(f32)
            bn["cur_mean"], bn["cur_var"] = mean.astype(f32), var.astype(f32)
            g = bn["gamma"] if rms is None else f32(rms)
            b = bn["beta"] if rms is None else f32(0)
            sc = (g / np.sqrt(var + f32(bn["eps"]))).astype(f32)
            return h(x32 * sc + (b - mean * sc).astype(f32))
        if rms is not None:
            return O.batchnorm_forward_rms(x, bn["mean"], bn["var"], rms, bn["eps"])
        return O.batchnorm_forward(x, bn["mean"], bn["var"], bn["gamma"], bn["beta"], bn["eps"])

    def _bn_scale(self, bn, rms=None):
        g = bn["gamma"] if rms is None else f32(rms)
        var = bn["cur_var"] if (self.train_bn and self.train and "cur_var" in bn) else bn["var"]
        return (g / np.sqrt(var + f32(bn["eps"]))).astype(f32)

    def _input_of(self, l, acts):
        parts = []
        for name in l.inputs:
            a = acts[name]
            if self.by_name[name].per_seq and not l.per_seq:
                a = np.repeat(a, self.L, axis=0)   # broadcast per sequence (Q6 fix)
            parts.append(a)
        return parts[0] if len(parts) == 1 else np.concatenate(parts, axis=1)

    def _patches(self, l, x):
        """im2col (forward.go:429-455): P[(t*Hout+ho), tap*Fin+f] = X[t+dt, (ho*sub+dh)*Fin+f], zero outside"""
        hin = int(l.kv["height-in"])
        hout = int(l.kv.get("height-out", hin))
        sub = int(l.kv.get("height-subsample-out", 1))
        fin = l.in_dim // hin
        taps = self.conv_taps(l)
        x4 = x.reshape(self.n_seq, self.L, hin, fin)
        P = np.zeros((self.n_seq, self.L, hout, len(taps), fin), f32)
        for k, (dt, dh) in enumerate(taps):
            for ho in range(hout):
                hs = ho * sub + dh
                if hs < 0 or hs >= hin:
                    continue
                t0, t1 = max(0, -dt), min(self.L, self.L - dt)
                P[:, t0:t1, ho, k, :] = x4[:, t0 + dt:t1 + dt, hs, :]
        return P.reshape(self.n_seq * self.L * hout, len(taps) * fin), (hin, hout, sub, fin, taps)

    def _patches_T(self, l, dP, geom):
        hin, hout, sub, fin, taps = geom
        dP5 = dP.reshape(self.n_seq, self.L, hout, len(taps), fin)
        dx = np.zeros((self.n_seq, self.L, hin, fin), f32)
        for k, (dt, dh) in enumerate(taps):
            for ho in range(hout):
                hs = ho * sub + dh
                if hs < 0 or hs >= hin:
                    continue
                t0, t1 = max(0, -dt), min(self.L, self.L - dt)
                dx[:, t0 + dt:t1 + dt, hs, :] += dP5[:, t0:t1, ho, k, :]
        return dx.reshape(self.n_seq * self.L, hin * fin)

    # ------------------------------------------------------------------ forward
    def forward(self, inputs: dict) -> dict:
        """inputs: name -> fp16-representable float32 dense rows.  Returns all activations."""
        acts, saved = {}, {}
        P = self.params
        for l in self.layers:
            t = l.type
            if t == "input":
                acts[l.name] = np.asarray(inputs[l.name], f32)
                continue
            x = self._input_of(l, acts)
            saved[l.name] = {"x": x}
            if t == "idct-layer":
                y = O.gemm(x, idct_matrix(l.out_dim, float(l.kv.get("cepstral-lifter", 22))))
            elif t == "linear-component":
                y = O.gemm(x, P[f"{l.name}.W"])
            elif t == "batchnorm-component":
                rms = float(l.kv.get("target-rms", 1.0))
                y = self._bn_fwd(x, self.bn[(l.name, "")], rms if rms != 1.0 else None)
            elif t == "spec-augment-layer":
                y = x.copy()
                if self.spec_augment and self.train:
                    keep = self.spec_augment_keep(l)
                    saved[l.name].update(keep=keep)
                    y = np.where(keep, y, f32(0))
            elif t == "combine-feature-maps-layer":
                y = O.combine_feature_maps(x, int(l.kv["height"]), int(l.kv.get("num-filters1", 1)), int(l.kv.get("num-filters2", 1)))
            elif t == "tdnnf-layer":
                s = int(l.kv.get("time-stride", 3))
                s1 = np.concatenate([self._shift(x, -s), x], 1) if s > 0 else x
                b = O.gemm(s1, P[f"{l.name}.LinearW"])
                s2 = np.concatenate([b, self._shift(b, s)], 1) if s > 0 else b
                z = O.gemm(s2, P[f"{l.name}.AffineW"])
                z = O.add_bias(z, P[f"{l.name}.AffineBias"])
                z = O.relu(z)
                relu_out = z
                z = self._bn_fwd(z, self.bn[(l.name, "AffBN")])
                mask = relu_out > 0
                pdrop = float(l.kv.get("dropout-proportion", 0.0)) if self.train else 0.0
                if pdrop > 0:
                    keep = self.dropout_keep(l, pdrop)
                    z = O.dropout_forward(z, keep, pdrop)
                    mask = mask & keep                      # gradient gate: ReLU active AND kept
                bypass = float(l.kv.get("bypass-scale", 0.66))
                y = O.add_scaled(z, x, bypass, 1.0) if (bypass > 0 and l.in_dim == l.out_dim) else z
                saved[l.name].update(s1=s1, s2=s2, mask=mask, pdrop=pdrop)
            elif t == "prefinal-layer":
                g = O.gemm(x, P[f"{l.name}.BigW"])
                g = O.add_bias(g, P[f"{l.name}.BigBias"])
                g = O.relu(g)
                mask = g > 0
                g = self._bn_fwd(g, self.bn[(l.name, "PfBN")])
                y = O.gemm(g, P[f"{l.name}.SmallW"])
                y = self._bn_fwd(y, self.bn[(l.name, "BN")])
                saved[l.name].update(g=g, mask=mask)
            elif t == "attention-relu-batchnorm-layer":
                y = self._attention_fwd(l, x, P, saved)
            elif t == "output-layer":
                y = O.gemm(x, P[f"{l.name}.W"])
                y = O.add_bias(y, P[f"{l.name}.Bias"])
                if l.kv.get("include-log-softmax", "true").lower() in ("true", "1", "yes"):
                    y = O.log_softmax(y)
            elif t == "conv-relu-batchnorm-layer":
                Pm, geom = self._patches(l, x)
                z = O.gemm(Pm, P[f"{l.name}.W"])
                z = O.add_bias(z, P[f"{l.name}.Bias"])
                z = O.relu(z)
                mask = z > 0
                z = self._bn_fwd(z, self.bn[(l.name, "BN")])          # per filter
                hout = geom[1]
                y = z.reshape(self.n_seq * self.L, hout * z.shape[1])  # stays height-major [T x H*F]
                saved[l.name].update(P=Pm, geom=geom, mask=mask)
            else:
                raise ValueError(t)
            acts[l.name] = y
        self.acts, self.saved = acts, saved
        return acts

    # ------------------------------------------------------------------ backward
    def backward(self, out_name: str, d_out: np.ndarray, masks: dict | None = None) -> tuple[dict, dict]:
        """returns (weight grads fp32 un-rounded accumulate of fp16 operands, activation grads).
        masks: optional layer -> bool ReLU mask overriding the oracle's own (a mask element that
        flips because a pre-activation is within rounding of 0 changes a gradient element by its
        full value; tests pass the kernel's masks and bound the flip rate separately)."""
        P = self.params
        if masks:
            for k, m in masks.items():
                self.saved[k]["mask"] = np.asarray(m, bool).reshape(self.saved[k]["mask"].shape)
        dact = {out_name: np.asarray(d_out, f32)}
        wg = {}
        for l in reversed(self.layers):
            if l.name not in dact or l.type == "input":
                continue
            dy = dact[l.name]
            sv = self.saved[l.name]
            x = sv["x"]
            t = l.type
            dx = None
            if t == "idct-layer":
                dx = O.gemm(dy, idct_matrix(l.out_dim, float(l.kv.get("cepstral-lifter", 22))), transB=True)
            elif t == "linear-component":
                dx = O.gemm(dy, P[f"{l.name}.W"], transB=True)
                wg[f"{l.name}.W"] = x.T.astype(f32) @ dy
            elif t == "batchnorm-component":
                rms = float(l.kv.get("target-rms", 1.0))
                dx = h(dy * self._bn_scale(self.bn[(l.name, "")], rms if rms != 1.0 else None))
            elif t == "spec-augment-layer":
                dx = dy.copy()
                if self.spec_augment and self.train:
                    dx = np.where(self.saved[l.name]["keep"], dx, f32(0))
            elif t == "combine-feature-maps-layer":
                H, f1, f2 = int(l.kv["height"]), int(l.kv.get("num-filters1", 1)), int(l.kv.get("num-filters2", 1))
                g3 = dy.reshape(-1, H, f1 + f2)
                dx = np.concatenate([g3[:, :, :f1].reshape(-1, H * f1), g3[:, :, f1:].reshape(-1, H * f2)], 1)
            elif t == "tdnnf-layer":
                s = int(l.kv.get("time-stride", 3))
                bnd = int(l.kv["bottleneck-dim"])
                bscale = self._bn_scale(self.bn[(l.name, "AffBN")])
                if sv.get("pdrop", 0.0) > 0:               # DropoutLayer.Backward folded into the batch-norm factor
                    bscale = (bscale * f32(1.0 / (1.0 - sv["pdrop"]))).astype(f32)
                dz = np.where(sv["mask"], h(dy * bscale), f32(0))
                wg[f"{l.name}.AffineBias"] = dz.sum(0, dtype=f32).reshape(1, -1)
                wg[f"{l.name}.AffineW"] = sv["s2"].T.astype(f32) @ dz
                Wa, Wl = P[f"{l.name}.AffineW"], P[f"{l.name}.LinearW"]
                if s > 0:
                    # one fused GEMM in the kernel: fp32 sum of both halves, one fp16 store
                    db = h(dz @ Wa[:bnd].T + self._shift_T(dz @ Wa[bnd:].T, s))
                else:
                    db = O.gemm(dz, Wa, transB=True)
                wg[f"{l.name}.LinearW"] = sv["s1"].T.astype(f32) @ db
                bypass = float(l.kv.get("bypass-scale", 0.66))
                use_bp = bypass > 0 and l.in_dim == l.out_dim
                if s > 0:
                    acc = self._shift_T(db @ Wl[:l.in_dim].T, -s) + db @ Wl[l.in_dim:].T
                else:
                    acc = db @ Wl.T
                dx = h(acc + (f32(bypass) * dy if use_bp else 0))
            elif t == "prefinal-layer":
                dys = h(dy * self._bn_scale(self.bn[(l.name, "BN")]))
                wg[f"{l.name}.SmallW"] = sv["g"].T.astype(f32) @ dys
                dg = (dys @ P[f"{l.name}.SmallW"].T).astype(f32) * self._bn_scale(self.bn[(l.name, "PfBN")])
                dg = np.where(sv["mask"], h(dg), f32(0))
                wg[f"{l.name}.BigBias"] = dg.sum(0, dtype=f32).reshape(1, -1)
                wg[f"{l.name}.BigW"] = x.T.astype(f32) @ dg
                dx = O.gemm(dg, P[f"{l.name}.BigW"], transB=True)
            elif t == "attention-relu-batchnorm-layer":
                dx = self._attention_bwd(l, dy, sv, P, wg)
            elif t == "output-layer":
                wg[f"{l.name}.W"] = x.T.astype(f32) @ dy
                wg[f"{l.name}.Bias"] = dy.sum(0, dtype=f32).reshape(1, -1)
                dx = O.gemm(dy, P[f"{l.name}.W"], transB=True)
            elif t == "conv-relu-batchnorm-layer":
                fout = int(l.kv["num-filters-out"])
                dzz = dy.reshape(-1, fout)
                dz = np.where(sv["mask"], h(dzz * self._bn_scale(self.bn[(l.name, "BN")])), f32(0))
                wg[f"{l.name}.Bias"] = dz.sum(0, dtype=f32).reshape(1, -1)
                wg[f"{l.name}.W"] = sv["P"].T.astype(f32) @ dz
                dP = (dz @ P[f"{l.name}.W"].T).astype(f32)
                dx = h(self._patches_T(l, dP, sv["geom"]))
            # route to producers
            col = 0
            for name in l.inputs:
                src = self.by_name[name]
                part = dx[:, col:col + src.out_dim]
                col += src.out_dim
                if src.per_seq and not l.per_seq:
                    part = h(part.reshape(self.n_seq, self.L, -1).sum(1, dtype=f32))
                dact[name] = h(dact[name] + part) if name in dact else np.array(part, f32)
        return wg, dact

    def sgd(self, state: dict, wg: dict, lr: float, momentum: float):
        """SGDOptimizer.Update (optimize.go:95-120) on every parameter: grad rounded to FP16 first."""
        for k, g in wg.items():
            if k not in state:
                state[k] = dict(w32=self.params[k].astype(f32), v=np.zeros_like(self.params[k], f32))
            st = state[k]
            st["w32"], w16, st["v"] = O.sgd_update(st["w32"], h(g), st["v"], lr, momentum)
            self.params[k] = w16
        return state
