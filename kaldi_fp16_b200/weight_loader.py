"""Kaldi nnet3 weight import: the text form `nnet3-copy --binary=false` prints -> the executor's parameters.

Host-side mirror of internal/nnet/weight_loader.go (the reference is Go; there is no Go toolchain in this image, so the
host side above the C ABI is Python -- same names, same mapping):

  ParseNnet3Text      weight_loader.go:617-733   component name -> KaldiComponent (LinearParams / BiasParams / StatsMean /
                                                 StatsVar + the scalar tags of the header line)
  LoadWeights         weight_loader.go:754-946   per layer type: which components feed which tensors.  Kaldi stores
                                                 [out x in]; the executor (like the reference, replaceMatrix 973-991)
                                                 takes [in x out] through the truncating FP32 -> FP16 converter.
  batch-norm          weight_loader.go:438-463   makeBN: mean = StatsMean, var = StatsVar, gamma = target-rms, beta = 0
                                                 (the consistent path; LoadWeights' replaceBN normalises twice: quirk Q5)
  ExportModelText     weight_loader.go:605-614   runs Kaldi's nnet3-copy (only where Kaldi is installed)

plus WriteNnet3Text, the inverse of the parser (round-trip tests, exporting a trained network back to the text form).
The arithmetic-free part of the path: nothing here touches the GPU except Network.SetParam / SetBN / SetIDCT.
"""
from __future__ import annotations

import re
import subprocess
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

_F32 = np.float32


@dataclass
class KaldiComponent:
    """weight_loader.go:25-62 KaldiComponent"""
    Name: str = ""
    Type: str = ""
    LinearParams: Optional[np.ndarray] = None      # [rows x cols] as Kaldi stores it: [out x in]
    BiasParams: Optional[np.ndarray] = None
    StatsMean: Optional[np.ndarray] = None
    StatsVar: Optional[np.ndarray] = None
    LearningRate: float = 0.0
    MaxChange: float = 0.0
    L2Regularize: float = 0.0
    Epsilon: float = 0.0
    TargetRms: float = 0.0
    Count: float = 0.0
    NumFiltersIn: int = 0
    NumFiltersOut: int = 0
    HeightIn: int = 0
    HeightOut: int = 0
    NumHeads: int = 0
    KeyDim: int = 0
    ValueDim: int = 0
    KeyScale: float = 0.0
    Offsets: list = field(default_factory=list)        # TimeHeightConvolution: [(time, height), ...]
    TimeOffsets: list = field(default_factory=list)    # TdnnComponent

    @property
    def LinearRows(self) -> int:
        return 0 if self.LinearParams is None else self.LinearParams.shape[0]

    @property
    def LinearCols(self) -> int:
        return 0 if self.LinearParams is None else self.LinearParams.shape[1]


_SCALAR_F = {"<LearningRate>": "LearningRate", "<MaxChange>": "MaxChange", "<L2Regularize>": "L2Regularize",
             "<Epsilon>": "Epsilon", "<TargetRms>": "TargetRms", "<Count>": "Count", "<KeyScale>": "KeyScale"}
_SCALAR_I = {"<NumFiltersIn>": "NumFiltersIn", "<NumFiltersOut>": "NumFiltersOut", "<HeightIn>": "HeightIn",
             "<HeightOut>": "HeightOut", "<NumHeads>": "NumHeads", "<KeyDim>": "KeyDim", "<ValueDim>": "ValueDim"}
_MATRIX_TAGS = {"<LinearParams>": "LinearParams", "<Params>": "LinearParams", "<BiasParams>": "BiasParams",
                "<StatsMean>": "StatsMean", "<StatsVar>": "StatsVar"}


def _parse_matrix(body: str) -> np.ndarray:
    """text between '[' and ']': rows separated by newlines (a vector is one row)"""
    rows = [r for r in (line.split() for line in body.strip().splitlines()) if r]
    if not rows:
        return np.zeros((0, 0), _F32)
    width = len(rows[0])
    if any(len(r) != width for r in rows):
        raise ValueError("ragged matrix in nnet3 text")
    return np.array(rows, dtype=np.float64).astype(_F32)


def ParseNnet3Text(text: str) -> Dict[str, KaldiComponent]:
    """nnet3-copy --binary=false output -> {component name: KaldiComponent}.  Everything before the first <ComponentName>
    (the config lines) is skipped; tags the loader does not use are ignored."""
    comps: Dict[str, KaldiComponent] = {}
    starts = [m.start() for m in re.finditer(r"<ComponentName>", text)]
    for i, st in enumerate(starts):
        chunk = text[st:starts[i + 1] if i + 1 < len(starts) else len(text)]
        head = chunk.split(None, 3)
        if len(head) < 3:
            continue
        c = KaldiComponent(Name=head[1], Type=head[2].strip("<>"))
        # matrices / vectors: <Tag> [ ... ]
        for tag, attr in _MATRIX_TAGS.items():
            m = re.search(re.escape(tag) + r"\s*\[([^\]]*)\]", chunk)
            if m and getattr(c, attr) is None:
                mat = _parse_matrix(m.group(1))
                if mat.size == 0:
                    continue
                setattr(c, attr, mat if attr == "LinearParams" else mat.reshape(-1))
        # scalar tags: the first occurrence wins, a tag followed by another tag or a bracket carries no value
        for tag, attr in {**_SCALAR_F, **_SCALAR_I}.items():
            m = re.search(re.escape(tag) + r"\s+([^\s<\[]+)", chunk)
            if m:
                try:
                    setattr(c, attr, int(m.group(1)) if tag in _SCALAR_I else float(m.group(1)))
                except ValueError:
                    pass
        m = re.search(r"<Offsets>\s*\[([^\]]*)\]", chunk)
        if m:
            c.Offsets = [tuple(int(v) for v in p.split(",")) for p in m.group(1).split()]
        m = re.search(r"<TimeOffsets>\s*\[([^\]]*)\]", chunk)
        if m:
            c.TimeOffsets = [int(v) for v in m.group(1).split()]
        comps[c.Name] = c
    return comps


def _fmt_matrix(a: np.ndarray) -> str:
    a = np.asarray(a, _F32)
    if a.ndim == 1:
        return "[ " + " ".join(repr(float(v)) for v in a) + " ]"
    return "[\n" + "\n".join("  " + " ".join(repr(float(v)) for v in row) for row in a) + " ]"


def WriteNnet3Text(comps: Dict[str, KaldiComponent]) -> str:
    """inverse of ParseNnet3Text for the fields the loader reads (one component per <ComponentName> block)"""
    out = []
    for c in comps.values():
        line = f"<ComponentName> {c.Name} <{c.Type}>"
        for tag, attr in {**_SCALAR_F, **_SCALAR_I}.items():
            v = getattr(c, attr)
            if v:
                line += f" {tag} {v!r}"
        if c.Offsets:
            line += " <Offsets> [ " + " ".join(f"{t},{h}" for t, h in c.Offsets) + " ]"
        if c.TimeOffsets:
            line += " <TimeOffsets> [ " + " ".join(str(t) for t in c.TimeOffsets) + " ]"
        if c.LinearParams is not None:
            line += " <LinearParams>  " + _fmt_matrix(c.LinearParams)
            out.append(line)
            line = "<BiasParams>  " + (_fmt_matrix(c.BiasParams) if c.BiasParams is not None else "[ ]")
        if c.StatsMean is not None:
            line += " <StatsMean>  " + _fmt_matrix(c.StatsMean)
            out.append(line)
            line = "<StatsVar>  " + _fmt_matrix(c.StatsVar if c.StatsVar is not None else np.zeros_like(c.StatsMean))
        out.append(line)
    return "\n".join(out) + "\n"


def ExportModelText(model_path: str) -> str:
    """weight_loader.go:605-614: `nnet3-copy --binary=false <model> -` (needs a Kaldi installation on PATH)"""
    try:
        return subprocess.run(["nnet3-copy", "--binary=false", model_path, "-"], check=True, capture_output=True, text=True).stdout
    except FileNotFoundError as e:
        raise RuntimeError("nnet3-copy not found: export the model to text where Kaldi is installed and pass the text to "
                           "ParseNnet3Text") from e


class WeightLoadError(RuntimeError):
    pass


def _need(comps, name) -> KaldiComponent:
    c = comps.get(name)
    if c is None:
        raise WeightLoadError(f"{name} not found")
    return c


def _matrix(net, param: str, comp: KaldiComponent) -> int:
    """replaceMatrix (weight_loader.go:973-991): Kaldi [out x in] -> [in x out], truncating converter inside SetParam"""
    if comp.LinearParams is None or comp.LinearParams.size == 0:
        return 0
    w = np.ascontiguousarray(comp.LinearParams.T, _F32)
    rows, cols, _ = net.params[param]
    if w.shape != (rows, cols):
        raise WeightLoadError(f"{comp.Name}: Kaldi matrix {comp.LinearParams.shape[0]}x{comp.LinearParams.shape[1]} (out x in) does not fit "
                              f"{param} [{rows} x {cols}] (in x out)")
    net.SetParam(param, w)
    return w.size


def _vector(net, param: str, comp: KaldiComponent) -> int:
    if comp.BiasParams is None or comp.BiasParams.size == 0:
        return 0
    rows, cols, _ = net.params[param]
    if comp.BiasParams.size != cols:
        raise WeightLoadError(f"{comp.Name}: bias of {comp.BiasParams.size} does not fit {param} [1 x {cols}]")
    net.SetParam(param, comp.BiasParams.reshape(1, -1).astype(_F32))
    return comp.BiasParams.size


def _bn(net, layer: str, which: str, comp: KaldiComponent, dim: int) -> int:
    """makeBN (weight_loader.go:438-463): gamma = target-rms, beta = 0 -> y = target_rms * (x - mean) / sqrt(var + eps).
    A conv layer's BatchNormComponent has <Dim> = heights*filters with <BlockDim> = filters: one statistic per filter."""
    if comp.StatsMean is None or comp.StatsMean.size == 0:
        raise WeightLoadError(f"{comp.Name}: empty StatsMean")
    mean = comp.StatsMean
    var = comp.StatsVar if comp.StatsVar is not None else np.zeros_like(mean)
    if mean.size != dim or var.size != dim:
        raise WeightLoadError(f"{comp.Name}: {mean.size} statistics for a batch-norm of dim {dim}")
    rms = comp.TargetRms if comp.TargetRms > 0 else 1.0
    eps = comp.Epsilon if comp.Epsilon > 0 else 1e-3
    net.SetBN(layer, which, mean, np.maximum(var, 0), np.full(dim, rms, _F32), np.zeros(dim, _F32), eps)
    return 4 * dim


def LoadWeights(net, components: Dict[str, KaldiComponent], strict: bool = True) -> dict:
    """weight_loader.go:754-946 on a kaldi_fp16_b200.nnet.Network: returns {"loaded": layers, "params": values, "skipped": [...]}.
    strict = False skips layers whose components are missing instead of raising."""
    loaded, total, skipped = 0, 0, []
    for name, ltype, dim in net.layers:
        try:
            if ltype == "idct-layer":
                c = _need(components, "idct")
                if c.LinearParams is not None:
                    net.SetIDCT(name, np.ascontiguousarray(c.LinearParams.T, _F32))
                    total += c.LinearParams.size
            elif ltype == "linear-component":
                total += _matrix(net, f"{name}.W", _need(components, name))
            elif ltype == "batchnorm-component":
                total += _bn(net, name, "", _need(components, name), dim)
            elif ltype == "conv-relu-batchnorm-layer":
                conv = _need(components, f"{name}.conv")
                total += _matrix(net, f"{name}.W", conv) + _vector(net, f"{name}.Bias", conv)
                fout = net.params[f"{name}.W"][1]
                total += _bn(net, name, "BN", _need(components, f"{name}.batchnorm"), fout)
            elif ltype == "tdnnf-layer":
                total += _matrix(net, f"{name}.LinearW", _need(components, f"{name}.linear"))
                aff = _need(components, f"{name}.affine")
                total += _matrix(net, f"{name}.AffineW", aff) + _vector(net, f"{name}.AffineBias", aff)
                total += _bn(net, name, "AffBN", _need(components, f"{name}.batchnorm"), dim)
            elif ltype == "prefinal-layer":
                # (the reference maps every prefinal layer to "prefinal-chain" / "prefinal-xent" by substring,
                #  weight_loader.go:873-876; the layer's own name is the same thing for Kaldi's standard recipes)
                prefix = name if f"{name}.affine" in components else ("prefinal-xent" if "xent" in name else "prefinal-chain")
                aff = _need(components, f"{prefix}.affine")
                total += _matrix(net, f"{name}.BigW", aff) + _vector(net, f"{name}.BigBias", aff)
                total += _matrix(net, f"{name}.SmallW", _need(components, f"{prefix}.linear"))
                big = net.params[f"{name}.BigW"][1]
                total += _bn(net, name, "PfBN", _need(components, f"{prefix}.batchnorm1"), big)
                if f"{prefix}.batchnorm2" in components:      # (the reference leaves the second batch-norm at identity)
                    total += _bn(net, name, "BN", components[f"{prefix}.batchnorm2"], dim)
            elif ltype == "output-layer":
                c = _need(components, f"{name}.affine")
                total += _matrix(net, f"{name}.W", c) + _vector(net, f"{name}.Bias", c)
            elif ltype == "attention-relu-batchnorm-layer":
                # weight_loader.go:253-275: <name>.affine (projection), <name>.attention (<KeyScale>; the executor takes key-scale=
                # from the xconfig, default 1/sqrt(key-dim) as the reference), <name>.batchnorm
                aff = _need(components, f"{name}.affine")
                total += _matrix(net, f"{name}.W", aff) + _vector(net, f"{name}.Bias", aff)
                total += _bn(net, name, "BN", _need(components, f"{name}.batchnorm"), dim)
            else:
                if ltype != "input":
                    skipped.append(name)
                continue
            loaded += 1
        except WeightLoadError:
            if strict:
                raise
            skipped.append(name)
    return {"loaded": loaded, "params": total, "skipped": skipped}


def LoadWeightsFromFile(net, model_path: str) -> dict:
    """weight_loader.go:735-751"""
    return LoadWeights(net, ParseNnet3Text(ExportModelText(model_path)))


def ComponentsFromNetwork(net, bn_stats: Optional[dict] = None) -> Dict[str, KaldiComponent]:
    """the network's current parameters as Kaldi components ([out x in] matrices): the export direction.  Batch-norm running
    statistics are read back from the device (GetBN); bn_stats may override them ({(layer, which): (mean, var, rms, eps)})."""
    comps: Dict[str, KaldiComponent] = {}

    def mat(param):
        return np.ascontiguousarray(net.GetParam(param).T, _F32)

    def vec(param):
        return net.GetParam(param).reshape(-1).astype(_F32)

    def bn(cname, layer, which, dim):
        if bn_stats and (layer, which) in bn_stats:
            mean, var, rms, eps = bn_stats[(layer, which)]
        else:
            (mean, var), rms, eps = net.GetBN(layer, which, dim), 1.0, 1e-3
        comps[cname] = KaldiComponent(Name=cname, Type="BatchNormComponent", StatsMean=np.asarray(mean, _F32),
                                      StatsVar=np.asarray(var, _F32), TargetRms=float(rms), Epsilon=float(eps))

    for name, ltype, dim in net.layers:
        if ltype == "linear-component":
            comps[name] = KaldiComponent(Name=name, Type="LinearComponent", LinearParams=mat(f"{name}.W"))
        elif ltype == "batchnorm-component":
            bn(name, name, "", dim)
        elif ltype == "conv-relu-batchnorm-layer":
            comps[f"{name}.conv"] = KaldiComponent(Name=f"{name}.conv", Type="TimeHeightConvolutionComponent",
                                                   LinearParams=mat(f"{name}.W"), BiasParams=vec(f"{name}.Bias"))
            bn(f"{name}.batchnorm", name, "BN", net.params[f"{name}.W"][1])
        elif ltype == "tdnnf-layer":
            comps[f"{name}.linear"] = KaldiComponent(Name=f"{name}.linear", Type="TdnnComponent", LinearParams=mat(f"{name}.LinearW"))
            comps[f"{name}.affine"] = KaldiComponent(Name=f"{name}.affine", Type="TdnnComponent", LinearParams=mat(f"{name}.AffineW"),
                                                     BiasParams=vec(f"{name}.AffineBias"))
            bn(f"{name}.batchnorm", name, "AffBN", dim)
        elif ltype == "prefinal-layer":
            comps[f"{name}.affine"] = KaldiComponent(Name=f"{name}.affine", Type="NaturalGradientAffineComponent",
                                                     LinearParams=mat(f"{name}.BigW"), BiasParams=vec(f"{name}.BigBias"))
            comps[f"{name}.linear"] = KaldiComponent(Name=f"{name}.linear", Type="LinearComponent", LinearParams=mat(f"{name}.SmallW"))
            bn(f"{name}.batchnorm1", name, "PfBN", net.params[f"{name}.BigW"][1])
            bn(f"{name}.batchnorm2", name, "BN", dim)
        elif ltype == "output-layer":
            comps[f"{name}.affine"] = KaldiComponent(Name=f"{name}.affine", Type="NaturalGradientAffineComponent",
                                                     LinearParams=mat(f"{name}.W"), BiasParams=vec(f"{name}.Bias"))
    return comps
