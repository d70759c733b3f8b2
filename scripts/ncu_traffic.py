"""DRAM traffic of one training step's kernels from an ncu --set full report:
    python scripts/ncu_traffic.py X.ncu-rep [workload] > profiles/r02_ncu_traffic_<workload>.json
Sums dram__bytes_read.sum + dram__bytes_write.sum over the captured launches (one eager step: every GEMM incl. the implicit
convolutions, the elementwise passes, the chain objective, the SGD update) and lists them per kernel."""
import collections
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "cnn_tdnn"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    i = col[name]
    v = float(r[i].replace(",", "")) if r[i] else 0.0
    u = units[i].lower()
    mul = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
    return v * mul


per = collections.OrderedDict()
tot = gemm = 0.0
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    b = num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    t = num(r, "gpu__time_duration.sum")
    e = per.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += b
    e["time_us"] += t
    tot += b
    if "gemm_f16_sm100" in name:
        gemm += b
out = {"workload": workload, "dram_bytes_per_step": gemm, "dram_bytes_per_step_all_kernels": tot,
       "launches": sum(e["launches"] for e in per.values()),
       "how": "ncu --set full --clock-control none, one eager step (scripts/profile_cnn_tdnn_step.py): dram__bytes_read.sum + dram__bytes_write.sum "
              "summed over the step's gemm_f16_sm100 launches (dram_bytes_per_step) and over every captured kernel (…_all_kernels)",
       "per_kernel": [dict(kernel=k, **v) for k, v in sorted(per.items(), key=lambda kv: -kv[1]["dram_bytes"])]}
print(json.dumps(out, indent=1))
