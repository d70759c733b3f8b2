"""Per-instruction stall summary from `ncu -i REP --page source --csv -k regex:NAME` output (first matching kernel).
usage: python scripts/ncu_stalls.py REP.ncu-rep KERNEL_REGEX [launch_index=0] [top=25]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# split into kernels: each starts with a "Kernel Name" row followed by the header row
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if not starts:
    raise SystemExit("no kernel matched")
s = starts[min(which, len(starts) - 1)]
e = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(rows)
print("kernel:", rows[s][1][:120])
hdr = rows[s + 1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data, tot = [], 0
for r in rows[s + 2:e]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    tot += n
    data.append((n, r))
agg = {h: sum(int(r[idx[h]] or 0) for _, r in data) for h in stalls}
print("total samples", tot, " instructions", len(data))
print("stall totals:", ", ".join(f"{h[6:]} {v}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in sorted(data, key=lambda x: -x[0])[:top_n]:
    t = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{n:6d} {100.0 * n / max(tot, 1):5.1f}%  {r[idx['Source']].strip()[:64]:64s} {t}")
