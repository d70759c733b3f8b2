"""Summarise a KFP16_PROFILE_DUMP stderr log: average device time per distinct GEMM launch shape."""
import collections
import sys

t = collections.defaultdict(list)
for line in open(sys.argv[1]):
    if not line.startswith("[kfp16 gemm]"):
        continue
    parts = line.split()
    us, tf = float(parts[2]), float(parts[4])
    key = " ".join(parts[6:])
    t[key].append((us, tf))
tot = 0.0
for k, v in sorted(t.items(), key=lambda kv: -sum(u for u, _ in kv[1])):
    us = sum(u for u, _ in v) / len(v)
    tf = sum(f for _, f in v) / len(v)
    tot += sum(u for u, _ in v)
    print(f"{us:9.2f} us avg x{len(v):4d} {tf:7.1f} TF  {k}")
print(f"total {tot / 1e3:.3f} ms over all launches in the log")
