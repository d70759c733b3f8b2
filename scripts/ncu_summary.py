"""Condense an ncu report into the table kept under profiles/:   python scripts/ncu_summary.py X.ncu-rep > profiles/...txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
M = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
     ("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%act"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("dram__bytes_read.sum", "dram_rd"),
     ("dram__bytes_write.sum", "dram_wr"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm_bytes"),
     ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l2->sm_rate"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
     ("sm__cycles_elapsed.max", "sm_cycles"), ("smsp__cycles_active.avg", "smsp_active_cyc")]
print(f"# ncu --set full --clock-control none   report: {rep}")
for r in rows[2:]:
    print("-" * 100)
    for name, short in M:
        if name in col:
            print(f"{short:18s} {r[col[name]]:>40s} {units[col[name]]}")
