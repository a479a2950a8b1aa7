"""Summarise every kernel launch of an ncu report (raw page) as one line each:  python tools/summarize_ncu_multi.py <rep> <out.txt> [title]"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
def val(r, name, scale=1.0):
    if name not in col or r[col[name]] in ("", "n/a"): return float("nan")
    x = float(r[col[name]].replace(",", "")); unit = u[col[name]]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(unit, 1.0)
    return x * mult * scale
lines = ["# " + title, "# ncu --set full --clock-control none; one line per captured launch (cold cache, serialised)",
         "%-44s %10s %9s %9s %8s %7s %7s %7s %6s %s" % ("kernel", "time_us", "dram_MB", "GB/s", "dram%", "issue%", "fp64%", "smem%", "regs", "grid x block")]
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("lwsb::", "").replace("<unnamed>::", "").replace("void ", "")
    t = val(r, "gpu__time_duration.sum")
    by = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    lines.append("%-44s %10.1f %9.2f %9.1f %8.1f %7.1f %7.1f %7.1f %6d %s x %s" % (
        name[:44], t * 1e6, by / 1e6, by / t / 1e9 if t > 0 else 0.0, val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), val(r, "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"),
        val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), int(val(r, "launch__registers_per_thread")),
        r[col["Grid Size"]], r[col["Block Size"]]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
