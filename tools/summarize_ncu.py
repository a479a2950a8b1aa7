"""Turn the ncu artefacts of tools/gpu_round_check.sh into the text summaries kept under profiles/.

    python tools/summarize_ncu.py <tag>      (reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_strips.ncu-rep)
"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, tag + "_launches.csv"))) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in data:
    v, u = float(r[vi].replace(",", "")), r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    name = re.sub(r"\(.*", "", r[ki]).replace("lwsb::<unnamed>::", "").replace("void ", "")
    tot[name] += ms; cnt[name] += 1
T = sum(tot.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 80, python bench.py --steps 2 --warmup 3 (cfg2, default thresholds)",
       "# per-launch times under ncu are serialised and cold-cache: compare shares, not absolutes",
       "%-60s %8s %14s %8s" % ("kernel", "launches", "total_ms", "share")]
out += ["%-60s %8d %14.3f %7.2f%%" % (k[:60], cnt[k], v, 100 * v / T) for k, v in tot.most_common()]
open(os.path.join(P, tag + "_launches_summary.txt"), "w").write("\n".join(out) + "\n")
subprocess.run(["cp", os.path.join(G, tag + "_launches.csv"), os.path.join(P, tag + "_launches.csv")])

# ---- full capture: raw page
rep = os.path.join(G, tag + "_strips.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, d = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__cluster_max_active", "launch__cluster_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio"]
want += [x for x in h if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio")]
out = ["# ncu --set full --clock-control none --import-source on, cluster strip kernel, bench.py cfg2 default thresholds, 1 launch (64 utterances, 100 sweeps)"]
out += ["%-95s %-16s %s" % (w, u[h.index(w)], d[h.index(w)]) for w in want if w in h]
open(os.path.join(P, tag + "_strips_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
gb = lambda name: float(d[h.index(name)]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[h.index(name)]]
rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
json.dump({"kernel": d[h.index("Kernel Name")], "workload": "cfg2 default thresholds (64 utterances, 100 sweeps)", "dram_bytes_read": rd,
           "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "source": "profiles/%s_strips_ncu_summary.txt" % tag},
          open(os.path.join(P, "traffic_cfg2.json"), "w"), indent=1)

# ---- full capture: source page, stall samples by opcode
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
idx = {x: i for i, x in enumerate(hdr)}
stalls = [x for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
tot, byop, ex, per = collections.Counter(), collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r in data:
    s = r[idx["Source"]].split()
    op = (s[0] if not s[0].startswith("@") else s[1]).split(".")[0]
    byop[op] += int(r[idx["# Samples"]] or 0); ex[op] += int(r[idx["Instructions Executed"]] or 0)
    for st in stalls:
        v = int(r[idx[st]] or 0); tot[st] += v; per[op][st] += v
total, te = sum(tot.values()), sum(ex.values())
out = ["# ncu --set full --import-source on, source page of the cluster strip kernel, bench.py cfg2 default thresholds",
       "# warp-state samples: %d; share by stall reason (all warps, the idle control warp included):" % total,
       "  " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v / total) for k, v in tot.most_common(10)),
       "# by opcode: share of samples, share of executed warp-instructions, top stall reasons (share of all samples)"]
out += ["  %-8s samples %5.1f%%  executed %5.1f%%   %s" % (op, 100 * v / total, 100 * ex[op] / te,
        ", ".join("%s %.1f" % (k.replace("stall_", ""), 100 * x / total) for k, x in per[op].most_common(3))) for op, v in byop.most_common(16)]
open(os.path.join(P, tag + "_stall_by_opcode.txt"), "w").write("\n".join(out) + "\n")
for f in ("default", "zero", "cfg5", "cfg1"):
    subprocess.run(["cp", os.path.join(G, "%s_bench_%s.json" % (tag, f)), os.path.join(P, "%s_bench_%s.json" % (tag, f))])
print(open(os.path.join(P, tag + "_launches_summary.txt")).read())
print("\n".join(open(os.path.join(P, tag + "_strips_ncu_summary.txt")).read().splitlines()[:22]))
