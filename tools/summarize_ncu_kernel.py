"""Summary files for one full ncu capture of a kernel:  python tools/summarize_ncu_kernel.py <rep> <out_prefix> <title> [traffic_json workload]"""
import collections, csv, json, subprocess, sys
rep, pre, title = sys.argv[1], sys.argv[2], sys.argv[3]
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
out = ["# " + title]
out += ["%-95s %-16s %s" % (w, u[h.index(w)], d[h.index(w)]) for w in want if w in h]
open(pre + "_ncu_summary.txt", "w").write("\n".join(out) + "\n")
gb = lambda name: float(d[h.index(name)]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[h.index(name)]]
rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
if len(sys.argv) > 5:
    json.dump({"kernel": d[h.index("Kernel Name")], "workload": sys.argv[5], "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_launch": rd + wr, "source": pre.split("/")[-1] + "_ncu_summary.txt"}, open(sys.argv[4], "w"), indent=1)
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
o2 = ["# source page (SASS) of the same capture: " + title,
      "# warp-state samples: %d; share by stall reason (all warps):" % total,
      "  " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v / total) for k, v in tot.most_common(10)),
      "# by opcode: share of samples, share of executed warp-instructions, top stall reasons (share of all samples)"]
o2 += ["  %-8s samples %5.1f%%  executed %5.1f%%   %s" % (op, 100 * v / total, 100 * ex[op] / te,
       ", ".join("%s %.1f" % (k.replace("stall_", ""), 100 * x / total) for k, x in per[op].most_common(3))) for op, v in byop.most_common(16)]
open(pre + "_stall_by_opcode.txt", "w").write("\n".join(o2) + "\n")
print("\n".join(out[:24])); print("\n".join(o2[:8]))
