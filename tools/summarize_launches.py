"""Kernel shares from an ncu launch list (--metrics gpu__time_duration.sum ... --csv):  python tools/summarize_launches.py <csv> <out.txt> <title>"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
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
out = ["# " + sys.argv[3], "# per-launch times under ncu are serialised and cold-cache: compare shares, not absolutes",
       "%-60s %8s %14s %8s" % ("kernel", "launches", "total_ms", "share")]
out += ["%-60s %8d %14.3f %7.2f%%" % (k[:60], cnt[k], v, 100 * v / T) for k, v in tot.most_common()]
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out))
