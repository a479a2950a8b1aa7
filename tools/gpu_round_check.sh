#!/bin/bash
# Round-end check on a B200 (run through gpurun from the repo root): smoke, GPU parity tests, the bench lines of all five
# BASELINE configurations (per-GPU shapes), the ncu launch list and one full capture of the dominant kernel.
# usage: tools/gpu_round_check.sh <tag> [noncu]
tag=${1:-check}
o=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg2 cfg1 cfg3 cfg4 cfg5; do
  python bench.py --workload $w --cpu-seconds 4 > $o/${tag}_bench_$w.json 2> $o/${tag}_bench_$w.err
done
python bench.py --thresholds zero --cpu-seconds 2 > $o/${tag}_bench_cfg2_zero.json 2>> $o/${tag}_bench_cfg2.err
python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench_cfg2.err
if [ "$2" != "noncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > $o/${tag}_ncu.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:k_batch_strips -s 3 -c 1 -o $o/${tag}_strips python bench.py --steps 1 --warmup 3 --cpu-seconds 1 >> $o/${tag}_ncu.log 2>&1
fi
python - <<PY
import json
for f in ("cfg1", "cfg2", "cfg2_zero", "cfg3", "cfg4", "cfg5"):
    d = json.load(open("$o/${tag}_bench_%s.json" % f)); r = d["roofline"]
    print(f, "%.4g" % d["value"], "%.1f ms" % d["ms_per_step"], {k: round(v, 1) for k, v in r["stage_ms"].items()}, "frac %.4f" % r["frac"], "fp64 %s" % r.get("fp64_frac"),
          "e2e %.4g %.1f ms plain %.1f" % (d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["plain_call_ms"]), "cpu %.4g" % d["cpu_baseline"]["value"], d["gpu_launches"])
d = json.load(open("$o/${tag}_bench_reference.json")); print("reference arm", "%.4g" % d["value"], d["cpu_baseline"]["cores"], "cores")
PY
tail -2 $o/${tag}_bench_cfg2.err
