#!/bin/bash
# Round-end check on a B200 (run through gpurun from the repo root): smoke, GPU parity tests, the bench lines of the
# BASELINE configurations that fit one GPU, the ncu launch list and one full capture of the dominant kernel.
# usage: tools/gpu_round_check.sh <tag>
tag=${1:-check}
o=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench.err
python bench.py --thresholds zero --cpu-seconds 2 > $o/${tag}_bench_zero.json 2>> $o/${tag}_bench.err
python bench.py --workload cfg5 --steps 2 --warmup 3 --cpu-seconds 2 > $o/${tag}_bench_cfg5.json 2>> $o/${tag}_bench.err
python bench.py --workload cfg1 --cpu-seconds 2 > $o/${tag}_bench_cfg1.json 2>> $o/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > $o/${tag}_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_batch_strips -s 3 -c 1 -o $o/${tag}_strips python bench.py --steps 1 --warmup 3 --cpu-seconds 1 >> $o/${tag}_ncu.log 2>&1
python - <<PY
import json
for f in ("default", "zero", "cfg5", "cfg1"):
    d = json.load(open("$o/${tag}_bench_%s.json" % f))
    print(f, "%.4g" % d["value"], "%.4g" % d["e2e"]["value"], "%.4f" % d["roofline"]["frac"], "%.2f" % d["roofline"]["kernel_ms"],
          "%.4g" % d["cpu_baseline"]["value"], d["gpu_launches"])
PY
tail -2 $o/${tag}_bench.err
