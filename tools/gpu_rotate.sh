# rotating lane order of the strip kernel (NS = 17): parity + bench with and without
o=gpurun_out; tag=${1:-h1}
timeout 900 python -m pytest tests -m gpu -x -q -k "strip or cfg2 or full_size or ragged or medium or results_leave or plan" 2>&1 | tail -5 > $o/${tag}_pytest.log
cat $o/${tag}_pytest.log
for nr in 0 1; do
  if [ $nr = 1 ]; then export LWSB_STRIP_NO_ROTATE=1; fi
  timeout 300 python bench.py --cpu-seconds 1 > $o/${tag}_bench_cfg2_nr$nr.json 2> $o/${tag}_bench.err
  timeout 300 python bench.py --thresholds zero --cpu-seconds 1 > $o/${tag}_bench_cfg2_zero_nr$nr.json 2>> $o/${tag}_bench.err
done
python - <<PY
import json
for f in ("cfg2_nr0","cfg2_nr1","cfg2_zero_nr0","cfg2_zero_nr1"):
    d=json.load(open("$o/${tag}_bench_%s.json"%f)); print(f, "%.1f ms"%d["ms_per_step"], d["roofline"]["stage_ms"], "e2e %.1f"%d["e2e"]["ms_per_step"], d["plan"]["sweep_fastest"] if d.get("plan") else None)
PY
