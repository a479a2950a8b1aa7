# rotating lane order of the strip kernel (NS = 17) with / without the extra frame of sweep lag: parity + bench
o=gpurun_out; tag=${1:-h1}
timeout 900 python -m pytest tests -m gpu -x -q -k "strip or cfg2 or full_size or ragged or medium or results_leave or plan or rotating or cfg3_cfg4" 2>&1 | tail -4 > $o/${tag}_pytest.log
cat $o/${tag}_pytest.log
for nr in 0 1; do
  if [ $nr = 1 ]; then export LWSB_STRIP_NO_EXTRA=1; fi
  timeout 300 python bench.py --cpu-seconds 1 --steps 6 > $o/${tag}_bench_cfg2_nx$nr.json 2> $o/${tag}_bench.err
  timeout 300 python bench.py --thresholds zero --cpu-seconds 1 > $o/${tag}_bench_cfg2_zero_nx$nr.json 2>> $o/${tag}_bench.err
done
python - <<PY
import json
for f in ("cfg2_nx0","cfg2_nx1","cfg2_zero_nx0","cfg2_zero_nx1"):
    d=json.load(open("$o/${tag}_bench_%s.json"%f)); print(f, "%.1f ms"%d["ms_per_step"], d["roofline"]["stage_ms"], "e2e %.1f"%d["e2e"]["ms_per_step"], d["plan"]["sweep_extra_from"], d["plan"]["load_lead"])
PY
