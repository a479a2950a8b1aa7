# quick check after a kernel change: a subset of the parity tests and one bench line
o=gpurun_out; tag=${1:-q}; sel=${2:-"nofuture or golden or cfg3 or medium"}; wl=${3:-cfg3}
timeout 900 python -m pytest tests -m gpu -x -q -k "$sel" 2>&1 | tail -5 > $o/${tag}_pytest.log
cat $o/${tag}_pytest.log
timeout 300 python bench.py --workload $wl --cpu-seconds 1 > $o/${tag}_bench_$wl.json 2> $o/${tag}_bench_$wl.err
python -c "
import json; d=json.load(open('$o/${tag}_bench_$wl.json')); print('%.4g'%d['value'], '%.1f ms'%d['ms_per_step'], d['roofline']['stage_ms'], 'e2e %.1f'%d['e2e']['ms_per_step'], 'plain %.1f'%d['e2e']['plain_call_ms'])"
