o=gpurun_out
for g in 4 8 6 3; do LWSB_EARLY_GROUPS=$g timeout 200 python bench.py --cpu-seconds 1 --steps 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('groups $g: value %.1f e2e %.1f plain %.1f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['plain_call_ms']))"; done | tee $o/groups.log
