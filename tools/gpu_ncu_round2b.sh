#!/bin/bash
# ncu captures of the end of round 2: the online chain kernel (k_online_flow) and the kernels around the sweeps (the shared-memory
# NoFuture_LWSQ4 kernel included); launch list of bench.py --workload cfg3.   usage: tools/gpu_ncu_round2b.sh <tag>
tag=${1:-r2z}
o=gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_online_ -s 1 -c 1 -o $o/${tag}_online python tools/gpu_ncu_online.py 16 > $o/${tag}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_stft|k_extend|k_stats|k_nofuture|k_crop|k_istft|k_overlap|k_sq_norms" -s 9 -c 12 -o $o/${tag}_others python tools/gpu_ncu_others.py >> $o/${tag}_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --cpu-seconds 1 >> $o/${tag}_ncu.log 2>&1
tail -4 $o/${tag}_ncu.log
