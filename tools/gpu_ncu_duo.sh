#!/bin/bash
# ncu captures for the kernel experiments of round 2: tools/gpu_ncu_duo.sh <tag>
tag=${1:-r2c}
o=gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lat tools/ubench/lat.cu && /tmp/lat > $o/${tag}_ubench.log 2>&1
# two lanes per task, plan C=2 G=7, 16 utterances, default and zero thresholds
ncu --set full --import-source on --clock-control none -k regex:k_batch_strips -s 1 -c 1 -o $o/${tag}_duo python tools/gpu_ncu_strips.py 16 3 2 7 zero > $o/${tag}_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_batch_strips -s 1 -c 1 -o $o/${tag}_scalar python tools/gpu_ncu_strips.py 16 2 2 7 zero >> $o/${tag}_ncu.log 2>&1
tail -5 $o/${tag}_ncu.log; cat $o/${tag}_ubench.log
