# online chain kernels: parity tests, timings over the knobs, one ncu capture
o=gpurun_out; tag=${1:-f2}
timeout 600 python -m pytest tests -m gpu -x -q -k "online or cfg3" 2>&1 | tail -5 > $o/${tag}_pytest.log
cat $o/${tag}_pytest.log
for cfg in "1" "2" "0"; do echo "MAP=$cfg"; LWSB_ONLINE_FLOW_MAP=$cfg timeout 120 python tools/gpu_ncu_online.py 64; done 2>&1 | tee $o/${tag}_times.log
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_online_ -s 1 -c 1 -o $o/${tag}_online python tools/gpu_ncu_online.py 16 > $o/${tag}_ncu.log 2>&1
tail -3 $o/${tag}_ncu.log
fi
