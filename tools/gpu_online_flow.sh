o=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "online or cfg3 or golden or anyq" 2>&1 | tail -15 > $o/f1_pytest.log
cat $o/f1_pytest.log
for cfg in "4 0" "0 0" "3 0" "2 0" "4 10" "4 12"; do set -- $cfg; echo "FLOW=$1 S=$2"; LWSB_ONLINE_FLOW=$1 LWSB_ONLINE_FLOW_S=$2 timeout 120 python tools/gpu_ncu_online.py 64; done 2>&1 | tee $o/f1_times.log
