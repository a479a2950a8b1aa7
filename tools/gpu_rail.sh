# experiments build (build_exp/liblws_b200.so): value warps + chain warps for the online chain
o=gpurun_out; tag=${1:-g2}
export LWSB_LIB_PATH=$PWD/build_exp/liblws_b200.so
timeout 600 python -m pytest tests -m gpu -x -q -k "online_kernel_choices" 2>&1 | tail -5 > $o/${tag}_pytest.log
cat $o/${tag}_pytest.log
for cfg in "1 9" "0 9"; do set -- $cfg; echo "RAIL=$1 S=$2"; LWSB_ONLINE_RAIL=$1 LWSB_ONLINE_RAIL_S=$2 timeout 120 python tools/gpu_ncu_online.py 64; done 2>&1 | tee $o/${tag}_times.log
if [ "$2" != "noncu" ]; then
LWSB_ONLINE_RAIL=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_online_ -s 1 -c 1 -o $o/${tag}_online python tools/gpu_ncu_online.py 16 > $o/${tag}_ncu.log 2>&1
tail -3 $o/${tag}_ncu.log
fi
