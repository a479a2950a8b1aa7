# the kernel-variant tests on the experiments build (build_exp/liblws_b200.so, built with LWSB_NVCC_EXTRA=-DLWSB_EXPERIMENTS)
o=gpurun_out; tag=${1:-x1}
for i in 1 2; do timeout 200 python bench.py --workload cfg1 --cpu-seconds 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg1 value %.1f ms e2e %.1f ms kernel %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stage_ms']))"; done
export LWSB_LIB_PATH=$PWD/build_exp/liblws_b200.so
timeout 1500 python -m pytest tests -m gpu -q -k "variant or two_lanes or four_bin or tensor_memory or online_kernel_choices or strip" 2>&1 | tail -6 | tee $o/${tag}_pytest_experiments.log
