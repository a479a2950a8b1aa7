# compute-sanitizer over small calls: tools/gpu_sanitize.sh <tag> "<tool> <ops> <shapes> [ENV=..]" ...
o=gpurun_out; tag=$1; shift
i=0
for spec in "$@"; do
  set -- $spec; tool=$1; ops=$2; shapes=$3; envs=$4; i=$((i+1))
  env $envs timeout 600 compute-sanitizer --tool $tool --print-limit 3 python tools/gpu_sanitize.py $ops $shapes > $o/${tag}_${i}_$tool.log 2>&1
  echo "== $spec: $(grep -c '^done' $o/${tag}_${i}_$tool.log) done; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $o/${tag}_${i}_$tool.log | tail -1); $(grep -m1 -E 'Barrier error|Race reported' $o/${tag}_${i}_$tool.log)"
done
