# compute-sanitizer over the i-vector kernels changed late in round 2 (SPD solve with async panel loads, float4 split kernels,
# warp-per-row softmax, the batched / split-K tensor-core contractions)
mkdir -p gpurun_out
T=${1:-ivsan}
run() { name=$1; tool=$2; lim=$3; shift 3; timeout $lim compute-sanitizer --tool $tool --print-limit 20 python -m pytest "$@" -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_${name}_${tool}.log 2>&1; echo "== $name $tool: $(grep -E 'passed|failed' gpurun_out/${T}_${name}_${tool}.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${name}_${tool}.log | tail -1)"; }
run iv racecheck 110 tests/test_gpu_iv.py -k "embed_backward_against_oracle and 45"
run iv memcheck 130 tests/test_gpu_iv.py -k "embed_backward_against_oracle or (embed_stages and 37)"
grep -E "Error|error|hazard" gpurun_out/${T}_*.log | grep -v "ERROR SUMMARY: 0\|0 hazards" | sort | uniq -c | sort -rn | head -12
