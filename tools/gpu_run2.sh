set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/r2b_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2b_pytest.log | tail -3; grep -E "^FAILED" gpurun_out/r2b_pytest.log | head -20
SGB200_CUDA_GRAPH=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-ladder --no-cpu-baseline > gpurun_out/r2b_bench_nograph.json 2> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_nograph.json')); print('nograph', d['value'], d['e2e']['value'], d['kernel_ms_per_step'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-ladder --no-cpu-baseline > gpurun_out/r2b_bench_graph.json 2>> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_graph.json')); print('graph', d['value'], d['e2e']['value'], d['gpu_launches'])"
for g in 0 1; do SGB200_CUDA_GRAPH=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-ladder --no-cpu-baseline --batch 128 > gpurun_out/r2b_bench_b128_g$g.json 2>> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_b128_g$g.json')); print('B128 graph=$g', d['value'], d['e2e']['value'], d['roofline']['per_layer'])" | cut -c1-1500; done
timeout 300 python bench.py --workload iv --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_iv_default.json 2>> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_iv_default.json')); print('iv default', d['value'], d['kernel_ms_per_step'])"
SGB200_TC_PAIR_BF16=4 timeout 300 python bench.py --workload iv --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_iv_pair4.json 2>> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_iv_pair4.json')); print('iv pair4', d['value'], d['kernel_ms_per_step'])"
SGB200_TC_PAIR_BF16=4 timeout 300 python bench.py --precision tf32 --steps 3 --warmup 2 --no-ladder --no-cpu-baseline > gpurun_out/r2b_bench_tf32_pair4.json 2>> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_tf32_pair4.json')); print('tf32 pair4', d['value'], d['roofline']['frac'])"
tail -5 gpurun_out/r2b_bench.err
timeout 900 compute-sanitizer --tool racecheck --print-limit 100000 python -m pytest tests/test_gpu_tc.py -q --no-header -p no:cacheprovider -k "tc_conv_bf16 or pool_adjoint or tc_conv_matches" > gpurun_out/r2b_racecheck_full.log 2>&1
grep -oE "in sg_[a-z_]+\.cu:[0-9]+" gpurun_out/r2b_racecheck_full.log | sort | uniq -c | sort -rn | head -20 > gpurun_out/r2b_racecheck_sites.txt; cat gpurun_out/r2b_racecheck_sites.txt; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2b_racecheck_full.log | tail -3
grep -E "Error: Race" -A3 gpurun_out/r2b_racecheck_full.log | grep -vE "sg_conv_tc.cu:260" | head -30 > gpurun_out/r2b_racecheck_other.txt; wc -l gpurun_out/r2b_racecheck_other.txt
rm -f gpurun_out/r2b_racecheck_full.log
