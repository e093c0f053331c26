set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2j_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/r2j_pytest.log | head -20
b() { tag=$1; shift; env "$@" > gpurun_out/r2j_bench_$tag.json 2>> gpurun_out/r2j_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_$tag.json')); n=d['config']['passes_per_step']; print('$tag', round(d['value']), d['gpu_launches'], {a: round(b / n, 3) for a, b in d['kernel_ms_per_step'].items()}, d['clocks']['sm_mhz'])"; }
for rep in 1 2; do
b rc1_$rep SGB200_ROW_COMPACTION=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
b rc0_$rep SGB200_ROW_COMPACTION=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
done
tail -3 gpurun_out/r2j_bench.err
