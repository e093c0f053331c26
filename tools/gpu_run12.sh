set -x
mkdir -p gpurun_out
T=r2n
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
b() { tag=$1; shift; env "$@" > gpurun_out/${T}_bench_$tag.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); n=d['config']['passes_per_step']; print('$tag', round(d['value']), d['gpu_launches'], {a: round(b / n, 3) for a, b in d['kernel_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['attack_metrics']['success_rate'])"; }
b v2 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
b v1 SGB200_FEAT_V2=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
b v2occ2 SGB200_FEAT2_OCC=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
b v2_b128 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak --batch 128
b v1_b128 SGB200_FEAT_V2=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak --batch 128
tail -3 gpurun_out/${T}_bench.err
