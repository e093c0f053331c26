set -x
mkdir -p gpurun_out
T=${1:-r2zx}
timeout 900 python -m pytest tests/test_gpu_xv.py tests/test_gpu_tc.py tests/test_gpu_shard.py tests/test_gpu_precision.py tests/test_gpu_api.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
b() { tag=$1; shift; env "$@" > gpurun_out/${T}_bench_$tag.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); n=d['config']['passes_per_step']; print('$tag', round(d['value']), {a: round(b / n, 3) for a, b in d['kernel_ms_per_step'].items() if a.startswith('mfcc')}, d['clocks']['sm_mhz'], d['attack_metrics']['success_rate'])"; }
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak"
b hyb1 SGB200_FEAT_BWD_HYBRID=1 $B
b hyb0 SGB200_FEAT_BWD_HYBRID=0 $B
b hyb1b SGB200_FEAT_BWD_HYBRID=1 $B
b hyb0b SGB200_FEAT_BWD_HYBRID=0 $B
b hyb1_b128 SGB200_FEAT_BWD_HYBRID=1 $B --batch 128
b hyb0_b128 SGB200_FEAT_BWD_HYBRID=0 $B --batch 128
tail -3 gpurun_out/${T}_bench.err
