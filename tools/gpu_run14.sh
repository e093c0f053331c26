set -x
mkdir -p gpurun_out
T=${1:-r2p}
timeout 900 python -m pytest tests/test_gpu_xv.py tests/test_gpu_tc.py tests/test_gpu_shard.py tests/test_gpu_precision.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
b() { tag=$1; shift; env "$@" > gpurun_out/${T}_bench_$tag.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); n=d['config']['passes_per_step']; print('$tag', round(d['value']), d['gpu_launches'], {a: round(b / n, 3) for a, b in d['kernel_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['attack_metrics']['success_rate'])"; }
b v2 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
b v1 SGB200_FEAT_V2=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
b v2occ2 SGB200_FEAT2_OCC=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
tail -3 gpurun_out/${T}_bench.err
export SGB200_CUDA_GRAPH=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc -s 6 -c 2 -f -o gpurun_out/${T}_mfcc2 python bench.py --steps 1 --warmup 0 --iters 6 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu.err
