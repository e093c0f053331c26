set -x
mkdir -p gpurun_out
T=${1:-r2s}
timeout 900 python -m pytest tests/test_gpu_xv.py tests/test_gpu_tc.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
b() { tag=$1; shift; env "$@" > gpurun_out/${T}_bench_$tag.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); n=d['config']['passes_per_step']; print('$tag', round(d['value']), {a: round(b / n, 3) for a, b in d['kernel_ms_per_step'].items() if a.startswith('mfcc')}, d['clocks']['sm_mhz'])"; }
B="timeout 300 python bench.py --steps 2 --warmup 2 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak"
b base $B
b fpc80 SGB200_FEAT2_FPC=80 $B
b fpc112 SGB200_FEAT2_FPC=112 $B
b fpc160 SGB200_FEAT2_FPC=160 $B
b fpc304 SGB200_FEAT2_FPC=304 $B
b ch2 SGB200_FEAT2_CHUNKS=2 $B
b ch4 SGB200_FEAT2_CHUNKS=4 $B
b ch5 SGB200_FEAT2_CHUNKS=5 $B
b base2 $B
tail -3 gpurun_out/${T}_bench.err
