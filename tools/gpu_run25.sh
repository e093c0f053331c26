set -x
mkdir -p gpurun_out
T=${1:-r3a}
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_shard.py tests/test_gpu_tc.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
b() { tag=$1; shift; env "$@" > gpurun_out/${T}_bench_$tag.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); n=d['config']['passes_per_step']; print('$tag', round(d['value']), {a: round(b / n, 3) for a, b in d['kernel_ms_per_step'].items() if a.startswith('tdnn')}, d['gpu_launches'])"; }
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak"
b ts1_b128 SGB200_TC_TAIL_SPLIT=1 $B --batch 128
b ts0_b128 SGB200_TC_TAIL_SPLIT=0 $B --batch 128
b ts1_b128b SGB200_TC_TAIL_SPLIT=1 $B --batch 128
b ts0_b128b SGB200_TC_TAIL_SPLIT=0 $B --batch 128
b ts1_b256 SGB200_TC_TAIL_SPLIT=1 $B --batch 256
b ts0_b256 SGB200_TC_TAIL_SPLIT=0 $B --batch 256
tail -3 gpurun_out/${T}_bench.err
