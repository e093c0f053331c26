set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2g_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/r2g_pytest.log | head -20
b() { tag=$1; shift; env "$@" > gpurun_out/r2g_bench_$tag.json 2>> gpurun_out/r2g_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_$tag.json')); print('$tag', round(d['value']), round(d['e2e']['value'] or 0), d['gpu_launches'], d['kernel_ms_per_step'])"; }
b b1024 timeout 300 python bench.py --steps 4 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0 --no-peak
b b128 timeout 300 python bench.py --steps 5 --warmup 3 --no-ladder --no-cpu-baseline --batch 128 --e2e-steps 0 --no-peak
tail -3 gpurun_out/r2g_bench.err
export SGB200_CUDA_GRAPH=0
K='regex:conv_tc|mfcc|cmvn|head|score|loss|pool|gather|ctl|splitk'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 112 -c 56 --csv --log-file gpurun_out/r2g_launches_b1024.csv python bench.py --steps 1 --warmup 0 --iters 8 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/r2g_ncu_l.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 112 -c 56 --csv --log-file gpurun_out/r2g_launches_b128.csv python bench.py --steps 1 --warmup 0 --iters 8 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak --batch 128 > /dev/null 2>> gpurun_out/r2g_ncu_l.err
python tools/launch_summary.py gpurun_out/r2g_launches_b128.csv 60 | tail -32
