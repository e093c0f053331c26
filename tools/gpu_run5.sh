set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2e_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/r2e_pytest.log | head -20
b() { tag=$1; shift; env "$@" > gpurun_out/r2e_bench_$tag.json 2>> gpurun_out/r2e_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_$tag.json')); print('$tag', round(d['value']), round(d['e2e']['value'] or 0), d['e2e'].get('ms_per_step_serial_copies'), d['gpu_launches'], d['kernel_ms_per_step'])"; }
b b1024 timeout 300 python bench.py --steps 4 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 3
b b1024_old SGB200_HEAD_MULTI=0 SGB200_TC_SMALL_BN=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 0
b b128 timeout 300 python bench.py --steps 5 --warmup 3 --no-ladder --no-cpu-baseline --batch 128 --e2e-steps 0
b b128_old SGB200_HEAD_MULTI=0 SGB200_TC_SMALL_BN=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-ladder --no-cpu-baseline --batch 128 --e2e-steps 0
tail -3 gpurun_out/r2e_bench.err
# launch list of one PGD pass (launch by launch), B = 1024 and B = 128
export SGB200_CUDA_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/r2e_launches_b1024.csv python bench.py --steps 1 --warmup 0 --iters 20 --e2e-steps 0 --no-ladder --no-cpu-baseline > /dev/null 2> gpurun_out/r2e_ncu_l.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/r2e_launches_b128.csv python bench.py --steps 1 --warmup 0 --iters 20 --e2e-steps 0 --no-ladder --no-cpu-baseline --batch 128 > /dev/null 2>> gpurun_out/r2e_ncu_l.err
python tools/launch_summary.py gpurun_out/r2e_launches_b128.csv 60 | tail -30
