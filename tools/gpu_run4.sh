set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2d_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/r2d_pytest.log | head -20
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2d_bench_$tag.json 2>> gpurun_out/r2d_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_$tag.json')); print('$tag', round(d['value']), round(d['e2e']['value']), d['e2e'].get('ms_per_step_serial_copies'), d['gpu_launches'], d['kernel_ms_per_step'])"; }
b cmvn0 SGB200_CMVN_FUSION=0
b cmvn1 SGB200_CMVN_FUSION=1
tail -5 gpurun_out/r2d_bench.err
export SGB200_CUDA_GRAPH=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 11 -f -o gpurun_out/r2d_conv_tc python bench.py --steps 1 --warmup 0 --iters 6 --e2e-steps 0 --no-ladder --no-cpu-baseline > /dev/null 2> gpurun_out/r2d_ncu3.err
ls -la gpurun_out/r2d*.ncu-rep | tail -5; tail -3 gpurun_out/r2d_ncu3.err
