set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2c_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/r2c_pytest.log | head -20
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-ladder --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2c_bench_$tag.json 2>> gpurun_out/r2c_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_$tag.json')); print('$tag', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], d['kernel_ms_per_step'])"; }
b base SGB200_CMVN_FUSION=0 SGB200_FEAT_OCC=3
b occ4 SGB200_CMVN_FUSION=0 SGB200_FEAT_OCC=4
b cmvn3 SGB200_CMVN_FUSION=1 SGB200_FEAT_OCC=3
b cmvn4 SGB200_CMVN_FUSION=1 SGB200_FEAT_OCC=4
tail -5 gpurun_out/r2c_bench.err
# source-level profiles of the three kernels furthest from their roofline (launch by launch: no graph replay under ncu)
export SGB200_CUDA_GRAPH=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc_fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2c_mfcc_fwd python bench.py --steps 1 --warmup 0 --iters 6 --e2e-steps 0 --no-ladder --no-cpu-baseline > /dev/null 2> gpurun_out/r2c_ncu1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc_bwd_kernel -s 3 -c 1 -f -o gpurun_out/r2c_mfcc_bwd python bench.py --steps 1 --warmup 0 --iters 6 --e2e-steps 0 --no-ladder --no-cpu-baseline > /dev/null 2> gpurun_out/r2c_ncu2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 33 -c 11 -f -o gpurun_out/r2c_conv_tc python bench.py --steps 1 --warmup 0 --iters 6 --e2e-steps 0 --no-ladder --no-cpu-baseline > /dev/null 2> gpurun_out/r2c_ncu3.err
ls -la gpurun_out/*.ncu-rep | tail -5; tail -3 gpurun_out/r2c_ncu3.err
