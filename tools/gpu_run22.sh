mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xv.py tests/test_gpu_iv.py -m gpu -q --no-header -p no:cacheprovider -k "cmvn or iv" > gpurun_out/r2zs_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2zs_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/r2zs_pytest.log | head
timeout 600 python bench.py --workload iv --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2zs_iv.json 2> gpurun_out/r2zs_iv.err
python -c "
import json; d=json.load(open('gpurun_out/r2zs_iv.json')); print('iv', round(d['value']), d['kernel_ms_per_step'])"
