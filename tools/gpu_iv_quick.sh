# i-vector tests + tensor-core tests + the iv workload (no ncu): the quick check after a change local to the i-vector path
mkdir -p gpurun_out
T=${1:-ivq}
timeout 600 python -m pytest tests/test_gpu_iv.py tests/test_gpu_tc.py tests/test_gpu_api.py -m gpu -q --no-header -p no:cacheprovider -x -s 2>&1 | grep -E "iv medium|passed|failed|Error" | tail -8
timeout 300 python bench.py --workload iv --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-peak > gpurun_out/${T}_iv.json 2> gpurun_out/${T}_iv.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_iv.json')); print('iv', round(d['value']), d['ms_per_step'])"
