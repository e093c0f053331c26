# i-vector path after a kernel change: its tests, the workload's throughput, and the launch times of the per-utterance kernels
set -x
mkdir -p gpurun_out
T=${1:-ivab}
timeout 600 python -m pytest tests/test_gpu_iv.py tests/test_gpu_audionet.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -5
timeout 300 python bench.py --workload iv --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-peak > gpurun_out/${T}_iv.json 2> gpurun_out/${T}_iv.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_iv.json')); print('iv', round(d['value']), d['ms_per_step'])"
export SGB200_CUDA_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file gpurun_out/${T}_launches_iv.csv python bench.py --workload iv --steps 1 --warmup 0 --iters 5 --e2e-steps 0 --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu_l.err
python tools/launch_summary.py gpurun_out/${T}_launches_iv.csv 60 | grep -E "quad|chol"
