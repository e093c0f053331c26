# fp32 SIMT GEMM change: full GPU tests, fp32 leg, i-vector workload
set -x
mkdir -p gpurun_out
T=${1:-fp32ab}
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
timeout 600 python bench.py --precision fp32 --steps 2 --warmup 1 --iters 10 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > gpurun_out/${T}_fp32.json 2> gpurun_out/${T}_fp32.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_fp32.json')); print('fp32', round(d['value']), d['ms_per_step'], d['clocks'])"
timeout 600 python bench.py --workload iv --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-peak > gpurun_out/${T}_iv.json 2> gpurun_out/${T}_iv.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_iv.json')); print('iv', round(d['value']), d['ms_per_step'])"
export SGB200_CUDA_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/${T}_launches_fp32.csv python bench.py --precision fp32 --steps 1 --warmup 0 --iters 4 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu_l.err
python tools/launch_summary.py gpurun_out/${T}_launches_fp32.csv 30 | grep simt
