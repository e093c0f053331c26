# i-vector workload: SIMT tile width A/B for the N = 80 statistics adjoint (launch lists under ncu, then plain timings)
set -x
mkdir -p gpurun_out
T=${1:-ivbn}
for bn in 0 64 128; do
  SGB200_SIMT_BN=$bn timeout 300 python bench.py --workload iv --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-peak > gpurun_out/${T}_iv_bn$bn.json 2> gpurun_out/${T}_iv_bn$bn.err
  python -c "
import json; d=json.load(open('gpurun_out/${T}_iv_bn$bn.json')); print('iv bn$bn', round(d['value']), d['ms_per_step'])"
done
export SGB200_CUDA_GRAPH=0
for bn in 64 128; do
  SGB200_SIMT_BN=$bn timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file gpurun_out/${T}_launches_bn$bn.csv python bench.py --workload iv --steps 1 --warmup 0 --iters 5 --e2e-steps 0 --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu_l.err
  python tools/launch_summary.py gpurun_out/${T}_launches_bn$bn.csv 60 | grep simt
done
