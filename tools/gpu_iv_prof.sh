# i-vector workload: launch list of one PGD pass and full ncu captures of the Cholesky and batched-statistics kernels
set -x
mkdir -p gpurun_out
T=${1:-ivp}
export SGB200_CUDA_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/${T}_launches_iv.csv python bench.py --workload iv --steps 1 --warmup 0 --iters 5 --e2e-steps 0 --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu_l.err
python tools/launch_summary.py gpurun_out/${T}_launches_iv.csv 80 | tail -70
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol -s 2 -c 2 -f -o gpurun_out/${T}_chol python bench.py --workload iv --steps 1 --warmup 0 --iters 3 --e2e-steps 0 --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu3.err
ls -la gpurun_out/${T}*.ncu-rep; tail -3 gpurun_out/${T}_ncu3.err
