# fp32 (reference-precision) mode: launch list of one PGD pass and a full ncu capture of the SIMT GEMM
set -x
mkdir -p gpurun_out
T=${1:-fp32}
export SGB200_CUDA_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/${T}_launches_fp32.csv python bench.py --precision fp32 --steps 1 --warmup 0 --iters 4 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu_l.err
python tools/launch_summary.py gpurun_out/${T}_launches_fp32.csv 60 | tail -45
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_simt -s 10 -c 10 -f -o gpurun_out/${T}_simt python bench.py --precision fp32 --steps 1 --warmup 0 --iters 3 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu3.err
ls -la gpurun_out/${T}*.ncu-rep; tail -3 gpurun_out/${T}_ncu3.err
