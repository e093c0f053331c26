set -x
mkdir -p gpurun_out
export SGB200_CUDA_GRAPH=0
K='regex:conv_tc|mfcc|cmvn|head|score|loss|pool|gather|ctl'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 104 -c 52 --csv --log-file gpurun_out/r2f_launches_b1024.csv python bench.py --steps 1 --warmup 0 --iters 8 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/r2f_ncu_l.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 104 -c 52 --csv --log-file gpurun_out/r2f_launches_b128.csv python bench.py --steps 1 --warmup 0 --iters 8 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak --batch 128 > /dev/null 2>> gpurun_out/r2f_ncu_l.err
python tools/launch_summary.py gpurun_out/r2f_launches_b128.csv 60 | tail -30
