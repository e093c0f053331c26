mkdir -p gpurun_out
T=${1:-ivp2}
export SGB200_CUDA_GRAPH=0
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"quad_expand_bwd|chol_factor" -s 0 -c 2 -f -o gpurun_out/${T}_qc python bench.py --workload iv --steps 1 --warmup 0 --iters 1 --e2e-steps 0 --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu.err
ls -la gpurun_out/${T}_qc.ncu-rep; tail -2 gpurun_out/${T}_ncu.err
