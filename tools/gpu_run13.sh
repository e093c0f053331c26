mkdir -p gpurun_out
export SGB200_CUDA_GRAPH=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc -s 6 -c 2 -f -o gpurun_out/r2o_mfcc2 python bench.py --steps 1 --warmup 0 --iters 6 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/r2o_ncu.err
ls -la gpurun_out/r2o*.ncu-rep; tail -3 gpurun_out/r2o_ncu.err
