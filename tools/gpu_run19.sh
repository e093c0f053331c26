mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/r2zg_cw2_launches.csv python bench.py --workload cw2 --precision tf32 --steps 1 --warmup 0 --iters 30 --search-steps 1 --e2e-steps 0 --no-cpu-baseline > /dev/null 2> gpurun_out/r2zg_ncu.err
python tools/launch_summary.py gpurun_out/r2zg_cw2_launches.csv 60 | tail -46
tail -3 gpurun_out/r2zg_ncu.err
