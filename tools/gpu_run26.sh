mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 160 --csv --log-file gpurun_out/r3b_iv_launches.csv python bench.py --workload iv --steps 1 --warmup 0 --iters 4 --e2e-steps 0 --no-cpu-baseline > /dev/null 2> gpurun_out/r3b_ncu.err
python tools/launch_summary.py gpurun_out/r3b_iv_launches.csv 400 | awk '{n[$1" "$2" "$3" "$4]++; t[$1" "$2" "$3" "$4]+=$NF} END {for (k in n) printf "%9.1f us  x%3d  %s\n", t[k], n[k], k}' | sort -rn | head -40
tail -2 gpurun_out/r3b_ncu.err
