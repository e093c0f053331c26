# duration of one kernel (regex $2) inside the iv workload
mkdir -p gpurun_out
export SGB200_CUDA_GRAPH=0
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$2" -s 1 -c 2 --csv --log-file gpurun_out/$1_k.csv python bench.py --workload iv --steps 1 --warmup 0 --iters 2 --e2e-steps 0 --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/$1_k.err
grep -v "^==" gpurun_out/$1_k.csv | awk -F'","' 'NR>1 {print $5, $(NF)}' | cut -c1-120
