set -x
mkdir -p gpurun_out
T=${1:-r2t}
( time timeout 600 python bench.py --workload iv --steps 2 --warmup 1 > gpurun_out/${T}_iv.json 2> gpurun_out/${T}_iv.err ) 2>&1 | grep real; cut -c1-900 gpurun_out/${T}_iv.json; tail -3 gpurun_out/${T}_iv.err
( time SGB200_TC_PAIR_BF16=4 timeout 600 python bench.py --workload iv --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_iv_pair4.json 2> gpurun_out/${T}_iv_pair4.err ) 2>&1 | grep real; cut -c1-400 gpurun_out/${T}_iv_pair4.json; tail -3 gpurun_out/${T}_iv_pair4.err
( time timeout 900 python bench.py --workload cw2 --steps 1 --warmup 1 > gpurun_out/${T}_cw2.json 2> gpurun_out/${T}_cw2.err ) 2>&1 | grep real; cut -c1-1200 gpurun_out/${T}_cw2.json; tail -3 gpurun_out/${T}_cw2.err
( time timeout 900 python tools/bench_configs.py > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err ) 2>&1 | grep real; cut -c1-700 gpurun_out/${T}_configs.jsonl; tail -3 gpurun_out/${T}_configs.err
