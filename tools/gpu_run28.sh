mkdir -p gpurun_out
T=r3d
timeout 600 python bench.py --workload iv --steps 2 --warmup 1 > gpurun_out/${T}_iv.json 2> gpurun_out/${T}_iv.err
timeout 900 python bench.py --workload cw2 --steps 1 --warmup 1 > gpurun_out/${T}_cw2.json 2> gpurun_out/${T}_cw2.err
timeout 900 python bench.py --workload cw2 --precision fp32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_cw2_fp32.json 2> gpurun_out/${T}_cw2_fp32.err
timeout 900 python tools/bench_configs.py > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err
SGB200_CFG4_B=256 timeout 900 python tools/bench_configs.py 2>/dev/null | grep "config\": \"4" > gpurun_out/${T}_config4_b256.jsonl
timeout 600 python bench.py --workload antrain --steps 3 --warmup 2 > gpurun_out/${T}_antrain.json 2> gpurun_out/${T}_antrain.err
for f in iv cw2 cw2_fp32 antrain; do python -c "
import json; d=json.loads([l for l in open('gpurun_out/${T}_$f.json') if l.startswith('{')][-1]); print('$f', round(d['value']), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}))"; done
cat gpurun_out/${T}_configs.jsonl gpurun_out/${T}_config4_b256.jsonl | cut -c1-300
