# round record: the driver's own bench command, the launch list and the ncu captures the profiles/ summaries are made from
set -x
mkdir -p gpurun_out
T=${1:-r2m}
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_full.json')); print('full', round(d['value']), round(d['e2e']['value'] or 0), d['roofline']['frac'], d['roofline'].get('step_frac'), d.get('cpu_baseline'), {k:(round(v['value']), v.get('roofline',{}).get('frac')) for k,v in d.get('precision_ladder',{}).items()})"
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err ) 2>&1 | grep real
cut -c1-600 gpurun_out/${T}_bench_reference.json
export SGB200_CUDA_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/${T}_launches_b1024.csv python bench.py --steps 1 --warmup 0 --iters 20 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu_l.err
python tools/launch_summary.py gpurun_out/${T}_launches_b1024.csv 60 | tail -32
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 16 -c 11 -f -o gpurun_out/${T}_conv_tc python bench.py --steps 1 --warmup 0 --iters 6 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2> gpurun_out/${T}_ncu3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc -s 6 -c 2 -f -o gpurun_out/${T}_mfcc python bench.py --steps 1 --warmup 0 --iters 6 --e2e-steps 0 --no-ladder --no-cpu-baseline --no-peak > /dev/null 2>> gpurun_out/${T}_ncu3.err
ls -la gpurun_out/${T}*.ncu-rep | tail -5; tail -3 gpurun_out/${T}_ncu3.err
