set -x
mkdir -p gpurun_out
T=${1:-r2u}
N=${2:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-ladder > gpurun_out/${T}_weak_n$N.json 2> gpurun_out/${T}_weak_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_weak_n$N.json')); print('weak', d['n_gpus'], round(d['value']), round(d['e2e']['value'] or 0), d['ms_per_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-cpu-baseline --no-ladder > gpurun_out/${T}_strong_n$N.json 2> gpurun_out/${T}_strong_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_strong_n$N.json')); print('strong', d['n_gpus'], round(d['value']), round(d['e2e']['value'] or 0), d['ms_per_step'], d['config'].get('batch_per_gpu'))"
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-ladder --e2e-steps 0 > gpurun_out/${T}_samebox_n1.json 2> gpurun_out/${T}_samebox_n1.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_samebox_n1.json')); print('n1', round(d['value']))"
tail -2 gpurun_out/${T}_weak_n$N.err gpurun_out/${T}_strong_n$N.err
