# two-GPU check of the driver's launch form: own arm and reference arm under torchrun
set -x
mkdir -p gpurun_out
T=${1:-chk2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_n2.json 2> gpurun_out/${T}_n2.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/${T}_n2.json')); print('n2', d['n_gpus'], round(d['value']), round(d['e2e']['value'] or 0), d['ms_per_step'], d['scaling'], d['clocks'])"
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_ref_n2.json 2> gpurun_out/${T}_ref_n2.err ) 2>&1 | grep real
cut -c1-300 gpurun_out/${T}_ref_n2.json; wc -l gpurun_out/${T}_ref_n2.json
timeout 600 python -m pytest tests/test_gpu_shard.py -q --no-header -p no:cacheprovider 2>&1 | tail -2
