mkdir -p gpurun_out
T=${1:-r2ze}
timeout 900 python -m pytest tests/test_gpu_audionet.py tests/test_gpu_antrain.py tests/test_gpu_tc.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )|tf32" gpurun_out/${T}_pytest.log | head -30
for prec in fp32 tf32; do
timeout 600 python bench.py --workload cw2 --precision $prec --steps 1 --warmup 1 --iters 200 --search-steps 2 --no-cpu-baseline > gpurun_out/${T}_cw2_$prec.json 2> gpurun_out/${T}_cw2_$prec.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_cw2_$prec.json')); print('$prec', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], d.get('kernel_ms_per_step'), d.get('attack_metrics'))"
tail -2 gpurun_out/${T}_cw2_$prec.err
done
