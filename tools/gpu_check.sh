# quick end-of-session check: full GPU test suite, smoke, the driver's bench command, the reference arm
set -x
mkdir -p gpurun_out
T=${1:-chk}
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i smoke
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print('full', round(d['value']), round(d['e2e']['value'] or 0), d['roofline']['frac'], d['roofline'].get('step_frac'), d['clocks'], {k:(round(v['value']), v.get('roofline',{}).get('frac')) for k,v in d.get('precision_ladder',{}).items()})"
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err ) 2>&1 | grep real
cut -c1-400 gpurun_out/${T}_bench_reference.json
