# full GPU test suite + smoke + a short driver-form bench (no ladder / reference arm): the last check when GPU minutes are short
mkdir -p gpurun_out
T=${1:-chkmin}
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i smoke
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-ladder > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print('bench', round(d['value']), round(d['e2e']['value'] or 0), d['roofline']['frac'], d['clocks']['sm_mhz'], d['cpu_baseline']['value'])"
