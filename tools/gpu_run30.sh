mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/r3g_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r3g_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/r3g_pytest.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -E "smoke|rror"
timeout 300 python bench.py --steps 3 --warmup 3 --no-ladder --no-cpu-baseline --no-peak > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3g_bench.json')); print('bench', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic'])"
