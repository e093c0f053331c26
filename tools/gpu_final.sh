# end-of-round record: full GPU test suite, smoke, the driver's bench command + reference arm, launch list, ncu captures
set -x
mkdir -p gpurun_out
T=${1:-r2fin}
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
bash tools/gpu_record.sh $T 2>&1 | grep -v "^+"
