mkdir -p gpurun_out
python tools/time_feco.py 1280 300 2>&1 | grep feco_kmeans
python tools/time_feco.py 1600 500 2>&1 | grep feco_kmeans
timeout 900 python -m pytest tests/test_gpu_feco.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2zo_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2zo_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/r2zo_pytest.log | head
python tools/prof_cfg4.py 256 5 2>/dev/null | tail -1
