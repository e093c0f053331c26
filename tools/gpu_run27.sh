mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_feco.py tests/test_gpu_fullsize.py tests/test_gpu_xv.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/r3c_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r3c_pytest.log | tail -3; grep -E "^(FAILED|E  )|identical" gpurun_out/r3c_pytest.log | head -20
