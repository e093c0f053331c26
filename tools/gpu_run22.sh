mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_audionet.py -m gpu -q --no-header -p no:cacheprovider -s -k "baseline_hyper" > gpurun_out/r2zt_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2zt_pytest.log | tail -3; grep -E "^(FAILED|E  )|targeted CW2" gpurun_out/r2zt_pytest.log | head
