mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/r2zy_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2zy_pytest.log | tail -3; grep -E "^(FAILED|E  )|MFCC adjoint|differs" gpurun_out/r2zy_pytest.log | head
