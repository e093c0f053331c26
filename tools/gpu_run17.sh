set -x
mkdir -p gpurun_out
T=${1:-r2w}
SGB200_KMEANS_V2=0 python tools/time_feco.py 1600 300 2>&1 | grep feco_kmeans
python tools/time_feco.py 1600 300 2>&1 | grep feco_kmeans
python tools/time_feco.py 256 300 2>&1 | grep feco_kmeans
python tools/time_feco.py 1600 500 2>&1 | grep feco_kmeans
timeout 600 python -m pytest tests/test_gpu_feco.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3; grep -E "^(FAILED|E  )" gpurun_out/${T}_pytest.log | head -20
timeout 600 python tools/bench_configs.py 2>/dev/null | grep "config\": \"4"
