set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
timeout 300 python tools/layer_table.py --out gpurun_out/r2a_layer_table.json > gpurun_out/r2a_layer_table.log 2>&1; tail -2 gpurun_out/r2a_layer_table.log | cut -c1-400
timeout 300 python bench.py --impl torch_gpu --steps 1 --warmup 1 --precision bf16 > gpurun_out/r2a_torch_gpu_bf16.json 2> gpurun_out/r2a_torch_gpu.err; cat gpurun_out/r2a_torch_gpu_bf16.json | cut -c1-300; tail -2 gpurun_out/r2a_torch_gpu.err
timeout 300 python bench.py --impl torch_gpu --steps 1 --warmup 1 --precision tf32 > gpurun_out/r2a_torch_gpu_tf32.json 2>> gpurun_out/r2a_torch_gpu.err; cat gpurun_out/r2a_torch_gpu_tf32.json | cut -c1-300
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_tc.py -q --no-header -p no:cacheprovider -k "tc_conv_bf16 or pool_adjoint" 2>&1 | tail -30 > gpurun_out/r2a_racecheck.log; tail -8 gpurun_out/r2a_racecheck.log
