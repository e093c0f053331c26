mkdir -p gpurun_out
python tools/prof_cfg4.py 256 5 2>/dev/null | tail -1
SGB200_CFG4_B=256 timeout 900 python tools/bench_configs.py 2>/dev/null | grep "config\": \"[14]" | cut -c1-260
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1; lscpu | grep "Model name\|MHz" | head -3
