mkdir -p gpurun_out
for cfg in "tf32 200" "bf16 300" "bf16 150"; do
timeout 300 compute-sanitizer --tool memcheck python tools/dbg_rc.py $cfg > gpurun_out/r2k_san.log 2>&1
echo "== $cfg"; grep -v "^=========     Host Frame\|^=========         in " gpurun_out/r2k_san.log | head -12
done
sed -i 's/r2j/r2l/g' tools/gpu_run10.sh
bash tools/gpu_run10.sh 2>&1 | grep -v "^+"
