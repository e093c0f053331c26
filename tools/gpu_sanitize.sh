mkdir -p gpurun_out
T=r2san
run() { name=$1; tool=$2; shift 2; timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest "$@" -q --no-header -p no:cacheprovider -x > gpurun_out/${T}_${name}_${tool}.log 2>&1; echo "== $name $tool: $(grep -E 'passed|failed' gpurun_out/${T}_${name}_${tool}.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${name}_${tool}.log | tail -1)"; }
run mfcc2 memcheck tests/test_gpu_xv.py -k "mfcc_forward_no_dither or mfcc_adjoint or philox"
run mfcc2 racecheck tests/test_gpu_xv.py -k "mfcc_forward_no_dither or mfcc_adjoint"
run feco memcheck tests/test_gpu_feco.py -k "lloyd or fused_feco_step or eot_copies"
run feco racecheck tests/test_gpu_feco.py -k "lloyd or fused_feco_step"
run an_tc memcheck tests/test_gpu_audionet.py -k "tf32_mode_shapes or tf32_mode_vs"
run rowc memcheck tests/test_gpu_tc.py -k "row_compaction"
grep -E "Error|error" gpurun_out/${T}_*.log | grep -v "ERROR SUMMARY: 0" | sort | uniq -c | sort -rn | head -20
