#!/bin/bash
# Same-box A/B of one engine switch: runs the headline bench once per value, twice over, and appends the JSON lines
# (tagged with the setting) to gpurun_out/ab_<VAR>.jsonl.  Box-to-box variation on the shared pool is +-3 %, larger than
# most single changes, so every kernel decision of round 1 was taken from runs of this form inside ONE gpurun call.
#   usage: tools/ab_env.sh SGB200_TC_PAIR_BF16 0 2        (switches: SGB200_POOL_FUSION, SGB200_FEAT_STASH,
#          SGB200_L1_TAP_FORM, SGB200_TC_PAIR_BF16, SGB200_TC_PAIR_XF, SGB200_TC_ISSUE, SGB200_TC_PREFETCH, SGB200_TC_DEEP_RING)
var=$1; shift
out=gpurun_out/ab_${var}.jsonl
for rep in 1 2; do
  for v in "$@"; do
    env "$var=$v" python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 0 2>/dev/null | grep '^{' | sed "s/^{/{\"$var\": \"$v\", /" >> "$out"
  done
done
python - "$out" "$var" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l); k = d["kernel_ms_per_step"]; n = d["config"]["passes_per_step"]
    print(d[sys.argv[2]], round(d["value"]), {a: round(b / n, 3) for a, b in k.items() if b > 0.5}, d["clocks"]["sm_mhz"])
PY
