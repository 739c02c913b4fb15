#!/bin/bash
# same-box A/B of one environment switch: bench.py (no side measurements) alternating VAR=a / VAR=b, R rounds
# usage (under gpurun): bash tools/ab_env.sh QV_OVERLAP_REDUCE 0 1 3 r02ar
VAR=$1; A=$2; B=$3; R=${4:-3}; TAG=${5:-ab}
mkdir -p gpurun_out
for r in $(seq 1 $R); do
  for v in $A $B; do
    env $VAR=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-side --no-cpu-baseline > gpurun_out/${TAG}_${VAR}_${v}_$r.json 2> gpurun_out/${TAG}_${VAR}_${v}_$r.err
    python - "$VAR=$v round $r" gpurun_out/${TAG}_${VAR}_${v}_$r.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "sm_mhz", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
  done
done
