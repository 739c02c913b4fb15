#!/bin/bash
# same-box A/B of two builds of the library: bench.py (no side measurements) alternating QV_LIB=<a> / <b>, R rounds
# usage (under gpurun): bash tools/ab_lib.sh qat-vit_b200/lib/libqatvit_b200_old.so qat-vit_b200/lib/libqatvit_b200.so 3 r02ax
A=$1; B=$2; R=${3:-3}; TAG=${4:-ablib}
mkdir -p gpurun_out
for r in $(seq 1 $R); do
  for v in $A $B; do
    n=$(basename $v .so)
    QV_LIB=$PWD/$v timeout 300 python bench.py --steps 30 --warmup 5 --no-side --no-cpu-baseline > gpurun_out/${TAG}_${n}_$r.json 2> gpurun_out/${TAG}_${n}_$r.err
    python - "$n round $r" gpurun_out/${TAG}_${n}_$r.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    b = d["roofline"]["breakdown_ms"]
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "attn_fwd", b.get("attn_fwd"), "sm_mhz", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
  done
done
