# CTA-pair GEMM: parity tests, then same-box A/B of the bench step (QV_GEMM_PAIR: bit 0 teacher, bit 1 student, bit 2 dgrad+gp)
cd $GRAFT_REPO_ROOT
timeout 150 python -m pytest tests/test_gemm_pair_gpu.py -m gpu -x -q 2>&1 | grep -v Warning | tail -15 > gpurun_out/pair_tests.log
cat gpurun_out/pair_tests.log | tail -8
grep -q "passed" gpurun_out/pair_tests.log && ! grep -q "failed\|error" gpurun_out/pair_tests.log || exit 1
for mode in ${PAIR_MODES:-0 1 7 0 7}; do
  QV_GEMM_PAIR=$mode timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/pair_bench_$mode.json 2> gpurun_out/pair_bench_$mode.err || { echo "bench mode $mode failed"; tail -5 gpurun_out/pair_bench_$mode.err; exit 2; }
  python - <<PY
import json
d = json.load(open("gpurun_out/pair_bench_$mode.json"))
f = d["roofline"]["families"]
print("mode $mode: %.2f ms/step %.0f img/s | teacher %.2f ms student %.2f dgrad+gp %.2f | sm %s MHz" % (d["ms_per_step"], d["value"], f["gemm[teacher linear]"]["ms"], f["gemm[student fwd/dgrad]"]["ms"], f["gemm[student dgrad+gp]"]["ms"], d["clocks"]["sm_mhz"]))
PY
done
