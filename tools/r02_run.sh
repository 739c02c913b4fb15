# round 2: full GPU suite + one bench line (run under gpurun)
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r02}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -v Warning gpurun_out/${TAG}_pytest_gpu.log | tail -${TAIL:-15}
if [ "${BENCH:-1}" = "1" ]; then
timeout 400 python bench.py --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("%.2f ms/step %.0f img/s e2e %.0f | launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
    print({k: v["ms"] for k, v in d["roofline"]["families"].items()})
    print(d["roofline"]["breakdown_ms"])
    for k in ("torch_cuda_eager", "int8_eval", "microbench", "kd_ce_sweep"):
        v = d.get(k)
        if v is None: continue
        if "error" in v: print(k, "ERROR", v["error"])
        elif k == "microbench": print(k, [(r["N"], r["K"], r["fwd_us"], r["ste_bwd_us"], r["fwd_gemm_frac"]) for r in v["shapes"]])
        elif k == "kd_ce_sweep": print(k, v["shapes"])
        else: print(k, {kk: vv for kk, vv in v.items() if kk in ("img_per_s", "ms_per_step", "ms_per_batch", "int8_linear", "cpu_reference", "logits_max_abs_diff_in_head_steps")})
    print("side s", d.get("side_measurements_s"), "clocks", d["clocks"])
except Exception as e:
    print("no bench line:", e)
    print(open("gpurun_out/${TAG}_bench.err").read()[-3000:])
PY
fi
