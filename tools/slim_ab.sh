# Round-2 opener: the experimental slim-epilogue build (make -C qat-vit_b200/csrc slim; six pipeline stages for the CTA-pair GEMMs).
# Run under gpurun AFTER building both libraries here:   gpurun --timeout 400 -- 'bash tools/slim_ab.sh'
cd $GRAFT_REPO_ROOT
SLIM=$GRAFT_REPO_ROOT/qat-vit_b200/lib/libqatvit_b200_slim.so
[ -f $SLIM ] || { echo "build it first: make -C qat-vit_b200/csrc slim"; exit 1; }
# 1. parity of every GEMM epilogue with the slim build (bit-identity vs oracle / one-CTA kernels is what these tests check)
QV_LIB=$SLIM timeout 200 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_mix_gpu.py tests/test_fused_gp_gpu.py \
    -m gpu -x -q 2>&1 | grep -v Warning | tail -6 > gpurun_out/slim_tests.log
tail -4 gpurun_out/slim_tests.log
grep -q "passed" gpurun_out/slim_tests.log && ! grep -q "failed\|error" gpurun_out/slim_tests.log || exit 1
# 2. isolated teacher-shape GEMMs, default vs slim
python tools/pair_probe.py 2>&1 | grep "mix" > gpurun_out/slim_probe_default.log
QV_LIB=$SLIM python tools/pair_probe.py 2>&1 | grep "mix" > gpurun_out/slim_probe_slim.log
paste -d'\n' gpurun_out/slim_probe_default.log gpurun_out/slim_probe_slim.log
# 3. same-box A/B of the step
for lib in default slim default slim; do
  if [ $lib = slim ]; then export QV_LIB=$SLIM; else unset QV_LIB; fi
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/slim_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); f = d['roofline']['families']
print('$lib: %.2f ms/step %.0f img/s | teacher %.2f ms' % (d['ms_per_step'], d['value'], f['gemm[teacher linear]']['ms']))"
done
