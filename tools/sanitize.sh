#!/bin/bash
# compute-sanitizer pass over the op-level GPU tests (mandatory for hand-written mbarrier / TMA / tcgen05 code).
# NOTE (round 1): this pool answers `compute-sanitizer is closed on this pool` (exit 86) -- the out-of-bounds check that
# runs here is tests/test_guard_bands_gpu.py (guard bands around every engine buffer).  Kept for pools where the tool is open.
# Run under gpurun, one GPU:   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Small-shape op tests only (the sanitizer serialises kernels and slows them 20-100x); each tool writes
# gpurun_out/sanitize_<tool>.log, a line "ERROR SUMMARY: 0 errors" at its end is the pass criterion.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
TESTS=${SANITIZE_TESTS:-"tests/test_fq_gpu.py tests/test_gemm_gpu.py tests/test_attention_gpu.py tests/test_errors_gpu.py"}
SEL=${SANITIZE_K:-"not full and not sweep"}
rc=0
for tool in ${SANITIZE_TOOLS:-memcheck synccheck racecheck}; do
    log=gpurun_out/sanitize_${tool}.log
    # --target-processes all: pytest may fork; --launch-timeout 0: the first import torch is slow on a fresh box
    timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool "$tool" --target-processes all --launch-timeout 0 \
        --error-exitcode 77 --print-limit 20 \
        python -m pytest $TESTS -m gpu -x -q -k "$SEL" > "$log" 2>&1
    r=$?
    echo "== $tool: exit $r ==" | tee -a "$log"
    grep -E "ERROR SUMMARY|passed|failed" "$log" | tail -4
    [ $r -ne 0 ] && rc=$r
done
exit $rc
