set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r01c_pytest_gpu.log 2>&1; tail -3 gpurun_out/r01c_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r01c_bench.json 2> gpurun_out/r01c_bench.err; tail -2 gpurun_out/r01c_bench.err
python tools/profile_step.py > gpurun_out/r01c_plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01c_ncu_launches_step_b256.csv python tools/profile_step.py > gpurun_out/r01c_ncu1.log 2>&1
tail -2 gpurun_out/r01c_ncu1.log
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:qv_gemm_kernel -s 1 -c 4 -o gpurun_out/r01c_teacher_gemm python tools/profile_step.py > gpurun_out/r01c_ncu2.log 2>&1
tail -2 gpurun_out/r01c_ncu2.log
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:qv_attn_bwd_kernel|qv_attn_fwd_kernel<\(int\)1>|resid_ln_fwd_kernel<\(int\)3>|ln_bwd_kernel|act_planes_kernel' -s 0 -c 8 -o gpurun_out/r01c_student_kernels python tools/profile_step.py > gpurun_out/r01c_ncu3.log 2>&1
tail -2 gpurun_out/r01c_ncu3.log
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:qv_gemm_kernel<\(int\)192, \(int\)2, \(int\)1, \(bool\)0, \(bool\)0, \(int\)2>|qv_gemm_kernel<\(int\)192, \(int\)2, \(int\)2, \(bool\)1, \(bool\)1' -s 0 -c 3 -o gpurun_out/r01c_dgrad_gp_wgrad python tools/profile_step.py > gpurun_out/r01c_ncu4.log 2>&1
tail -2 gpurun_out/r01c_ncu4.log
ls -la gpurun_out/r01c_*
