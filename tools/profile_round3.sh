# round 1, session e: mixed-format teacher GEMMs + persistent resid_ln_fwd (run under gpurun, after the plain run exits 0)
cd $GRAFT_REPO_ROOT
python tools/profile_step.py > gpurun_out/r01e_plain_step.log 2>&1 || { tail -5 gpurun_out/r01e_plain_step.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01e_ncu_launches_step_b256.csv python tools/profile_step.py > gpurun_out/r01e_ncu1.log 2>&1
tail -2 gpurun_out/r01e_ncu1.log
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:qv_gemm_kernel<\(int\)192, \(int\)2, \(int\)2, \(bool\)0, \(bool\)0, \(int\)[01], \(int\)1>' -s 0 -c 4 -o gpurun_out/r01e_teacher_gemm_mix python tools/profile_step.py > gpurun_out/r01e_ncu2.log 2>&1
tail -2 gpurun_out/r01e_ncu2.log
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:resid_ln_fwd_kernel' -s 1 -c 4 -o gpurun_out/r01e_resid_ln python tools/profile_step.py > gpurun_out/r01e_ncu3.log 2>&1
tail -2 gpurun_out/r01e_ncu3.log
ls -la gpurun_out/r01e_*
