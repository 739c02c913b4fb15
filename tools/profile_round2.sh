# round 2 (run under gpurun, after the suite and the bench have passed without ncu): ncu launch list of ONE step with the current build
# + one `--set full` capture of the first teacher GEMMs and of the student-side kernels; TAG names the output files
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r02}
python tools/profile_step.py > gpurun_out/${TAG}_plain_step.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain_step.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_ncu_launches_step_b256.csv python tools/profile_step.py > gpurun_out/${TAG}_ncu1.log 2>&1
tail -1 gpurun_out/${TAG}_ncu1.log; wc -l gpurun_out/${TAG}_ncu_launches_step_b256.csv
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:qv_gemm_kernel -s 1 -c 4 -o gpurun_out/${TAG}_teacher_gemm -f python tools/profile_step.py > gpurun_out/${TAG}_ncu2.log 2>&1
tail -1 gpurun_out/${TAG}_ncu2.log
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:qv_attn_bwd_kernel|qv_attn_fwd_kernel|resid_ln_fwd_kernel<\(int\)3>|ln_bwd_kernel|act_planes_kernel' -s 0 -c 8 -o gpurun_out/${TAG}_student_kernels -f python tools/profile_step.py > gpurun_out/${TAG}_ncu3.log 2>&1
tail -1 gpurun_out/${TAG}_ncu3.log
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:qv_gemm_kernel<\(int\)192, \(int\)2, \(int\)1, \(bool\)0, \(bool\)0, \(int\)[02]|qv_gemm_kernel<\(int\)192, \(int\)2, \(int\)2, \(bool\)1, \(bool\)1' -s 0 -c 6 -o gpurun_out/${TAG}_student_gemms -f python tools/profile_step.py > gpurun_out/${TAG}_ncu4.log 2>&1
tail -1 gpurun_out/${TAG}_ncu4.log
ls -la gpurun_out/${TAG}_*
