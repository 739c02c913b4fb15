cd $GRAFT_REPO_ROOT
python tools/profile_step.py > gpurun_out/r01d_plain_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:qv_attn_bwd_kernel|ln_bwd_kernel|qv_fq_weight_grouped_kernel|colsum_reduce_kernel|qv_splitk_reduce_kernel' -s 0 -c 12 -o gpurun_out/r01d_bwd_kernels python tools/profile_step.py > gpurun_out/r01d_ncu.log 2>&1
tail -2 gpurun_out/r01d_ncu.log
