# L2 -> SM traffic of the GEMM kernels inside one step (is the teacher GEMM bound by the L2 slice throughput rather than by
# the tensor pipe?).  Run under gpurun after tools/profile_step.py exits 0 without ncu.
cd $GRAFT_REPO_ROOT
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpc__cycles_elapsed.avg.per_second,lts__cycles_elapsed.avg.per_second,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none --profile-from-start off --kernel-name-base demangled -k 'regex:qv_gemm_kernel' -c ${QV_NCU_COUNT:-40} --csv --log-file gpurun_out/r01f_ncu_l2_gemm.csv python tools/profile_step.py > gpurun_out/r01f_ncu_l2.log 2>&1
tail -2 gpurun_out/r01f_ncu_l2.log
wc -l gpurun_out/r01f_ncu_l2_gemm.csv
