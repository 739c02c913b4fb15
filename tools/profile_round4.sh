# round 1, session g: CTA-pair teacher GEMMs (run under gpurun): full GPU suite, bench line, then ncu (tensor pipe / L2 -> SM bytes) of
# the first GEMM launches of one step with the final build
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r01g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01g_pytest_gpu.log
tail -3 gpurun_out/r01g_pytest_gpu.log
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/r01g_bench.json 2> gpurun_out/r01g_bench.err; echo "bench rc=$?"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpc__cycles_elapsed.avg.per_second,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__cluster_dim_x
timeout 120 ncu --metrics $M --clock-control none --profile-from-start off --kernel-name-base demangled -k 'regex:qv_gemm_kernel' -c ${QV_NCU_COUNT:-14} --csv --log-file gpurun_out/r01g_ncu_pair_gemm.csv python tools/profile_step.py > gpurun_out/r01g_ncu.log 2>&1
tail -1 gpurun_out/r01g_ncu.log; wc -l gpurun_out/r01g_ncu_pair_gemm.csv
