# constant-time evidence (tools/ct_counts.py): per-launch instruction / sector counters of the secret-handling kernels under ncu
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
M="smsp__inst_executed.sum,smsp__thread_inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,smsp__inst_executed_op_branch.sum,smsp__sass_branch_targets.sum,smsp__sass_branch_targets_threads_divergent.sum"
timeout 900 ncu --metrics $M --clock-control none -k regex:'refund_sign_seq_kernel|issue_mode_kernel|spend_head_kernel|finalize_ctx_kernel' --csv --log-file gpurun_out/ct_counts.csv python tools/ct_counts.py run > gpurun_out/ct_counts.log 2>&1
echo ncu rc=$?; tail -3 gpurun_out/ct_counts.log
python tools/ct_counts.py summarise gpurun_out/ct_counts.csv gpurun_out/ct_counts_summary.txt | tail -60
