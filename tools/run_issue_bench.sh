set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/bin/issue_bench > gpurun_out/issue_bench.json
cat gpurun_out/issue_bench.json
timeout 300 ncu --metrics smsp__inst_executed.sum,smsp__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.avg.per_cycle_active --clock-control none --csv --log-file gpurun_out/issue_bench_ncu.csv ./tools/bin/issue_bench > /dev/null 2>&1
wc -l gpurun_out/issue_bench_ncu.csv
