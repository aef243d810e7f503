# round 2, step e: ncu evidence at the product launch shapes + constant-time counters
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
# 1. the dominant kernel, full set (persistent grid of 592 blocks; 4736 proofs = 32 units per resident warp)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spend_range -s 1 -c 1 -o gpurun_out/range_r02e -f python tools/prof_spend.py 4736 2 > gpurun_out/prof_e1.log 2>&1; tail -1 gpurun_out/prof_e1.log
# 2. the thread-per-proof stages and the issue kernel at the product launch shape (65 536 proofs per launch; 262 144 requests per launch)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spend_head_kernel|refund_sign_kernel|spend_encode_kernel|^issue_kernel' -c 4 -o gpurun_out/side_r02e -f python tools/prof_spend.py 65536 1 > gpurun_out/prof_e2.log 2>&1; tail -1 gpurun_out/prof_e2.log
# 3. launch list of the bench command (reduced size): per-launch durations, the range kernel's share of the step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02e_launches_bench.csv python bench.py --n-spend 131072 --n-issue 131072 --mixed-frac 0 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/prof_e3.json 2> gpurun_out/prof_e3.err; tail -2 gpurun_out/prof_e3.err
# 4. peak microbenchmark under ncu: pipe utilisation of the measurement itself
timeout 120 ncu --set full --clock-control none -k regex:int_mul_peak -s 2 -c 1 -o gpurun_out/peak_r02e -f python -c "import importlib; a = importlib.import_module('anonymous-credit-tokens_b200'); print(a.measure_int_mul_peak(0))" > gpurun_out/prof_e4.log 2>&1; tail -1 gpurun_out/prof_e4.log
# 5. constant-time counters
bash tools/run_ct_counts.sh 2>&1 | tail -70
# summaries are made here on the box (gpurun copies back at most 64 MiB): only the range kernel's report travels
python tools/ncu_summary.py gpurun_out/range_r02e.ncu-rep gpurun_out/r02e_spend_range.txt "spend_range_kernel (bucket form), persistent grid, 4736 proofs; ncu --set full --clock-control none" > /dev/null
python tools/ncu_summary.py gpurun_out/side_r02e.ncu-rep gpurun_out/r02e_side_kernels.txt "head / encode / sign at 65 536 proofs per launch and issue_kernel at 262 144 requests per launch (the product's launch shapes); ncu --set full --clock-control none" > /dev/null
python tools/ncu_summary.py gpurun_out/peak_r02e.ncu-rep gpurun_out/r02e_int_mul_peak.txt "int_mul_peak_kernel (act_measure_int_mul_peak); ncu --set full --clock-control none" > /dev/null
for r in side_r02e peak_r02e; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page details --csv --print-details all 2>/dev/null | grep -iE "IMAD|Executed Ipc|inst_executed|Issue Slots|Registers|Theoretical|Achieved Occ|Local" | head -80 > gpurun_out/$r.details.txt
  rm -f gpurun_out/$r.ncu-rep
done
ncu -i gpurun_out/range_r02e.ncu-rep --page raw --csv > gpurun_out/range_r02e.raw.csv 2>/dev/null
rm -f gpurun_out/*r1i.ncu-rep gpurun_out/*r1j.ncu-rep gpurun_out/*r1k.ncu-rep
ls -la gpurun_out/ | tail -15; du -sh gpurun_out
