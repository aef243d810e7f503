# round 2: compute-sanitizer (memcheck, initcheck, racecheck) over every entry point incl. the round-2 ones; then one ncu capture of the
# encode kernel with 64-point batches (executed IMAD.WIDE count = its work constant)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
for tool in memcheck initcheck racecheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_run.py"
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_run ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|Invalid|Uninitialized" | head -12
done > gpurun_out/r02k_compute_sanitizer.txt 2>&1
cat gpurun_out/r02k_compute_sanitizer.txt
k=spend_encode_kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -c 1 -o gpurun_out/$k -f python tools/prof_spend.py 65536 1 > gpurun_out/prof_k_$k.log 2>&1; tail -1 gpurun_out/prof_k_$k.log
python tools/ncu_summary.py gpurun_out/$k.ncu-rep gpurun_out/r02k_$k.txt "$k with 64 points per inversion at the product launch shape (65 536 proofs per launch); ncu --set full --clock-control none" > /dev/null
rm -f gpurun_out/$k.ncu-rep
grep -E "gpu__time_duration|fmaheavy|IMAD.WIDE|^total" gpurun_out/r02k_$k.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
