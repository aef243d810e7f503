# round 2, step f: one ncu --set full capture PER stage kernel at the product launch shape, with the executed-opcode table of each
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
for k in spend_head_kernel refund_sign_kernel spend_encode_kernel issue_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -c 1 -o gpurun_out/$k -f python tools/prof_spend.py 65536 1 > gpurun_out/prof_f_$k.log 2>&1; tail -1 gpurun_out/prof_f_$k.log
  python tools/ncu_summary.py gpurun_out/$k.ncu-rep gpurun_out/r02f_$k.txt "$k at the product launch shape (65 536 proofs per launch; issue: 262 144 requests per launch); ncu --set full --clock-control none" > /dev/null
  rm -f gpurun_out/$k.ncu-rep
done
grep -h -E "^## kernel|gpu__time_duration|fmaheavy|IMAD.WIDE|^total" gpurun_out/r02f_*.txt
