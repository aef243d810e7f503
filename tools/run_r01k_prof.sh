set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:spend_range -s 1 -c 1 -o gpurun_out/range_r1k -f python tools/prof_spend.py 2368 2 > gpurun_out/prof_k.log 2>&1; tail -2 gpurun_out/prof_k.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01k_launches.csv python tools/prof_spend.py 16384 1 > gpurun_out/prof_k2.log 2>&1
ls -la gpurun_out/*.ncu-rep
