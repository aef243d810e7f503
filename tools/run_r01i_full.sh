set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spend_range -s 1 -c 1 -o gpurun_out/range_r1i -f python tools/prof_spend.py 2368 2 > gpurun_out/prof_i.log 2>&1; tail -2 gpurun_out/prof_i.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01i_launches.csv python tools/prof_spend.py 16384 1 > gpurun_out/prof_i2.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-600
ls -la gpurun_out/*.ncu-rep
