set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01j_launches_bench.csv python bench.py --steps 2 --warmup 1 --n-spend 131072 --n-issue 131072 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
tail -c 300 gpurun_out/bench_under_ncu.err; wc -l gpurun_out/r01j_launches_bench.csv
