# round 2, step c: reduced-size bench (debug of the new legs), then GPU tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python bench.py --n-spend 131072 --n-issue 131072 --mixed-n 262144 --steps 2 --warmup 1 > gpurun_out/r02c_bench_small.json 2> gpurun_out/r02c_bench_small.err
echo rc=$?; tail -5 gpurun_out/r02c_bench_small.err; cut -c1-3000 gpurun_out/r02c_bench_small.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
