# round 2, step d: the full default bench as the driver runs it (N = 1), both arms
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02d_bench_1gpu.json 2> gpurun_out/r02d_bench_1gpu.err ) 2>&1 | tail -3
echo rc=$?; tail -4 gpurun_out/r02d_bench_1gpu.err; cut -c1-600 gpurun_out/r02d_bench_1gpu.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r02d_bench_reference.json 2> gpurun_out/r02d_bench_reference.err; cut -c1-400 gpurun_out/r02d_bench_reference.json
