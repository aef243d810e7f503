# round 2, step g: two GPUs -- the GPU test suite (multi-device engine on two real devices) and the bench at N = 2 (strong + multi-ABI legs)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; tail -4 gpurun_out/pytest_2gpu.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err ) 2>&1 | tail -3
tail -6 gpurun_out/r02g_bench_2gpu.err; cut -c1-300 gpurun_out/r02g_bench_2gpu.json
