set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 300 gpurun_out/bench_2gpu.err; cut -c1-300 gpurun_out/bench_2gpu.json
timeout 600 python -m pytest tests -m gpu -q -k two_gpu > gpurun_out/pytest_2gpu.log 2>&1; tail -2 gpurun_out/pytest_2gpu.log
