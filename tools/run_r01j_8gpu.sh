set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
nvidia-smi -L | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -c 400 gpurun_out/bench_8gpu.err; cut -c1-500 gpurun_out/bench_8gpu.json
