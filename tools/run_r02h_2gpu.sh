cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; tail -4 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 --n-spend 524288 --n-issue 131072 --no-strong --mixed-frac 0 > gpurun_out/r02h_bench_2gpu_small.json 2> gpurun_out/r02h_bench_2gpu_small.err
tail -3 gpurun_out/r02h_bench_2gpu_small.err; python -c "
import json; d=json.load(open('gpurun_out/r02h_bench_2gpu_small.json')); print(d['value'], d['e2e']['value']); print(d['multi_abi'])"
