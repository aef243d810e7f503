cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 1 --warmup 1 --no-strong --mixed-frac 0 > gpurun_out/r02n_bench_4gpu.json 2> gpurun_out/r02n_bench_4gpu.err; echo torchrun rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02n_bench_4gpu.json')); print('value', d['value'], 'e2e', d['e2e']['value']); print(d['multi_abi'])"
