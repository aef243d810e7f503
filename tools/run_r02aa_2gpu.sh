cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 1 --warmup 1 --n-spend 131072 --n-issue 131072 --no-strong --mixed-frac 0 > gpurun_out/r02aa_bench_2gpu.json 2> gpurun_out/r02aa_bench_2gpu.err; echo torchrun rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02aa_bench_2gpu.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'issue', d['issue']['value']); print(d['multi_abi'])"
