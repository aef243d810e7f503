# round 2, final code at N = 8 as the driver launches it (fewer steps)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02u_bench_8gpu.json 2> gpurun_out/r02u_bench_8gpu.err; echo torchrun rc=$? ) 2>&1 | tail -4
grep -v "^\[bench\]" gpurun_out/r02u_bench_8gpu.err | tail -4 | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02u_bench_8gpu.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'issue', d['issue']['value'], d['issue']['e2e']['value']); print(d['multi_abi']); print(d['mixed_adversarial']['value'])"
