# round 2, step i: the bench at N = 8 as the driver launches it (fewer steps), one box
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
nvidia-smi -L | wc -l
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 2 > gpurun_out/r02i_bench_8gpu.json 2> gpurun_out/r02i_bench_8gpu.err ) 2>&1 | tail -3
tail -5 gpurun_out/r02i_bench_8gpu.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02i_bench_8gpu.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'issue', d['issue']['value']); print(d['multi_abi']); print(d['strong_scaling']); print(d['mixed_adversarial']['value'])"
