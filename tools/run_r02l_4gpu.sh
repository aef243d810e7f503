# round 2, step l: N = 4 (weak + the strong form of configs[3]: 8M proofs split over 4 GPUs) and the GPU suite incl. the differential fuzz
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_4gpu.log 2>&1; tail -3 gpurun_out/pytest_4gpu.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 1 > gpurun_out/r02l_bench_4gpu.json 2> gpurun_out/r02l_bench_4gpu.err ) 2>&1 | tail -3
tail -3 gpurun_out/r02l_bench_4gpu.err | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_4gpu.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'issue', d['issue']['value']); print(d['multi_abi']); print(d['strong_scaling']); print(d['mixed_adversarial']['value'])"
