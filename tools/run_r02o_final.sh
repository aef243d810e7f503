# round 2, final check on one GPU: the GPU suite, smoke(), and the bench exactly as the driver runs it (both arms)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02o_bench_reference.json 2> gpurun_out/r02o_bench_reference.err; cut -c1-200 gpurun_out/r02o_bench_reference.json
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02o_bench_1gpu.json 2> gpurun_out/r02o_bench_1gpu.err ) 2>&1 | tail -3
tail -3 gpurun_out/r02o_bench_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_1gpu.json')); r=d['roofline']
print('value', d['value'], 'e2e', d['e2e']['value'], 'issue', d['issue']['value'], d['issue']['roofline_frac'])
print('frac', r['frac'], r['frac_of_model'], r['other_kernels_frac'], r['step_over_range_kernel'], r['whole_step'])
print(d['mixed_adversarial']['value'], d['mixed_adversarial']['oracle_sample_equal_status_refund_nullifier'], d['oracle_checks'])"
