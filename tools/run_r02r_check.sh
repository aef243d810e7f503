cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --n-spend 262144 --n-issue 262144 --mixed-n 524288 --steps 2 --warmup 1 > gpurun_out/r02r_bench_small.json 2> gpurun_out/r02r_bench_small.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/r02r_bench_small.json')); print(d['value'], d['e2e']['value'], d['mixed_adversarial']['oracle_sample_equal_status_refund_nullifier'])"
