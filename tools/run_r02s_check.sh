cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --n-spend 131072 --n-issue 1048576 --mixed-frac 0 --no-cpu-baseline --steps 5 --warmup 2 > gpurun_out/r02s_bench_issue.json 2> gpurun_out/r02s_bench_issue.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_issue.json')); print('issue', d['issue']['value'], 'e2e', d['issue']['e2e']['value'])"
