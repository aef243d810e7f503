set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 300 python -m pytest tests -m gpu -q -k "same_token or lifecycles" > gpurun_out/pytest_gpu_k.log 2>&1; tail -5 gpurun_out/pytest_gpu_k.log
timeout 600 python bench.py --n-spend 131072 --n-issue 131072 --steps 1 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 1500 gpurun_out/bench_small.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_small.json').read().strip().splitlines()[-1])
print(json.dumps(d["mixed_adversarial"]))
print(d["value"], d["e2e"]["value"])
PY
