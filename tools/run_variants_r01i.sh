set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
for v in b128 b32 b64 b128 b32; do
  timeout 600 python tools/bench_variant.py $v --n-spend 524288 --n-issue 131072 --no-cpu-baseline --mixed-frac 0 --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms']['spend_range'])"
done > gpurun_out/variants18.txt 2>&1
cat gpurun_out/variants18.txt
