# A/B of the spend pipeline's chunk size (ACT_SPEND_CHUNK): device-resident step rate and e2e at 262144 proofs.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
for c in 16384 32768 65536 131072; do
  ACT_SPEND_CHUNK=$c timeout 600 python bench.py --n-spend 262144 --n-issue 65536 --mixed-frac 0 --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r02_chunk_$c.json 2> gpurun_out/r02_chunk_$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_chunk_$c.json"))
print($c, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "range_frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel_ms"])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
