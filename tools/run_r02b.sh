# round 2, step b: full GPU suite on the new ABI (two-pass, multi-device, screened), smoke, and the measured integer peak
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python - <<'PY' 2>&1 | tee gpurun_out/r02b_peak.txt
import importlib, subprocess
act = importlib.import_module("anonymous-credit-tokens_b200")
for i in range(3):
    print("act_measure_int_mul_peak: %.3f Tlimb-MAC/s" % (act.measure_int_mul_peak(0) / 1e12))
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout)
PY
