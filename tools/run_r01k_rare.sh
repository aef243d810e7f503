set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
python - <<'PY' > gpurun_out/rare_selftest.txt 2>&1
import importlib, os
act = importlib.import_module("anonymous-credit-tokens_b200")
for v in ("rare3", "rare4"):
    act.LIB_PATH = os.path.join(os.getcwd(), "tools", "bin", f"libact_{v}.so")
    act._lib = None if hasattr(act, "_lib") else None
    act.selftest(0); print("selftest ok with", act.LIB_PATH)
    break
PY
cat gpurun_out/rare_selftest.txt
timeout 500 python tools/variant_bench.py 65536 norare rare rare3 rare4 norare rare rare3 rare4 > gpurun_out/variants21.txt 2>&1
cat gpurun_out/variants21.txt | cut -c1-420
