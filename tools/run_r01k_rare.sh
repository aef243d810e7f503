set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
for v in rare rareg; do
python - $v <<'PY' >> gpurun_out/rare_selftest.txt 2>&1
import importlib, os, sys
act = importlib.import_module("anonymous-credit-tokens_b200")
act.LIB_PATH = os.path.join(os.getcwd(), "tools", "bin", f"libact_{sys.argv[1]}.so")
act.selftest(0); print("selftest ok with", act.LIB_PATH)
PY
done
cat gpurun_out/rare_selftest.txt
timeout 300 python tools/variant_bench.py 65536 rare rareg rare rareg > gpurun_out/variants22.txt 2>&1
cat gpurun_out/variants22.txt | cut -c1-420
