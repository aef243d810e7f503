set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 120 python -c "
import importlib; act=importlib.import_module('anonymous-credit-tokens_b200'); act.selftest(0); print('selftest ok')"
timeout 900 python tools/variant_bench.py 65536 c0 cap c0 cap > gpurun_out/variants12.txt 2>&1
cat gpurun_out/variants12.txt | cut -c1-330
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-400
