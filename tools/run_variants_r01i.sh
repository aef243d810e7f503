set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python tools/variant_bench.py 65536 base hd base hd > gpurun_out/variants15.txt 2>&1
cat gpurun_out/variants15.txt | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-300
