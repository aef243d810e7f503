set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python tools/variant_bench.py 65536 base h2 base h2 > gpurun_out/variants17.txt 2>&1
cat gpurun_out/variants17.txt | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
