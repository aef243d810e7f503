set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python tools/variant_bench.py 65536 c0 c4 c0 c4 > gpurun_out/variants11.txt 2>&1
cat gpurun_out/variants11.txt | cut -c1-330
