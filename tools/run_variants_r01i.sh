set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python tools/variant_bench.py 65536 base addn base addn > gpurun_out/variants16.txt 2>&1
cat gpurun_out/variants16.txt | cut -c1-400
