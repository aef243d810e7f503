cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 400 python tools/variant_bench.py 65536 default issold default > gpurun_out/r02z_variants_head_pub.txt 2>&1
cut -c1-330 gpurun_out/r02z_variants_head_pub.txt
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
