cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 400 python tools/variant_bench.py 65536 default issold default issold > gpurun_out/r02y_variants_issue_pub.txt 2>&1
cut -c1-330 gpurun_out/r02y_variants_issue_pub.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "issue or lifecycles or golden or sequential or two_pass or fuzz" 2>&1 | tail -2
