cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 300 python tools/latency_probe.py 2>&1 | grep -v Warning | tee gpurun_out/r02q_latency.txt
