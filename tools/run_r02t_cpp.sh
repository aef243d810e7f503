cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 600 python -m pytest tests/test_cpp_host.py -x -q 2>&1 | tail -15
