cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
nvidia-smi topo -m | head -8
for mode in torch default; do timeout 600 python tools/multi_abi_bench.py 4 262144 $mode 2>&1 | grep -v Warning; done | tee gpurun_out/r02m_multi_abi.txt
