cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
{
timeout 300 python tools/variant_bench.py 131072 default skip0
for r in 0.3 0.5 0.7; do echo "ACT_L2_PERSIST=$r"; ACT_L2_PERSIST=$r timeout 300 python tools/variant_bench.py 131072 default skip0; done
} > gpurun_out/r02p_variants_skip0_l2.txt 2>&1
cut -c1-330 gpurun_out/r02p_variants_skip0_l2.txt
