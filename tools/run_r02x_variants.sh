cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 600 python tools/variant_bench.py 131072 default ib64 ib256 hb128 hb32 default > gpurun_out/r02x_variants_block_sizes.txt 2>&1
cut -c1-330 gpurun_out/r02x_variants_block_sizes.txt
