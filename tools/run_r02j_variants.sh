cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python tools/variant_bench.py 131072 default enc32 enc64 sign2 sign1 default > gpurun_out/r02j_variants_enc_sign.txt 2>&1
cut -c1-420 gpurun_out/r02j_variants_enc_sign.txt
