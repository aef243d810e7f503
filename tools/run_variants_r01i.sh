set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 900 python tools/variant_bench.py 65536 b128x4 b32x16 r144 r136 r120 r112 r104 b32x20 b32x21 b64x9 b32x12 b32x14 c8k c32k b128x4 > gpurun_out/variants10.txt 2>&1
cat gpurun_out/variants10.txt | cut -c1-330
