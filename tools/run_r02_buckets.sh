# Next-round first step for the bucket form of the range-proof pair.
# Before: tools/build_variants.sh buckets:"-DACT_RANGE_BUCKETS=1"    (here, on the CPU; the .so travels with the snapshot)
# Then:   gpurun --timeout 900 -- 'bash tools/run_r02_buckets.sh'
# A/B timing, then the FULL GPU parity suite and the smoke run with the bucket library in place of the product library
# (only in the GPU box's scratch copy of the repo).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
timeout 300 python tools/variant_bench.py 65536 default buckets default buckets > gpurun_out/r02_variants_buckets.txt 2>&1
cut -c1-330 gpurun_out/r02_variants_buckets.txt
cp anonymous-credit-tokens_b200/libact_b200.so /tmp/libact_default.so
cp tools/bin/libact_buckets.so anonymous-credit-tokens_b200/libact_b200.so
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_buckets.log 2>&1; tail -3 gpurun_out/pytest_gpu_buckets.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_buckets.log 2>&1; tail -1 gpurun_out/smoke_buckets.log
cp /tmp/libact_default.so anonymous-credit-tokens_b200/libact_b200.so
