set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s > /dev/null 2>&1
for tool in memcheck initcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_run ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|error" | head -8
done > gpurun_out/sanitizer_r01j.txt 2>&1
cat gpurun_out/sanitizer_r01j.txt
