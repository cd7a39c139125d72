set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2e_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_multi.log
tail -15 gpurun_out/r2e_multi.log
