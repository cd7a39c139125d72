set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2d_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_multi.log
tail -15 gpurun_out/r2d_multi.log
( time timeout 900 python -m pytest tests/test_gpu_workloads.py -x -q -m gpu -k "packed or cli or pipelined" ) > gpurun_out/r2d_packed.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_packed.log
tail -15 gpurun_out/r2d_packed.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo rc=$?
tail -3 gpurun_out/r2d_bench.err
