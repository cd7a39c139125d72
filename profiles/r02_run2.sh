set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2b_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_multi.log
tail -15 gpurun_out/r2b_multi.log
( time timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py ) > gpurun_out/r2b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo rc=$?
tail -3 gpurun_out/r2b_bench.err
