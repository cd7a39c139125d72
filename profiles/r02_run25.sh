set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_workloads.py tests/test_gpu_multi.py -x -q -m gpu -k "packed or cli or loopback or devices or pipelined" ) > gpurun_out/r2z_packed.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_packed.log
tail -6 gpurun_out/r2z_packed.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2z_bench.json')); print(d['e2e'])"
