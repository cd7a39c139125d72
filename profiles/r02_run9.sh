set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2i_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_multi.log
tail -12 gpurun_out/r2i_multi.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dpx_min profiles/microbench/dpx_min.cu && /tmp/dpx_min > gpurun_out/r2i_dpx.txt 2>&1
cat gpurun_out/r2i_dpx.txt
