set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2f_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_multi.log
tail -8 gpurun_out/r2f_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --trace > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo rc=$?
tail -12 gpurun_out/r2f_bench_n2.err
cat gpurun_out/r2f_bench_n2.json | head -c 1500
