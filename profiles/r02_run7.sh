set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo8.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2g_multi.log 2>&1 || { tail -20 gpurun_out/r2g_multi.log; exit 1; }
tail -3 gpurun_out/r2g_multi.log
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --trace > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; echo rc=$?
grep -E "trace|Error|error" gpurun_out/r2g_bench_n$N.err | head -12
head -c 400 gpurun_out/r2g_bench_n$N.json; echo
done
