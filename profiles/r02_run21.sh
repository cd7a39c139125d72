set -x
mkdir -p gpurun_out
for N in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 --trace > gpurun_out/r2v_bench_n${N}.json 2> gpurun_out/r2v_bench_n${N}.err; echo rc=$?
grep -E "trace|Error|error|raise" gpurun_out/r2v_bench_n${N}.err | cut -c1-200 | head -10
done
