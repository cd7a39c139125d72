set -x
mkdir -p gpurun_out
for N in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5 --trace > gpurun_out/r2ab_bench_n${N}.json 2> gpurun_out/r2ab_bench_n${N}.err; echo rc=$?
grep -E "trace|Error|error|raise" gpurun_out/r2ab_bench_n${N}.err | cut -c1-200 | head -10
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2ab_ref_n8.json 2> gpurun_out/r2ab_ref_n8.err; echo rc=$?
head -c 300 gpurun_out/r2ab_ref_n8.json
