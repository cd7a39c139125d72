set -x
mkdir -p gpurun_out
for CFG in "8 " "8 --no-rebalance" "4 " "2 "; do
set -- $CFG; N=$1; F=$2; T=${F:+_nb}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 --trace $F > gpurun_out/r2u_bench_n${N}$T.json 2> gpurun_out/r2u_bench_n${N}$T.err; echo rc=$?
grep -E "trace|Error|error|raise" gpurun_out/r2u_bench_n${N}$T.err | cut -c1-200 | head -10
head -c 300 gpurun_out/r2u_bench_n${N}$T.json; echo
done
