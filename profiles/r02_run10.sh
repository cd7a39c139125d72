set -x
mkdir -p gpurun_out
for CFG in "8 2" "8 3" "4 2" "2 2"; do
set -- $CFG; N=$1; P=$2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 --trace --phases $P > gpurun_out/r2j_bench_n${N}_p$P.json 2> gpurun_out/r2j_bench_n${N}_p$P.err; echo rc=$?
grep -E "trace|Error|error|raise" gpurun_out/r2j_bench_n${N}_p$P.err | head -10
head -c 300 gpurun_out/r2j_bench_n${N}_p$P.json; echo
done
