set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_b.json 2> gpurun_out/r2p_b.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"crf::" -c 200 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_ncu_launch.log 2>&1
for W in sr s22; do
python profiles/prof_scan.py --workload $W --scale 1 --reps 3 > gpurun_out/r2p_prof_$W.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/r2p_${W}_scan python profiles/prof_scan.py --workload $W --scale 1 --reps 3 > gpurun_out/r2p_ncu_$W.log 2>&1
done
python bench.py --workload sr --no-cpu-baseline > gpurun_out/r2p_bench_sr.json 2>/dev/null
python bench.py --workload s22 --no-cpu-baseline > gpurun_out/r2p_bench_s22.json 2>/dev/null
