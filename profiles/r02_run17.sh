set -x
mkdir -p gpurun_out
for S in 0.25 1; do
CRF_SCAN_WEAK_FILTERS=1 timeout 300 python profiles/prof_scan.py --scale $S --reps 4 > gpurun_out/r2q_prof_weak_$S.txt 2>&1; tail -1 gpurun_out/r2q_prof_weak_$S.txt
timeout 300 python profiles/prof_scan.py --scale $S --reps 4 > gpurun_out/r2q_prof_strong_$S.txt 2>&1; tail -1 gpurun_out/r2q_prof_strong_$S.txt
done
for W in sr s22; do
CRF_SCAN_WEAK_FILTERS=1 timeout 300 python profiles/prof_scan.py --workload $W --scale 1 --reps 4 > gpurun_out/r2q_prof_${W}_weak.txt 2>&1; tail -1 gpurun_out/r2q_prof_${W}_weak.txt
timeout 300 python profiles/prof_scan.py --workload $W --scale 1 --reps 4 > gpurun_out/r2q_prof_${W}_strong.txt 2>&1; tail -1 gpurun_out/r2q_prof_${W}_strong.txt
done
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/r2q_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_parity.log
tail -3 gpurun_out/r2q_parity.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_b.json 2> gpurun_out/r2q_b.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|gather_kernel|tile_offsets|translate_kernel|spill_sort|pack_kernel" -c 200 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_ncu_launch.log 2>&1
