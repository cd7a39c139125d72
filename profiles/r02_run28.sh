set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/r2ac_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2ac_parity.log; tail -3 gpurun_out/r2ac_parity.log
for S in 0.25 1; do
CRF_LIB_PATH=profiles/ab/libcrf_prev.so timeout 300 python profiles/prof_scan.py --scale $S --reps 4 > gpurun_out/r2ac_prof_prev_$S.txt 2>&1; tail -1 gpurun_out/r2ac_prof_prev_$S.txt
timeout 300 python profiles/prof_scan.py --scale $S --reps 4 > gpurun_out/r2ac_prof_new_$S.txt 2>&1; tail -1 gpurun_out/r2ac_prof_new_$S.txt
done
for W in sr s22; do
timeout 300 python profiles/prof_scan.py --workload $W --scale 1 --reps 4 > gpurun_out/r2ac_prof_${W}.txt 2>&1; tail -1 gpurun_out/r2ac_prof_${W}.txt
done
