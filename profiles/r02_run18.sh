set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/r2r_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_parity.log; tail -3 gpurun_out/r2r_parity.log
for S in 0.25 1; do
CRF_SCAN_WEAK_FILTERS=1 timeout 300 python profiles/prof_scan.py --scale $S --reps 4 > gpurun_out/r2r_prof_weak_$S.txt 2>&1; tail -1 gpurun_out/r2r_prof_weak_$S.txt
timeout 300 python profiles/prof_scan.py --scale $S --reps 4 > gpurun_out/r2r_prof_strong_$S.txt 2>&1; tail -1 gpurun_out/r2r_prof_strong_$S.txt
done
for K in warp1 warp2 block2; do
CRF_SCAN_KERNEL=$K timeout 300 python profiles/prof_scan.py --scale 0.25 --reps 4 > gpurun_out/r2r_prof_$K.txt 2>&1; tail -1 gpurun_out/r2r_prof_$K.txt
done
for W in sr s22; do
timeout 300 python profiles/prof_scan.py --workload $W --scale 1 --reps 4 > gpurun_out/r2r_prof_${W}.txt 2>&1; tail -1 gpurun_out/r2r_prof_${W}.txt
done


