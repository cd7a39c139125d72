set -x
mkdir -p gpurun_out
for S in 0.25 1; do
CRF_LIB_PATH=profiles/ab/libcrf_base.so timeout 300 python profiles/prof_scan.py --scale $S --reps 4 > gpurun_out/r2k_prof_base_$S.txt 2>&1; tail -1 gpurun_out/r2k_prof_base_$S.txt
for K in block block2; do
  CRF_SCAN_KERNEL=$K timeout 300 python profiles/prof_scan.py --scale $S --reps 4 > gpurun_out/r2k_prof_${K}_$S.txt 2>&1
  tail -1 gpurun_out/r2k_prof_${K}_$S.txt
done
done
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/r2k_parity_block.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_parity_block.log
tail -3 gpurun_out/r2k_parity_block.log
( CRF_SCAN_KERNEL=block2 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/r2k_parity_block2.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_parity_block2.log
tail -3 gpurun_out/r2k_parity_block2.log
for W in sr s22; do for K in block block2; do
CRF_SCAN_KERNEL=$K timeout 300 python profiles/prof_scan.py --workload $W --scale 1 --reps 4 > gpurun_out/r2k_prof_${W}_$K.txt 2>&1; tail -1 gpurun_out/r2k_prof_${W}_$K.txt
done; done
( timeout 600 python -m pytest tests/test_ref_python.py -x -q -m gpu ) > gpurun_out/r2k_ref.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_ref.log
tail -5 gpurun_out/r2k_ref.log
