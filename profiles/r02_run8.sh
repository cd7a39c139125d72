set -x
mkdir -p gpurun_out
for K in block warp1 warp2; do
  CRF_SCAN_KERNEL=$K timeout 300 python profiles/prof_scan.py --scale 0.25 --reps 4 > gpurun_out/r2h_prof_$K.txt 2>&1
  tail -2 gpurun_out/r2h_prof_$K.txt
done
( CRF_SCAN_KERNEL=warp2 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/r2h_parity_warp2.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_parity_warp2.log
tail -5 gpurun_out/r2h_parity_warp2.log
( CRF_SCAN_KERNEL=warp1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_vectors or fuzz" ) > gpurun_out/r2h_parity_warp1.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_parity_warp1.log
tail -5 gpurun_out/r2h_parity_warp1.log
