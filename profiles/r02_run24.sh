set -x
mkdir -p gpurun_out
for V in prev 3cta; do
CRF_LIB_PATH=profiles/ab/libcrf_$V.so timeout 300 python profiles/prof_scan.py --scale 1 --reps 4 > gpurun_out/r2y_prof_$V.txt 2>&1; tail -1 gpurun_out/r2y_prof_$V.txt
done
