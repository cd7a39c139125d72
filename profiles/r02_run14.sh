set -x
mkdir -p gpurun_out
python profiles/prof_scan.py --scale 1 --reps 3 > gpurun_out/r2n_prof_s38.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/r2n_s38_scan python profiles/prof_scan.py --scale 1 --reps 3 > gpurun_out/r2n_ncu_s38.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_b.json 2> gpurun_out/r2n_b.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_ncu_launch.log 2>&1
ls -la gpurun_out | tail -5
