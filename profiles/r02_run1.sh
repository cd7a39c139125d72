set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
nproc
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo rc=$?
python bench.py --workload sr --no-cpu-baseline > gpurun_out/r2a_bench_sr.json 2>/dev/null
python bench.py --workload s22 --no-cpu-baseline > gpurun_out/r2a_bench_s22.json 2>/dev/null
python profiles/prof_scan.py --workload sr --scale 1 --reps 3 > gpurun_out/r2a_prof_sr.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/r2a_sr_scan python profiles/prof_scan.py --workload sr --scale 1 --reps 3 > gpurun_out/r2a_ncu_sr.log 2>&1
python profiles/prof_scan.py --workload s22 --scale 1 --reps 3 > gpurun_out/r2a_prof_s22.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/r2a_s22_scan python profiles/prof_scan.py --workload s22 --scale 1 --reps 3 > gpurun_out/r2a_ncu_s22.log 2>&1
python profiles/prof_scan.py --workload s38 --scale 0.25 --reps 3 > gpurun_out/r2a_prof_s38.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/r2a_s38_scan python profiles/prof_scan.py --workload s38 --scale 0.25 --reps 3 > gpurun_out/r2a_ncu_s38.log 2>&1
ls -la gpurun_out
