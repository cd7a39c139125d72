set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_pytest.log
tail -6 gpurun_out/r2s_pytest.log
python profiles/prof_scan.py --scale 1 --reps 4 > gpurun_out/r2s_prof_s38.txt 2>&1; tail -1 gpurun_out/r2s_prof_s38.txt
python profiles/prof_scan.py --scale 1 --reps 3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/r2s_s38_scan python profiles/prof_scan.py --scale 1 --reps 3 > gpurun_out/r2s_ncu_s38.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_b.json 2> gpurun_out/r2s_b.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|gather_kernel|tile_offsets|translate_kernel|spill_sort|pack_kernel" -c 200 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_ncu_launch.log 2>&1
python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err; echo rc=$?
python bench.py --workload sr --no-cpu-baseline > gpurun_out/r2s_bench_sr.json 2>/dev/null
python bench.py --workload s22 --no-cpu-baseline > gpurun_out/r2s_bench_s22.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()"
