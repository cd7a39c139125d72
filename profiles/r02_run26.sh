set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2aa_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2aa_pytest.log
tail -6 gpurun_out/r2aa_pytest.log
python bench.py > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo rc=$?
python bench.py --workload sr --no-cpu-baseline > gpurun_out/r2aa_bench_sr.json 2>/dev/null
python bench.py --workload s22 --no-cpu-baseline > gpurun_out/r2aa_bench_s22.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()"
