set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest.log
tail -8 gpurun_out/r2m_pytest.log
python bench.py --workload s22 --no-cpu-baseline > gpurun_out/r2m_bench_s22.json 2> gpurun_out/r2m_bench_s22.err; echo rc=$?
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo rc=$?
python - <<'PY'
import json
for f in ['_s22','']:
    try:
        d=json.load(open(f'gpurun_out/r2m_bench{f}.json'))
        print(f, 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'scan_ms', round(d['scan_stats']['scan_ms'],4), 'e2e', d['e2e']['ms_per_step'], d['parity'])
    except Exception as e: print(f,'ERR',e)
PY
