set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest.log
tail -15 gpurun_out/r2l_pytest.log
for W in s22 sr; do
python bench.py --workload $W --no-cpu-baseline > gpurun_out/r2l_bench_$W.json 2> gpurun_out/r2l_bench_$W.err; echo rc=$?
CRF_NO_GRAPH=1 python bench.py --workload $W --no-cpu-baseline > gpurun_out/r2l_bench_${W}_nograph.json 2>/dev/null; echo rc=$?
done
python - <<'PY'
import json
for f in ['s22','s22_nograph','sr','sr_nograph']:
    try:
        d=json.load(open(f'gpurun_out/r2l_bench_{f}.json'))
        print(f, 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'scan_ms', round(d['scan_stats']['scan_ms'],4), 'tiles', d['scan_stats']['tiles'])
    except Exception as e: print(f,'ERR',e)
PY
