# full parity suite + the default bench line (as the driver runs it) + reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_pytest.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],2))
print('match',d['roofline']['match_ms_per_view'],'build',d['roofline']['build_ms_per_view'])
print('cpu',d['cpu_baseline'] and {k:d['cpu_baseline'][k] for k in ('value','cores','parity')})
print('ref',d['cpu_baseline'] and d['cpu_baseline'].get('reference_itself'))
print('like',d['like_for_like']); print('job',d['job'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-400
