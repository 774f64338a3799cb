timeout 1500 python bench.py --workload cfg3 --steps 1 --warmup 1 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; echo rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_cfg3.json') if l.startswith('{')][-1])
print('cfg3 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],1))
print(d['cpu_baseline']); print(d['like_for_like'])
PY
