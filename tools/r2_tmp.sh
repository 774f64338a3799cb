for wlk in cfg3 cfg5; do
timeout 1200 python bench.py --workload $wlk --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/w_$wlk.json 2> gpurun_out/w_$wlk.err; echo "$wlk rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/w_$wlk.json') if l.startswith('{')][-1])
    print('$wlk', 'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1),'match',round(d['roofline']['match_ms_per_view'],2),'build',round(d['roofline']['build_ms_per_view'],2), 'launches', d['gpu_launches'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/w_$wlk.err').read()[-1500:])
PY
done
