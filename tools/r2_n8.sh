# 8-GPU evidence: multi-GPU parity test at 2/4/8, cfg5 row-sharded (BASELINE configs[4]) with parity vs 1 GPU
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload cfg5 --partition rows --steps 2 --warmup 1 --no-cpu > gpurun_out/n${N}_cfg5_rows.json 2> gpurun_out/n${N}_cfg5_rows.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/n${N}_cfg5_rows.json') if l.startswith('{')][-1])
    print('cfg5 rows N=$N', 'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1))
    print(' job',d['job'] and {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['job'].items() if k!='what'})
    print(' parity',d['parity_vs_1gpu'] and {k:v for k,v in d['parity_vs_1gpu'].items() if k!='checked'})
    s=d['secondary']; print(' secondary', s and (s['partition'], round(s['value'],1), s['parity_vs_1gpu'] and s['parity_vs_1gpu']['identical']))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/n${N}_cfg5_rows.err').read()[-2000:])
PY
