# N-GPU legs: multi-GPU parity test, bench with both partitions, reference arm under torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
for part in views rows; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --partition $part > gpurun_out/n${N}_${part}.json 2> gpurun_out/n${N}_${part}.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/n${N}_${part}.json') if l.startswith('{')][-1])
    print('$part', 'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'job',d['job'] and {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['job'].items() if k!='what'})
    print(' parity',d['parity_vs_1gpu'])
    s=d['secondary']; print(' secondary', s and (s['partition'], round(s['value'],1), s['parity_vs_1gpu']))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/n${N}_${part}.err').read()[-1500:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-300
