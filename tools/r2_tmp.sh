export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
for e in SR_LANES=2 SR_LANES=3 SR_LANES=4; do
env $e python bench.py --steps 3 --warmup 2 --no-cpu --no-extras 2> gpurun_out/l_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', '$e', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
done
