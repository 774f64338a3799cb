timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/h_pytest.log
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras 2> gpurun_out/h_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'e2e s', round(d['e2e']['seconds_per_step'],4))"
