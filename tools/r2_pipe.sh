# pipeline kernel: parity tests, then bench with and without it, lag sweep
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
run() { env "$@" timeout 600 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2> gpurun_out/${tag}_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', '$*', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'match', round(d['roofline']['match_ms_per_view'],3), 'build', round(d['roofline']['build_ms_per_view'],3))"; }
run SR_PIPELINE=0
run SR_PIPELINE=1
run SR_PIPE_LAG=32
run SR_PIPE_LAG=64
run SR_PIPE_LAG=256
run SR_PIPE_LAG=512
run SR_PIPE_LAG=64 SR_PIPE_RING_MB=160
SR_MATCH_STATS=1 timeout 600 python bench.py --steps 1 --warmup 1 --views 2 --no-cpu 2>&1 | grep "stats:"
