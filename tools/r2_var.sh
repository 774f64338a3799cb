# parity tests + bench + stats with one library variant: tools/r2_var.sh <tag> <lib.so|intree>
tag=$1; v=$2; [ "$v" = "intree" ] && v=""
mkdir -p gpurun_out
export SR_LIB=$v
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
timeout 600 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2> gpurun_out/${tag}_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', '${v:-in-tree}', round(d['value'],1), 'match', round(d['roofline']['match_ms_per_view'],3), 'build', round(d['roofline']['build_ms_per_view'],3))"
SR_MATCH_STATS=1 timeout 600 python bench.py --steps 1 --warmup 1 --views 2 --no-cpu 2>&1 | grep "stats:"
