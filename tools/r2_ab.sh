# Round 2 A/B on the GPU box: parity tests with the in-tree library, then bench A/B of library variants.
#   tools/r2_ab.sh <tag> [variant.so ...]      ("" = the in-tree library)
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
for v in "" "$@"; do
  SR_LIB=$v timeout 600 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2> gpurun_out/${tag}_err_$(basename "${v:-intree}").log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', '${v:-in-tree}', round(d['value'],1), 'match', round(d['roofline']['match_ms_per_view'],3), 'build', round(d['roofline']['build_ms_per_view'],3))"
done
SR_MATCH_STATS=1 timeout 600 python bench.py --steps 1 --warmup 1 --views 2 --no-cpu 2>&1 | grep "stats:" 
