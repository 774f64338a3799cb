# bench-only A/B of library variants (no parity tests):  tools/r2_bench_ab.sh <tag> variant.so ...
tag=$1; shift
mkdir -p gpurun_out
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
for v in "$@"; do
  [ "$v" = "intree" ] && v=""
  SR_LIB=$v timeout 600 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2> gpurun_out/${tag}_err_$(basename "${v:-intree}").log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', '${v:-in-tree}', round(d['value'],1), 'match', round(d['roofline']['match_ms_per_view'],3), 'build', round(d['roofline']['build_ms_per_view'],3))"
done
