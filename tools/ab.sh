# A/B of kernel variants on the GPU box: tools/ab.sh "" build_variants/libsr_x.so ...   ("" = the in-tree library)
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
for v in "$@"; do
  SR_LIB=$v python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', '${v:-in-tree}', round(d['value'],1), 'match', round(d['roofline']['match_ms_per_view'],3), 'build', round(d['roofline']['build_ms_per_view'],3))"
done
