# bench A/B over environment settings with the in-tree library: tools/r2_envab.sh <tag> "ENV=.. ENV=.." ...
tag=$1; shift
mkdir -p gpurun_out
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
for e in "$@"; do
  env $e timeout 600 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2> gpurun_out/${tag}_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', '$e', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'match', round(d['roofline']['match_ms_per_view'],3), 'build', round(d['roofline']['build_ms_per_view'],3))"
done
