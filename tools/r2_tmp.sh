timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py tests/test_gpu_random.py -m gpu -x -q 2>&1 | tail -2
bash tools/r2_bench_ab.sh bs build_variants/libsr_head.so intree build_variants/libsr_s6.so build_variants/libsr_s8.so intree build_variants/libsr_head.so
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
for v in "" build_variants/libsr_s6.so build_variants/libsr_s8.so; do
SR_LIB=$v SR_MATCH_STATS=1 timeout 600 python bench.py --steps 1 --warmup 1 --views 2 --no-cpu 2>&1 | grep "stats:"
done
