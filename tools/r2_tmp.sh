bash tools/r2_bench_ab.sh bs intree build_variants/libsr_fix.so intree build_variants/libsr_fix.so
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
SR_LIB=build_variants/libsr_fix.so SR_MATCH_STATS=1 timeout 600 python bench.py --steps 1 --warmup 1 --views 2 --no-cpu 2>&1 | grep "build stats:"
SR_LIB=build_variants/libsr_fix.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py tests/test_gpu_random.py tests/test_bunny_full.py -m gpu -x -q 2>&1 | tail -2
