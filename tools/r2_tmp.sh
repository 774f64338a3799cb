bash tools/r2_bench_ab.sh bs build_variants/libsr_rect1.so intree build_variants/libsr_t5r136.so build_variants/libsr_t3r136.so intree build_variants/libsr_rect1.so
SR_LIB=build_variants/libsr_t5r136.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -2
