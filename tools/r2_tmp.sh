timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py tests/test_gpu_random.py tests/test_gpu_screen_hardening.py -m gpu -x -q 2>&1 | tail -2
bash tools/r2_bench_ab.sh bs build_variants/libsr_rect1.so intree build_variants/libsr_rect1.so intree
