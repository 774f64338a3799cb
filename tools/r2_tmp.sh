timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py tests/test_gpu_random.py -m gpu -x -q 2>&1 | tail -2
bash tools/r2_bench_ab.sh bs build_variants/libsr_diet0.so intree build_variants/libsr_geo6.so build_variants/libsr_geo8.so intree
PYTHONPATH=. timeout 300 python tools/curve_time.py 2>&1 | tail -3
