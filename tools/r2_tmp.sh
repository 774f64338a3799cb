for v in build_variants/libsr_wstage.so ""; do
SR_LIB=$v timeout 600 python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3', '${v:-in-tree}', round(d['value'],1), 'ms/step', round(d['ms_per_step'],1))"
done
bash tools/r2_bench_ab.sh bs build_variants/libsr_wstage.so intree
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_random.py tests/test_bunny.py tests/test_bunny_full.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
