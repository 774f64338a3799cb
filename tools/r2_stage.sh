# TMA-staging experiment: parity of the staged variant, then A/B at the same occupancy (12 warps/SM)
mkdir -p gpurun_out
SR_LIB=build_variants/libsr_stage.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_screen_hardening.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
for v in build_variants/libsr_m3.so build_variants/libsr_stage.so ""; do
  SR_LANES=1 SR_LIB=$v timeout 300 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu --no-extras 2> gpurun_out/s_err_$(basename "${v:-intree}").log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', '${v:-in-tree}', round(d['value'],1), 'match', round(d['roofline']['match_ms_per_view'],3), 'build', round(d['roofline']['build_ms_per_view'],3))"
done
