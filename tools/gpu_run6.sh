set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in mb4 mb5 mb6; do SR_LIB=$PWD/stereoreconstruction_b200/variants/lib_$v.so SR_MATCH_STATS=1 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2>gpurun_out/err_$v.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"; done
grep "match stats" gpurun_out/err_mb4.log
