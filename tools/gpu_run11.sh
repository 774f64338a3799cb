SR_LIB=$PWD/stereoreconstruction_b200/variants/lib_s8.so SR_MATCH_STATS=1 python bench.py --steps 1 --warmup 1 --views 2 --no-cpu 2>gpurun_out/err_s8.log | tail -1 > /dev/null; echo "stride 8:"; grep "build stats" gpurun_out/err_s8.log
SR_MATCH_STATS=1 python bench.py --steps 1 --warmup 1 --views 2 --no-cpu 2>gpurun_out/err_s4.log | tail -1 > /dev/null; echo "stride 4:"; grep "build stats" gpurun_out/err_s4.log
python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('s4', d['value'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
python -m pytest tests -m gpu -q -x -k "build or fullsize or mvs" 2>&1 | tail -2
