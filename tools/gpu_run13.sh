python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 2 --warmup 2 --views 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('single', d['value'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
SR_LIB=$PWD/stereoreconstruction_b200/variants/lib_pair.so python bench.py --steps 2 --warmup 2 --views 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pair', d['value'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
SR_LIB=$PWD/stereoreconstruction_b200/variants/lib_pair.so python -m pytest tests -m gpu -q -x -k "mvs or fullsize or random or bunny" 2>&1 | tail -2
