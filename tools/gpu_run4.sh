set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in 32 64 128; do SR_BUILD_CHUNK=$c python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk $c', d['value'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"; done
SR_BUILD_REFR=0 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old build', d['value'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
