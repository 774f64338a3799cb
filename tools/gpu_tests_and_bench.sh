python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 2 --warmup 2 --views 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
