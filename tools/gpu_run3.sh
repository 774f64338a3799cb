set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
SR_MATCH_STATS=1 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu > gpurun_out/b_stats.log 2>&1; grep "match stats" gpurun_out/b_stats.log; tail -1 gpurun_out/b_stats.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
