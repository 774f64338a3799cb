set -x
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits 2>&1 | head -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 2 --warmup 1 --views 2 --no-cpu > gpurun_out/b_screen.log 2>&1; tail -1 gpurun_out/b_screen.log
SR_MATCH_SCREEN=0 python bench.py --steps 2 --warmup 1 --views 2 --no-cpu > gpurun_out/b_noscreen.log 2>&1; tail -1 gpurun_out/b_noscreen.log
