set -x
python bench.py --steps 1 --warmup 1 --views 1 --no-cpu > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:build_refr -s 3 -c 1 -o gpurun_out/prof_r1j python bench.py --steps 1 --warmup 1 --views 1 --no-cpu > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c1-200
