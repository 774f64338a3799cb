# Round-end evidence: the plain bench line, the ncu launch list of the same command, and one
# `--set full` capture of the two dominant kernels (B200_PROFILING.md recipe).
set -x
CMD="python bench.py --steps 1 --warmup 1 --views 1 --no-cpu"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-300
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz  # the profiled runs below render the scene once
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'match_mvs|build_refr' -s 3 -c 2 -o gpurun_out/prof_r1_final $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
