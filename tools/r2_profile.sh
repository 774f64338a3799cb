# Round-2 evidence (B200_PROFILING.md recipe): the plain default bench line, the ncu launch list of a
# one-view step, one `--set full` capture of the dominant kernels (cfg4) and of the two-view kernel (cfg3).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 400 gpurun_out/r2_bench_default.json
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
export SR_LANES=1   # one view at a time on the context stream: per-launch times are each kernel's own
CMD="python bench.py --steps 1 --warmup 1 --views 1 --no-cpu --no-extras"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_raw.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'match_mvs_screen2|build_refr|weights_geodesic' -s 4 -c 3 -f -o gpurun_out/prof_r2_final $CMD > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log | cut -c1-200
unset SR_BENCH_IMAGE_CACHE
CMD3="python bench.py --workload cfg3 --steps 1 --warmup 0 --no-cpu --no-extras"
$CMD3 > gpurun_out/r2_plain_cfg3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'match_kernel' -s 0 -c 1 -f -o gpurun_out/prof_r2_cfg3 $CMD3 > gpurun_out/r2_ncu_cfg3.log 2>&1
tail -2 gpurun_out/r2_ncu_cfg3.log | cut -c1-200
python bench.py --workload cfg3 --steps 1 --warmup 1 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; tail -c 300 gpurun_out/r2_bench_cfg3.json
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
