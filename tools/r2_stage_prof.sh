# ncu A/B pair of the TMA-staging experiment (same occupancy: 12 warps/SM)
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
export SR_LANES=1
CMD="python bench.py --steps 1 --warmup 1 --views 1 --no-cpu --no-extras"
for v in m3 stage; do
  export SR_LIB=build_variants/libsr_$v.so
  $CMD > gpurun_out/sp_plain_$v.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:match_mvs_screen2 -s 1 -c 1 -f -o gpurun_out/prof_r2_$v $CMD > gpurun_out/sp_ncu_$v.log 2>&1
  tail -1 gpurun_out/sp_ncu_$v.log | cut -c1-120
done
