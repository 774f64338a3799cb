bash tools/r2_bench_ab.sh bs build_variants/libsr_diet0.so intree build_variants/libsr_mb5.so intree
export SR_LANES=1
bash tools/r2_prof.sh diet2 build_refr 1 1
