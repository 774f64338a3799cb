# one --set full capture of kernels matching $2 (skip $3 launches) -> gpurun_out/prof_$1.ncu-rep
tag=$1; pat=$2; skip=${3:-1}; cnt=${4:-1}
export SR_BENCH_IMAGE_CACHE=/tmp/sr_bench_cfg4.npz
CMD="python bench.py --steps 1 --warmup 1 --views 1 --no-cpu"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log | cut -c1-200
