# Builds a library variant for A/B runs (SR_LIB=...):  tools/build_variant.sh <name> [src_dir] [-Dflags ...]
#   src_dir: a tree holding stereoreconstruction_b200/csrc (default: this repository), e.g. a `git archive` of an older commit
name=$1; shift
src=.
if [ -d "$1" ]; then src=$1; shift; fi
mkdir -p build_variants
nvcc -std=c++17 -O3 -fmad=false -DSR_FEW_RADII -gencode arch=compute_100a,code=sm_100a -lineinfo --shared -Xcompiler -fPIC "$@" \
  -o build_variants/libsr_$name.so $src/stereoreconstruction_b200/csrc/sr_capi.cu -lcudart -ldl && echo built build_variants/libsr_$name.so
