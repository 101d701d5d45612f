#!/bin/bash
# Builds kernel variants of libsnag_b200.so for A/B power/throughput measurements (scripts/gpu_variants.py).
set -e
cd "$(dirname "$0")/.."
mkdir -p snag_b200/_variants
build() {  # name, extra flags
  name=$1; shift
  objs=""
  for f in abi bw_kernels sim_kernels; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
      -c snag_b200/csrc/$f.cu -o snag_b200/_variants/${name}_$f.o &
    objs="$objs snag_b200/_variants/${name}_$f.o"
  done
  wait
  nvcc -shared -o snag_b200/_variants/lib_$name.so $objs -gencode arch=compute_100a,code=sm_100a
  rm -f $objs
}
build base
build wg2 -DSNAG_EPI_WG=2
build single -DSNAG_CTRL_CONVERGED=0
build hint1000 -DSNAG_TRYWAIT_HINT_NS=1000
build wg2_hint -DSNAG_EPI_WG=2 -DSNAG_TRYWAIT_HINT_NS=1000
ls -la snag_b200/_variants/
