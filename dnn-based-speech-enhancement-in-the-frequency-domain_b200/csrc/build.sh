#!/bin/bash
# Build libsefd.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../sefd/libsefd.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p build
pids=()
for f in api dccrn crn cbn lms pmsqe fsn fsnet lstm_seq lstm_step_tc lstm_cluster nccl_dp tapgemm_simt tapgemm_tc wgrad_tc skinny elementwise stft lstm prof; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer build/$f.o)" ] || [ ../../include/sefd.h -nt build/$f.o ]; then
    nvcc $FLAGS -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/*.o -lcudart -ldl
echo "built $OUT"
