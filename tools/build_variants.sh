#!/bin/bash
# Development aid: build libflux3d_b200 variants with experiment macros into build/variants/ (timed by tools/time_variants.py).
# usage: tools/build_variants.sh name1:"-DFOO -DBAR" name2:"-DBAZ" ...
set -e
cd "$(dirname "$0")/../flux3d.jl_b200/csrc"
mkdir -p ../../build/variants
rm -f ../../build/variants/*.so
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  d=$(mktemp -d)
  for f in capi chamfer chamfer_tc chamfer_pipe chamfer_bwd knn knn_tc knn_gram mesh sampling comm; do
    if [ "$f" = chamfer ] || [ "$f" = chamfer_bwd ] || [ "$f" = chamfer_tc ] || [ "$f" = knn ] || [ "$f" = knn_tc ] || [ "$f" = knn_gram ]; then
      /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -I../../include $defs -c -o $d/$f.o $f.cu &
    else
      cp $f.o $d/$f.o
    fi
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/$name.so $d/*.o -ldl
  rm -rf $d
  echo built $name
done
