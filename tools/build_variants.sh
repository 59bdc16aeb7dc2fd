#!/bin/bash
# builds crixus_b200/build_exp/lib_<tag>.so for kernel experiments (tools/time_vv.py picks one through CX_LIB_PATH)
# usage: tools/build_variants.sh tag1="vert_volume:-DVV_LB_LINES=8" tag2="fill:-DFILL_LB=3" ...     (source file : nvcc flags)
set -e
cd "$(dirname "$0")/../crixus_b200"
mkdir -p build_exp
(cd .. && python -m crixus_b200.build >/dev/null 2>&1)
for spec in "$@"; do
  tag="${spec%%=*}"; rest="${spec#*=}"; src="${rest%%:*}"; flags="${rest#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-fno-fast-math $flags -c csrc/$src.cu -o build_exp/${src}_$tag.o &
done
wait
for spec in "$@"; do
  tag="${spec%%=*}"; rest="${spec#*=}"; src="${rest%%:*}"
  objs=$(ls build/*.o | grep -v "/$src.cu.o")
  nvcc -shared -o build_exp/lib_$tag.so $objs build_exp/${src}_$tag.o -lcudart -ldl 2>/dev/null
done
ls build_exp/*.so
