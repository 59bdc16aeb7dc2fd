#!/bin/bash
# builds crixus_b200/build_exp/lib_<k>.so with -DVV_EXP=<k> for kernel experiments (tools/time_vv.py picks one through CX_LIB_PATH)
set -e
cd "$(dirname "$0")/../crixus_b200"
mkdir -p build_exp
python -m crixus_b200.build >/dev/null 2>&1 || (cd .. && python -m crixus_b200.build >/dev/null)
for k in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-fno-fast-math -DVV_EXP=$k -c csrc/vert_volume.cu -o build_exp/vv_$k.o &
done
wait
for k in "$@"; do
  objs=$(ls build/*.o | grep -v vert_volume)
  nvcc -shared -o build_exp/lib_$k.so $objs build_exp/vv_$k.o -lcudart -ldl 2>/dev/null
done
ls -la build_exp/*.so
