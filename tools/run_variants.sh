#!/bin/bash
# usage: tools/run_variants.sh "<configs>" k1 k2 ...   (default library first)
cfgs="$1"; shift
for c in $cfgs; do
  python tools/time_vv.py ${c/:/ }
  for k in "$@"; do CX_LIB_PATH=$PWD/crixus_b200/build_exp/lib_$k.so python tools/time_vv.py ${c/:/ }; done
done
