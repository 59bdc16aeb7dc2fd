#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the CPU restatement oracle/crixus_oracle.c -> oracle/liboracle.so
# -ffp-contract=off: every fused multiply-add in the oracle is an explicit fmaf() (see the file header).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
gcc -std=c11 -O2 -fopenmp -mfma -ffp-contract=off -fno-fast-math -fPIC -shared -Wall -Wno-unused-function \
    "$HERE/crixus_oracle.c" -o "$HERE/liboracle.so" -lm
g++ -std=c++17 -O1 -fPIC -shared -Wall "$HERE/stl_reader.cpp" -o "$HERE/liboracle_stl.so"
echo "built $HERE/liboracle.so $HERE/liboracle_stl.so"
