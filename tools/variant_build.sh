#!/bin/bash
# A/B builds of the CUDA library with different -D knobs: tools/variant_build.sh name "-DHS_PRE_BAR=1" -> build/libhs_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $2 \
     -o build/libhs_$1.so cpp-optical-flow_b200/csrc/hs_api.cu
