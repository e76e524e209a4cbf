#!/bin/bash
# build_variant.sh NAME SRC.cu [extra nvcc flags]: libmfgp variant with gpr_small_v4.cu replaced by SRC.cu -> variants/libmfgp_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p variants
P=multi_fidelity_gpflow_b200
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I$P/csrc "$@" -c "$src" -o variants/v4_$name.o
objs=$(ls $P/build/*.o | grep -v gpr_small_v4.o)
nvcc -shared -o variants/libmfgp_$name.so $objs variants/v4_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo variants/libmfgp_$name.so
