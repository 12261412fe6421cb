#!/bin/bash
# Build a variant of libg3cuda.so with extra nvcc flags (profiling / experiment builds) into gadget3_b200/csrc/variants/.
# Usage: tools/build_variant.sh NAME "-DFLAG ..."
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=/tmp/g3v_$1
mkdir -p "$W/gadget3_b200" "$W/include" "$ROOT/gadget3_b200/csrc/variants"
cp "$ROOT"/include/*.h "$W/include/"
mkdir -p "$W/gadget3_b200/csrc"
cp "$ROOT"/gadget3_b200/csrc/*.cu "$ROOT"/gadget3_b200/csrc/*.cuh "$ROOT"/gadget3_b200/csrc/*.h "$ROOT"/gadget3_b200/csrc/Makefile "$W/gadget3_b200/csrc/"
make -C "$W/gadget3_b200/csrc" -B -j8 NVCCFLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -diag-suppress 177 $2" > "$W/build.log" 2>&1
cp "$W/gadget3_b200/csrc/libg3cuda.so" "$ROOT/gadget3_b200/csrc/variants/libg3cuda_$1.so"
