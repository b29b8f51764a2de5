#!/bin/bash
# Builds the library of a given commit (default HEAD) as path_tracer_python_b200/libb200pt_base.so, the
# "A" side of tools/ab.py when the working tree holds the "B" side (development aid).
set -e
REV=${1:-HEAD}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
git -C "$ROOT" archive "$REV" path_tracer_python_b200/csrc include | tar -x -C "$T"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -ftz=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC -Xcompiler -fno-fast-math -Xcompiler -ffp-contract=off -shared -cudart shared -ccbin /usr/bin/g++ \
  -o "$ROOT/path_tracer_python_b200/libb200pt_base.so" "$T/path_tracer_python_b200/csrc/b200pt.cu" "$T/path_tracer_python_b200/csrc/pt_host.cpp"
rm -rf "$T"
echo "$ROOT/path_tracer_python_b200/libb200pt_base.so"
