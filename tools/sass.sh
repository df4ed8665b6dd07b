#!/bin/bash
# usage: tools/sass.sh <file.cu under csrc> -> /tmp/t/<name>.sass (+ ptxas -v summary)
set -e
name=$(basename "$1" .cu)
mkdir -p /tmp/t
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v \
  -c /root/repo/generativedensification_b200/csrc/$name.cu -o /tmp/t/$name.o 2>&1 | grep -E "error|registers|spill" || true
cuobjdump -sass /tmp/t/$name.o | grep -E "^\s*/\*[0-9a-f]{4}\*/|Function" | sed 's#/\* 0x[0-9a-f]* \*/##' | awk '{$1=$1};1' > /tmp/t/$name.sass
wc -l /tmp/t/$name.sass
