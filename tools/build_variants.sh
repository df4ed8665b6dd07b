#!/bin/bash
# Build variants of libgdr.so for A/B runs on the GPU box (they travel with the snapshot; select one with GDR_LIB=path).
#   usage: tools/build_variants.sh name1 "-DFLAG=1 ..." name2 "-D..." ...
# Leaves generativedensification_b200/build/variants/<name>.so and rebuilds the default library at the end.
set -e
cd "$(dirname "$0")/.."
out=generativedensification_b200/build/variants
mkdir -p $out
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  GDR_NVCC_FLAGS="$flags" python -m generativedensification_b200.build --force > /dev/null
  cp generativedensification_b200/libgdr.so $out/$name.so
  echo "built $out/$name.so  ($flags)"
done
python -m generativedensification_b200.build --force > /dev/null
