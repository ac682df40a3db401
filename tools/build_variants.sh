#!/bin/bash
# Builds kernel variants of libfolp_b200.so into scratch/ (git-ignored, travels with gpurun) for
# tools/probe_kernels.py:  tools/build_variants.sh name "DEFS" [name "DEFS" ...]
set -e
cd "$(dirname "$0")/../firstorderlp.jl_b200/csrc"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  make -s -j8 OUT=../../scratch/libfolp_$name.so OBJDIR=/tmp/folp_variant_$name DEFS="$defs" > /tmp/folp_variant_$name.log 2>&1 \
    || { echo "variant $name failed"; tail -20 /tmp/folp_variant_$name.log; exit 1; }
  echo "built scratch/libfolp_$name.so ($defs)"
done
