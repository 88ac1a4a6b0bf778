#!/usr/bin/env bash
# profiles/tools/ab.sh "<sides>" <variant> [variant ...] — on the GPU box: for every exp_libs/libpbf_<variant>.so, copy it over
# the product library and run size_sweep.py on the given dam-break sides; prints one line per (variant, side).
set -u
SIDES=$1; shift
cp pbf_sph_b200/libpbf_cuda.so /tmp/libpbf_cuda.orig.so
for v in "$@"; do
  cp exp_libs/libpbf_$v.so pbf_sph_b200/libpbf_cuda.so
  for s in $SIDES; do
    echo -n "$v "; timeout 300 python profiles/tools/size_sweep.py $s 2>&1 | tail -1
  done
done
cp /tmp/libpbf_cuda.orig.so pbf_sph_b200/libpbf_cuda.so
