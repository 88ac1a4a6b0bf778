#!/usr/bin/env bash
# profiles/tools/build_variant.sh <name> "<-D flags>" [file.cu ...] — A/B experiments: rebuild the named sources (default:
# neighbour_list.cu) with extra -D flags and link them with the product's other objects into exp_libs/libpbf_<name>.so.
# On the GPU box an experiment copies that file over pbf_sph_b200/libpbf_cuda.so (the box copy is scratch).  Never shipped.
set -eu
NAME=$1; FLAGS=$2; shift 2
FILES=${@:-neighbour_list.cu}
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
CS=$ROOT/pbf_sph_b200/csrc
make -C $CS -j8 >/dev/null
mkdir -p $ROOT/exp_libs $CS/build/var_$NAME
OBJS=""
for o in $CS/build/*.o; do
  b=$(basename $o .o)
  if [[ " $FILES " == *" $b.cu "* ]]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin g++ \
      -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -I$ROOT/include -I$CS $FLAGS -c $CS/$b.cu -o $CS/build/var_$NAME/$b.o
    OBJS="$OBJS $CS/build/var_$NAME/$b.o"
  else
    OBJS="$OBJS $o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/exp_libs/libpbf_$NAME.so $OBJS -lcudart -ldl
echo built exp_libs/libpbf_$NAME.so
