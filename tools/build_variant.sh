#!/bin/bash
# A/B library builds: tools/build_variant.sh NAME "-DFLAG1 -DFLAG2=3" [source.cu ...] -> build/lib_NAME.so (the listed sources,
# default assemble_items.cu, recompiled with the flags; every other object taken from the regular build).
# Select it at run time with REYNA_B200_LIB=$PWD/build/lib_NAME.so.
set -e
cd "$(dirname "$0")/../reyna_b200/csrc"
make -s > /dev/null
mkdir -p ../../build
NAME=$1; FLAGS=$2; shift 2 || true
SRCS=${@:-assemble_items.cu}
OBJS=""
for o in *.o; do
  keep=1; for s in $SRCS; do [ "$o" = "${s%.cu}.o" ] && keep=0; done
  [ $keep = 1 ] && OBJS="$OBJS $o"
done
for s in $SRCS; do
  b=${s%.cu}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-Wall -Xptxas -v --expt-relaxed-constexpr $FLAGS \
    -c $s -o ../../build/${b}_$NAME.o 2> ../../build/${b}_$NAME.ptxas.log
  OBJS="$OBJS ../../build/${b}_$NAME.o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/lib_$NAME.so $OBJS -lcudart -ldl
echo "built build/lib_$NAME.so"
