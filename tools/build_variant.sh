#!/bin/bash
# Builds an A/B variant of the library: tools/build_variant.sh NAME FILE.cu "-DFLAG=..." -> lattice_symmetries_b200/variants/NAME.so
# (recompiles FILE.cu with the extra flags, links it with the objects of the main build). Load with LS_B200_LIBRARY=...
set -e
NAME=$1; SRC=$2; FLAGS=$3
cd "$(dirname "$0")/../lattice_symmetries_b200/csrc"
mkdir -p ../variants build/variant_$NAME
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcudafe --diag_suppress=177 $FLAGS -c $SRC -o build/variant_$NAME/${SRC%.cu}.o
OBJS=""
for f in runtime group index basis_build matvec operator_apply peaks abi_layout; do
  if [ "$f.cu" == "$SRC" ]; then OBJS="$OBJS build/variant_$NAME/$f.o"; else OBJS="$OBJS build/$f.o"; fi
done
nvcc $ARCH -shared -o ../variants/$NAME.so $OBJS -lcudart_static -lpthread -ldl -lrt
echo built ../variants/$NAME.so
