#!/bin/sh
# Builds a variant of libogrecu.so: the named sources recompiled with extra nvcc defines, every other object taken from the
# regular build, into build/variants/<name>/libogrecu.so (load with OGRECU_LIB=...).  Kernel tuning A/B only.
# usage: sh tools/build_variant.sh name "-DLN_MAXB_N=20 -DLN_NBUF_N=1" ["skin_lane.cu ..."]
set -e
NAME=$1; EXTRA=$2; FILES=${3:-skin_lane.cu}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build/variants/$NAME
mkdir -p $OUT
cd $ROOT/ogre_b200/csrc
cp *.o $OUT/
for f in $FILES; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -ccbin /usr/bin/g++ $EXTRA -c -o $OUT/${f%.cu}.o $f &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libogrecu.so $OUT/*.o -Xlinker -soname,libogrecu.so
rm -f $OUT/*.o
echo "variant $NAME -> $OUT/libogrecu.so"
