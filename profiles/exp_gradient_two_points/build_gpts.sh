#!/bin/bash
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
S=$ROOT/spectralkit.jl_b200/csrc
NAME=$1; shift
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -diag-suppress 128 -Xcompiler -fPIC -ccbin /usr/bin/g++ -I$ROOT/include -I$S -I$ROOT/exp"
nvcc $NVF "$@" -Xptxas -v -c $ROOT/exp/gpts_variant.cu -o $ROOT/exp/obj/gpts_$NAME.o 2> $ROOT/exp/obj/gpts_$NAME.ptxas &
if [ ! -f $ROOT/exp/obj/leaf_hook.o ]; then nvcc $NVF -DSK_GPTS_HOOK -c $S/sk_leaf.cu -o $ROOT/exp/obj/leaf_hook.o; fi &
wait
OBJS=$(ls $S/*.o | grep -v "sk_leaf.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $ROOT/exp/lib_gpts_$NAME.so $OBJS $ROOT/exp/obj/leaf_hook.o $ROOT/exp/obj/gpts_$NAME.o -lquadmath -ldl
grep -E "Used|spill" $ROOT/exp/obj/gpts_$NAME.ptxas | paste - - | head -3
