#!/bin/bash
# Builds a slim variant of the library for kernel experiments: exp/lib_<name>.so holds everything except the unrolled-leaf
# translation units, of which only the InteriorGrid2 gradient one (sk_leaf_g2g) is compiled — with the extra nvcc flags
# given — and the others are stubs.  Use with SK_B200_LIB=exp/lib_<name>.so SK_LEAF_NO_SINGLE=1.
# usage: profiles/build_variant.sh <name> [extra nvcc flags...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
S=$ROOT/spectralkit.jl_b200/csrc
NAME=$1; shift
mkdir -p $ROOT/exp/obj
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -diag-suppress 128 -Xcompiler -fPIC -ccbin /usr/bin/g++ -I$ROOT/include -I$S"
SRC=${SK_VARIANT_SRC:-$S/sk_leaf_g2g.cu}
nvcc $NVF "$@" -c $SRC -o $ROOT/exp/obj/g2g_$NAME.o
if [ ! -f $ROOT/exp/obj/stubs.o ]; then
cat > $ROOT/exp/obj/stubs.cu <<'EOS'
#include "sk_leaf_impl.cuh"
namespace skk {
#define STUB1(n) cudaError_t n(int, const LeafParams &, cudaStream_t) { return cudaErrorNotSupported; }
#define STUB2(n) cudaError_t n(int, int, const LeafParams &, cudaStream_t) { return cudaErrorNotSupported; }
STUB1(dispatch_leaf_g0v) STUB1(dispatch_leaf_g0g) STUB1(dispatch_leaf_g1v) STUB1(dispatch_leaf_g1g) STUB1(dispatch_leaf_g2v)
STUB2(dispatch_leaf_s0v) STUB2(dispatch_leaf_s0g) STUB2(dispatch_leaf_s1v) STUB2(dispatch_leaf_s1g) STUB2(dispatch_leaf_s2v) STUB2(dispatch_leaf_s2g)
STUB2(dispatch_leaf_t1v) STUB2(dispatch_leaf_t1g) STUB2(dispatch_leaf_t2v) STUB2(dispatch_leaf_t2g)
}
EOS
nvcc $NVF -c $ROOT/exp/obj/stubs.cu -o $ROOT/exp/obj/stubs.o
fi
nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $ROOT/exp/lib_$NAME.so \
  $S/sk_host.o $S/sk_kernels.o $S/sk_nested.o $S/sk_general.o $S/sk_block.o $S/sk_basis.o $S/sk_solve.o $S/sk_experimental.o $S/sk_multi.o $S/sk_leaf.o $S/sk_api.o \
  $ROOT/exp/obj/g2g_$NAME.o $ROOT/exp/obj/stubs.o -lquadmath -ldl
ls -la $ROOT/exp/lib_$NAME.so
