#include "sk_gpts_impl.cuh"
#ifndef GNT
#define GNT 192
#endif
#ifndef GTREG
#define GTREG 2
#endif
#ifndef GLOOPD
#define GLOOPD 6
#endif
namespace skk {
cudaError_t smolyak_gpts_hook(const LeafParams &q, cudaStream_t s) {
    return launch_gpts<SK_GRID_INTERIOR2, 6, 4, GTREG, GLOOPD, GNT, 1>(q, s);
}
}
