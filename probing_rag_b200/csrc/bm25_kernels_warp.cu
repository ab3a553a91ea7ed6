// Instantiations of prw::bm25_warp_kernel (tuning.mode 3/4).
#include "bm25_warp.cuh"
#include "bm25_kernels.h"

namespace prk {

template <int NW, bool LAZY>
static warp_fn_t pick_warp(int E)
{
    if (E == 1) return prw::bm25_warp_kernel<NW, 1, LAZY>;
    if (E == 2) return prw::bm25_warp_kernel<NW, 2, LAZY>;
    return prw::bm25_warp_kernel<NW, 4, LAZY>;
}

warp_fn_t pick_warp_fn(int nw, int E, bool lazy)
{
    if (lazy) {
        if (nw == 4) return pick_warp<4, true>(E);
        if (nw == 9) return pick_warp<9, true>(E);
        if (nw == 13) return pick_warp<13, true>(E);
        if (nw == 16) return pick_warp<16, true>(E);
        return pick_warp<8, true>(E);
    }
    if (nw == 4) return pick_warp<4, false>(E);
    if (nw == 9) return pick_warp<9, false>(E);
    if (nw == 13) return pick_warp<13, false>(E);
    if (nw == 16) return pick_warp<16, false>(E);
    return pick_warp<8, false>(E);
}

}  // namespace prk
