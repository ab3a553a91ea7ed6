// Instantiations of prf::bm25_flat_kernel (tuning.mode 5/6/7).
#include "bm25_flat.cuh"
#include "bm25_kernels.h"

namespace prk {

template <int NW, bool SKIP>
static warp_fn_t pick_flat(int E)
{
    if (E == 1) return prf::bm25_flat_kernel<NW, 1, SKIP>;
    if (E == 2) return prf::bm25_flat_kernel<NW, 2, SKIP>;
    return prf::bm25_flat_kernel<NW, 4, SKIP>;
}

warp_fn_t pick_flat_fn(int nw, int E, bool skip)
{
    if (skip) {
        if (nw == 4) return pick_flat<4, true>(E);
        if (nw == 12) return pick_flat<12, true>(E);
        return pick_flat<8, true>(E);
    }
    if (nw == 4) return pick_flat<4, false>(E);
    if (nw == 10) return pick_flat<10, false>(E);   // 2 CTAs of 10 warps: 20 warps per SM, 96 registers per thread
    if (nw == 12) return pick_flat<12, false>(E);
    return pick_flat<8, false>(E);
}

}  // namespace prk
