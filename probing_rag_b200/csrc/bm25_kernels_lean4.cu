// Instantiations of prl::bm25_lean_kernel for CTAs of 4 warps.
#include "bm25_lean.cuh"
#include "bm25_kernels.h"

namespace prk {

template <bool R>
static score_fn_t pick(int E)
{
    if (E == 1) return prl::bm25_lean_kernel<4, 1, R>;
    if (E == 2) return prl::bm25_lean_kernel<4, 2, R>;
    return prl::bm25_lean_kernel<4, 4, R>;
}

score_fn_t pick_lean_fn_nw4(int E, bool refresh) { return refresh ? pick<true>(E) : pick<false>(E); }

}  // namespace prk
