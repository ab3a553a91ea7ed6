// Instantiations of prl::bm25_lean_kernel for CTAs of 12 warps.
#include "bm25_lean.cuh"
#include "bm25_kernels.h"

namespace prk {

template <bool R>
static score_fn_t pick(int E)
{
    if (E == 1) return prl::bm25_lean_kernel<12, 1, R>;
    if (E == 2) return prl::bm25_lean_kernel<12, 2, R>;
    return prl::bm25_lean_kernel<12, 4, R>;
}

score_fn_t pick_lean_fn_nw12(int E, bool refresh) { return refresh ? pick<true>(E) : pick<false>(E); }

}  // namespace prk
