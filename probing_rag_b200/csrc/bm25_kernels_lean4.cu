// Instantiations of prl::bm25_lean_kernel for CTAs of 4 warps.
#include "bm25_lean.cuh"
#include "bm25_kernels.h"

namespace prk {

template <int VAR>
static score_fn_t pick(int E)
{
    if (E == 1) return prl::bm25_lean_kernel<4, 1, VAR>;
    if (E == 2) return prl::bm25_lean_kernel<4, 2, VAR>;
    return prl::bm25_lean_kernel<4, 4, VAR>;
}

score_fn_t pick_lean_fn_nw4(int E, int var) { return var == 1 ? pick<1>(E) : var == 2 ? pick<2>(E) : pick<0>(E); }

}  // namespace prk
