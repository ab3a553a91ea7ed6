// Instantiations of prl::bm25_lean_kernel for the non-default CTA shapes (4, 10, 12 warps).
#include "bm25_lean.cuh"
#include "bm25_kernels.h"

namespace prk {

template <int NW>
static score_fn_t pick_lean(int E)
{
    if (E == 1) return prl::bm25_lean_kernel<NW, 1>;
    if (E == 2) return prl::bm25_lean_kernel<NW, 2>;
    return prl::bm25_lean_kernel<NW, 4>;
}

score_fn_t pick_lean_fn(int nw, int E)
{
    if (nw == 4) return pick_lean<4>(E);
    if (nw == 10) return pick_lean<10>(E);
    if (nw == 12) return pick_lean<12>(E);
    return pick_lean_fn_nw8(E);
}

}  // namespace prk
