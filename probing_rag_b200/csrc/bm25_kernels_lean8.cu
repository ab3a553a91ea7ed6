// Instantiations of prl::bm25_lean_kernel for the default CTA shape (8 warps, 3 CTAs per SM).
#include "bm25_lean.cuh"
#include "bm25_kernels.h"

namespace prk {

score_fn_t pick_lean_fn_nw8(int E)
{
    if (E == 1) return prl::bm25_lean_kernel<8, 1>;
    if (E == 2) return prl::bm25_lean_kernel<8, 2>;
    return prl::bm25_lean_kernel<8, 4>;
}

}  // namespace prk
