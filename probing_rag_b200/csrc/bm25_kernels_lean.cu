// Picker over the CTA shapes of prl::bm25_lean_kernel (instantiated in bm25_kernels_lean{4,8,12}.cu).
#include "bm25_kernels.h"

namespace prk {

score_fn_t pick_lean_fn(int nw, int E, int var)
{
    if (nw == 4) return pick_lean_fn_nw4(E, var);
    if (nw == 12) return pick_lean_fn_nw12(E, var);
    return pick_lean_fn_nw8(E, var);
}

}  // namespace prk
