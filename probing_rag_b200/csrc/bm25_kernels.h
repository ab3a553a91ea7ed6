// Picker of the templated scoring kernel.  The instantiations live in their own translation units
// (bm25_kernels_lean*.cu, one per CTA shape) so the library compiles in parallel.
#pragma once

#include "bm25_tables.cuh"

namespace prk {

typedef void (*score_fn_t)(const prw::ScoreArgs);

// nw = warps per CTA (4, 8, 12), E = ceil(k / 32) rounded up to 1, 2, 4, var = kernel variant (bm25_lean.cuh: 0 large
// batches / two tile epochs, 1 small batches, 2 large batches / four tile epochs)
score_fn_t pick_lean_fn(int nw, int E, int var);
score_fn_t pick_lean_fn_nw4(int E, int var);
score_fn_t pick_lean_fn_nw8(int E, int var);
score_fn_t pick_lean_fn_nw12(int E, int var);

}  // namespace prk
