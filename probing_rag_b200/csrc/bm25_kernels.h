// Picker of the templated scoring kernel.  The instantiations live in their own translation units
// (bm25_kernels_lean*.cu) so the library compiles in parallel.
#pragma once

#include "bm25_tables.cuh"

namespace prk {

typedef void (*score_fn_t)(const prw::ScoreArgs);

score_fn_t pick_lean_fn(int nw, int E);   // nw = warps per CTA (4, 8, 10, 12), E = ceil(k / 32) rounded up to 1, 2, 4
score_fn_t pick_lean_fn_nw8(int E);       // (bm25_kernels_lean8.cu)

}  // namespace prk
