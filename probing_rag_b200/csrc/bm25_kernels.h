// Pickers of the templated warp-autonomous scoring kernels.  Each family is instantiated in its own translation
// unit (bm25_kernels_warp.cu / _flat.cu / _lean.cu) so the ~60 kernel instantiations compile in parallel.
#pragma once

#include "bm25_warp.cuh"

namespace prk {

typedef void (*warp_fn_t)(const prw::WarpArgs);

warp_fn_t pick_warp_fn(int nw, int E, bool lazy);   // segment-loop kernel, tuning.mode 3/4
warp_fn_t pick_flat_fn(int nw, int E, bool skip);   // flat-step kernel, modes 5/6 (skip: mode 7)
warp_fn_t pick_lean_fn(int nw, int E);              // lean-step kernel, mode 8

}  // namespace prk
