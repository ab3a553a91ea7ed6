// Flat-step BM25 scoring kernel (tuning.mode 5/6): warp-autonomous like bm25_warp.cuh -- one
// warp owns a (query, document-range) work item and a private 2048-document fp32 score tile in
// shared memory, applies the query's terms in query-token order (one rounded fp32 add per
// posting: bit-identical to the reference's dense accumulator), keeps its running top-k in
// registers -- but with all per-term control flow taken out of the hot loop:
//
//   * per sub-tile every lane (= one query term) turns its posting range into STEP descriptors;
//     a warp prefix sum lays the steps of all terms out in query-token order in a small
//     shared-memory list;
//   * ONE uniform loop runs the steps with their loads issued kPipe steps ahead (register ring,
//     statically indexed by unrolling the ring).  The step stream is continuous over the
//     sub-tiles of a work item: the list of sub-tile g+1 is produced before the list of g is
//     drained, "end of sub-tile" (select + re-zero) is itself a step, so the first loads of a
//     sub-tile are in flight while the previous one finishes;
//   * frequent ("hot") terms are read from the hot posting stream (bm25_hot.cuh): mask-free,
//     padded, bank-aware wide (128-slot) and narrow (32-slot) steps with pre-scaled tile offsets;
//   * all other terms produce narrow steps straight from the CSR: tabulated terms through the
//     boundary table `tp`, rare terms through a per-lane cursor, so a term emits a step only for
//     the sub-tiles where it really has postings.
//
//
// Mode 7 adds RANK-SAFE TERM SKIPPING (MaxScore-style) on top of mode 6.  Once a query has a
// running k-th score theta (from the launches before), the tabulated terms are sorted by the
// largest weight of their list (term_maxw) and the longest prefix whose bounds add up to M < theta
// is not accumulated at all: a document matching only skipped terms scores <= M < theta and can
// never enter the top-k.  The tile then holds the partial sum `a(d)` of the remaining terms, the
// push threshold becomes theta - M, and every pushed candidate is RESCORED exactly: each lane
// binary-searches its term's segment for the document and the weights are added in query-token
// order in fp32, so what enters the list is bit-identical to the exhaustive modes.  Safety
// margins (1e-5 relative) cover the fp32 rounding of the partial sums (<= 32 terms).
//
// ncu history of this kernel is under profiles/r01 (v3 = segment-loop kernel it replaces).
#pragma once

#include "bm25_hot.cuh"
#include "bm25_warp.cuh"

#ifndef PR_PIPE
#define PR_PIPE 3
#endif
#ifndef PR_FLAT_CTAS
#define PR_FLAT_CTAS 3
#endif

#ifdef PR_STATS  // instrumented variant build only (tools/skip_stats.py): per-launch event counters
__device__ unsigned long long pr_stats_dev[512 * 16];
__device__ int pr_stats_launch;
#define PR_STAT(i, v) st_acc[i] += (v)
#else
#define PR_STAT(i, v)
#endif

namespace prf {

using prw::kSub;
using prw::kSubShift;
using prw::kWarpCand;
using prw::WarpArgs;

constexpr int kListCap = 64;    // step descriptors per list; every warp owns two lists (produce one, drain the other)
constexpr int kPipe = PR_PIPE;  // steps whose loads are in flight
constexpr int kScanLimit = 4;   // lane-local forward scan of a rare term before the warp search
constexpr int kTileWords = kSub + 32;  // + one dummy word per lane for the padding slots

// step kinds (descriptor .y bits 1..0); a special step is a no-op, or with kStepEnd the end of a
// sub-tile (select from the tile and re-zero it; .x = sub-tile index)
enum : uint32_t { kStepGen = 0, kStepWide = 1, kStepNarrow = 2, kStepSpecial = 3, kStepEnd = 4 };

__host__ __device__ inline size_t flat_smem_bytes(int nw) { return (size_t)nw * (kTileWords * 4 + kWarpCand * 4 + 2 * kListCap * 8); }

struct StepBuf {
    uint4 d;        // tile byte offsets (hot) / raw doc id in .x (general)
    float4 w;
    uint32_t meta;  // kind | valid lanes << 2 (general)
};

__device__ __forceinline__ float lds_f32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint2 lds_u2(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const void *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u1(const void *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f1(const void *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// Exact scores of up to 32/nq candidate documents of sub-tile g for a query of nq <= 32 terms.
// Lane l serves (candidate l / nq, term l % nq): it looks the term's weight for the document up
// (tabulated terms: binary search inside the sub-tile's segment; other terms: the whole list,
// at most the table threshold long); then every lane adds its candidate's nq weights in
// query-token order in fp32 (absent terms add +0.0f, which is exact), i.e. exactly what the
// exhaustive modes accumulate.  `offs` = the candidates' tile offsets in shared memory, n of them
// (n * nq <= 32).  Returns the score of candidate l / nq.  Warp-collective; rare, kept out of line.
static __device__ __noinline__ float flat_rescore(const int32_t *__restrict__ q_terms, int nq, const int64_t *__restrict__ indptr,
                                           const int32_t *__restrict__ heavy_row, const uint32_t *__restrict__ tp,
                                           const int32_t *__restrict__ doc_ids, const float *__restrict__ weights,
                                           int n_terms, int n_sub, int g, const int32_t *offs, int n)
{
    const int lane = threadIdx.x & 31;
    const int ci = lane / nq, tj = lane - ci * nq;
    float w = 0.f;
    if (ci < n) {
        const int doc = (g << kSubShift) + offs[ci];
        const int32_t t = q_terms[tj];
        if (t >= 0 && t < n_terms) {
            const int64_t b0 = indptr[t];
            int64_t lo = 0, hi = indptr[t + 1] - b0;
            const int row = heavy_row[t];
            if (row >= 0) {
                const uint32_t *r = tp + (size_t)row * ((size_t)n_sub + 1) + g;
                lo = r[0];
                hi = r[1];
            }
            const int64_t end = hi;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(doc_ids + b0 + mid) < doc) lo = mid + 1;
                else hi = mid;
            }
            if (lo < end && __ldg(doc_ids + b0 + lo) == doc) w = __ldg(weights + b0 + lo);
        }
    }
    float s = 0.f;
    const int base = min(ci, 31 / nq) * nq;  // idle lanes read a valid group
    for (int j = 0; j < nq; ++j) s += __shfl_sync(PR_FULL_MASK, w, base + j);
    return s;
}

// ---- mode-7 planner: which terms of a query to skip, given its running k-th score theta ------
// One warp per query (nq <= 32, lane j <-> term j), run after every merge.  Only tabulated terms
// can be skipped; they are taken in ascending (term_maxw, position) order and every prefix whose
// bounds add up to M < theta (margins included) is a SAFE choice: a document matching only
// skipped terms scores <= M < theta.  Among the safe prefixes the planner takes the cheapest
// under a cost model in units of postings: the lists that stay cost df each, and every posting of
// a remaining list that alone reaches the push threshold theta - M costs `rescore_cost` more
// (fraction estimated from the list's max / 1% / 10% weight levels).  The choice only affects
// speed: any safe prefix gives bit-identical results.
static __global__ void __launch_bounds__(128) bm25_plan_kernel(const int64_t *__restrict__ q_indptr, const int32_t *__restrict__ q_terms,
                                                       const int64_t *__restrict__ indptr, const int32_t *__restrict__ heavy_row,
                                                       const float *__restrict__ term_maxw, const float *__restrict__ row_q,
                                                       const float *__restrict__ run_theta, float *plan_theta,
                                                       uint32_t *plan_mask, float *plan_m, int B, int n_terms, float rescore_cost)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= B) return;
    const float theta = run_theta[q];
    if (theta == plan_theta[q]) return;  // unchanged since the last plan
    const int64_t qb = q_indptr[q];
    const int nq = (int)(q_indptr[q + 1] - qb);
    uint32_t best_mask = 0;
    float best_m = 0.f;
    if (theta > 0.f && nq > 1 && nq <= 32) {
        const float inf = __int_as_float(0x7f800000);
        float df = 0.f, mw = 0.f, q99 = 0.f, q90 = 0.f, key = inf;
        if (lane < nq) {
            const int32_t t = q_terms[qb + lane];
            if (t >= 0 && t < n_terms) {
                df = (float)(indptr[t + 1] - indptr[t]);
                mw = term_maxw[t];
                q99 = q90 = mw;
                const int row = heavy_row[t];
                if (row >= 0) {
                    key = mw;
                    q99 = row_q[2 * row];
                    q90 = row_q[2 * row + 1];
                }
            }
        }
        float sum_le = 0.f;  // bounds of the skippable terms ordered before or at this lane's
        for (int i = 0; i < nq; ++i) {
            const float ki = __shfl_sync(PR_FULL_MASK, key, i);
            if (ki < key || (ki == key && i <= lane)) sum_le += ki;
        }
        const float lim = theta * 0.99998f;
        const bool feasible = key < inf && sum_le * 1.00001f < lim;
        // cost of skipping nothing: the exhaustive pass over every list
        float c0 = df;
#pragma unroll
        for (int o = 16; o; o >>= 1) c0 += __shfl_xor_sync(PR_FULL_MASK, c0, o);
        float best_cost = c0;
        unsigned fm = __ballot_sync(PR_FULL_MASK, feasible);
        while (fm) {
            const int i = __ffs(fm) - 1;  // candidate prefix: everything ordered before or at lane i
            fm &= fm - 1;
            const float ki = __shfl_sync(PR_FULL_MASK, key, i);
            const float m = __shfl_sync(PR_FULL_MASK, sum_le, i) * 1.00001f;
            const float push = lim - m;
            const bool skipped = key < ki || (key == ki && lane <= i);
            float c = 0.f;
            if (!skipped && df > 0.f) {
                const float frac = push > mw ? 0.f : push >= q99 ? 0.01f : push >= q90 ? 0.1f : 1.f;
                c = df * (1.f + rescore_cost * frac);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(PR_FULL_MASK, c, o);
            if (c < best_cost) {
                best_cost = c;
                best_mask = __ballot_sync(PR_FULL_MASK, skipped);
                best_m = m;
            } else {
                __ballot_sync(PR_FULL_MASK, skipped);
            }
        }
    }
    if (lane == 0) {
        plan_mask[q] = best_mask;
        plan_m[q] = best_m;
        plan_theta[q] = theta;
    }
}

template <int NW, int E, bool SKIP>
__global__ void __launch_bounds__(NW * 32, (NW <= 4 ? 2 * PR_FLAT_CTAS : NW <= 8 ? PR_FLAT_CTAS : NW <= 12 ? 2 : 1))
    bm25_flat_kernel(const WarpArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tile = reinterpret_cast<float *>(smem_raw) + warp * kTileWords;
    int32_t *cand = reinterpret_cast<int32_t *>(smem_raw + (size_t)NW * kTileWords * 4) + warp * kWarpCand;
    uint2 *desc = reinterpret_cast<uint2 *>(smem_raw + (size_t)NW * (kTileWords * 4 + kWarpCand * 4)) + warp * 2 * kListCap;
    float4 *tile4 = reinterpret_cast<float4 *>(tile);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned lt_mask = (1u << lane) - 1u;
    // 32-bit shared-window addresses, kept opaque so they live in registers instead of being
    // rematerialised from the generic pointers inside the step loop
    uint32_t tile_sa = (uint32_t)__cvta_generic_to_shared(tile);
    uint32_t desc_sa = (uint32_t)__cvta_generic_to_shared(desc);
    asm volatile("" : "+r"(tile_sa), "+r"(desc_sa));
    const uint32_t dummy_off = (uint32_t)(kSub + lane) * 4u;

#pragma unroll
    for (int v = lane; v < kTileWords / 4; v += 32) tile4[v] = zero4;
    __syncwarp();

    const int K = a.K, C = a.n_chunks_launch, G = a.subs_per_item;
    const int64_t n_items = (int64_t)a.n_queries * C;
    const size_t tab_stride = (size_t)a.n_sub + 1;
    WarpTopK<E> item;

    while (true) {
        int item_i = 0;
        if (lane == 0) item_i = atomicAdd(a.counter, 1);
        item_i = __shfl_sync(PR_FULL_MASK, item_i, 0);
        if ((int64_t)item_i >= n_items) break;
        const int q = item_i / C, c = item_i % C;
        const int64_t qb = a.q_indptr[q];
        const int nq = (int)(a.q_indptr[q + 1] - qb);
        const float theta_run = a.run_theta[q];
        const bool update_mode = (a.mode >= 6) && (theta_run > 0.f);
#ifdef PR_STATS
        unsigned st_acc[16] = {0};
        PR_STAT(0, 1);
        PR_STAT(8, nq);
#endif
        item.reset();
        float thr = fmaxf(theta_run, PR_DENORM_MIN);  // warp-uniform filter for candidates
        float iks = PR_SENT_SCORE;
        int ikd = PR_SENT_DOC;
        const int sub0 = (a.chunk0 + c) * G;
        const int sub1 = min(sub0 + G, a.n_sub);
        const bool single = nq <= 32;
        int cnt = 0;  // candidates pushed for the sub-tile being drained (warp-uniform)
        // mode 7: this lane's term is skipped (t_skip); skip_m = upper bound of what the skipped
        // terms can add to any document (0: nothing skipped, the tile holds exact scores)
        bool t_skip = false;
        float skip_m = 0.f;
        float thr_push = update_mode ? thr : __int_as_float(0x7f800000);

        // ---- per-lane description of one query term (lane j <-> term p0+j of the current pass)
        // class: 2 = hot (steps from the hot stream, boundaries hot_off[t_row][g]), 1 = tabulated
        // (CSR, boundaries tp[t_row][g]), 0 = rare (CSR, cursor), -1 = no term
        int t_class = -1;
        int64_t t_b0 = 0;                 // start of the term's posting list (classes 0, 1)
        int32_t t_row = 0;                // row of hot_off / tp
        uint32_t tb_cur = 0, tb_next = 0; // single pass, classes 1, 2: table entries g+1, g+2
        // class 0: [t_pos, t_le) = postings not yet consumed inside the item's (single pass) or the
        // sub-tile's (several passes) document range, relative to t_b0; t_nd = document at t_pos
        int32_t t_pos = 0, t_le = 0, t_nd = 0x7fffffff;

        auto load_info = [&](int p0, int np, int dlo, int dhi) {
            t_class = -1;
            t_b0 = 0;
            t_row = 0;
            t_pos = 0;
            t_le = 0;
            t_nd = 0x7fffffff;
            int32_t df = 0;
            if (lane < np) {
                const int32_t t = a.q_terms[qb + p0 + lane];
                if (t < 0 || t >= a.n_terms) {
                    atomicOr(a.status, 1);
                } else {
                    t_b0 = a.indptr[t];
                    df = (int32_t)(a.indptr[t + 1] - t_b0);
                    const int row = a.heavy_row[t];
                    t_class = row >= 0 ? 1 : 0;
                    t_row = row;
                    if (row >= 0 && a.hot_of_row) {
                        const int h = a.hot_of_row[row];
                        if (h >= 0) {
                            t_class = 2;
                            t_row = h;
                        }
                    }
                }
            }
            const bool rare = t_class == 0 && df > 0;
            unsigned sm = __ballot_sync(PR_FULL_MASK, rare);
            while (sm) {  // locate [dlo, dhi) in the list by a warp-collective search
                const int j = __ffs(sm) - 1;
                sm &= sm - 1;
                const int64_t b0 = __shfl_sync(PR_FULL_MASK, t_b0, j);
                const int64_t e0 = b0 + __shfl_sync(PR_FULL_MASK, df, j);
                const int64_t lo = pr_lower_bound_warp(a.doc_ids, b0, e0, dlo, lane);
                int64_t lim = lo + (dhi - dlo);
                if (lim > e0) lim = e0;
                const int64_t hi = pr_lower_bound_warp(a.doc_ids, lo, lim, dhi, lane);
                if (lane == j) {
                    t_pos = (int32_t)(lo - b0);
                    t_le = (int32_t)(hi - b0);
                }
            }
            if (rare && t_pos < t_le) t_nd = __ldg(a.doc_ids + t_b0 + t_pos);
        };

        if (single && nq > 0) load_info(0, nq, sub0 << kSubShift, min(sub1 << kSubShift, a.n_docs));
        if (SKIP && update_mode && single) {  // the plan of bm25_plan_kernel for this query (made for a theta <= theta_run)
            t_skip = (a.plan_mask[q] >> lane) & 1u;
            skip_m = a.plan_m[q];
            if (skip_m > 0.f) thr_push = thr * 0.99998f - skip_m;
            PR_STAT(7, __popc(a.plan_mask[q]));
        }

        // ---- producer: the next list of step descriptors of this item, in (sub-tile, pass, chunk) order
        int it_g = nq > 0 ? sub0 : sub1, it_p0 = 0, it_w0 = 0;
        bool it_touched = false;          // a step was emitted for sub-tile it_g
        uint32_t seg_x = 0;               // per lane, for (it_g, it_p0): first table unit / posting (relative)
        int32_t seg_len = 0;              // ... and how many

        // this lane's step descriptors k in [k0, k1) of its current segment (seg_x, seg_len), from list address `la` on
        auto write_steps = [&](uint32_t la, int k0, int k1, int n_wide) {
            if (t_class == 2) {
#pragma unroll 1
                for (int k = k0; k < k1; ++k, la += 8u) {
                    const bool wide = k < n_wide;
                    const uint32_t x = wide ? seg_x + 4u * (uint32_t)k : seg_x + 3u * (uint32_t)n_wide + (uint32_t)k;
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(la), "r"(x), "r"(wide ? kStepWide : kStepNarrow) : "memory");
                }
            } else {
                int64_t p = t_b0 + seg_x + 32 * (int64_t)k0;
                int left = seg_len - 32 * k0;
#pragma unroll 1
                for (int k = k0; k < k1; ++k, la += 8u, p += 32, left -= 32) {
                    const uint32_t y = kStepGen | ((uint32_t)min(32, left) << 2) | ((uint32_t)(p >> 32) << 8);
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(la), "r"((uint32_t)p), "r"(y) : "memory");
                }
            }
        };

        auto produce = [&](uint32_t list_sa) -> int {
            while (it_g < sub1) {
                const int g = it_g;
                if (it_w0 == 0) {  // new (sub-tile, pass): what each term has inside this sub-tile
                    const int sub_lo = g << kSubShift;
                    const int sub_hi = sub_lo + min(kSub, a.n_docs - sub_lo);
                    if (!single) load_info(it_p0, min(32, nq - it_p0), sub_lo, sub_hi);
                    uint32_t sb = 0, se = 0;
                    int32_t scan_e = 0;
                    bool unresolved = false;
                    if (t_class >= 1 && !(SKIP && t_skip)) {
                        const uint32_t *tab = (t_class == 2 ? a.hot_off : a.tp) + (size_t)t_row * tab_stride + g;
                        if (single && g > sub0) {  // carried from the previous sub-tile / prefetched
                            sb = tb_cur;
                            se = tb_next;
                        } else {
                            sb = __ldg(tab);
                            se = __ldg(tab + 1);
                        }
                        if (single) {  // entry g+2, needed by the next sub-tile: load it now
                            tb_cur = se;
                            if (g + 2 <= a.n_sub) tb_next = __ldg(tab + 2);
                        }
                    } else if (t_class == 0) {
                        if (!single) {
                            sb = (uint32_t)t_pos;  // located for exactly this sub-tile
                            se = (uint32_t)t_le;
                        } else if (t_nd < sub_hi) {  // cursor: the term has a posting in this sub-tile
                            sb = (uint32_t)t_pos;
                            scan_e = t_pos + 1;
                            int probe = 0x7fffffff;
                            unresolved = true;
#pragma unroll 1
                            for (int it = 0; it < kScanLimit; ++it) {
                                probe = scan_e < t_le ? __ldg(a.doc_ids + t_b0 + scan_e) : 0x7fffffff;
                                if (probe >= sub_hi) {
                                    unresolved = false;
                                    break;
                                }
                                ++scan_e;
                            }
                            if (!unresolved) {
                                se = (uint32_t)scan_e;
                                t_pos = scan_e;
                                t_nd = probe;
                            }
                        }
                    }
                    unsigned um = __ballot_sync(PR_FULL_MASK, unresolved);
                    while (um) {  // clustered rare term: finish with a warp-collective search
                        const int j = __ffs(um) - 1;
                        um &= um - 1;
                        const int64_t b0 = __shfl_sync(PR_FULL_MASK, t_b0, j);
                        const int64_t from = b0 + __shfl_sync(PR_FULL_MASK, scan_e, j);
                        const int64_t lim = b0 + __shfl_sync(PR_FULL_MASK, t_le, j);
                        const int64_t hi = pr_lower_bound_warp(a.doc_ids, from, lim, sub_hi, lane);
                        if (lane == j) {
                            se = (uint32_t)(hi - b0);
                            t_pos = (int32_t)se;
                        }
                    }
                    if (unresolved) t_nd = t_pos < t_le ? __ldg(a.doc_ids + t_b0 + t_pos) : 0x7fffffff;
                    seg_x = sb;
                    seg_len = (int32_t)(se - sb);
                }
                // ---- steps of this (sub-tile, pass), laid out in term order
                int n_wide = 0, n = 0;
                if (t_class == 2) {
                    n_wide = seg_len >> 2;
                    n = n_wide + (seg_len & 3);
                } else if (t_class >= 0) {
                    n = (seg_len + 31) >> 5;
                }
                int incl = n, S = 0;
                if (!SKIP || __any_sync(PR_FULL_MASK, n > 0)) {  // (most sub-tiles are empty once the frequent terms are skipped)
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(PR_FULL_MASK, incl, o);
                        if (lane >= o) incl += v;
                    }
                    S = __shfl_sync(PR_FULL_MASK, incl, 31);
                }
                const int pre = incl - n;
                const bool last_pass = it_p0 + 32 >= nq;
                it_touched = it_touched || S > 0;
                const int w0 = it_w0;
                const int chunk = min(S - w0, kListCap - kPipe);
                const bool fin = w0 + chunk >= S;
                const bool end = fin && last_pass && it_touched;
                // advance the iterator
                if (!fin) {
                    it_w0 = w0 + chunk;
                } else {
                    it_w0 = 0;
                    if (last_pass) {
                        it_g = g + 1;
                        it_p0 = 0;
                        it_touched = false;
                    } else {
                        it_p0 += 32;
                    }
                }
                if (chunk == 0 && !end) continue;  // nothing in this pass / untouched sub-tile
                // ---- this lane's entries k in [k0, k1) -> list positions pre + k - w0
                write_steps(list_sa + 8u * (uint32_t)(pre + max(0, w0 - pre) - w0), max(0, w0 - pre), min(n, w0 + chunk - pre), n_wide);
                // ---- no-ops up to a multiple of kPipe; an END step always sits in the last ring slot
                int len = chunk;
                const int pad = (kPipe - ((len + (end ? 1 : 0)) % kPipe)) % kPipe;
                if (lane < pad) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list_sa + 8u * (uint32_t)(len + lane)), "r"(0u), "r"((uint32_t)kStepSpecial) : "memory");
                len += pad;
                if (end) {
                    if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list_sa + 8u * (uint32_t)len), "r"((uint32_t)g), "r"((uint32_t)(kStepSpecial | kStepEnd)) : "memory");
                    ++len;
                }
                __syncwarp();
                return len;
            }
            return 0;
        };

        // ---- fast producer: queries of <= 16 terms whose rare terms have <= 4 postings in the item's range
        // (almost every round-0 query).  The lanes are re-mapped to (sub-tile slot s, term j) = (lane / TPL,
        // lane % TPL), TPL = 8 or 16, so ONE pass of table look-ups, prefix sums and descriptor stores lays out
        // 32 / TPL consecutive sub-tiles; a rare term's few documents sit in registers (tb_cur, tb_next, t_nd,
        // t_le re-used), so there is no cursor.  A sub-tile whose steps overflow the list is cut into chunks.
        const int tpl_shift = nq <= 8 ? 3 : 4;
        const bool fast = single && nq > 0 && nq <= 16 && !__any_sync(PR_FULL_MASK, t_class == 0 && t_le - t_pos > 4);
        if (fast) {
            const int j = lane & ((1 << tpl_shift) - 1);
            const int c_ = __shfl_sync(PR_FULL_MASK, t_class, j), row_ = __shfl_sync(PR_FULL_MASK, t_row, j);
            const int pos_ = __shfl_sync(PR_FULL_MASK, t_pos, j), le_ = __shfl_sync(PR_FULL_MASK, t_le, j);
            const int64_t b0_ = __shfl_sync(PR_FULL_MASK, t_b0, j);
            const bool sk_ = __shfl_sync(PR_FULL_MASK, (int)t_skip, j) != 0;
            t_class = c_;
            t_row = row_;
            t_b0 = b0_;
            t_pos = pos_;
            t_skip = sk_;
            int dd[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
            if (c_ == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (pos_ + i < le_) dd[i] = __ldg(a.doc_ids + b0_ + pos_ + i);
            }
            tb_cur = (uint32_t)dd[0];
            tb_next = (uint32_t)dd[1];
            t_nd = dd[2];
            t_le = dd[3];
        }

        auto produce_fast = [&](uint32_t list_sa) -> int {
            const int NS = 32 >> tpl_shift, TPL = 1 << tpl_shift;
            const int s = lane >> tpl_shift, j = lane & (TPL - 1);
            while (it_g < sub1) {
                const int g0 = it_g, g = g0 + s;
                if (it_w0 == 0) {  // what each term has inside each of the NS sub-tiles
                    uint32_t sb = 0, se = 0;
                    if (g < sub1 && !(SKIP && t_skip)) {
                        if (t_class >= 1) {
                            const uint32_t *tab = (t_class == 2 ? a.hot_off : a.tp) + (size_t)t_row * tab_stride + g;
                            sb = __ldg(tab);
                            se = __ldg(tab + 1);
                        } else if (t_class == 0) {
                            const int lo = g << kSubShift, hi = a.n_docs - lo > kSub ? lo + kSub : a.n_docs;
                            const int d0 = (int)tb_cur, d1 = (int)tb_next, d2 = t_nd, d3 = t_le;
                            sb = (uint32_t)(t_pos + (d0 < lo) + (d1 < lo) + (d2 < lo) + (d3 < lo));
                            se = (uint32_t)(t_pos + (d0 < hi) + (d1 < hi) + (d2 < hi) + (d3 < hi));
                        }
                    }
                    seg_x = sb;
                    seg_len = (int32_t)(se - sb);
                }
                int n_wide = 0, n = 0;
                if (it_w0 == 0 || s == 0) {  // a chunked sub-tile continues with slot 0 only
                    if (t_class == 2) {
                        n_wide = seg_len >> 2;
                        n = n_wide + (seg_len & 3);
                    } else if (t_class >= 0) {
                        n = (seg_len + 31) >> 5;
                    }
                }
                if (!__any_sync(PR_FULL_MASK, n > 0)) {  // nothing in these sub-tiles
                    it_g = g0 + NS;
                    continue;
                }
                int incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(PR_FULL_MASK, incl, o);
                    if (lane >= o) incl += v;
                }
                int e_prev = __shfl_sync(PR_FULL_MASK, incl, max((s << tpl_shift) - 1, 0));
                if (s == 0) e_prev = 0;
                const int pre = incl - n - e_prev;  // steps of this lane's sub-tile before its own
                // ---- lay the sub-tiles out one after the other: steps, no-ops up to the last ring slot, END
                int base = 0, my_base = -1, my_t = 0, my_total = 0, ns_eff = 0, prev_e = 0, t0 = 0;
                bool chunked = it_w0 > 0;
                for (int ss = 0; ss < NS && !chunked; ++ss) {
                    const int e = __shfl_sync(PR_FULL_MASK, incl, (ss << tpl_shift) + TPL - 1);
                    const int t = e - prev_e;
                    prev_e = e;
                    if (ss == 0) t0 = t;
                    if (t > 0) {
                        const int total = (t + kPipe) / kPipe * kPipe;  // t steps + END, rounded up to whole rings
                        if (base + total > kListCap) {
                            chunked = ss == 0;
                            break;
                        }
                        if (ss == s) {
                            my_base = base;
                            my_t = t;
                            my_total = total;
                        }
                        base += total;
                    }
                    ns_eff = ss + 1;
                }
                if (!chunked) {
                    it_g = g0 + ns_eff;
                    if (base == 0) continue;  // the first non-empty sub-tile did not fit behind empty ones: next round
                    if (my_base >= 0) {
                        write_steps(list_sa + 8u * (uint32_t)(my_base + pre), 0, n, n_wide);
                        const int pad = my_total - my_t - 1;
                        if (j < pad) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list_sa + 8u * (uint32_t)(my_base + my_t + j)), "r"(0u), "r"((uint32_t)kStepSpecial) : "memory");
                        if (j == TPL - 1) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list_sa + 8u * (uint32_t)(my_base + my_total - 1)), "r"((uint32_t)g), "r"((uint32_t)(kStepSpecial | kStepEnd)) : "memory");
                    }
                    __syncwarp();
                    return base;
                }
                // ---- one long sub-tile (g0), a chunk of its steps per call
                if (it_w0 > 0) t0 = __shfl_sync(PR_FULL_MASK, incl, TPL - 1);
                const int w0 = it_w0;
                const int chunk = min(t0 - w0, kListCap - kPipe);
                const bool fin = w0 + chunk >= t0;
                if (!fin) {
                    it_w0 = w0 + chunk;
                } else {
                    it_w0 = 0;
                    it_g = g0 + 1;
                }
                if (s == 0) write_steps(list_sa + 8u * (uint32_t)(pre + max(0, w0 - pre) - w0), max(0, w0 - pre), min(n, w0 + chunk - pre), n_wide);
                int len = chunk;
                const int pad = (kPipe - ((len + (fin ? 1 : 0)) % kPipe)) % kPipe;
                if (lane < pad) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list_sa + 8u * (uint32_t)(len + lane)), "r"(0u), "r"((uint32_t)kStepSpecial) : "memory");
                len += pad;
                if (fin) {
                    if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list_sa + 8u * (uint32_t)len), "r"((uint32_t)g0), "r"((uint32_t)(kStepSpecial | kStepEnd)) : "memory");
                    ++len;
                }
                __syncwarp();
                return len;
            }
            return 0;
        };

        StepBuf buf[kPipe];
        auto issue = [&](uint32_t entry_sa, StepBuf &b) {
            const uint2 ds = lds_u2(entry_sa);
            const uint32_t kind = ds.y & 3u;
            b.meta = ds.y;
            if (kind == kStepWide) {
                const unsigned char *p = a.hot_stream + (size_t)ds.x * prh::kUnitBytes + lane * 16;
                b.d = ldg_stream_u4(p);
                b.w = ldg_stream_f4(p + 512);
            } else if (kind == kStepNarrow) {
                const unsigned char *p = a.hot_stream + (size_t)ds.x * prh::kUnitBytes + lane * 8;  // interleaved (offset, weight) pairs
                b.d.x = ldg_stream_u1(p);
                b.w.x = ldg_stream_f1(p + 4);
            } else if (kind == kStepGen) {
                const int64_t p = ((int64_t)(ds.y >> 8) << 32 | ds.x) + lane;
                if (lane < (int)((ds.y >> 2) & 63u)) {
                    b.d.x = ldg_stream_u1(a.doc_ids + p);
                    b.w.x = ldg_stream_f1(a.weights + p);
                }
            } else {
                b.d.x = ds.x;
            }
        };
        auto process = [&](const StepBuf &b, const bool may_end) {
            const uint32_t kind = b.meta & 3u;
            PR_STAT(1 + kind, 1);
            if (kind == kStepWide) {
                const uint32_t oo[4] = {b.d.x, b.d.y, b.d.z, b.d.w};
                const float ww[4] = {b.w.x, b.w.y, b.w.z, b.w.w};
                float v[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] = lds_f32(tile_sa + oo[x]);
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] += ww[x];
#pragma unroll
                for (int x = 0; x < 4; ++x) sts_f32(tile_sa + oo[x], v[x]);
                const bool hit = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])) >= thr_push;
                if (__any_sync(PR_FULL_MASK, hit)) {  // rare: remember candidates for the end of the sub-tile
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const bool h = v[x] >= thr_push;
                        const unsigned pm = __ballot_sync(PR_FULL_MASK, h);
                        const int slot = cnt + __popc(pm & lt_mask);
                        if (h && slot < kWarpCand) cand[slot] = (int32_t)(oo[x] >> 2);
                        cnt += __popc(pm);
                    }
                }
                __syncwarp();  // order this step's stores before the next step's loads
            } else if (kind != kStepSpecial) {
                uint32_t o = b.d.x;
                float w = b.w.x;
                if (kind == kStepGen) {
                    const bool m = lane < (int)((b.meta >> 2) & 63u);
                    o = m ? (o & (uint32_t)(kSub - 1)) * 4u : dummy_off;  // sub_lo is a multiple of kSub
                    w = m ? w : 0.f;
                }
                const float v = lds_f32(tile_sa + o) + w;
                sts_f32(tile_sa + o, v);
                const bool hit = v >= thr_push;
                const unsigned pm = __ballot_sync(PR_FULL_MASK, hit);
                if (pm) {
                    const int slot = cnt + __popc(pm & lt_mask);
                    if (hit && slot < kWarpCand) cand[slot] = (int32_t)(o >> 2);
                    cnt += __popc(pm);
                }
                __syncwarp();
            } else if (may_end && (b.meta & kStepEnd)) {
                // ---- select from the finished sub-tile (padding slots only ever hold +0.0f) and re-zero it.
                // With skipped terms (skip_m > 0) the tile holds partial sums: every word >= thr_push is
                // rescored exactly (flat_rescore, 32/nq documents at a time) before it meets the list.
                const int g_end = (int)b.d.x;
                const int base_doc = (g_end << kSubShift) + a.doc_id_base;
                const bool skipping = SKIP && skip_m > 0.f;
                const float thr_sel = skipping ? thr_push : thr;  // fixed while this sub-tile is selected from
                auto consider = [&](float bs, int off) {           // warp-uniform arguments, exact score
                    const int bd = base_doc + off;
                    if (bs >= thr && bs > theta_run && pr_beats(bs, bd, iks, ikd)) {
                        item.insert(bs, bd, lane);
                        item.kth(K, iks, ikd);
                        thr = fmaxf(thr, iks);
                    }
                };
                PR_STAT(5, 1);
                auto rescore_list = [&](int n) {  // the n tile offsets in cand[]
                    PR_STAT(6, n);
                    const int gsz = 32 / nq;
                    for (int c0 = 0; c0 < n; c0 += gsz) {
                        const int nb = min(gsz, n - c0);
                        const float sc = flat_rescore(a.q_terms + qb, nq, a.indptr, a.heavy_row, a.tp, a.doc_ids, a.weights,
                                                      a.n_terms, a.n_sub, g_end, cand + c0, nb);
                        for (int ci = 0; ci < nb; ++ci) consider(__shfl_sync(PR_FULL_MASK, sc, ci * nq), cand[c0 + ci]);
                    }
                };
                if (update_mode && cnt <= kWarpCand) {
                    if (cnt > 0) {
                        float cs = -1.f;
                        int co = 0;
                        if (lane < cnt) {
                            co = cand[lane];
                            // a doc pushed twice reads a cleared word (= 0) the second time
                            cs = atomicExch(&tile[co], 0.f);
                        }
                        unsigned mm = __ballot_sync(PR_FULL_MASK, cs >= thr_sel);
                        if (skipping) {
                            __syncwarp();
                            if (cs >= thr_sel) cand[__popc(mm & lt_mask)] = co;
                            __syncwarp();
                            rescore_list(__popc(mm));
                        } else {
                            while (mm) {
                                const int l = __ffs(mm) - 1;
                                mm &= mm - 1;
                                consider(__shfl_sync(PR_FULL_MASK, cs, l), __shfl_sync(PR_FULL_MASK, co, l));
                            }
                        }
                    }
#pragma unroll
                    for (int vv = lane; vv < kSub / 4; vv += 32) tile4[vv] = zero4;
                } else {
                    PR_STAT(10, 1);
                    int n_c = 0;  // skipping: offsets waiting in cand[] for their rescoring
#pragma unroll 4
                    for (int vv = lane; vv < kSub / 4; vv += 32) {
                        const float4 xb = tile4[vv];
                        tile4[vv] = zero4;
                        const float xs[4] = {xb.x, xb.y, xb.z, xb.w};
                        const bool any = (xs[0] >= thr_sel) || (xs[1] >= thr_sel) || (xs[2] >= thr_sel) || (xs[3] >= thr_sel);
                        if (__any_sync(PR_FULL_MASK, any)) {
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) {
                                unsigned mm = __ballot_sync(PR_FULL_MASK, xs[cc] >= thr_sel);
                                if (skipping) {
                                    if (mm) {
                                        if (n_c + __popc(mm) > kWarpCand) {
                                            rescore_list(n_c);
                                            n_c = 0;
                                            __syncwarp();
                                        }
                                        if (xs[cc] >= thr_sel) cand[n_c + __popc(mm & lt_mask)] = 4 * vv + cc;
                                        n_c += __popc(mm);
                                        __syncwarp();
                                    }
                                } else {
                                    while (mm) {
                                        const int l = __ffs(mm) - 1;
                                        mm &= mm - 1;
                                        consider(__shfl_sync(PR_FULL_MASK, xs[cc], l), 4 * (vv - lane + l) + cc);
                                    }
                                }
                            }
                        }
                    }
                    if (skipping) rescore_list(n_c);
                }
                cnt = 0;
                thr_push = !update_mode ? __int_as_float(0x7f800000) : skipping ? thr * 0.99998f - skip_m : thr;
                __syncwarp();
            }
        };

        // ---- consumer: drain list `cur` while list `cur ^ 1` is already produced, so the ring never runs dry
        // (one producer call site: the first round produces into `nxt` and drains an empty `cur`)
        uint32_t cur_sa = desc_sa, nxt_sa = desc_sa + 8u * kListCap;
        int n_cur = 0;
        bool first = true;
        while (true) {
            const int n_next = fast ? produce_fast(nxt_sa) : produce(nxt_sa);
            if (first) {
                first = false;
#pragma unroll
                for (int d = 0; d < kPipe; ++d) {
                    if (d < n_next) issue(nxt_sa + 8u * d, buf[d]);
                    else buf[d].meta = kStepSpecial;
                }
            }
#pragma unroll 1
            for (int s0 = 0; s0 < n_cur; s0 += kPipe) {
#pragma unroll
                for (int d = 0; d < kPipe; ++d) {
                    process(buf[d], d == kPipe - 1);
                    const int nx = s0 + d + kPipe;
                    if (nx < n_cur) issue(cur_sa + 8u * (uint32_t)nx, buf[d]);
                    else if (nx - n_cur < n_next) issue(nxt_sa + 8u * (uint32_t)(nx - n_cur), buf[d]);
                    else buf[d].meta = kStepSpecial;
                }
            }
            const uint32_t t = cur_sa;
            cur_sa = nxt_sa;
            nxt_sa = t;
            n_cur = n_next;
            if (n_cur == 0) break;
        }

        float *ps = a.part_s + ((size_t)q * C + c) * K;
        int32_t *pdst = a.part_d + ((size_t)q * C + c) * K;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = e * 32 + lane;
            if (i < K) {
                ps[i] = item.s[e];
                pdst[i] = item.d[e];
            }
        }
#ifdef PR_STATS
        if (lane == 0) {
            const int li = min(pr_stats_launch, 511);
            for (int i = 0; i < 16; ++i)
                if (st_acc[i]) atomicAdd(&pr_stats_dev[li * 16 + i], (unsigned long long)st_acc[i]);
        }
#endif
    }
}

}  // namespace prf
