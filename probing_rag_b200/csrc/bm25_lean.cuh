// Lean-step BM25 scoring kernel -- the kernel behind pr_bm25_topk.  One warp owns a (query, document-range) work
// item and a private 2048-document fp32 score tile in shared memory, applies the query's terms in query-token order
// (one rounded fp32 add per posting: bit-identical to the reference's dense accumulator, /root/reference/exp_rag.py:426
// -> bm25s `np.add.at`) and runs ONE loop over step descriptors with the posting loads issued kPipe
// steps ahead.  History and ncu numbers of the kernels it replaced (CTA-cooperative, segment-loop, flat-step) are in
// DESIGN.md 4.1 and profiles/r01.
//
//   * ONE posting address space.  Next to the hot stream (bm25_hot.cuh) the index keeps a COLD stream: every CSR
//     posting as an interleaved (pre-scaled tile byte offset, weight) pair, 8 bytes, at granule index = posting index;
//     hot narrow units are interleaved pairs too.  A descriptor is (granule index u32, flags): address = base + 8 *
//     (granule + lane [* 2]).  Two step shapes: WIDE (4 slots per lane: a 64-bit load of four 16-bit offsets and a
//     128-bit load of four weights) and NARROW (one 64-bit load
//     for the lanes below `cnt`; the others add +0.0f to a dummy word).  No-op and END steps are narrow steps with
//     cnt = 0, so there is no third kind, and every per-step branch tests one flag bit of a word all lanes hold.
//   * One 128-entry descriptor ring per warp; each produced list ends with a flagged descriptor (no loop counter) and
//     is followed by kPipe no-op descriptors, so the look-ahead never needs a bounds check.
//   * Threshold-on-update without candidate lists: each lane tracks the largest accumulator value it wrote; one vote
//     at the end of a sub-tile decides whether the tile has to be scanned at all.
//   * LIVE THRESHOLDS.  theta[q] is the best known lower bound of query q's final k-th score.  A warp reads it at the
//     start of an item and again in front of every tile scan, and RAISES it with an
//     atomicMax whenever the k-th score of its own item list exceeds it -- k documents with at least that score
//     exist, so the final k-th score cannot be lower.  Items of the same query running at the same time on other
//     SMs (and, through pr_bm25_topk_range + an all-reduce(MAX), on other GPUs) therefore filter with each other's
//     thresholds inside one launch; no warm-up launches are needed.  The test against theta is NON-strict (a document
//     that ties with a bound found elsewhere may still win on doc id); the merge applies the exact total order.
//   * Sign epochs: every other sub-tile accumulates negated sums on top of the previous sub-tile's, so the tile is
//     re-zeroed half as often.
//   * Posting loads carry an L2 evict_last policy; items are handed out chunk-major.
#pragma once

#include "bm25_hot.cuh"

#ifndef PR_LEAN_FOLD
#define PR_LEAN_FOLD 6  // bit v: kernel variant v folds the END of a sub-tile into its last step
#endif
#ifndef PR_LEAN_PIPE
#define PR_LEAN_PIPE 2
#endif
#ifndef PR_LEAN_CTAS
#define PR_LEAN_CTAS 3
#endif
#ifndef PR_LEAN_SIGN_EPOCH
#define PR_LEAN_SIGN_EPOCH 1
#endif
#ifndef PR_LEAN_L2HINT
// 1: wide-step loads carry an L2 evict_last policy, 2: narrow-step loads too.  Measured at 21M x 64k: L2 hit rate
// 60% -> 94%, DRAM reads per launch 4.6 GB -> 0.5 GB, +11% queries/s (the default policy lets the streaming .nc
// loads of one query evict the slice that the next thousand queries are about to read)
#define PR_LEAN_L2HINT 2
#endif
#include <type_traits>

namespace prl {

using prw::kSub;
using prw::kSubShift;
using prw::ScoreArgs;

constexpr int kScanLimit = 4;          // lane-local forward scan of a rare term before the warp search
// + 32 dummy words (one per bank: what the idle lanes of a narrow step and the pad slots of the hot stream add +0.0f
// to) + 4 words of per-warp scratch (word kSub + 32: the bound known at the item's start, see the kernel)
constexpr int kTileWords = kSub + 32 + 4;

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
using prw::ld_theta;
using prw::raise_theta;

constexpr int kRing = 128;    // step descriptors in the per-warp ring (two lists + the trailing no-ops)
constexpr int kListCap = 60;  // longest list one producer call lays out
constexpr int kPipe = PR_LEAN_PIPE;
constexpr int kCursorCap = 1024;  // query terms whose posting cursors fit a warp's scratch (longer queries search per sub-tile)
constexpr int kCursorWords = 5;   // scratch words per term: cursor + (list start lo/hi, class|row, df)
static_assert(2 * kListCap + kPipe <= kRing, "ring too small");
static_assert(kListCap % kPipe == 0, "lists are whole rings of kPipe steps");

// descriptor .y: bits 5..0 valid lanes of a narrow step, bit 6 end of sub-tile (sub-tile index in bits 30..8), bit 7 last
// step of its list, bit 31 wide step (the sign bit: one ISETP tests it).  END and LAST are flags on whatever step sits
// in the last slot of a round -- a real step or the no-op that pads the round
enum : uint32_t { kFlagWide = 0x80000000u, kFlagEnd = 0x40, kFlagLast = 0x80 };  // kFlagLast: last step of a produced list
constexpr int kCntShift = 0, kSubIdxShift = 8;

// rings first (the block is re-aligned to 1 KB inside the kernel so a ring slot is `base | offset`), then the tiles
__host__ __device__ inline size_t lean_smem_bytes(int nw) { return (size_t)nw * (kTileWords * 4 + kRing * 8) + 1024; }

struct StepBuf {
    uint2 d;   // wide: four 16-bit tile byte offsets; narrow: (.x, .y) = (tile byte offset, weight bits)
    float4 w;  // wide: four weights
    uint32_t meta;
};
static_assert((prw::kSub + 64) * 4 < 65536, "tile byte offsets of the hot stream are 16 bits wide");

__device__ __forceinline__ uint2 ldg_stream_u2(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_hint_u4(const void *p, uint64_t pol)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ldg_hint_f4(const void *p, uint64_t pol)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint32_t ldg_hint_u1(const void *p, uint64_t pol)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void sts_u2(uint32_t a, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
// warp-uniform copy of a value all lanes agree on (REDUX writes a uniform register)
__device__ __forceinline__ int uni(int v) { return __reduce_max_sync(PR_FULL_MASK, v); }

// cold stream: posting p of the CSR as (tile byte offset, weight bits)
static __global__ void __launch_bounds__(256) cold_fill_kernel(const int32_t *__restrict__ doc_ids, const float *__restrict__ weights,
                                                        int64_t nnz, uint2 *__restrict__ cold)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += stride)
        cold[p] = make_uint2((uint32_t)(doc_ids[p] & (kSub - 1)) * 4u, __float_as_uint(weights[p]));
}

// VAR selects the kernel variant:
//   0  large batches, two tile epochs (sign)                -- any weights; also the first launches of a call, where
//      bounds are still weak and most sub-tiles are scanned
//   1  small batches (REFRESH: re-read theta[q] in front of tile scans, see end_subtile), two epochs
//   2  large batches, four tile epochs (sign x 2^80 scale)  -- indexes whose weights lie in [2^-30, 2^8] (every real
//      BM25 index): the tile is re-zeroed every FOURTH sub-tile
// A template parameter, not a run-time flag: with a flag ptxas put a YIELD into the step loop (-2% queries/s;
// tests/test_capi.py checks the built kernels for it).
constexpr float kEpochScale = 1.2089258196146292e+24f;    // 2^80
constexpr float kEpochUnscale = 8.271806125530277e-25f;   // 2^-80
constexpr float kEpochCut = 8.881784197001252e-16f;       // 2^-50: above every unscaled stale word, below every weight
constexpr float kEpochMaxSum = 16777216.f;                // 2^24: four epochs only for queries whose scores stay below

template <int NW, int E, int VAR>
__global__ void __launch_bounds__(NW * 32, (NW <= 4 ? 2 * PR_LEAN_CTAS : NW <= 8 ? PR_LEAN_CTAS : NW <= 12 ? 2 : 1))
    bm25_lean_kernel(const ScoreArgs a)
{
    constexpr bool REFRESH = VAR == 1;
    constexpr int kEpochs = VAR == 2 ? 4 : 2;
    // the END of a sub-tile folded into its last step (see finish_list) -- not in variant 0, the one that runs while the
    // bounds are weak: with most sub-tiles scanned the folded layout measured 2-4% slower per launch, 3% faster otherwise
    constexpr bool kFold = (PR_LEAN_FOLD >> VAR) & 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    asm volatile("" : "+r"(lane));  // opaque: kept in a register instead of being rematerialised by S2R in the step loop
    unsigned char *smem_al = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    uint2 *desc = reinterpret_cast<uint2 *>(smem_al) + warp * kRing;
    float *tile = reinterpret_cast<float *>(smem_al + (size_t)NW * kRing * 8) + warp * kTileWords;
    float4 *tile4 = reinterpret_cast<float4 *>(tile);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t tile_sa = (uint32_t)__cvta_generic_to_shared(tile);
    uint32_t desc_sa = (uint32_t)__cvta_generic_to_shared(desc);  // 1 KB aligned
    asm volatile("" : "+r"(tile_sa), "+r"(desc_sa));
    const uint32_t dummy_off = (uint32_t)kSub * 4u;  // all idle lanes of a narrow step share one dummy word (a broadcast)
    const unsigned char *const sbase = a.stream_base;
    // per-lane bases of the two step shapes, so a step's address is one IMAD.WIDE (granule * 8 + base)
    const unsigned char *wide_base = sbase + lane * 16, *narrow_base = sbase + lane * 8;
    asm volatile("" : "+l"(wide_base), "+l"(narrow_base));
#if PR_LEAN_L2HINT
    uint64_t l2_keep;  // the launch's posting slice is re-read by every query of the batch: keep it in L2
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(l2_keep));
#endif
    const uint32_t hot_base_g = a.hot_base_g;

#pragma unroll
    for (int v = lane & 31; v < kTileWords / 4; v += 32) tile4[v] = zero4;
    __syncwarp();

    const int K = a.K, C = a.n_chunks_launch, G = a.subs_per_item;
    const int64_t n_items = (int64_t)a.n_queries * C;
    const size_t tab_stride = (size_t)a.n_sub + 1;
    WarpTopK<E> item;

    auto slot = [&](int pos) -> uint32_t { return desc_sa | (((uint32_t)pos << 3) & (uint32_t)(kRing * 8 - 8)); };

    while (true) {
        int item_i = 0;
        if (lane == 0) item_i = atomicAdd(a.counter, 1);
        item_i = __shfl_sync(PR_FULL_MASK, item_i, 0);
        if ((int64_t)item_i >= n_items) break;
        // chunk-major item order: all queries score document chunk 0 of the launch, then chunk 1, ... so the postings
        // the resident warps read at any moment span one chunk (subs_per_item sub-tiles), not the whole launch slice
        // (ncu: 52% L2 hit rate with query-major order -- the slice's ~90 MB footprint overflows one L2 partition)
        const int c = item_i / a.n_queries, q = item_i - c * a.n_queries;
        const int64_t qb = a.q_indptr[q];
        int nq = (int)min(a.q_indptr[q + 1] - qb, (int64_t)0x7fffffff);
        // raw pointers crossed the C ABI: a query whose CSR slice is not inside q_terms (or a batch whose offsets do not
        // start at 0) is flagged -- pr_bm25_status reports PR_EINVAL -- and scores nothing; nothing is read out of bounds
        if (qb < 0 || nq < 0 || qb + nq > a.n_q_terms || (q == 0 && qb != 0)) {
            if (lane == 0) atomicOr(a.status, 2);
            nq = 0;
        }
        float *const theta_q = a.theta + q;
        float theta_pub = ld_theta(theta_q);  // what this warp knows to be published (-1: nothing yet)
        // Large batches publish once, at the item's end, and only if the item's k-th score beats what was known at its
        // start -- kept in a spare word of the tile meanwhile: one more live register in the step loop cost 1.3%
        // queries/s on the full batch and 3.4% on a 2.6M-document shard.
        const uint32_t stash_sa = tile_sa + (uint32_t)(kSub + 32) * 4u;
        if (!REFRESH) sts_f32(stash_sa, theta_pub);
        item.reset();
        // warp-uniform candidate filter, NON-strict: max(known bound on the final k-th score, k-th score of this item's
        // list).  Without any bound it is the smallest positive float: every touched sub-tile is scanned.
        float thr = fmaxf(theta_pub, PR_DENORM_MIN);
        float iks = PR_SENT_SCORE;
        int ikd = PR_SENT_DOC;
        const int sub0 = (a.chunk0 + c) * G;
        const int sub1 = min(sub0 + G, a.n_sub);
        const bool single = nq <= 32;
        float mx = 0.f;  // per lane: largest accumulator value written for the sub-tile being drained

        // ---- per-lane description of one query term (lane j <-> term p0+j of the current pass)
        // class: 2 = hot (steps from the hot stream, boundaries hot_off[t_row][g]), 1 = tabulated
        // (cold stream, boundaries tp[t_row][g]), 0 = rare (cold stream, cursor), -1 = no term
        int t_class = -1;
        int64_t t_b0 = 0;                 // start of the term's posting list (classes 0, 1)
        int32_t t_row = 0;                // row of hot_off / tp
        uint32_t tb_cur = 0, tb_next = 0; // single pass, classes 1, 2: table entries g+1, g+2
        // class 0: [t_pos, t_le) = postings not yet consumed inside the item's (single pass) or the
        // sub-tile's (several passes) document range, relative to t_b0; t_nd = document at t_pos
        int32_t t_pos = 0, t_le = 0, t_nd = 0x7fffffff;

        int32_t t_df = 0;  // df of the lane's term (set by load_info; used when the cursors are set up)
        auto load_info = [&](int p0, int np, int dlo, int dhi) {
            t_class = -1;
            t_b0 = 0;
            t_row = 0;
            t_pos = 0;
            t_le = 0;
            t_nd = 0x7fffffff;
            int32_t df = 0;
            t_df = 0;
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
            t_df = df;
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
        // Queries of more than 32 terms (LM transcripts, exp_rag.py:428) take several 32-term passes per sub-tile.
        // Resolving every term of every pass in every sub-tile from scratch (term id -> indptr -> row tables, then a
        // search for every rare term) cost ~300 us per (query, sub-tile).  Instead each term's static description and
        // a CURSOR (its next unconsumed posting) live in this warp's global scratch: written once per item here, read
        // back with one coalesced load per pass, the cursor advanced by a short forward scan per sub-tile.
        uint32_t *const cur_w = a.cursors ? a.cursors + (size_t)(blockIdx.x * NW + warp) * (kCursorCap * kCursorWords) : nullptr;
        uint4 *const info_w = reinterpret_cast<uint4 *>(cur_w + kCursorCap);
        const bool multi = !single && cur_w != nullptr && nq <= kCursorCap;
        if (multi) {
            const int item_lo = sub0 << kSubShift, item_hi = min(sub1 << kSubShift, a.n_docs);
            for (int p0 = 0; p0 < nq; p0 += 32) {
                load_info(p0, min(32, nq - p0), item_lo, item_hi);
                if (p0 + lane < nq) {
                    cur_w[p0 + lane] = (uint32_t)t_pos;
                    info_w[p0 + lane] = make_uint4((uint32_t)t_b0, (uint32_t)((uint64_t)t_b0 >> 32),
                                                   ((uint32_t)(t_class + 1) << 28) | ((uint32_t)t_row & 0x0fffffffu), (uint32_t)t_df);  // (row is -1 for rare terms)
                }
            }
        }
        auto load_cached = [&](int p0, int np) {  // the pass's 32 terms from the scratch
            t_class = -1;
            t_b0 = 0;
            t_row = 0;
            t_pos = 0;
            t_le = 0;
            t_nd = 0x7fffffff;
            if (lane < np) {
                const uint4 w = info_w[p0 + lane];
                t_class = (int)(w.z >> 28) - 1;
                t_row = (int32_t)(w.z & 0x0fffffffu);
                t_b0 = (int64_t)(((uint64_t)w.y << 32) | w.x);
                if (t_class == 0 && w.w > 0u) {
                    t_le = (int32_t)w.w;  // the list's end: the forward scan stops at the sub-tile's last document anyway
                    t_pos = (int32_t)cur_w[p0 + lane];
                    if (t_pos < t_le) t_nd = __ldg(a.doc_ids + t_b0 + t_pos);
                }
            }
        };

        // ---- producer: the next list of step descriptors of this item, in (sub-tile, pass, chunk) order
        int it_g = nq > 0 ? sub0 : sub1, it_p0 = 0, it_w0 = 0;
        bool it_touched = false;          // a step was emitted for sub-tile it_g
        uint32_t seg_x = 0;               // per lane, for (it_g, it_p0): first table unit / posting (relative)
        int32_t seg_len = 0;              // ... and how many

        // this lane's step descriptors k in [k0, k1) of its current segment (seg_x, seg_len), from ring position `pos` on
        auto write_steps = [&](int pos, int k0, int k1, int n_wide) {
            if (t_class == 2) {
#pragma unroll 1
                for (int k = k0; k < k1; ++k, ++pos) {
                    const bool wide = k < n_wide;
                    const uint32_t unit = wide ? seg_x + (uint32_t)prh::kWideUnits * (uint32_t)k
                                               : seg_x + (uint32_t)(prh::kWideUnits - 1) * (uint32_t)n_wide + (uint32_t)k;
                    sts_u2(slot(pos), hot_base_g + unit * 32u, wide ? (uint32_t)kFlagWide : (32u << kCntShift));
                }
            } else {
                uint32_t p = (uint32_t)t_b0 + seg_x + 32u * (uint32_t)k0;  // granule = posting index (< 2^32: checked by the host)
                int left = seg_len - 32 * k0;
#pragma unroll 1
                for (int k = k0; k < k1; ++k, ++pos, p += 32u, left -= 32) sts_u2(slot(pos), p, (uint32_t)min(32, left) << kCntShift);
            }
        };
        // what follows every list: kPipe no-ops for the look-ahead of the consumer (overwritten by the next list)
        // (`pos` = ring position behind the list, `len` = its length); the list's last step gets kFlagLast, which
        // ends the consumer's drain loop without a counter
        // `fold`: END flags for the list's last step when that step is a real one (see or_flags)
        auto or_flags = [&](int pos, uint32_t flags) {  // one lane: descriptor .y at ring position pos |= flags
            const uint32_t la = slot(pos) + 4u;
            uint32_t y;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(y) : "r"(la) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(la), "r"(y | flags) : "memory");
        };
        auto finish_list = [&](int pos, int len, uint32_t fold = 0u) {
            if (lane < kPipe) sts_u2(slot(pos + lane), 0u, 0u);
            __syncwarp();
            if (len > 0 && lane == 0) or_flags(pos - 1, (uint32_t)kFlagLast | fold);
            __syncwarp();
        };
        // THE END OF A SUB-TILE IS A FLAG, NOT A STEP, whenever the sub-tile's last real step falls into the last slot
        // of a round (the consumer looks at the flags of that slot only): the END rides on that step.  Otherwise the
        // no-op that pads the round carries it, as a count-0 narrow step.  (An END step of its own behind every
        // sub-tile was 11% of all steps of a round-0 batch.)

        auto produce = [&](const int tl) -> int {
            while (it_g < sub1) {
                const int g = it_g;
                if (it_w0 == 0) {  // new (sub-tile, pass): what each term has inside this sub-tile
                    const int sub_lo = g << kSubShift;
                    const int sub_hi = sub_lo + min(kSub, a.n_docs - sub_lo);
                    if (multi) load_cached(it_p0, min(32, nq - it_p0));
                    else if (!single) load_info(it_p0, min(32, nq - it_p0), sub_lo, sub_hi);
                    uint32_t sb = 0, se = 0;
                    int32_t scan_e = 0;
                    bool unresolved = false;
                    if (t_class >= 1) {
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
                        if (!single && !multi) {
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
                    if (multi && t_class == 0 && t_le > 0) cur_w[it_p0 + lane] = (uint32_t)t_pos;
                    seg_x = sb;
                    seg_len = (int32_t)(se - sb);
                }
                // ---- steps of this (sub-tile, pass), laid out in term order
                int n_wide = 0, n = 0;
                if (t_class == 2) {
                    n_wide = seg_len / prh::kWideUnits;
                    n = n_wide + seg_len % prh::kWideUnits;
                } else if (t_class >= 0) {
                    n = (seg_len + 31) >> 5;
                }
                int incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(PR_FULL_MASK, incl, o);
                    if (lane >= o) incl += v;
                }
                const int S = __shfl_sync(PR_FULL_MASK, incl, 31);
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
                write_steps(tl + pre + max(0, w0 - pre) - w0, max(0, w0 - pre), min(n, w0 + chunk - pre), n_wide);
                // ---- no-ops up to a multiple of kPipe; the END sits in the last ring slot
                int len = chunk;
                uint32_t fold = 0u;
                if (kFold && end && len > 0 && len % kPipe == 0) {
                    fold = (uint32_t)kFlagEnd | ((uint32_t)g << kSubIdxShift);
                } else {
                    const int pad = (kPipe - ((len + (end ? 1 : 0)) % kPipe)) % kPipe;
                    if (lane < pad) sts_u2(slot(tl + len + lane), 0u, 0u);
                    len += pad;
                    if (end) {
                        if (lane == 0) sts_u2(slot(tl + len), 0u, (uint32_t)kFlagEnd | ((uint32_t)g << kSubIdxShift));
                        ++len;
                    }
                }
                finish_list(tl + len, len, fold);
                return len;
            }
            finish_list(tl, 0);
            return 0;
        };

        // ---- fast producer: queries of <= 16 terms whose rare terms have <= 4 postings in the item's range
        // (almost every round-0 query).  The lanes are re-mapped to (sub-tile slot s, term j) = (lane / TPL,
        // lane % TPL), TPL = 4, 8 or 16, so ONE pass of table look-ups, prefix sums and descriptor stores lays out
        // 32 / TPL consecutive sub-tiles; a rare term's few documents sit in registers (tb_cur, tb_next, t_nd,
        // t_le re-used), so there is no cursor.  A sub-tile whose steps overflow the list is cut into chunks.
        const int tpl_shift = nq <= 4 ? 2 : nq <= 8 ? 3 : 4;
        const bool fast = single && nq > 0 && nq <= 16 && !__any_sync(PR_FULL_MASK, t_class == 0 && t_le - t_pos > 4);
        if (fast) {
            const int j = lane & ((1 << tpl_shift) - 1);
            const int c_ = __shfl_sync(PR_FULL_MASK, t_class, j), row_ = __shfl_sync(PR_FULL_MASK, t_row, j);
            const int pos_ = __shfl_sync(PR_FULL_MASK, t_pos, j), le_ = __shfl_sync(PR_FULL_MASK, t_le, j);
            const int64_t b0_ = __shfl_sync(PR_FULL_MASK, t_b0, j);
            t_class = c_;
            t_row = row_;
            t_b0 = b0_;
            t_pos = pos_;
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

        auto produce_fast = [&](const int tl) -> int {
            const int NS = 32 >> tpl_shift, TPL = 1 << tpl_shift;
            const int s = lane >> tpl_shift, j = lane & (TPL - 1);
            while (it_g < sub1) {
                const int g0 = it_g, g = g0 + s;
                if (it_w0 == 0) {  // what each term has inside each of the NS sub-tiles
                    uint32_t sb = 0, se = 0;
                    if (g < sub1) {
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
                        n_wide = seg_len / prh::kWideUnits;
                        n = n_wide + seg_len % prh::kWideUnits;
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
                        // t steps rounded up to whole rings, the END riding on the last slot (or, unfolded, a step of its own)
                        const int total = (t + (kFold ? kPipe - 1 : kPipe)) / kPipe * kPipe;
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
                    // the END of the last sub-tile laid out is the list's last step: it carries kFlagLast itself
                    const uint32_t end_flags = (uint32_t)kFlagEnd | ((uint32_t)g << kSubIdxShift) | (my_base + my_total == base ? (uint32_t)kFlagLast : 0u);
                    const int pad = my_total - my_t;  // slots behind the last real step
                    if (my_base >= 0) {
                        write_steps(tl + my_base + pre, 0, n, n_wide);
                        if (j < pad - 1) sts_u2(slot(tl + my_base + my_t + j), 0u, 0u);
                        if (j == TPL - 1 && pad > 0) sts_u2(slot(tl + my_base + my_total - 1), 0u, end_flags);
                    }
                    __syncwarp();
                    if (my_base >= 0 && j == TPL - 1 && pad == 0) or_flags(tl + my_base + my_total - 1, end_flags);
                    finish_list(tl + base, 0);  // (kFlagLast already set above)
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
                if (s == 0) write_steps(tl + pre + max(0, w0 - pre) - w0, max(0, w0 - pre), min(n, w0 + chunk - pre), n_wide);
                int len = chunk;
                uint32_t fold = 0u;
                if (kFold && fin && len > 0 && len % kPipe == 0) {
                    fold = (uint32_t)kFlagEnd | ((uint32_t)g0 << kSubIdxShift);
                } else {
                    const int pad = (kPipe - ((len + (fin ? 1 : 0)) % kPipe)) % kPipe;
                    if (lane < pad) sts_u2(slot(tl + len + lane), 0u, 0u);
                    len += pad;
                    if (fin) {
                        if (lane == 0) sts_u2(slot(tl + len), 0u, (uint32_t)kFlagEnd | ((uint32_t)g0 << kSubIdxShift));
                        ++len;
                    }
                }
                finish_list(tl + len, len, fold);
                return len;
            }
            finish_list(tl, 0);
            return 0;
        };

        StepBuf buf[kPipe];
        // ring position of the next step to process; bits 30..29 = epoch of the tile (see rmw below: non-zero = the tile
        // still holds sums of earlier sub-tiles) -- slot() masks the high bits away, and the epoch costs no register
        // of its own
        int hd = 0;
        constexpr int kEpShift = 29, kEpMask = 3 << kEpShift;
        // four epochs need every stale word to vanish in the rounding of the first scaled add: sums below 2^24 against
        // scaled weights >= 2^50 (half an ulp: 2^26); a query that could exceed that (thousands of terms times a huge
        // weight) runs its item with two epochs
        const int ep_last = (kEpochs == 4 && (float)nq * a.max_weight < kEpochMaxSum) ? 3 : 1;
        auto issue = [&](const uint2 ds, StepBuf &b) {
            b.meta = ds.y;
            if ((int32_t)ds.y < 0) {  // wide flag; the same word in every lane: a uniform branch, cheaper than a vote + guard
                // 256 B of 16-bit offsets (8 bytes per lane), then 512 B of weights (16 bytes per lane)
                const size_t off = (size_t)ds.x << 3;
#if PR_LEAN_L2HINT
                asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;"
                             : "=r"(b.d.x), "=r"(b.d.y) : "l"(narrow_base + off), "l"(l2_keep));
                b.w = ldg_hint_f4(wide_base + off + 256, l2_keep);
#else
                b.d = ldg_stream_u2(narrow_base + off);
                b.w = ldg_stream_f4(wide_base + off + 256);
#endif
            } else {
                // lanes at or past `cnt` keep (dummy word, +0.0f).  ONE 64-bit load into (d.x, d.y) -- an aligned register
                // pair -- so a narrow step's weight travels in d.y.  (Loading the
                // pair into (d.x, w.x) made ptxas copy the words right behind the load: a full L2 latency stall per
                // step; two 32-bit loads avoid that too but cost two extra L1TEX wavefronts per step.)
                uint2 v = make_uint2(dummy_off, 0u);
                if ((uint32_t)lane < (ds.y & 63u)) {
                    const unsigned char *p = narrow_base + ((size_t)ds.x << 3);
#if PR_LEAN_L2HINT >= 2
                    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;"
                                 : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(l2_keep));
#else
                    v = ldg_stream_u2(p);
#endif
                }
                b.d.x = v.x;
                b.d.y = v.y;
            }
        };
        // ---- one step applied to the tile.  No warp barrier between steps: the warp is converged here (every branch
        // of the loop is warp-uniform) and the shared-memory pipe runs one warp's instructions in order, so a step's
        // stores land before the next step's loads; the asm statements carry "memory" clobbers, so the compiler keeps
        // the order too.
        // TILE EPOCHS (re-zeroing 8 KB per sub-tile was 39% of the shared-memory wavefronts in ncu).  A sub-tile that
        // follows an UNSCANNED one does not start from a zeroed tile; it accumulates in a form that makes every stale
        // word read as zero, at no extra instruction:
        //   epoch 0   v = x + w                         tile zeroed before
        //   epoch 1   v = min(x, -0) - w                stale words are positive: they read as -0
        //   epoch 2   v = fma(-w, 2^80, min(x, -0))     stale positives read as -0; stale NEGATIVE sums (epoch 1, > -2^24)
        //                                               vanish in the rounding of the first add (-w 2^80 <= -2^50)
        //   epoch 3   v = fma(w, 2^80, max(x, +0))      stale negatives (small or scaled) read as +0, small positives vanish
        // Negation and scaling by a power of two commute with fp32 rounding (no overflow: sums stay below 2^104; no
        // underflow: weights >= 2^-30), so the de-scaled sum is bit-identical to the plain one.  The tile is re-zeroed
        // at the end of the last epoch (every 4th sub-tile; every 2nd in the two-epoch variants) or by a scan.
        auto acc = [&](float x, float w, auto ep_c) -> float {
            constexpr int EP = decltype(ep_c)::value;
            if (EP == 0) return x + w;
            if (EP == 1) return fminf(x, -0.f) - w;
            if (EP == 2) return fmaf(-w, kEpochScale, fminf(x, -0.f));
            return fmaf(w, kEpochScale, fmaxf(x, 0.f));
        };
        auto rmw = [&](const StepBuf &b, auto ep_c) {
            constexpr bool NEG = decltype(ep_c)::value == 1 || decltype(ep_c)::value == 2;   // sums are negative
            if ((int32_t)b.meta < 0) {
                const uint32_t oo[4] = {b.d.x & 0xffffu, b.d.x >> 16, b.d.y & 0xffffu, b.d.y >> 16};
                const float ww[4] = {b.w.x, b.w.y, b.w.z, b.w.w};
                float v[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] = lds_f32(tile_sa + oo[x]);
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] = acc(v[x], ww[x], ep_c);
#pragma unroll
                for (int x = 0; x < 4; ++x) sts_f32(tile_sa + oo[x], v[x]);
                mx = NEG ? fminf(fminf(mx, v[0]), fminf(fminf(v[1], v[2]), v[3])) : fmaxf(fmaxf(mx, v[0]), fmaxf(fmaxf(v[1], v[2]), v[3]));
            } else {
                const uint32_t o = b.d.x;
                const float x = lds_f32(tile_sa + o);
                const float w = __uint_as_float(b.d.y);  // narrow step: (offset, weight) = (d.x, d.y)
                const float v = acc(x, w, ep_c);
                sts_f32(tile_sa + o, v);
                mx = NEG ? fminf(mx, v) : fmaxf(mx, v);
            }
        };
        // ---- end of a sub-tile: select from it (padding words only ever hold +-0) and, if needed, re-zero it
        auto end_subtile = [&](const uint32_t meta) {
            const int g_end = (int)((meta & ~(uint32_t)kFlagWide) >> kSubIdxShift);  // (the END may ride on a wide step)
            const int base_doc = (g_end << kSubShift) + a.doc_id_base;
            auto consider = [&](float bs, int off) {  // warp-uniform arguments, exact score
                const int bd = base_doc + off;
                if (bs >= thr && pr_beats(bs, bd, iks, ikd)) {
                    item.insert(bs, bd, lane);
                    item.kth(K, iks, ikd);
                    thr = fmaxf(thr, iks);
                }
            };
            // threshold-on-update: scores only grow and weights are >= 0, so a document can enter the list only if
            // one of its updates reached the running k-th score; that happens in well under 1% of the sub-tiles
            // once a threshold exists
            const int ep = (hd >> kEpShift) & 3;
            // de-scaling factor of this epoch's sums (negative: the sums are stored negated)
            const float unscale = (ep == 1 || ep == 2 ? -1.f : 1.f) * (ep >= 2 ? kEpochUnscale : 1.f);
            const float peak = mx * unscale;
            bool scan = __any_sync(PR_FULL_MASK, peak >= thr);
            if (scan) {
                // about to scan: first look at what other warps published since this item started -- only when items of
                // one query run side by side (a batch smaller than the resident warps; otherwise the query's previous
                // item finished long before this one started and the read at the item's start saw all there is).
                // Read here, synchronously, in front of a scan: a load kept in flight across the sub-tile would share a
                // scoreboard with the posting loads of the step loop (measured: -4% queries/s on the full batch), and
                // an unconditional read costs an L2 round trip per scanned sub-tile while bounds are still weak
                // (+5% on the early launches of a short shard).  The other variant reads the tile's dummy word (+-0: a
                // no-op) so that both have the same shape -- without any read here ptxas puts a YIELD into the step loop.
                const float t = REFRESH ? ld_theta(theta_q) : lds_f32(tile_sa + dummy_off);
                if (t > thr) {
                    thr = t;
                    scan = __any_sync(PR_FULL_MASK, peak >= thr);
                }
                if (REFRESH) theta_pub = fmaxf(theta_pub, t);
            }
            if (!scan) {
                if (ep == ep_last || !PR_LEAN_SIGN_EPOCH) {
#pragma unroll
                    for (int vv = lane & 31; vv < kSub / 4; vv += 32) tile4[vv] = zero4;
                    hd &= ~kEpMask;
                } else {
                    hd += 1 << kEpShift;  // leave the sums where they are: the next sub-tile accumulates in the next form
                }
            } else {
                // `thr` is read live: it rises with every insert (the k-th score of the item's list), and a document below
                // it can no longer enter the list -- in a launch without thresholds this cuts the candidates visited in an
                // item's first sub-tile from every positive score (~450) to a few dozen.  The visiting order does not
                // matter: the list is the top-k under a total order (score desc, doc id asc).
                // de-scaled words: this epoch's sums come out as the plain positive sums; stale words come out negative
                // (never selected) or, in the scaled epochs, as positive dust below 2^-56 that the cut sets to zero.
                // (Two copies of this loop, one without the cut for the unscaled epochs, measured slower: code size.)
                const float cut = ep >= 2 ? kEpochCut : -1.f;
#pragma unroll 4
                for (int vv = lane & 31; vv < kSub / 4; vv += 32) {
                    const float4 xb = tile4[vv];
                    tile4[vv] = zero4;
                    float xs[4] = {xb.x * unscale, xb.y * unscale, xb.z * unscale, xb.w * unscale};
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) xs[cc] = xs[cc] >= cut ? xs[cc] : 0.f;
                    const float m4 = fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3]));
                    if (__any_sync(PR_FULL_MASK, m4 >= thr)) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            unsigned mm = __ballot_sync(PR_FULL_MASK, xs[cc] >= thr);
                            while (mm) {
                                const int l = __ffs(mm) - 1;
                                mm &= mm - 1;
                                consider(__shfl_sync(PR_FULL_MASK, xs[cc], l), 4 * (vv - (lane & 31) + l) + cc);
                            }
                        }
                    }
                }
                hd &= ~kEpMask;
            }
            mx = 0.f;
            if (REFRESH && iks > theta_pub) {  // this item alone holds k documents scoring >= iks: tell everyone scoring this query
                raise_theta(a.theta, a.peer_theta, a.n_peers, q, iks, lane);
                theta_pub = iks;
            }
            __syncwarp();
        };
        // ---- one round of the ring: kPipe steps, each followed by the load of the step kPipe ahead
        auto round = [&](auto ep_c) -> uint32_t {
            uint32_t meta_last = 0u;
#pragma unroll
            for (int d = 0; d < kPipe; ++d) {
                const uint2 ds = lds_u2(slot(hd + d + kPipe));  // read early: its latency hides behind this step
                if (d == kPipe - 1) meta_last = buf[d].meta;
                rmw(buf[d], ep_c);
                issue(ds, buf[d]);
            }
            hd += kPipe;
            return meta_last;
        };

        // ---- consumer: drain the list produced one round earlier while the next one is already in the ring, so the
        // look-ahead never runs dry.  hd = ring position of the next step to process, tl = end of what is produced.
        int tl = 0;
        bool have_cur = false, first = true;
        while (true) {
            const int n_next = uni(fast ? produce_fast(tl) : produce(tl));
            tl += n_next;
            if (first) {
                first = false;
#pragma unroll
                for (int d = 0; d < kPipe; ++d) issue(lds_u2(slot(d)), buf[d]);
            }
            if (have_cur) {
                // The epoch only changes at the end of a sub-tile, so it is dispatched once per sub-tile, not once per
                // round: each epoch has its own tight loop of rounds that runs until a round ends on an END or LAST step
                // (both only ever sit in the last slot of a round).
                uint32_t meta_last;
                auto rounds = [&](auto ep_c) -> uint32_t {
                    uint32_t m;
#pragma unroll 1
                    do {
                        m = round(ep_c);
                    } while (!(m & (kFlagEnd | kFlagLast)));
                    return m;
                };
#pragma unroll 1
                do {
                    if (kEpochs == 2) {
                        meta_last = (hd & kEpMask) ? rounds(std::integral_constant<int, 1>{}) : rounds(std::integral_constant<int, 0>{});
                    } else if (hd & (2 << kEpShift)) {
                        meta_last = (hd & (1 << kEpShift)) ? rounds(std::integral_constant<int, 3>{}) : rounds(std::integral_constant<int, 2>{});
                    } else {
                        meta_last = (hd & (1 << kEpShift)) ? rounds(std::integral_constant<int, 1>{}) : rounds(std::integral_constant<int, 0>{});
                    }
                    if (meta_last & kFlagEnd) end_subtile(meta_last);
                } while (!(meta_last & kFlagLast));
            }
            have_cur = n_next > 0;
            if (!have_cur) break;
        }
        if (hd & kEpMask) {  // the item ended on unscanned sub-tiles: leave a clean tile for the next item
#pragma unroll
            for (int vv = lane & 31; vv < kSub / 4; vv += 32) tile4[vv] = zero4;
            __syncwarp();
        }

        if (!REFRESH && iks > lds_f32(stash_sa)) raise_theta(a.theta, a.peer_theta, a.n_peers, q, iks, lane);

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
    }
}

}  // namespace prl
