// Warp-autonomous BM25 scoring kernel (tuning.mode 3/4) and the index-side tables it needs.
//
// The CTA-cooperative kernel in bm25.cu spends most of its time in __syncthreads(): a CTA
// barrier after every query term, with one exposed global-load latency per term (ncu:
// 57% barrier stalls, 19% long-scoreboard, shared-memory pipe 28% busy).  Here every WARP
// owns a private 2048-document fp32 accumulator tile in shared memory and scores
// (query, document-range) work items on its own:
//
//   * no CTA barrier anywhere -- a document is only ever touched by the warp that owns its
//     sub-tile, terms are applied in query-token order by that warp (fp32, one rounded add
//     per posting: still bit-identical to the reference's accumulator);
//   * 24-32 such warps per SM, all at different phases: the load latency of one warp is
//     covered by the others (plus a one-term-ahead register prefetch);
//   * posting ranges per (term, sub-tile) come from a table built once per index
//     (`tp`, for terms with df above a threshold) instead of per-tile searches; rarer terms
//     are located once per work item by a warp-collective 32-ary search, and lists of
//     <= 128 postings are simply filtered by document range;
//   * the warp keeps its running top-k in registers across the whole item, so there is no
//     cross-warp list merge; candidates are found by threshold-on-update (mode 4) or by
//     scanning the 2048 accumulators (mode 3) while re-zeroing them.
#pragma once

#include "common.cuh"

#ifndef PR_SUB_SHIFT
#define PR_SUB_SHIFT 11
#endif
#ifndef PR_PF
#define PR_PF 4
#endif

namespace prw {

constexpr int kSubShift = PR_SUB_SHIFT;
constexpr int kSub = 1 << kSubShift;  // documents per warp sub-tile (8 KB of fp32)
constexpr int kWarpCand = 32;         // candidate slots per warp (threshold-on-update)
constexpr int kLightDf = 128;         // lists this short are taken whole and range-filtered

struct WarpArgs {
    const int64_t *__restrict__ indptr;
    const int32_t *__restrict__ doc_ids;
    const float *__restrict__ weights;
    const int32_t *__restrict__ heavy_row;  // [n_terms] row of `tp`, or -1
    const uint32_t *__restrict__ tp;        // [n_rows][n_sub+1] postings with doc < s*kSub
    const int32_t *__restrict__ hot_of_row; // [n_rows] row of `hot_off`, or -1 (null: no hot stream)
    const uint32_t *__restrict__ hot_off;   // [n_hot][n_sub+1] 256-byte units of the hot stream before (row, sub-tile)
    const unsigned char *__restrict__ hot_stream;
    const unsigned char *__restrict__ stream_base;  // lean kernel: cold stream (8-byte granule p = CSR posting p), hot stream behind it
    uint32_t hot_base_g;                    // granule index of the hot stream's first byte relative to stream_base
    uint32_t *cursors;                      // lean kernel: per-warp scratch of kCursorCap * kCursorWords words (long queries), or null
    const float *__restrict__ term_maxw;    // [n_terms] largest weight of the term's list (mode 7), or null
    const uint32_t *__restrict__ plan_mask; // [n_queries] mode 7: bit j = term j of the query is skipped
    const float *__restrict__ plan_m;       // [n_queries] mode 7: upper bound of what the skipped terms add to a document
    const int64_t *__restrict__ q_indptr;
    const int32_t *__restrict__ q_terms;
    const float *__restrict__ run_theta;
    float *part_s;
    int32_t *part_d;
    int32_t *counter;
    int32_t *status;
    int64_t nnz;
    int32_t n_docs, n_terms, doc_id_base, n_queries, K;
    int32_t n_sub, subs_per_item, chunk0, n_chunks_launch;
    int32_t mode;  // 3/5 scan, 4/6 threshold-on-update, 7 threshold-on-update with rank-safe term skipping
};

__host__ __device__ inline size_t warp_smem_bytes(int nw) { return (size_t)nw * (kSub * 4 + kWarpCand * 4 + 32 * 8); }

// ---- lazily re-zeroed accumulators -------------------------------------------------------
// Re-zeroing 8 KB per (query, sub-tile) costs more shared-memory bandwidth than the scatter
// itself.  Instead each accumulator word carries a 3-bit epoch tag in bits 31..29 and a word
// whose tag is not the current epoch reads as 0.  The tag bits are free because scores are
// kept scaled by 2^-96: every partial sum of weights in [2^-30, 2^32) then lies in
// [2^-126, 2^-64), i.e. is a positive normal float whose top three bits are 000.  Scaling by a
// power of two commutes with fp32 rounding (no underflow, no overflow in that window), so the
// unscaled sums are bit-identical to adding the raw weights.  The tile is zeroed densely only
// when the epoch wraps (every 7th sub-tile).  pr_index_create checks the weight range; an
// index outside it uses the plain (LAZY=false) instantiation.
constexpr uint32_t kValLimit = 1u << 29;
constexpr float kLazyScale = 1.262177448353619e-29f;    // 2^-96
constexpr float kLazyUnscale = 7.922816251426434e+28f;  // 2^96

template <bool LAZY>
__device__ __forceinline__ float acc_decode(uint32_t bits, uint32_t ep_bits)
{
    if (LAZY) {
        const uint32_t t = bits ^ ep_bits;
        return t < kValLimit ? __uint_as_float(t) : 0.f;
    }
    return __uint_as_float(bits);
}

template <bool LAZY>
__device__ __forceinline__ uint32_t acc_encode(float v, uint32_t ep_bits)
{
    return LAZY ? (__float_as_uint(v) | ep_bits) : __float_as_uint(v);
}

__device__ __forceinline__ int2 pr_ldg_stream_i2(const int32_t *p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 pr_ldg_stream_f2(const float *p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// One warp-wide step over 32*W consecutive posting slots of a (term, sub-tile) segment:
// slot i0 + W*lane + x for x < W.  `pd`/`pw` point at the 4-aligned slot 0; slots in
// [head, head+len) hold the segment.  The documents of one term are distinct, so the plain
// load-add-store needs no atomics.
template <int W>
struct Slots {
    int d[W];
    float w[W];
};

template <int W>
__device__ __forceinline__ Slots<W> load_slots(const int32_t *pd, const float *pw, int i0, int lane, int lastgrp)
{
    Slots<W> s;
    const int ic = min(i0 + W * lane, lastgrp);  // lanes past the segment re-read its last group (masked later)
    if (W == 4) {
        const int4 dd = pr_ldg_stream_i4(pd + ic);
        const float4 ww = pr_ldg_stream_f4(pw + ic);
        s.d[0] = dd.x; s.d[1] = dd.y; s.d[W - 2] = dd.z; s.d[W - 1] = dd.w;
        s.w[0] = ww.x; s.w[1] = ww.y; s.w[W - 2] = ww.z; s.w[W - 1] = ww.w;
    } else {
        const int2 dd = pr_ldg_stream_i2(pd + ic);
        const float2 ww = pr_ldg_stream_f2(pw + ic);
        s.d[0] = dd.x; s.d[1] = dd.y;
        s.w[0] = ww.x; s.w[1] = ww.y;
    }
    return s;
}

template <int W, bool FILTER, bool LAZY>
__device__ __forceinline__ void apply_slots(uint32_t *tile, const Slots<W> &s, int i0, int lane, int head, int len,
                                            int sub_lo, int sub_n, float thr_push, uint32_t ep_bits, int &cnt,
                                            int32_t *cand, unsigned lt_mask)
{
    const int r = i0 + W * lane - head;
    unsigned o[W];
    bool m[W];
    float v[W];
#pragma unroll
    for (int x = 0; x < W; ++x) {
        m[x] = (unsigned)(r + x) < (unsigned)len;
        if (FILTER) m[x] = m[x] && (unsigned)(s.d[x] - sub_lo) < (unsigned)sub_n;
        o[x] = (unsigned)s.d[x] & (kSub - 1);  // sub_lo is a multiple of kSub
    }
#pragma unroll
    for (int x = 0; x < W; ++x) {
        const float cur = acc_decode<LAZY>(tile[o[x]], ep_bits);
        v[x] = LAZY ? __fmaf_rn(s.w[x], kLazyScale, cur) : cur + s.w[x];
    }
#pragma unroll
    for (int x = 0; x < W; ++x)
        if (m[x]) tile[o[x]] = acc_encode<LAZY>(v[x], ep_bits);
    bool hit = false;
#pragma unroll
    for (int x = 0; x < W; ++x) hit = hit || (m[x] && v[x] >= thr_push);
    if (__any_sync(PR_FULL_MASK, hit)) {  // rare: remember candidates for the end of the sub-tile
#pragma unroll
        for (int x = 0; x < W; ++x) {
            const bool h = m[x] && v[x] >= thr_push;
            const unsigned pm = __ballot_sync(PR_FULL_MASK, h);
            const int slot = cnt + __popc(pm & lt_mask);
            if (h && slot < kWarpCand) cand[slot] = (int32_t)o[x];
            cnt += __popc(pm);
        }
    }
}

// Segment descriptor in shared memory: list position (40 bits) | length (23 bits) | filter flag.
__device__ __forceinline__ uint2 seg_pack(int64_t B, int len, bool filt)
{
    return make_uint2((uint32_t)B, (uint32_t)(B >> 32) | ((uint32_t)len << 8) | (filt ? 0x80000000u : 0u));
}
__device__ __forceinline__ void seg_unpack(uint2 d, int64_t &B, int &len)
{
    B = (int64_t)d.x | ((int64_t)(d.y & 0xffu) << 32);
    len = (int)((d.y >> 8) & 0x7fffffu);
}

constexpr int kPF = PR_PF;  // segments whose first 64 slots are in flight together

// One (term, sub-tile) segment whose first 64 slots are already loaded (`first`), then 128
// slots per step while more than 64 remain, and a 64-slot tail: most segments are short, and
// a half-empty warp-wide shared-memory access costs as much as a full one.
template <bool FILTER, bool LAZY>
__device__ __forceinline__ void rmw_segment(uint32_t *tile, const Slots<2> &first, const int32_t *pd, const float *pw,
                                            int lane, int head, int len, int sub_lo, int sub_n, float thr_push,
                                            uint32_t ep_bits, int &cnt, int32_t *cand, unsigned lt_mask)
{
    const int total = head + len;
    apply_slots<2, FILTER, LAZY>(tile, first, 0, lane, head, len, sub_lo, sub_n, thr_push, ep_bits, cnt, cand, lt_mask);
    int i0 = 64;
#pragma unroll 1
    for (; total - i0 > 64; i0 += 128) {
        const Slots<4> s4 = load_slots<4>(pd, pw, i0, lane, (total - 1) & ~3);
        apply_slots<4, FILTER, LAZY>(tile, s4, i0, lane, head, len, sub_lo, sub_n, thr_push, ep_bits, cnt, cand, lt_mask);
    }
    if (i0 < total) {
        const Slots<2> s2 = load_slots<2>(pd, pw, i0, lane, (total - 1) & ~1);
        apply_slots<2, FILTER, LAZY>(tile, s2, i0, lane, head, len, sub_lo, sub_n, thr_push, ep_bits, cnt, cand, lt_mask);
    }
}

// Up to kPF consecutive (term, sub-tile) segments, in query-token order.  All first-step loads
// are issued before any accumulate, so one L2 latency is paid per group, not per term.  The
// accumulate loop is NOT unrolled (the prefetched registers rotate instead): an unrolled body
// overflows the instruction cache (ncu: 79% stall_no_inst).
template <bool LAZY>
__device__ __forceinline__ void rmw_group(uint32_t *tile, const uint2 *desc, int n, const int32_t *doc_ids,
                                          const float *weights, int lane, int sub_lo, int sub_n, float thr_push,
                                          uint32_t ep_bits, int &cnt, int32_t *cand, unsigned lt_mask)
{
    Slots<2> pf[kPF];
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
        pf[u].d[0] = pf[u].d[1] = 0;
        pf[u].w[0] = pf[u].w[1] = 0.f;
        if (u < n) {
            int64_t B;
            int len;
            seg_unpack(desc[u], B, len);
            const int head = (int)(B & 3);
            // the arrays are readable up to nnz rounded up to 4 (pr_index_create contract)
            pf[u] = load_slots<2>(doc_ids + (B - head), weights + (B - head), 0, lane, (head + len - 1) & ~1);
        }
    }
#pragma unroll 1
    for (int u = 0; u < n; ++u) {
        const uint2 dsc = desc[u];
        int64_t B;
        int len;
        seg_unpack(dsc, B, len);
        const int head = (int)(B & 3);
        const int32_t *pd = doc_ids + (B - head);
        const float *pw = weights + (B - head);
        if (dsc.y & 0x80000000u)
            rmw_segment<true, LAZY>(tile, pf[0], pd, pw, lane, head, len, sub_lo, sub_n, thr_push, ep_bits, cnt, cand,
                                    lt_mask);
        else
            rmw_segment<false, LAZY>(tile, pf[0], pd, pw, lane, head, len, sub_lo, sub_n, thr_push, ep_bits, cnt, cand,
                                     lt_mask);
#pragma unroll
        for (int k = 0; k + 1 < kPF; ++k) pf[k] = pf[k + 1];
        __syncwarp();  // order this term's stores before the next term's loads
    }
}

template <int NW, int E, bool LAZY>
__global__ void __launch_bounds__(NW * 32, (NW <= 4 ? 6 : NW <= 9 ? 3 : NW <= 13 ? 2 : 1)) bm25_warp_kernel(const WarpArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem_raw) + warp * kSub;
    int32_t *cand = reinterpret_cast<int32_t *>(smem_raw + (size_t)NW * kSub * 4) + warp * kWarpCand;
    uint2 *desc = reinterpret_cast<uint2 *>(smem_raw + (size_t)NW * (kSub * 4 + kWarpCand * 4)) + warp * 32;
    uint4 *tile4 = reinterpret_cast<uint4 *>(tile);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const unsigned lt_mask = (1u << lane) - 1u;
    const float scale = LAZY ? kLazyScale : 1.f, unscale = LAZY ? kLazyUnscale : 1.f;

#pragma unroll
    for (int v = lane; v < kSub / 4; v += 32) tile4[v] = zero4;
    __syncwarp();
    uint32_t ep = 1;  // current epoch tag (1..7); an all-zero word is stale in every epoch

    const int K = a.K, C = a.n_chunks_launch, G = a.subs_per_item;
    const int64_t n_items = (int64_t)a.n_queries * C;
    const size_t tp_stride = (size_t)a.n_sub + 1;
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
        const bool update_mode = (a.mode == 4) && (theta_run > 0.f);
        item.reset();
        float thr = fmaxf(theta_run, PR_DENORM_MIN);  // warp-uniform filter for candidates (unscaled)
        float iks = PR_SENT_SCORE;
        int ikd = PR_SENT_DOC;
        const int sub0 = (a.chunk0 + c) * G;
        const int sub1 = min(sub0 + G, a.n_sub);
        const bool single = nq <= 32;

        // per-lane description of one query term (lane j <-> term p0+j of the current pass)
        int64_t t_b0 = 0;  // start of the term's posting list
        int32_t t_lb = 0, t_le = 0, t_row = -1;
        int32_t tp_cur = 0, tp_next = 0;  // tabulated boundaries g+1, g+2 of a single-pass query
        bool t_ok = false;

        auto load_info = [&](int p0, int np, int dlo, int dhi) {
            t_ok = false;
            t_row = -1;
            t_b0 = 0;
            t_lb = 0;
            t_le = 0;
            int32_t df = 0;
            if (lane < np) {
                const int32_t t = a.q_terms[qb + p0 + lane];
                if (t < 0 || t >= a.n_terms) {
                    atomicOr(a.status, 1);
                } else {
                    t_ok = true;
                    t_b0 = a.indptr[t];
                    df = (int32_t)(a.indptr[t + 1] - t_b0);
                    t_row = a.heavy_row[t];
                }
            }
            const bool light = t_ok && t_row < 0;
            if (light && df <= kLightDf) t_le = df;  // whole list, filtered by doc range
            unsigned sm = __ballot_sync(PR_FULL_MASK, light && df > kLightDf);
            while (sm) {  // rarer terms: locate [dlo, dhi) once, by a warp-collective search
                const int j = __ffs(sm) - 1;
                sm &= sm - 1;
                const int64_t b0 = __shfl_sync(PR_FULL_MASK, t_b0, j);
                const int64_t e0 = b0 + __shfl_sync(PR_FULL_MASK, df, j);
                const int64_t lo = pr_lower_bound_warp(a.doc_ids, b0, e0, dlo, lane);
                int64_t lim = lo + (dhi - dlo);
                if (lim > e0) lim = e0;
                const int64_t hi = pr_lower_bound_warp(a.doc_ids, lo, lim, dhi, lane);
                if (lane == j) {
                    t_lb = (int32_t)(lo - b0);
                    t_le = (int32_t)(hi - b0);
                }
            }
        };

        if (single && nq > 0) load_info(0, nq, sub0 << kSubShift, min(sub1 << kSubShift, a.n_docs));

        for (int g = (nq > 0 ? sub0 : sub1); g < sub1; ++g) {
            const int sub_lo = g << kSubShift;
            const int sub_n = min(kSub, a.n_docs - sub_lo);
            const uint32_t ep_bits = ep << 29;
            int cnt = 0;  // candidates pushed for this sub-tile (warp-uniform)

            for (int p0 = 0; p0 < nq; p0 += 32) {
                const int np = min(32, nq - p0);
                if (!single) load_info(p0, np, sub_lo, sub_lo + sub_n);
                // segment of each term in this sub-tile: [sb, se) relative to the list start
                int32_t sb = 0, se = 0;
                bool filt = false;
                if (lane < np && t_ok) {
                    if (t_row >= 0) {
                        if (single && g > sub0) {  // carried from the previous sub-tile / prefetched
                            sb = tp_cur;
                            se = tp_next;
                        } else {
                            const uint32_t *r = a.tp + (size_t)t_row * tp_stride + g;
                            sb = (int32_t)__ldg(r);
                            se = (int32_t)__ldg(r + 1);
                        }
                        if (single) {  // boundary g+2, needed by the next sub-tile: load it now
                            tp_cur = se;
                            if (g + 2 <= a.n_sub) tp_next = (int32_t)__ldg(a.tp + (size_t)t_row * tp_stride + g + 2);
                        }
                    } else {
                        sb = t_lb;  // not tabulated: item-level range, filtered by document range
                        se = t_le;
                        filt = true;
                    }
                }
                // compact the non-empty segments, in term order, into the warp's descriptor list
                const bool mine = se > sb;
                const unsigned ne = __ballot_sync(PR_FULL_MASK, mine);
                if (mine) desc[__popc(ne & lt_mask)] = seg_pack(t_b0 + sb, se - sb, filt);
                __syncwarp();
                const int n_ne = __popc(ne);
                const float thr_push = update_mode ? thr * scale : __int_as_float(0x7f800000);
                for (int s0 = 0; s0 < n_ne; s0 += kPF) {
                    const int n = min(kPF, n_ne - s0);
                    rmw_group<LAZY>(tile, desc + s0, n, a.doc_ids, a.weights, lane, sub_lo, sub_n, thr_push, ep_bits, cnt,
                                    cand, lt_mask);
                }
                __syncwarp();  // the descriptor list is rewritten by the next pass / sub-tile
            }

            // ---- select from the finished sub-tile
            const int base_doc = sub_lo + a.doc_id_base;
            bool zeroed = false;
            if (update_mode && cnt <= kWarpCand) {
                if (cnt > 0) {
                    float cs = -1.f;
                    int cd = 0;
                    if (lane < cnt) {
                        const int off = cand[lane];
                        // a doc pushed twice reads a cleared word (= 0) the second time
                        cs = acc_decode<LAZY>(atomicExch(&tile[off], 0u), ep_bits) * unscale;
                        cd = base_doc + off;
                    }
                    unsigned mm = __ballot_sync(PR_FULL_MASK, cs >= thr);
                    while (mm) {
                        const int l = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const float bs = __shfl_sync(PR_FULL_MASK, cs, l);
                        const int bd = __shfl_sync(PR_FULL_MASK, cd, l);
                        if (bs > theta_run && pr_beats(bs, bd, iks, ikd)) {
                            item.insert(bs, bd, lane);
                            item.kth(K, iks, ikd);
                            thr = fmaxf(thr, iks);
                        }
                    }
                }
            } else {
                // prefilter in the accumulators' domain (thr may rise while scanning; the exact test
                // is below).  Every touched accumulator is >= FLT_MIN when scaled.
                const float thr_s = LAZY ? fmaxf(thr * scale, 1.17549435e-38f) : thr;
#pragma unroll 4
                for (int vv = lane; vv < kSub / 4; vv += 32) {
                    const uint4 xb = tile4[vv];
                    tile4[vv] = zero4;
                    const float xs[4] = {acc_decode<LAZY>(xb.x, ep_bits), acc_decode<LAZY>(xb.y, ep_bits),
                                         acc_decode<LAZY>(xb.z, ep_bits), acc_decode<LAZY>(xb.w, ep_bits)};
                    const bool any = (xs[0] >= thr_s) || (xs[1] >= thr_s) || (xs[2] >= thr_s) || (xs[3] >= thr_s);
                    if (__any_sync(PR_FULL_MASK, any)) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            unsigned mm = __ballot_sync(PR_FULL_MASK, xs[cc] >= thr_s);
                            while (mm) {
                                const int l = __ffs(mm) - 1;
                                mm &= mm - 1;
                                const float bs = __shfl_sync(PR_FULL_MASK, xs[cc], l) * unscale;
                                const int bd = base_doc + 4 * (vv - lane + l) + cc;
                                if (bs >= thr && bs > theta_run && pr_beats(bs, bd, iks, ikd)) {
                                    item.insert(bs, bd, lane);
                                    item.kth(K, iks, ikd);
                                    thr = fmaxf(thr, iks);
                                }
                            }
                        }
                    }
                }
                zeroed = true;
            }
            // ---- retire the sub-tile's accumulators
            if (LAZY) {
                if (zeroed) {
                    ep = 1;
                } else if (++ep == 8) {
#pragma unroll
                    for (int vv = lane; vv < kSub / 4; vv += 32) tile4[vv] = zero4;
                    ep = 1;
                }
            } else if (!zeroed) {
#pragma unroll
                for (int vv = lane; vv < kSub / 4; vv += 32) tile4[vv] = zero4;
            }
            __syncwarp();
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
    }
}

// ---------------------------------------------------------------- index-side tables (aux)
static __global__ void count_heavy_kernel(const int64_t *indptr, int n_terms, int64_t min_df, int32_t *count)
{
    int local = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_terms; t += (int64_t)gridDim.x * blockDim.x)
        local += (indptr[t + 1] - indptr[t]) > min_df;
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(PR_FULL_MASK, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// heavy_row[t] = exclusive count of heavy terms before t (block-local scan + block offsets)
static __global__ void heavy_block_count_kernel(const int64_t *indptr, int n_terms, int64_t min_df, int32_t *block_cnt)
{
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int flag = t < n_terms && (indptr[t + 1] - indptr[t]) > min_df;
    const unsigned m = __ballot_sync(PR_FULL_MASK, flag);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = s;
}

static __global__ void heavy_block_scan_kernel(int32_t *block_cnt, int n_blocks)  // one thread: n_blocks is a few thousand
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < n_blocks; ++i) {
            const int v = block_cnt[i];
            block_cnt[i] = acc;
            acc += v;
        }
    }
}

static __global__ void heavy_assign_kernel(const int64_t *indptr, int n_terms, int64_t min_df, const int32_t *block_off,
                                    int32_t *heavy_row, int32_t *row_term)
{
    __shared__ int warp_cnt[32];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int flag = t < n_terms && (indptr[t + 1] - indptr[t]) > min_df;
    const unsigned m = __ballot_sync(PR_FULL_MASK, flag);
    if (lane == 0) warp_cnt[w] = __popc(m);
    __syncthreads();
    int before = block_off[blockIdx.x];
    for (int i = 0; i < w; ++i) before += warp_cnt[i];
    const int row = before + __popc(m & ((1u << lane) - 1u));
    if (t < n_terms) {
        heavy_row[t] = flag ? row : -1;
        if (flag) row_term[row] = (int32_t)t;
    }
}

static __global__ void tp_fill_kernel(const int64_t *indptr, const int32_t *doc_ids, const int32_t *row_term, int n_rows,
                               int n_sub, uint32_t *tp)
{
    const int64_t total = (int64_t)n_rows * (n_sub + 1);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / (n_sub + 1)), s = (int)(idx % (n_sub + 1));
        const int t = row_term[r];
        const int64_t b0 = indptr[t], e0 = indptr[t + 1];
        const int64_t target = (int64_t)s << kSubShift;
        int64_t lo = b0, hi = e0;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (doc_ids[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        tp[idx] = (uint32_t)(lo - b0);
    }
}

// row_q[2r], row_q[2r+1] = weights of tabulated row r that at most ~1% / ~10% of its postings reach
// (lower edges of a 256-bin histogram over [0, term_maxw]).  Only the cost model of the mode-7
// planner reads them, so they need not be exact.  One CTA per row.
static __global__ void __launch_bounds__(256) row_quantile_kernel(const int64_t *indptr, const float *weights, const int32_t *row_term,
                                                          const float *term_maxw, int n_rows, float *row_q)
{
    __shared__ int hist[256];
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int t = row_term[r];
        const int64_t b = indptr[t], e = indptr[t + 1];
        const float mx = term_maxw[t];
        const float inv = mx > 0.f ? 256.f / mx : 0.f;
        hist[threadIdx.x] = 0;
        __syncthreads();
        for (int64_t p = b + threadIdx.x; p < e; p += 256) atomicAdd(&hist[min(255, (int)(weights[p] * inv))], 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const int64_t df = e - b;
            int64_t acc = 0;
            float q99 = mx, q90 = mx;
            bool have99 = false;
            for (int bin = 255; bin >= 0; --bin) {
                acc += hist[bin];
                if (!have99 && acc * 100 > df) {
                    q99 = (bin + 1) * mx / 256.f;
                    have99 = true;
                }
                if (acc * 10 > df) {
                    q90 = (bin + 1) * mx / 256.f;
                    break;
                }
            }
            row_q[2 * r] = q99;
            row_q[2 * r + 1] = q90;
        }
        __syncthreads();
    }
}

// term_maxw[t] = largest weight in the list of term t (0 for an empty list); one warp per term.
// Upper bound of any one posting's contribution, used by the rank-safe term skipping of mode 7.
static __global__ void __launch_bounds__(256) term_maxw_kernel(const int64_t *indptr, const float *weights, int n_terms, float *term_maxw)
{
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_terms; t += n_warps) {
        const int64_t b = indptr[t], e = indptr[t + 1];
        float m = 0.f;
        for (int64_t p = b + lane; p < e; p += 32) m = fmaxf(m, weights[p]);
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(PR_FULL_MASK, m, o));
        if (lane == 0) term_maxw[t] = m;
    }
}

}  // namespace prw
