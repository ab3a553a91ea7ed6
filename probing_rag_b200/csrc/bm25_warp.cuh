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

namespace prw {

constexpr int kSubShift = 11;
constexpr int kSub = 1 << kSubShift;  // documents per warp sub-tile (8 KB of fp32)
constexpr int kWarpCand = 32;         // candidate slots per warp (threshold-on-update)
constexpr int kLightDf = 128;         // lists this short are taken whole and range-filtered

struct WarpArgs {
    const int64_t *__restrict__ indptr;
    const int32_t *__restrict__ doc_ids;
    const float *__restrict__ weights;
    const int32_t *__restrict__ heavy_row;  // [n_terms] row of `tp`, or -1
    const uint32_t *__restrict__ tp;        // [n_rows][n_sub+1] postings with doc < s*kSub
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
    int32_t mode;  // 3 scan, 4 threshold-on-update
};

__host__ __device__ inline size_t warp_smem_bytes(int nw) { return (size_t)nw * (kSub * 4 + kWarpCand * 4); }

template <int NW, int E>
__global__ void __launch_bounds__(NW * 32) bm25_warp_kernel(const WarpArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tile = reinterpret_cast<float *>(smem_raw) + warp * kSub;
    int32_t *cand = reinterpret_cast<int32_t *>(smem_raw + (size_t)NW * kSub * 4) + warp * kWarpCand;
    float4 *tile4 = reinterpret_cast<float4 *>(tile);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned lt_mask = (1u << lane) - 1u;

#pragma unroll
    for (int v = lane; v < kSub / 4; v += 32) tile4[v] = zero4;
    __syncwarp();

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
        float thr = fmaxf(theta_run, PR_DENORM_MIN);  // warp-uniform filter for candidates
        float iks = PR_SENT_SCORE;
        int ikd = PR_SENT_DOC;
        const int sub0 = (a.chunk0 + c) * G;
        const int sub1 = min(sub0 + G, a.n_sub);
        const bool single = nq <= 32;

        // per-lane description of one query term (lane j <-> term p0+j of the current pass)
        int64_t t_b0 = 0;  // start of the term's posting list
        int32_t t_lb = 0, t_le = 0, t_row = -1;
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
            int cnt = 0;  // candidates pushed for this sub-tile (warp-uniform)

            for (int p0 = 0; p0 < nq; p0 += 32) {
                const int np = min(32, nq - p0);
                if (!single) load_info(p0, np, sub_lo, sub_lo + sub_n);
                int32_t sb = 0, se = 0;
                if (lane < np && t_ok) {
                    if (t_row >= 0) {
                        const uint32_t *r = a.tp + (size_t)t_row * tp_stride + g;
                        sb = (int32_t)__ldg(r);
                        se = (int32_t)__ldg(r + 1);
                    } else {
                        sb = t_lb;  // not tabulated: item-level range, filtered by doc range below
                        se = t_le;
                    }
                }
                const float thr_push = update_mode ? thr : __int_as_float(0x7f800000);
                for (int j = 0; j < np; ++j) {
                    const int32_t b = __shfl_sync(PR_FULL_MASK, sb, j);
                    const int32_t e = __shfl_sync(PR_FULL_MASK, se, j);
                    if (e <= b) continue;  // warp-uniform
                    const int64_t B = __shfl_sync(PR_FULL_MASK, t_b0, j) + b;
                    const int head = (int)(B & 3);
                    const int len = e - b;
                    const int total = head + len;
                    const int lastgrp = (total - 1) & ~3;  // last 4-aligned group holding a posting
                    // the arrays are readable up to nnz rounded up to 4 (pr_index_create contract)
                    const int32_t *pd = a.doc_ids + (B - head);
                    const float *pw = a.weights + (B - head);
                    for (int i0 = 0; i0 < total; i0 += 128) {
                        const int i = i0 + 4 * lane;
                        // lanes past the segment re-read its last group; their elements are masked
                        const int ic = min(i, lastgrp);
                        const int4 dd = pr_ldg_stream_i4(pd + ic);
                        const float4 ww = pr_ldg_stream_f4(pw + ic);
                        const int ds[4] = {dd.x, dd.y, dd.z, dd.w};
                        const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
                        const int r = i - head;
                        unsigned o[4];
                        bool m[4];
                        float v[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            m[x] = (unsigned)(r + x) < (unsigned)len && (unsigned)(ds[x] - sub_lo) < (unsigned)sub_n;
                            o[x] = (unsigned)ds[x] & (kSub - 1);  // sub_lo is a multiple of kSub
                        }
                        // distinct documents (one posting per doc and term): load all, add, store all
#pragma unroll
                        for (int x = 0; x < 4; ++x) v[x] = tile[o[x]] + wv[x];
#pragma unroll
                        for (int x = 0; x < 4; ++x)
                            if (m[x]) tile[o[x]] = v[x];
                        const bool hit = (m[0] && v[0] >= thr_push) || (m[1] && v[1] >= thr_push) ||
                                         (m[2] && v[2] >= thr_push) || (m[3] && v[3] >= thr_push);
                        if (__any_sync(PR_FULL_MASK, hit)) {  // rare: remember candidates
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                const bool h = m[x] && v[x] >= thr_push;
                                const unsigned pm = __ballot_sync(PR_FULL_MASK, h);
                                const int slot = cnt + __popc(pm & lt_mask);
                                if (h && slot < kWarpCand) cand[slot] = (int32_t)o[x];
                                cnt += __popc(pm);
                            }
                        }
                    }
                    __syncwarp();  // order this term's stores before the next term's loads
                }
            }

            // ---- select from the finished sub-tile, re-zero it
            const int base_doc = sub_lo + a.doc_id_base;
            if (update_mode && cnt <= kWarpCand) {
                if (cnt > 0) {
                    float cs = -1.f;
                    int cd = 0;
                    if (lane < cnt) {
                        const int off = cand[lane];
                        cs = atomicExch(&tile[off], 0.f);  // a doc pushed twice reads 0 the second time
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
#pragma unroll
                for (int vv = lane; vv < kSub / 4; vv += 32) tile4[vv] = zero4;
            } else {
#pragma unroll 4
                for (int vv = lane; vv < kSub / 4; vv += 32) {
                    const float4 x = tile4[vv];
                    tile4[vv] = zero4;
                    const bool any = (x.x >= thr) || (x.y >= thr) || (x.z >= thr) || (x.w >= thr);
                    if (__any_sync(PR_FULL_MASK, any)) {
                        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            unsigned mm = __ballot_sync(PR_FULL_MASK, xs[cc] >= thr);
                            while (mm) {
                                const int l = __ffs(mm) - 1;
                                mm &= mm - 1;
                                const float bs = __shfl_sync(PR_FULL_MASK, xs[cc], l);
                                const int bd = base_doc + 4 * (vv - lane + l) + cc;
                                if (bs > theta_run && pr_beats(bs, bd, iks, ikd)) {
                                    item.insert(bs, bd, lane);
                                    item.kth(K, iks, ikd);
                                    thr = fmaxf(thr, iks);
                                }
                            }
                        }
                    }
                }
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
__global__ void count_heavy_kernel(const int64_t *indptr, int n_terms, int64_t min_df, int32_t *count)
{
    int local = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_terms; t += (int64_t)gridDim.x * blockDim.x)
        local += (indptr[t + 1] - indptr[t]) > min_df;
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(PR_FULL_MASK, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// heavy_row[t] = exclusive count of heavy terms before t (block-local scan + block offsets)
__global__ void heavy_block_count_kernel(const int64_t *indptr, int n_terms, int64_t min_df, int32_t *block_cnt)
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

__global__ void heavy_block_scan_kernel(int32_t *block_cnt, int n_blocks)  // one thread: n_blocks is a few thousand
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

__global__ void heavy_assign_kernel(const int64_t *indptr, int n_terms, int64_t min_df, const int32_t *block_off,
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

__global__ void tp_fill_kernel(const int64_t *indptr, const int32_t *doc_ids, const int32_t *row_term, int n_rows,
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

}  // namespace prw
