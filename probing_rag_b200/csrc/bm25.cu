// Batched BM25 scoring + top-k over a term-major CSR inverted index in HBM (sm_100a).
//
// Replaces, at token-id level, what BM25Retriever.retrieve does on the CPU for one query at a
// time (/root/reference/exp_rag.py:426,428,492; utils.py:640 -> bm25s `_compute_relevance_
// from_scores` + `selection.topk`): a dense fp32 accumulator over all
// documents, one `scores[doc] += w` per posting of every query token in query order, then
// top-k.  Here:
//
//   * the document range is cut into tiles of `tile_docs` fp32 accumulators that live in
//     SHARED MEMORY; the dense N-float array of the reference never exists in HBM;
//   * a work item is (query, chunk of `tiles_per_item` consecutive tiles); CTAs are persistent
//     and pull items from an atomic counter;
//   * inside a tile the query's terms are applied ONE TERM AT A TIME with a CTA barrier in
//     between.  A document occurs at most once in a term's posting list, so the plain
//     read-add-write into shared memory needs no atomics, and every document's score is
//     summed in query-token order in fp32 -- bit-identical to the reference's accumulator;
//   * postings are read with 128-bit streaming loads (4 doc ids + 4 weights per lane);
//     the (term, tile) posting range comes from a warp-collective 32-ary search;
//   * selection: mode 1 scans the tile (fused with re-zeroing it) into per-warp register
//     top-k lists; mode 2 tests each updated accumulator against the query's running k-th
//     score (scores only grow, weights are >= 0) and touches only the few candidates;
//   * launches walk the document range in ascending order, so the part of the index a
//     launch reads (tens of MB) stays L2-resident across the whole query batch; after each
//     launch a merge kernel folds the per-item lists into the per-query running top-k.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "bm25_warp.cuh"
#include "bm25_flat.cuh"
#include "bm25_lean.cuh"
#include "bm25_kernels.h"

namespace {

constexpr int kMaxPassTerms = 256;  // query terms planned per pass (shared-memory plan arrays)
constexpr int kLightDf = 128;       // posting lists this short are scanned whole, no search

struct ScoreArgs {
    const int64_t *__restrict__ indptr;
    const int32_t *__restrict__ doc_ids;
    const float *__restrict__ weights;
    const int64_t *__restrict__ q_indptr;
    const int32_t *__restrict__ q_terms;
    const float *__restrict__ run_theta;  // [B] k-th score of the running list, -1 if not full
    float *part_s;                        // [B][C][K]
    int32_t *part_d;
    int32_t *counter;
    int32_t *status;
    int64_t nnz;
    int32_t n_docs, n_terms, doc_id_base, n_queries, K;
    int32_t tile_docs, tiles_per_item, chunk0, n_chunks_launch;
    int32_t mode;      // 1 scan, 2 threshold-on-update
    int32_t cand_cap;
};

__host__ __device__ inline size_t score_smem_bytes(int tile_docs, int cand_cap, int nw, int K)
{
    size_t b = (size_t)tile_docs * 4;          // tile
    b += (size_t)cand_cap * 8;                 // cand_off, cand_s
    b += (size_t)nw * K * 8;                   // wl_s, wl_d
    b += (size_t)kMaxPassTerms * 16;           // seg_b, seg_e (int64)
    b += (size_t)nw * 4 + 64;                  // wl_n + misc
    return (b + 15) & ~(size_t)15;
}

template <int NT, int E, int MINB>
__global__ void __launch_bounds__(NT, MINB) bm25_score_kernel(const ScoreArgs a)
{
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = a.tile_docs;
    const int K = a.K;
    float *tile = reinterpret_cast<float *>(smem_raw);
    int64_t *seg_b = reinterpret_cast<int64_t *>(tile + T);
    int64_t *seg_e = seg_b + kMaxPassTerms;
    int32_t *cand_off = reinterpret_cast<int32_t *>(seg_e + kMaxPassTerms);
    float *cand_s = reinterpret_cast<float *>(cand_off + a.cand_cap);
    float *wl_s = cand_s + a.cand_cap;
    int32_t *wl_d = reinterpret_cast<int32_t *>(wl_s + NW * K);
    int32_t *wl_n = wl_d + NW * K;
    int32_t *s_item = wl_n + NW;
    int32_t *s_cand_n = s_item + 1;
    float *s_thr = reinterpret_cast<float *>(s_cand_n + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4 *tile4 = reinterpret_cast<float4 *>(tile);
    const int nv = T >> 2;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int v = tid; v < nv; v += NT) tile4[v] = zero4;

    const int C = a.n_chunks_launch;
    const int64_t n_items = (int64_t)a.n_queries * C;
    WarpTopK<E> item;  // warp 0 only: best K of this work item so far

    while (true) {
        __syncthreads();
        if (tid == 0) *s_item = atomicAdd(a.counter, 1);
        __syncthreads();
        const int64_t item_id = *s_item;
        if (item_id >= n_items) break;
        const int q = (int)(item_id / C), c = (int)(item_id % C);
        const int64_t qb = a.q_indptr[q];
        const int nq = (int)(a.q_indptr[q + 1] - qb);
        const float theta_run = a.run_theta[q];
        // threshold-on-update needs a full running list to test against
        const bool update_mode = (a.mode == 2) && (theta_run > 0.f);
        if (warp == 0) item.reset();
        if (tid == 0) *s_thr = fmaxf(theta_run, PR_DENORM_MIN);
        const int tile0 = (a.chunk0 + c) * a.tiles_per_item;

        // an empty query scores nothing: its list stays empty and is zero-filled at the end
        const int n_tiles_item = nq > 0 ? a.tiles_per_item : 0;
        for (int ts = 0; ts < n_tiles_item; ++ts) {
            const int64_t tile_lo64 = (int64_t)(tile0 + ts) * T;
            if (tile_lo64 >= a.n_docs) break;
            const int tile_lo = (int)tile_lo64;
            const int tile_n = min(T, a.n_docs - tile_lo);
            const int tile_hi = tile_lo + tile_n;
            if (tid == 0) *s_cand_n = 0;

            for (int p0 = 0; p0 < nq; p0 += kMaxPassTerms) {
                const int np = min(kMaxPassTerms, nq - p0);
                const bool carry = (nq <= kMaxPassTerms) && (ts > 0);
                // ---- plan: posting range of every term of this pass inside the tile
                for (int j = warp; j < np; j += NW) {
                    const int32_t t = a.q_terms[qb + p0 + j];
                    int64_t b = 0, e = 0;
                    if (t < 0 || t >= a.n_terms) {
                        if (lane == 0) atomicOr(a.status, 1);
                    } else {
                        const int64_t b0 = a.indptr[t], e0 = a.indptr[t + 1];
                        if (e0 - b0 <= kLightDf) {
                            b = b0;  // short list: take it whole, RMW filters by doc range
                            e = e0;
                        } else {
                            b = carry ? seg_e[j]
                                      : pr_lower_bound_warp(a.doc_ids, b0, e0, tile_lo, lane);
                            int64_t lim = b + tile_n;  // a tile holds <= tile_n postings of a term
                            if (lim > e0) lim = e0;
                            e = pr_lower_bound_warp(a.doc_ids, b, lim, tile_hi, lane);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        seg_b[j] = b;
                        seg_e[j] = e;
                    }
                }
                __syncthreads();
                const float thr_push = update_mode ? *((volatile float *)s_thr) : __int_as_float(0x7f800000);

                // ---- scatter-accumulate, one term at a time (fp32, query-token order)
                for (int j = 0; j < np; ++j) {
                    const int64_t b = seg_b[j], e = seg_e[j];
                    if (e <= b) continue;  // uniform: same shared-memory values for all threads
                    for (int64_t i = (b & ~(int64_t)3) + 4 * (int64_t)tid; i < e; i += 4 * NT) {
                        int4 dd;
                        float4 ww;
                        if (i + 4 <= a.nnz) {
                            dd = pr_ldg_stream_i4(a.doc_ids + i);
                            ww = pr_ldg_stream_f4(a.weights + i);
                        } else {
                            dd.x = i + 0 < a.nnz ? a.doc_ids[i + 0] : -1;
                            dd.y = i + 1 < a.nnz ? a.doc_ids[i + 1] : -1;
                            dd.z = i + 2 < a.nnz ? a.doc_ids[i + 2] : -1;
                            dd.w = -1;
                            ww.x = i + 0 < a.nnz ? a.weights[i + 0] : 0.f;
                            ww.y = i + 1 < a.nnz ? a.weights[i + 1] : 0.f;
                            ww.z = i + 2 < a.nnz ? a.weights[i + 2] : 0.f;
                            ww.w = 0.f;
                        }
                        const unsigned o0 = (unsigned)(dd.x - tile_lo), o1 = (unsigned)(dd.y - tile_lo);
                        const unsigned o2 = (unsigned)(dd.z - tile_lo), o3 = (unsigned)(dd.w - tile_lo);
                        const bool m0 = (i + 0 >= b) && (i + 0 < e) && o0 < (unsigned)tile_n;
                        const bool m1 = (i + 1 >= b) && (i + 1 < e) && o1 < (unsigned)tile_n;
                        const bool m2 = (i + 2 >= b) && (i + 2 < e) && o2 < (unsigned)tile_n;
                        const bool m3 = (i + 3 >= b) && (i + 3 < e) && o3 < (unsigned)tile_n;
                        // the four documents are distinct (one posting per doc and term):
                        // load all, add, store all
                        float v0 = m0 ? tile[o0] : 0.f;
                        float v1 = m1 ? tile[o1] : 0.f;
                        float v2 = m2 ? tile[o2] : 0.f;
                        float v3 = m3 ? tile[o3] : 0.f;
                        v0 += ww.x;
                        v1 += ww.y;
                        v2 += ww.z;
                        v3 += ww.w;
                        if (m0) tile[o0] = v0;
                        if (m1) tile[o1] = v1;
                        if (m2) tile[o2] = v2;
                        if (m3) tile[o3] = v3;
                        if ((m0 && v0 >= thr_push) || (m1 && v1 >= thr_push) ||
                            (m2 && v2 >= thr_push) || (m3 && v3 >= thr_push)) {
                            const unsigned oo[4] = {o0, o1, o2, o3};
                            const float vv[4] = {v0, v1, v2, v3};
                            const bool mm[4] = {m0, m1, m2, m3};
#pragma unroll
                            for (int x = 0; x < 4; ++x)
                                if (mm[x] && vv[x] >= thr_push &&
                                    *((volatile int32_t *)s_cand_n) < a.cand_cap) {
                                    const int pos = atomicAdd(s_cand_n, 1);
                                    if (pos < a.cand_cap) cand_off[pos] = (int32_t)oo[x];
                                }
                        }
                    }
                    __syncthreads();
                }
                __syncthreads();  // plan arrays are rewritten by the next pass / tile
            }

            // ---- select from the finished tile
            const int cand_n = *((volatile int32_t *)s_cand_n);
            const bool scan = !update_mode || cand_n >= a.cand_cap;
            if (!scan) {
                if (cand_n > 0) {
                    for (int i = tid; i < cand_n; i += NT)
                        cand_s[i] = atomicExch(&tile[cand_off[i]], 0.f);  // duplicates read 0
                    __syncthreads();
                    if (warp == 0) {
                        float thr = *((volatile float *)s_thr);
                        float ks;
                        int kd;
                        item.kth(K, ks, kd);
                        for (int base = 0; base < cand_n; base += 32) {
                            const int i = base + lane;
                            const float cs = i < cand_n ? cand_s[i] : -1.f;
                            const int cd = i < cand_n ? tile_lo + cand_off[i] + a.doc_id_base : 0;
                            unsigned m = __ballot_sync(PR_FULL_MASK, cs >= thr);
                            while (m) {
                                const int l = __ffs(m) - 1;
                                m &= m - 1;
                                const float bs = __shfl_sync(PR_FULL_MASK, cs, l);
                                const int bd = __shfl_sync(PR_FULL_MASK, cd, l);
                                if (bs > theta_run && pr_beats(bs, bd, ks, kd)) {
                                    item.insert(bs, bd, lane);
                                    item.kth(K, ks, kd);
                                    thr = fmaxf(thr, ks);
                                }
                            }
                        }
                        if (lane == 0) *s_thr = thr;
                    }
                }
                for (int v = tid; v < nv; v += NT) tile4[v] = zero4;
            } else {
                WarpTopK<E> wl;
                wl.reset();
                int nins = 0;
                float thr_w = *((volatile float *)s_thr);
                float ks = PR_SENT_SCORE;
                int kd = PR_SENT_DOC;
                for (int v = tid; v < nv; v += NT) {
                    const float4 x = tile4[v];
                    tile4[v] = zero4;
                    const bool any = (x.x >= thr_w) || (x.y >= thr_w) || (x.z >= thr_w) || (x.w >= thr_w);
                    if (__any_sync(PR_FULL_MASK, any)) {
                        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            unsigned m = __ballot_sync(PR_FULL_MASK, xs[cc] >= thr_w);
                            while (m) {
                                const int l = __ffs(m) - 1;
                                m &= m - 1;
                                const float bs = __shfl_sync(PR_FULL_MASK, xs[cc], l);
                                const int bd = tile_lo + a.doc_id_base + 4 * (v - lane + l) + cc;
                                if (bs > theta_run && pr_beats(bs, bd, ks, kd)) {
                                    wl.insert(bs, bd, lane);
                                    ++nins;
                                    wl.kth(K, ks, kd);
                                    thr_w = fmaxf(thr_w, ks);
                                }
                            }
                        }
                    }
                }
                const int nvalid = min(nins, K);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = e * 32 + lane;
                    if (i < nvalid) {
                        wl_s[warp * K + i] = wl.s[e];
                        wl_d[warp * K + i] = wl.d[e];
                    }
                }
                if (lane == 0) wl_n[warp] = nvalid;
                __syncthreads();
                if (warp == 0) {
                    float thr = *((volatile float *)s_thr);
                    float iks;
                    int ikd;
                    item.kth(K, iks, ikd);
                    for (int w = 0; w < NW; ++w) {
                        const int n = wl_n[w];
                        for (int i = 0; i < n; ++i) {
                            const float bs = wl_s[w * K + i];
                            const int bd = wl_d[w * K + i];
                            if (!pr_beats(bs, bd, iks, ikd)) break;  // the list is sorted: the rest lose too
                            item.insert(bs, bd, lane);
                            item.kth(K, iks, ikd);
                        }
                    }
                    thr = fmaxf(thr, iks);
                    if (lane == 0) *s_thr = thr;
                }
            }
            __syncthreads();
        }

        if (warp == 0) {
            float *ps = a.part_s + ((size_t)q * C + c) * K;
            int32_t *pd = a.part_d + ((size_t)q * C + c) * K;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = e * 32 + lane;
                if (i < K) {
                    ps[i] = item.s[e];
                    pd[i] = item.d[e];
                }
            }
        }
    }
}

// One warp per query: running list  <-  running list  U  the C per-item lists of one launch.
// `finalize` also writes the caller's output, filling a short list with the shard's lowest
// doc ids at score 0 (canonical zero-score tail).
template <int E>
__global__ void __launch_bounds__(128) bm25_merge_kernel(
    const float *__restrict__ part_s, const int32_t *__restrict__ part_d, int C,
    int64_t stride_q, int64_t stride_c, float *run_s, int32_t *run_d, float *run_theta, int B,
    int K, int finalize, float *out_s, int32_t *out_d, int doc_id_base, int n_docs)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= B) return;
    WarpTopK<E> L;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        L.s[e] = (run_s && i < K) ? run_s[(size_t)q * K + i] : PR_SENT_SCORE;
        L.d[e] = (run_d && i < K) ? run_d[(size_t)q * K + i] : PR_SENT_DOC;
    }
    float ks;
    int kd;
    L.kth(K, ks, kd);
    // 32 lists at a time: every lane reads the head of one list, and only the lists whose head beats the current
    // k-th entry are walked (a small batch has thousands of per-item lists per query -- one dependent global load per
    // list made the merge of a single query take milliseconds)
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int cl = c0 + lane;
        float hs = -1.f;
        int hd = -1;
        if (cl < C) {
            hs = part_s[q * stride_q + cl * stride_c];
            hd = part_d[q * stride_q + cl * stride_c];
        }
        unsigned m = __ballot_sync(PR_FULL_MASK, hs >= 0.f && hd >= 0 && pr_beats(hs, hd, ks, kd));
        while (m) {
            const int c = c0 + __ffs(m) - 1;
            m &= m - 1;
            const float *ps = part_s + q * stride_q + c * stride_c;
            const int32_t *pd = part_d + q * stride_q + c * stride_c;
            for (int i = 0; i < K; ++i) {
                const float bs = ps[i];
                const int bd = pd[i];
                if (bs < 0.f || bd < 0) break;           // empty slot / missing entry: list ends
                if (!pr_beats(bs, bd, ks, kd)) break;    // sorted: the rest lose too
                L.insert(bs, bd, lane);
                L.kth(K, ks, kd);
            }
        }
    }
    if (run_s) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = e * 32 + lane;
            if (i < K) {
                run_s[(size_t)q * K + i] = L.s[e];
                run_d[(size_t)q * K + i] = L.d[e];
            }
        }
        if (lane == 0) run_theta[q] = ks;
    }
    if (finalize) {
        int nvalid = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = e * 32 + lane;
            const bool valid = i < K && L.s[e] >= 0.f;
            nvalid += __popc(__ballot_sync(PR_FULL_MASK, valid));
            if (i < K) {
                out_s[(size_t)q * K + i] = valid ? L.s[e] : -INFINITY;
                out_d[(size_t)q * K + i] = valid ? L.d[e] : -1;
            }
        }
        __syncwarp();
        if (n_docs >= 0 && nvalid < K && lane == 0) {
            // zero-score tail: lowest local doc ids not already listed
            int cand = 0;
            for (int pos = nvalid; pos < K; ++pos) {
                while (cand < n_docs) {
                    bool listed = false;
                    for (int i = 0; i < nvalid; ++i)
                        if (out_d[(size_t)q * K + i] == cand + doc_id_base) listed = true;
                    if (!listed) break;
                    ++cand;
                }
                if (cand >= n_docs) break;
                out_s[(size_t)q * K + pos] = 0.f;
                out_d[(size_t)q * K + pos] = cand + doc_id_base;
                ++cand;
            }
        }
    }
}

__global__ void bm25_init_kernel(float *run_s, int32_t *run_d, float *run_theta, float *plan_theta, uint32_t *plan_mask,
                                 float *plan_m, int64_t n_run, int B, int32_t *counters, int n_counters, int32_t *status)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_run) {
        run_s[i] = PR_SENT_SCORE;
        run_d[i] = PR_SENT_DOC;
    }
    if (i < B) {
        run_theta[i] = PR_SENT_SCORE;
        plan_theta[i] = -2.f;   // no plan yet (mode 7)
        plan_mask[i] = 0u;
        plan_m[i] = 0.f;
    }
    if (i < n_counters) counters[i] = 0;
    if (i == 0) *status = 0;
}

// index validation (pr_index_create): indptr monotone and consistent, doc ids in range and
// strictly ascending inside a term, weights finite and >= 0.
__global__ void bm25_validate_kernel(const int64_t *indptr, const int32_t *doc_ids,
                                     const float *weights, int n_terms, int64_t nnz, int n_docs,
                                     int32_t *bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t t = i0; t < n_terms; t += stride) {
        const int64_t b = indptr[t], e = indptr[t + 1];
        if (b > e || b < 0 || e > nnz) atomicOr(bad, 1);
    }
    if (i0 == 0 && (indptr[0] != 0 || indptr[n_terms] != nnz)) atomicOr(bad, 1);
    uint32_t wmin = 0xffffffffu, wmax = 0u;
    for (int64_t p = i0; p < nnz; p += stride) {
        const int32_t d = doc_ids[p];
        const float w = weights[p];
        if (d < 0 || d >= n_docs) atomicOr(bad, 2);
        if (!(w >= 0.f) || w > 3.0e38f) atomicOr(bad, 8);
        else {  // non-negative floats order like their bit patterns
            wmin = min(wmin, __float_as_uint(w));
            wmax = max(wmax, __float_as_uint(w));
        }
        if (p > 0 && d <= doc_ids[p - 1]) {
            // a descent is only legal where a new term's list starts: p must be in indptr
            int64_t lo = 0, hi = n_terms;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (indptr[mid] < p) lo = mid + 1;
                else hi = mid;
            }
            if (indptr[lo] != p) atomicOr(bad, 4);
        }
    }
    for (int o = 16; o; o >>= 1) {
        wmin = min(wmin, __shfl_xor_sync(PR_FULL_MASK, wmin, o));
        wmax = max(wmax, __shfl_xor_sync(PR_FULL_MASK, wmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(reinterpret_cast<uint32_t *>(bad) + 1, wmin);
        atomicMax(reinterpret_cast<uint32_t *>(bad) + 2, wmax);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------ host
struct pr_index {
    int device;
    int64_t n_docs_global;
    int32_t doc_id_base, n_docs, n_terms;
    int64_t nnz;
    const int64_t *indptr;
    const int32_t *doc_ids;
    const float *weights;
    pr_bm25_tuning_t tuning;
    int num_sms;
    int64_t last_launches;
    // optional per-kernel timing (pr_index_set_profiling): events around every score launch
    int profiling;
    std::vector<cudaEvent_t> ev;   // start/stop pairs
    int ev_used;
    // tables of the warp-autonomous kernel (pr_index_build_aux), in caller-owned memory
    const int32_t *heavy_row;
    const uint32_t *tp;
    int32_t n_rows, n_sub;
    int64_t heavy_min_df;
    bool lazy_ok;
    // hot posting stream of the flat-step kernel (bm25_hot.cuh), also in the aux buffer
    const int32_t *hot_of_row;
    const uint32_t *hot_off;
    const unsigned char *hot_stream;
    int32_t n_hot;
    int64_t hot_min_df, hot_stream_bytes;
    // largest weight per term (rank-safe term skipping, mode 7), also in the aux buffer
    const float *term_maxw;
    const float *row_q;   // [n_rows][2] weight levels ~1% / ~10% of a tabulated row's postings reach (planner cost model)
    // cold stream of the lean kernel (bm25_lean.cuh, mode 8): every CSR posting as an (offset, weight) pair, in the
    // aux buffer in front of the hot stream; lean_ok = it exists and the whole stream space fits 32-bit granule indices
    const unsigned char *cold_stream;
    uint32_t hot_base_g;
    bool lean_ok;
};

namespace {

struct Layout {
    int n_chunks, C, L;
    int G;  // sub-tiles per work item actually used (warp modes): tuning.subs_per_item, halved for small batches
    std::vector<int> launch_chunk0, launch_chunks;  // launch li covers chunks [chunk0, chunk0 + chunks)
    size_t off_status, off_counters, off_theta, off_plan_theta, off_plan_mask, off_plan_m, off_run_s, off_run_d, off_part_s, off_part_d, off_cursors, total;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

Layout make_layout(const pr_index *ix, int32_t B, int32_t K)
{
    const pr_bm25_tuning_t &t = ix->tuning;
    Layout l;
    int64_t c = B > 0 ? ((int64_t)t.min_items + B - 1) / B : 1;
    l.G = t.subs_per_item;
    if (t.mode >= 3) {
        // a small batch (the reference calls retrieve() with ONE query) has too few (query, chunk) items to fill
        // 148 SMs x 24 warps: cut the items shorter until there are a few thousand of them
        while (l.G > 1 && (int64_t)B * ((ix->n_sub + l.G - 1) / l.G) < 4096) l.G = (l.G + 1) / 2;
        l.n_chunks = (ix->n_sub + l.G - 1) / l.G;
        const int64_t chunk_docs = (int64_t)l.G * prw::kSub;
        const int64_t c2 = (t.docs_per_launch + chunk_docs - 1) / chunk_docs;
        if (c2 > c) c = c2;
        // docs_per_launch keeps a big batch's posting slice L2-resident; a small batch reads each posting a few times
        // at most, and every extra launch costs it a merge and a grid ramp: give each launch >= 32k items
        const int64_t c3 = B > 0 ? (32768 + (int64_t)B - 1) / B : 1;
        if (c3 > c) c = c3;
    } else {
        const int64_t n_tiles = ((int64_t)ix->n_docs + t.tile_docs - 1) / t.tile_docs;
        l.n_chunks = (int)((n_tiles + t.tiles_per_item - 1) / t.tiles_per_item);
    }
    if (c < 1) c = 1;
    if (c > l.n_chunks) c = l.n_chunks > 0 ? l.n_chunks : 1;
    l.C = (int)c;
    // Launch plan.  The first launch has no k-th score to filter with and scans every sub-tile; each later launch
    // filters with the scores of everything before it.  So a large batch (enough work items per chunk to fill the
    // GPU) starts with a SMALL launch and doubles -- c0, c0, 2 c0, 4 c0, ... up to C chunks -- which keeps the share
    // of documents scored without a useful threshold small even on a short shard (8 GPUs: 2.6M documents each).
    {
        int64_t c0 = B > 0 ? (4096 + (int64_t)B - 1) / B : l.C;
        if (c0 < 1) c0 = 1;
        if (c0 > l.C || t.mode < 3) c0 = l.C;
        int pos = 0, cl = (int)c0;
        while (pos < l.n_chunks) {
            const int take = cl < l.n_chunks - pos ? cl : l.n_chunks - pos;
            l.launch_chunk0.push_back(pos);
            l.launch_chunks.push_back(take);
            pos += take;
            cl = pos < l.C ? pos : l.C;
        }
        l.L = (int)l.launch_chunk0.size();
    }
    size_t o = 0;
    l.off_status = o;   o = align_up(o + 64, 256);
    l.off_counters = o; o = align_up(o + (size_t)(l.L + 1) * 4, 256);
    l.off_theta = o;    o = align_up(o + (size_t)B * 4, 256);
    l.off_plan_theta = o; o = align_up(o + (size_t)B * 4, 256);
    l.off_plan_mask = o;  o = align_up(o + (size_t)B * 4, 256);
    l.off_plan_m = o;     o = align_up(o + (size_t)B * 4, 256);
    l.off_run_s = o;    o = align_up(o + (size_t)B * K * 4, 256);
    l.off_run_d = o;    o = align_up(o + (size_t)B * K * 4, 256);
    l.off_part_s = o;   o = align_up(o + (size_t)B * l.C * K * 4, 256);
    l.off_part_d = o;   o = align_up(o + (size_t)B * l.C * K * 4, 256);
    // lean kernel: posting cursors of long queries, kCursorCap per resident warp (at most 32 warps per SM)
    l.off_cursors = o;
    if (t.mode == 8) o = align_up(o + (size_t)ix->num_sms * 32 * prl::kCursorCap * prl::kCursorWords * 4, 256);
    l.total = o;
    return l;
}

typedef void (*score_fn_t)(const ScoreArgs);

template <int NT, int MINB>
score_fn_t pick_score(int E)
{
    if (E == 1) return bm25_score_kernel<NT, 1, MINB>;
    if (E == 2) return bm25_score_kernel<NT, 2, MINB>;
    return bm25_score_kernel<NT, 4, MINB>;
}

score_fn_t pick_score_fn(int threads, int E)
{
    if (threads == 256) return pick_score<256, 4>(E);
    if (threads == 1024) return pick_score<1024, 1>(E);
    return pick_score<512, 2>(E);
}

// The templated warp-autonomous kernels are instantiated in their own translation units (bm25_kernels_*.cu) so the
// library compiles in parallel; bm25_kernels.h declares the pickers.
using prk::pick_flat_fn;
using prk::pick_lean_fn;
using prk::pick_warp_fn;
using prk::warp_fn_t;

int launch_merge(int E, dim3 grid, cudaStream_t st, const float *ps, const int32_t *pd, int C,
                 int64_t sq, int64_t sc, float *rs, int32_t *rd, float *rt, int B, int K, int fin,
                 float *os, int32_t *od, int base, int n_docs)
{
    if (E == 1) bm25_merge_kernel<1><<<grid, 128, 0, st>>>(ps, pd, C, sq, sc, rs, rd, rt, B, K, fin, os, od, base, n_docs);
    else if (E == 2) bm25_merge_kernel<2><<<grid, 128, 0, st>>>(ps, pd, C, sq, sc, rs, rd, rt, B, K, fin, os, od, base, n_docs);
    else bm25_merge_kernel<4><<<grid, 128, 0, st>>>(ps, pd, C, sq, sc, rs, rd, rt, B, K, fin, os, od, base, n_docs);
    PR_CUDA_CHECK(cudaGetLastError());
    return PR_OK;
}

void default_tuning(pr_bm25_tuning_t *t)
{
    t->tile_docs = 24576;
    t->tiles_per_item = 4;
    t->threads = 512;
    t->mode = 8;
    t->min_items = 2048;
    t->cand_cap = 1024;
    t->subs_per_item = 24;
    t->warps_per_cta = 8;
    t->docs_per_launch = 393216;
    t->lazy_zero = 2;
    t->rescore_cost = 64;
}

int check_tuning(const pr_bm25_tuning_t &t)
{
    if (t.threads != 256 && t.threads != 512 && t.threads != 1024) {
        pr_set_error("tuning.threads must be 256, 512 or 1024 (got %d)", t.threads);
        return PR_EINVAL;
    }
    if (t.tile_docs <= 0 || t.tile_docs % (4 * t.threads) != 0) {
        pr_set_error("tuning.tile_docs must be a positive multiple of 4*threads (got %d)", t.tile_docs);
        return PR_EINVAL;
    }
    if (t.subs_per_item < 1 || t.docs_per_launch < 1 ||
        (t.warps_per_cta != 4 && t.warps_per_cta != 8 && t.warps_per_cta != 9 && t.warps_per_cta != 10 && t.warps_per_cta != 12 &&
         t.warps_per_cta != 13 && t.warps_per_cta != 16) ||
        (t.mode >= 5 && t.warps_per_cta != 4 && t.warps_per_cta != 8 && t.warps_per_cta != 10 && t.warps_per_cta != 12) ||
        (t.mode < 5 && t.warps_per_cta == 10) || (t.mode == 7 && t.warps_per_cta == 10)) {
        pr_set_error("bad tuning (subs_per_item=%d docs_per_launch=%d warps_per_cta=%d; warps_per_cta is 4, 8, 9, 12, 13 or 16; "
                     "4, 8 or 12 for modes 5/6)",
                     t.subs_per_item, t.docs_per_launch, t.warps_per_cta);
        return PR_EINVAL;
    }
    if (t.tiles_per_item < 1 || t.mode < 1 || t.mode > 8 || t.min_items < 1 || t.cand_cap < 32) {
        pr_set_error("bad tuning (tiles_per_item=%d mode=%d min_items=%d cand_cap=%d)",
                     t.tiles_per_item, t.mode, t.min_items, t.cand_cap);
        return PR_EINVAL;
    }
    if (score_smem_bytes(t.tile_docs, t.cand_cap, t.threads / 32, PR_MAX_K) > 227 * 1024) {
        pr_set_error("tuning needs more than 227 KB of shared memory per CTA");
        return PR_EINVAL;
    }
    return PR_OK;
}

}  // namespace

extern "C" int pr_index_create(pr_index_t **out, int device, int64_t n_docs_global,
                               int32_t doc_id_base, int32_t n_docs, int32_t n_terms, int64_t nnz,
                               const int64_t *indptr_dev, const int32_t *doc_ids_dev,
                               const float *weights_dev)
{
    if (!out || !indptr_dev || n_docs < 0 || n_terms < 0 || nnz < 0 || n_docs_global < n_docs ||
        doc_id_base < 0 || (nnz > 0 && (!doc_ids_dev || !weights_dev))) {
        pr_set_error("pr_index_create: bad argument");
        return PR_EINVAL;
    }
    if (((uintptr_t)doc_ids_dev & 15) || ((uintptr_t)weights_dev & 15)) {
        pr_set_error("pr_index_create: doc_ids_dev and weights_dev must be 16-byte aligned");
        return PR_EINVAL;
    }
    if ((int64_t)doc_id_base + n_docs > 0x7fffffffLL) {
        pr_set_error("pr_index_create: global doc ids exceed int32");
        return PR_EINVAL;
    }
    PR_CUDA_CHECK(cudaSetDevice(device));
    int32_t *bad = nullptr;
    int32_t h_bad3[3] = {0, -1, 0};  // flags, min weight bits (start at 0xffffffff), max weight bits
    // one-time validation scratch; freed before returning (not on the query path)
    PR_CUDA_CHECK(cudaMalloc(&bad, 12));
    cudaError_t e = cudaMemcpy(bad, h_bad3, 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        bm25_validate_kernel<<<1184, 256>>>(indptr_dev, doc_ids_dev, weights_dev, n_terms, nnz, n_docs, bad);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(h_bad3, bad, 12, cudaMemcpyDeviceToHost);
    cudaFree(bad);
    const int32_t h_bad = h_bad3[0];
    if (e != cudaSuccess) {
        pr_set_error("pr_index_create: validation failed to run: %s", cudaGetErrorString(e));
        return PR_ECUDA;
    }
    if (h_bad) {
        pr_set_error("pr_index_create: invalid index (%s%s%s%s)", (h_bad & 1) ? "indptr not monotone/consistent; " : "",
                     (h_bad & 2) ? "doc id outside [0, n_docs); " : "",
                     (h_bad & 4) ? "doc ids not ascending within a term; " : "",
                     (h_bad & 8) ? "weight negative or not finite; " : "");
        return PR_EINVAL;
    }
    pr_index *ix = new pr_index();
    ix->device = device;
    ix->n_docs_global = n_docs_global;
    ix->doc_id_base = doc_id_base;
    ix->n_docs = n_docs;
    ix->n_terms = n_terms;
    ix->nnz = nnz;
    ix->indptr = indptr_dev;
    ix->doc_ids = doc_ids_dev;
    ix->weights = weights_dev;
    ix->last_launches = 0;
    ix->profiling = 0;
    ix->ev_used = 0;
    ix->heavy_row = nullptr;
    ix->tp = nullptr;
    ix->n_rows = 0;
    ix->n_sub = (int32_t)(((int64_t)n_docs + prw::kSub - 1) >> prw::kSubShift);
    ix->heavy_min_df = 0;
    ix->hot_of_row = nullptr;
    ix->hot_off = nullptr;
    ix->hot_stream = nullptr;
    ix->n_hot = 0;
    ix->hot_min_df = 0;
    ix->hot_stream_bytes = 0;
    ix->cold_stream = nullptr;
    ix->hot_base_g = 0;
    ix->lean_ok = false;
    ix->term_maxw = nullptr;
    ix->row_q = nullptr;
    {   // lazily re-zeroed accumulators need every weight in [2^-30, 2^10] (bm25_warp.cuh)
        float wmin, wmax;
        memcpy(&wmin, &h_bad3[1], 4);
        memcpy(&wmax, &h_bad3[2], 4);
        ix->lazy_ok = nnz > 0 && wmin >= 9.313225746154785e-10f && wmax <= 1024.f;
    }
    default_tuning(&ix->tuning);
    cudaDeviceProp prop;
    PR_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    ix->num_sms = prop.multiProcessorCount;
    *out = ix;
    return PR_OK;
}

extern "C" int pr_index_destroy(pr_index_t *index)
{
    if (index)
        for (cudaEvent_t e : index->ev) cudaEventDestroy(e);
    delete index;
    return PR_OK;
}

extern "C" size_t pr_index_aux_bytes(const pr_index_t *index, size_t table_budget_bytes)
{
    if (!index) return 0;
    // heavy_row[n_terms] + block counts + row_term/tp within the budget
    const size_t fixed = 2 * align_up((size_t)index->n_terms * 4, 256) + align_up(((size_t)index->n_terms / 1024 + 2) * 4, 256) + 256;
    // + the cold stream of the lean kernel: 8 bytes per posting
    return fixed + align_up(table_budget_bytes, 256) + align_up((size_t)index->nnz * 8 + 256, 256);
}

extern "C" int pr_index_build_aux(pr_index_t *index, void *aux_dev, size_t aux_bytes, pr_stream_t stream)
{
    if (!index || !aux_dev) {
        pr_set_error("pr_index_build_aux: null argument");
        return PR_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = index->n_terms;
    const int n_blocks = (nt + 1023) / 1024;
    unsigned char *p = (unsigned char *)aux_dev;
    size_t o = 0;
    int32_t *heavy_row = (int32_t *)(p + o);  o = align_up(o + (size_t)nt * 4, 256);
    float *term_maxw = (float *)(p + o);      o = align_up(o + (size_t)nt * 4, 256);
    int32_t *block_cnt = (int32_t *)(p + o);  o = align_up(o + ((size_t)nt / 1024 + 2) * 4, 256);
    int32_t *count = (int32_t *)(p + o);      o += 256;
    if (o > aux_bytes) {
        pr_set_error("pr_index_build_aux: aux buffer of %zu bytes, need at least %zu", aux_bytes, o);
        return PR_EWORKSPACE;
    }
    // the cold stream (8 bytes per posting) is carved first when the buffer was sized by pr_index_aux_bytes
    // with a budget that leaves room for it; a smaller buffer simply has none (mode 8 then runs as mode 6)
    const size_t cold_bytes = align_up((size_t)index->nnz * 8 + 256, 256);
    unsigned char *cold = nullptr;
    if (index->nnz > 0 && aux_bytes >= o + cold_bytes) {
        cold = p + o;
        o += cold_bytes;
    }
    // the boundary table gets a fifth of the table budget (at least 16 MB of it), the hot stream the rest
    const size_t budget = aux_bytes - o;
    size_t table_bytes = budget / 5;
    if (table_bytes < ((size_t)16 << 20)) table_bytes = budget < ((size_t)16 << 20) ? budget : ((size_t)16 << 20);
    const size_t per_row = ((size_t)index->n_sub + 1) * 4 + 12;  // tp row + row_term entry + row_q pair
    // smallest df threshold (doubling from kLightDf) whose table fits the budget
    int64_t min_df = prw::kLightDf;
    int32_t rows = 0;
    for (;;) {
        PR_CUDA_CHECK(cudaMemsetAsync(count, 0, 4, st));
        if (nt > 0) prw::count_heavy_kernel<<<296, 256, 0, st>>>(index->indptr, nt, min_df, count);
        PR_CUDA_CHECK(cudaGetLastError());
        PR_CUDA_CHECK(cudaMemcpyAsync(&rows, count, 4, cudaMemcpyDeviceToHost, st));
        PR_CUDA_CHECK(cudaStreamSynchronize(st));
        if ((size_t)rows * per_row + 1024 <= table_bytes || rows == 0) break;
        min_df *= 2;
    }
    int32_t *row_term = (int32_t *)(p + o);   o = align_up(o + (size_t)(rows > 0 ? rows : 1) * 4, 256);
    float *row_q = (float *)(p + o);          o = align_up(o + (size_t)(rows > 0 ? rows : 1) * 8, 256);
    uint32_t *tp = (uint32_t *)(p + o);       o = align_up(o + (size_t)rows * ((size_t)index->n_sub + 1) * 4, 256);
    if (nt > 0) {
        prw::term_maxw_kernel<<<2368, 256, 0, st>>>(index->indptr, index->weights, nt, term_maxw);
        prw::heavy_block_count_kernel<<<n_blocks, 1024, 0, st>>>(index->indptr, nt, min_df, block_cnt);
        prw::heavy_block_scan_kernel<<<1, 32, 0, st>>>(block_cnt, n_blocks);
        prw::heavy_assign_kernel<<<n_blocks, 1024, 0, st>>>(index->indptr, nt, min_df, block_cnt, heavy_row, row_term);
        if (rows > 0) {
            prw::tp_fill_kernel<<<2368, 256, 0, st>>>(index->indptr, index->doc_ids, row_term, rows, index->n_sub, tp);
            prw::row_quantile_kernel<<<rows < 2368 ? rows : 2368, 256, 0, st>>>(index->indptr, index->weights, row_term, term_maxw,
                                                                               rows, row_q);
        }
    }
    PR_CUDA_CHECK(cudaGetLastError());
    PR_CUDA_CHECK(cudaStreamSynchronize(st));
    index->heavy_row = heavy_row;
    index->term_maxw = term_maxw;
    index->row_q = row_q;
    index->tp = tp;
    index->n_rows = rows;
    index->heavy_min_df = min_df;
    index->hot_of_row = nullptr;
    index->hot_off = nullptr;
    index->hot_stream = nullptr;
    index->n_hot = 0;
    index->hot_min_df = 0;
    index->hot_stream_bytes = 0;
    index->cold_stream = nullptr;
    index->hot_base_g = 0;
    index->lean_ok = false;
    if (cold) {
        prl::cold_fill_kernel<<<2368, 256, 0, st>>>(index->doc_ids, index->weights, index->nnz, (uint2 *)cold);
        PR_CUDA_CHECK(cudaGetLastError());
        PR_CUDA_CHECK(cudaStreamSynchronize(st));
        index->cold_stream = cold;
        index->lean_ok = (uint64_t)index->nnz + 64 < ((uint64_t)1 << 32);  // without a hot stream
    }

    // ---- hot posting stream (bm25_hot.cuh) for the tabulated terms with >= kHotMinSeg postings
    // per sub-tile, as many of them as the rest of the buffer holds (threshold doubles until it fits)
    const int n_sub = index->n_sub;
    size_t oh = o;
    int32_t *hot_of_row = (int32_t *)(p + oh);  oh = align_up(oh + (size_t)(rows > 0 ? rows : 1) * 4, 256);
    int32_t *hot_rows = (int32_t *)(p + oh);    oh = align_up(oh + (size_t)(rows > 0 ? rows : 1) * 4, 256);
    int32_t *n_hot_dev = (int32_t *)(p + oh);
    uint32_t *total_dev = (uint32_t *)(p + oh + 64);
    oh += 256;
    if (rows == 0 || n_sub == 0 || oh + 4096 > aux_bytes) return PR_OK;
    int64_t hot_df = ((int64_t)prh::kHotMinSeg * index->n_docs + prw::kSub - 1) / prw::kSub;
    if (hot_df < min_df + 1) hot_df = min_df + 1;   // hot rows are a subset of the tabulated rows (df > min_df)
    for (;; hot_df *= 2) {
        int32_t H = 0;
        prh::hot_assign_kernel<<<1, 32, 0, st>>>(index->indptr, row_term, rows, hot_df, hot_of_row, hot_rows, n_hot_dev);
        PR_CUDA_CHECK(cudaGetLastError());
        PR_CUDA_CHECK(cudaMemcpyAsync(&H, n_hot_dev, 4, cudaMemcpyDeviceToHost, st));
        PR_CUDA_CHECK(cudaStreamSynchronize(st));
        if (H == 0) return PR_OK;
        const int64_t n_seg = (int64_t)H * n_sub;
        const int64_t nb = (n_seg + prh::kScanChunk - 1) / prh::kScanChunk;
        size_t os = oh;
        uint32_t *hot_off = (uint32_t *)(p + os);    os = align_up(os + (size_t)H * ((size_t)n_sub + 1) * 4, 256);
        uint32_t *block_sum = (uint32_t *)(p + os);  os = align_up(os + (size_t)nb * 4, 256);
        if (os > aux_bytes || nb > 0x7fffffff) continue;
        uint32_t total = 0;
        prh::hot_units_kernel<<<(unsigned)nb, 256, 0, st>>>(tp, hot_rows, n_sub, n_seg, block_sum);
        prh::hot_block_scan_kernel<<<1, 32, 0, st>>>(block_sum, (int)nb, total_dev);
        PR_CUDA_CHECK(cudaGetLastError());
        PR_CUDA_CHECK(cudaMemcpyAsync(&total, total_dev, 4, cudaMemcpyDeviceToHost, st));
        PR_CUDA_CHECK(cudaStreamSynchronize(st));
        if (os + (size_t)total * prh::kUnitBytes > aux_bytes) continue;
        unsigned char *stream_dev = p + os;
        prh::hot_offsets_kernel<<<(unsigned)nb, 256, 0, st>>>(tp, hot_rows, n_sub, n_seg, block_sum, hot_off);
        prh::hot_fill_kernel<<<2368, 256, 0, st>>>(index->indptr, index->doc_ids, index->weights, row_term, tp, hot_rows, n_sub,
                                                  n_seg, hot_off, stream_dev);
        PR_CUDA_CHECK(cudaGetLastError());
        PR_CUDA_CHECK(cudaStreamSynchronize(st));
        index->hot_of_row = hot_of_row;
        index->hot_off = hot_off;
        index->hot_stream = stream_dev;
        index->n_hot = H;
        index->hot_min_df = hot_df;
        index->hot_stream_bytes = (int64_t)total * prh::kUnitBytes;
        if (cold) {  // one granule space: cold stream, tables, hot stream
            const uint64_t g0 = (uint64_t)(stream_dev - cold) >> 3;
            index->hot_base_g = (uint32_t)g0;
            index->lean_ok = g0 + (uint64_t)total * 32 + 64 < ((uint64_t)1 << 32);
        }
        return PR_OK;
    }
}

extern "C" int pr_index_hot_info(const pr_index_t *index, int32_t *n_hot, int64_t *min_df, int64_t *stream_bytes)
{
    if (!index || !n_hot || !min_df || !stream_bytes) {
        pr_set_error("pr_index_hot_info: null argument");
        return PR_EINVAL;
    }
    *n_hot = index->n_hot;
    *min_df = index->hot_min_df;
    *stream_bytes = index->hot_stream_bytes;
    return PR_OK;
}

extern "C" int pr_index_lean_info(const pr_index_t *index, int32_t *lean_ok, int64_t *cold_bytes)
{
    if (!index || !lean_ok || !cold_bytes) {
        pr_set_error("pr_index_lean_info: null argument");
        return PR_EINVAL;
    }
    *lean_ok = index->lean_ok ? 1 : 0;
    *cold_bytes = index->cold_stream ? index->nnz * 8 : 0;
    return PR_OK;
}

extern "C" int pr_index_aux_info(const pr_index_t *index, int32_t *n_rows, int64_t *min_df)
{
    if (!index || !n_rows || !min_df) {
        pr_set_error("pr_index_aux_info: null argument");
        return PR_EINVAL;
    }
    *n_rows = index->n_rows;
    *min_df = index->heavy_min_df;
    return PR_OK;
}

extern "C" int pr_index_set_profiling(pr_index_t *index, int enable)
{
    if (!index) {
        pr_set_error("pr_index_set_profiling: null index");
        return PR_EINVAL;
    }
    index->profiling = enable ? 1 : 0;
    index->ev_used = 0;
    return PR_OK;
}

extern "C" int pr_bm25_profile(pr_index_t *index, float *score_ms, int32_t *score_launches)
{
    if (!index || !score_ms || !score_launches) {
        pr_set_error("pr_bm25_profile: null argument");
        return PR_EINVAL;
    }
    float total = 0.f;
    for (int i = 0; i + 1 < index->ev_used; i += 2) {
        PR_CUDA_CHECK(cudaEventSynchronize(index->ev[i + 1]));
        float ms = 0.f;
        PR_CUDA_CHECK(cudaEventElapsedTime(&ms, index->ev[i], index->ev[i + 1]));
        total += ms;
    }
    *score_ms = total;
    *score_launches = index->ev_used / 2;
    return PR_OK;
}

extern "C" int pr_index_set_tuning(pr_index_t *index, const pr_bm25_tuning_t *tuning)
{
    if (!index || !tuning) {
        pr_set_error("pr_index_set_tuning: null argument");
        return PR_EINVAL;
    }
    pr_bm25_tuning_t t = index->tuning;
    if (tuning->tile_docs) t.tile_docs = tuning->tile_docs;
    if (tuning->tiles_per_item) t.tiles_per_item = tuning->tiles_per_item;
    if (tuning->threads) t.threads = tuning->threads;
    if (tuning->mode) t.mode = tuning->mode;
    if (tuning->min_items) t.min_items = tuning->min_items;
    if (tuning->cand_cap) t.cand_cap = tuning->cand_cap;
    if (tuning->subs_per_item) t.subs_per_item = tuning->subs_per_item;
    if (tuning->warps_per_cta) t.warps_per_cta = tuning->warps_per_cta;
    if (tuning->docs_per_launch) t.docs_per_launch = tuning->docs_per_launch;
    if (tuning->lazy_zero) t.lazy_zero = tuning->lazy_zero;
    if (tuning->rescore_cost) t.rescore_cost = tuning->rescore_cost;
    const int rc = check_tuning(t);
    if (rc != PR_OK) return rc;
    index->tuning = t;
    return PR_OK;
}

extern "C" int pr_index_get_tuning(const pr_index_t *index, pr_bm25_tuning_t *tuning)
{
    if (!index || !tuning) {
        pr_set_error("pr_index_get_tuning: null argument");
        return PR_EINVAL;
    }
    *tuning = index->tuning;
    return PR_OK;
}

extern "C" size_t pr_bm25_workspace_bytes(const pr_index_t *index, int32_t n_queries, int32_t k)
{
    if (!index || n_queries < 0 || k < 1 || k > PR_MAX_K) return 0;
    return make_layout(index, n_queries, k).total;
}

extern "C" int64_t pr_bm25_last_launches(const pr_index_t *index) { return index ? index->last_launches : 0; }

extern "C" int pr_bm25_topk(pr_index_t *index, int32_t n_queries, const int64_t *q_indptr_dev,
                            const int32_t *q_terms_dev, int32_t k, float *out_scores_dev,
                            int32_t *out_doc_ids_dev, void *workspace_dev, size_t workspace_bytes,
                            pr_stream_t stream)
{
    if (!index || n_queries < 0 || !q_indptr_dev || !out_scores_dev || !out_doc_ids_dev || !workspace_dev) {
        pr_set_error("pr_bm25_topk: bad argument");
        return PR_EINVAL;
    }
    if (k < 1 || k > PR_MAX_K) {
        pr_set_error("pr_bm25_topk: k must be in [1, %d] (got %d)", PR_MAX_K, k);
        return PR_EINVAL;
    }
    if ((int64_t)k > index->n_docs_global) {
        pr_set_error("k of %d is larger than the number of documents %lld", k, (long long)index->n_docs_global);
        return PR_ERANGE;
    }
    const Layout l = make_layout(index, n_queries, k);
    if (workspace_bytes < l.total) {
        pr_set_error("pr_bm25_topk: workspace of %zu bytes, need %zu", workspace_bytes, l.total);
        return PR_EWORKSPACE;
    }
    index->last_launches = 0;
    index->ev_used = 0;
    if (n_queries == 0) return PR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const pr_bm25_tuning_t &t = index->tuning;
    const int E = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
    unsigned char *ws = (unsigned char *)workspace_dev;
    int32_t *status = (int32_t *)(ws + l.off_status);
    int32_t *counters = (int32_t *)(ws + l.off_counters);
    float *theta = (float *)(ws + l.off_theta);
    float *plan_theta = (float *)(ws + l.off_plan_theta);
    uint32_t *plan_mask = (uint32_t *)(ws + l.off_plan_mask);
    float *plan_m = (float *)(ws + l.off_plan_m);
    float *run_s = (float *)(ws + l.off_run_s);
    int32_t *run_d = (int32_t *)(ws + l.off_run_d);
    float *part_s = (float *)(ws + l.off_part_s);
    int32_t *part_d = (int32_t *)(ws + l.off_part_d);

    const int64_t n_run = (int64_t)n_queries * k;
    int64_t n_init = n_run > l.L + 1 ? n_run : l.L + 1;
    if (n_init < n_queries) n_init = n_queries;
    bm25_init_kernel<<<(unsigned)((n_init + 255) / 256), 256, 0, st>>>(run_s, run_d, theta, plan_theta, plan_mask, plan_m, n_run,
                                                                       n_queries, counters, l.L + 1, status);
    PR_CUDA_CHECK(cudaGetLastError());
    index->last_launches++;

    const bool warp_mode = t.mode >= 3;
    if (warp_mode && (!index->heavy_row || (index->n_rows > 0 && !index->tp))) {
        pr_set_error("pr_bm25_topk: tuning.mode %d needs the index tables: call pr_index_build_aux first", t.mode);
        return PR_EINVAL;
    }
    const int nw = warp_mode ? t.warps_per_cta : t.threads / 32;
    const int threads = nw * 32;
    const bool flat_mode = t.mode >= 5;
    const bool lean_mode = t.mode == 8 && index->lean_ok;  // mode 8 without a cold stream runs the flat kernel (mode 6)
    const size_t smem = lean_mode ? prl::lean_smem_bytes(nw) : flat_mode ? prf::flat_smem_bytes(nw)
                                  : warp_mode ? prw::warp_smem_bytes(nw) : score_smem_bytes(t.tile_docs, t.cand_cap, nw, k);
    score_fn_t fn = nullptr;
    warp_fn_t wfn = nullptr;
    const void *kfn = nullptr;
    if (warp_mode) {
        wfn = lean_mode ? pick_lean_fn(nw, E) : flat_mode ? pick_flat_fn(nw, E, t.mode == 7) : pick_warp_fn(nw, E, index->lazy_ok && t.lazy_zero == 1);
        kfn = (const void *)wfn;
    } else {
        fn = pick_score_fn(t.threads, E);
        kfn = (const void *)fn;
    }
    PR_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PR_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 0;
    PR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, threads, smem));
    if (occ < 1) {
        pr_set_error("pr_bm25_topk: kernel does not fit (threads=%d smem=%zu)", threads, smem);
        return PR_EINVAL;
    }
    const dim3 mgrid((unsigned)((n_queries + 3) / 4));

    ScoreArgs a;
    a.indptr = index->indptr;
    a.doc_ids = index->doc_ids;
    a.weights = index->weights;
    a.q_indptr = q_indptr_dev;
    a.q_terms = q_terms_dev;
    a.run_theta = theta;
    a.part_s = part_s;
    a.part_d = part_d;
    a.status = status;
    a.nnz = index->nnz;
    a.n_docs = index->n_docs;
    a.n_terms = index->n_terms;
    a.doc_id_base = index->doc_id_base;
    a.n_queries = n_queries;
    a.K = k;
    a.tile_docs = t.tile_docs;
    a.tiles_per_item = t.tiles_per_item;
    a.cand_cap = t.cand_cap;

    prw::WarpArgs w;
    w.indptr = index->indptr;
    w.doc_ids = index->doc_ids;
    w.weights = index->weights;
    w.heavy_row = index->heavy_row;
    w.tp = index->tp;
    w.hot_of_row = index->hot_of_row;
    w.hot_off = index->hot_off;
    w.hot_stream = index->hot_stream;
    w.stream_base = index->cold_stream;
    w.hot_base_g = index->hot_base_g;
    w.cursors = lean_mode ? (uint32_t *)(ws + l.off_cursors) : nullptr;
    w.term_maxw = index->term_maxw;
    w.plan_mask = plan_mask;
    w.plan_m = plan_m;
    w.q_indptr = q_indptr_dev;
    w.q_terms = q_terms_dev;
    w.run_theta = theta;
    w.part_s = part_s;
    w.part_d = part_d;
    w.status = status;
    w.nnz = index->nnz;
    w.n_docs = index->n_docs;
    w.n_terms = index->n_terms;
    w.doc_id_base = index->doc_id_base;
    w.n_queries = n_queries;
    w.K = k;
    w.n_sub = index->n_sub;
    w.subs_per_item = l.G;

    if (l.L == 0) {  // shard without documents: only the (empty) finalisation
        return launch_merge(E, mgrid, st, part_s, part_d, 0, 0, 0, run_s, run_d, theta, n_queries, k, 1,
                            out_scores_dev, out_doc_ids_dev, index->doc_id_base, index->n_docs);
    }
    for (int li = 0; li < l.L; ++li) {
        const int Cl = l.launch_chunks[li], chunk0 = l.launch_chunk0[li];
        const int64_t items = (int64_t)n_queries * Cl;
        int64_t grid = (int64_t)occ * index->num_sms;
        const int64_t need = warp_mode ? (items + nw - 1) / nw : items;
        if (grid > need) grid = need;
        if (index->profiling) {
            while ((int)index->ev.size() < index->ev_used + 2) {
                cudaEvent_t ev;
                PR_CUDA_CHECK(cudaEventCreate(&ev));
                index->ev.push_back(ev);
            }
            PR_CUDA_CHECK(cudaEventRecord(index->ev[index->ev_used], st));
        }
#ifdef PR_STATS
        PR_CUDA_CHECK(cudaMemcpyToSymbolAsync(pr_stats_launch, &li, sizeof(int), 0, cudaMemcpyHostToDevice, st));
#endif
        if (warp_mode) {
            w.chunk0 = chunk0;
            w.n_chunks_launch = Cl;
            w.counter = counters + li;
            w.mode = li == 0 ? (flat_mode ? 5 : 3) : (t.mode == 8 ? 6 : t.mode);  // the first launch has no running k-th score yet
            wfn<<<(unsigned)grid, threads, smem, st>>>(w);
        } else {
            a.chunk0 = chunk0;
            a.n_chunks_launch = Cl;
            a.counter = counters + li;
            a.mode = li == 0 ? 1 : t.mode;
            fn<<<(unsigned)grid, threads, smem, st>>>(a);
        }
        PR_CUDA_CHECK(cudaGetLastError());
        if (index->profiling) {
            PR_CUDA_CHECK(cudaEventRecord(index->ev[index->ev_used + 1], st));
            index->ev_used += 2;
        }
        const int rc = launch_merge(E, mgrid, st, part_s, part_d, Cl, (int64_t)Cl * k, k, run_s, run_d, theta,
                                    n_queries, k, li == l.L - 1, out_scores_dev, out_doc_ids_dev,
                                    index->doc_id_base, index->n_docs);
        if (rc != PR_OK) return rc;
        index->last_launches += 2;
        if (t.mode == 7 && li + 1 < l.L) {  // which terms the next launch may skip, given the new k-th scores
            prf::bm25_plan_kernel<<<mgrid, 128, 0, st>>>(q_indptr_dev, q_terms_dev, index->indptr, index->heavy_row,
                                                        index->term_maxw, index->row_q, theta, plan_theta, plan_mask, plan_m,
                                                        n_queries, index->n_terms, (float)t.rescore_cost);
            PR_CUDA_CHECK(cudaGetLastError());
            index->last_launches++;
        }
    }
    return PR_OK;
}

#ifdef PR_STATS
// instrumented variant build only: copies the per-launch counters out and clears them
extern "C" int pr_debug_stats(unsigned long long *out_host, int n)
{
    PR_CUDA_CHECK(cudaDeviceSynchronize());
    PR_CUDA_CHECK(cudaMemcpyFromSymbol(out_host, pr_stats_dev, (size_t)n * 8));
    static unsigned long long zeros[512 * 16];
    PR_CUDA_CHECK(cudaMemcpyToSymbol(pr_stats_dev, zeros, sizeof(zeros)));
    return PR_OK;
}
#endif

extern "C" int pr_bm25_status(const void *workspace_dev, pr_stream_t stream)
{
    if (!workspace_dev) {
        pr_set_error("pr_bm25_status: null workspace");
        return PR_EINVAL;
    }
    int32_t h = 0;
    PR_CUDA_CHECK(cudaMemcpyAsync(&h, workspace_dev, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PR_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    if (h & 1) {
        pr_set_error("query term id outside [0, n_terms)");
        return PR_ERANGE;
    }
    return PR_OK;
}

extern "C" int pr_topk_merge(int32_t n_queries, int32_t k, int32_t n_lists, const float *scores_dev,
                             const int32_t *ids_dev, float *out_scores_dev, int32_t *out_ids_dev,
                             pr_stream_t stream)
{
    if (n_queries < 0 || n_lists < 1 || !scores_dev || !ids_dev || !out_scores_dev || !out_ids_dev) {
        pr_set_error("pr_topk_merge: bad argument");
        return PR_EINVAL;
    }
    if (k < 1 || k > PR_MAX_K) {
        pr_set_error("pr_topk_merge: k must be in [1, %d] (got %d)", PR_MAX_K, k);
        return PR_EINVAL;
    }
    if (n_queries == 0) return PR_OK;
    const int E = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
    // lists are [n_lists, n_queries, k]: query stride k, list stride n_queries*k
    return launch_merge(E, dim3((unsigned)((n_queries + 3) / 4)), (cudaStream_t)stream, scores_dev, ids_dev,
                        n_lists, k, (int64_t)n_queries * k, nullptr, nullptr, nullptr, n_queries, k, 1,
                        out_scores_dev, out_ids_dev, 0, -1);
}
