// Batched BM25 scoring + top-k over a term-major CSR inverted index in HBM (sm_100a): host side of
// pr_bm25_topk (launch plan, workspace layout, C ABI) and the small kernels around the scoring kernel
// (bm25_lean.cuh): workspace initialisation, the k-way merge of ranked lists, index validation.
//
// Replaces, at token-id level, what BM25Retriever.retrieve does on the CPU for one query at a time
// (/root/reference/exp_rag.py:426,428,492; utils.py:640 -> bm25s `_compute_relevance_from_scores` +
// `selection.topk`): a dense fp32 accumulator over all documents, one
// `scores[doc] += w` per posting of every query token in query order, then top-k.  Here the document
// range is cut into 2048-document tiles that live in shared memory, a work item is (query, run of
// tiles), one warp owns an item, and per-item ranked lists are merged per query.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "bm25_lean.cuh"
#include "bm25_kernels.h"

namespace {

// Canonical zero-score tail: fewer than K positive scores -> the lowest local doc ids not listed yet.
__device__ __forceinline__ void fill_zero_tail(float *out_s, int32_t *out_d, int K, int nvalid, int doc_id_base, int n_docs)
{
    int cand = 0;
    for (int pos = nvalid; pos < K; ++pos) {
        while (cand < n_docs) {
            bool listed = false;
            for (int i = 0; i < nvalid; ++i)
                if (out_d[i] == cand + doc_id_base) listed = true;
            if (!listed) break;
            ++cand;
        }
        if (cand >= n_docs) break;
        out_s[pos] = 0.f;
        out_d[pos] = cand + doc_id_base;
        ++cand;
    }
}

// Walk the sorted lists c = c_begin, c_begin + c_step, ... < C of one query.  The lists are taken 32 at a time -- every
// lane reads the head of one list, and only lists whose head can still enter are walked -- and the heads of kHeadAhead
// such groups are loaded before the first is looked at (a single query leaves thousands of lists: one exposed L2 round
// trip per group made the merge the longest kernel of the call).  `floor` = known lower bound of the final k-th score:
// entries below it are never needed.
constexpr int kHeadAhead = 4;
template <int E>
__device__ __forceinline__ void merge_lists(WarpTopK<E> &L, float &ks, int &kd, const float *__restrict__ ps0,
                                            const int32_t *__restrict__ pd0, int64_t stride_c, int c_begin, int c_step, int C,
                                            int K, float floor, int lane)
{
    for (int cb = c_begin; cb < C; cb += kHeadAhead * 32 * c_step) {
        float hs[kHeadAhead];
        int hd[kHeadAhead];
#pragma unroll
        for (int u = 0; u < kHeadAhead; ++u) {
            const int cl = cb + (u * 32 + lane) * c_step;
            hs[u] = -1.f;
            hd[u] = -1;
            if (cl < C) {
                hs[u] = ps0[cl * stride_c];
                hd[u] = pd0[cl * stride_c];
            }
        }
#pragma unroll
        for (int u = 0; u < kHeadAhead; ++u) {
            const int c0 = cb + u * 32 * c_step;
            unsigned m = __ballot_sync(PR_FULL_MASK, hs[u] >= 0.f && hd[u] >= 0 && hs[u] >= floor && pr_beats(hs[u], hd[u], ks, kd));
            while (m) {
                const int c = c0 + (__ffs(m) - 1) * c_step;
                m &= m - 1;
                const float *ps = ps0 + c * stride_c;
                const int32_t *pd = pd0 + c * stride_c;
                // the whole list in one coalesced load (lane i holds entry 32 e + i), then walked from registers: entry by
                // entry it was up to K dependent L2 round trips per list
                bool done = false;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i0 = e * 32;
                    if (done || i0 >= K) break;
                    const float es = i0 + lane < K ? ps[i0 + lane] : -1.f;
                    const int ed = i0 + lane < K ? pd[i0 + lane] : -1;
                    const int n = min(32, K - i0);
                    for (int i = 0; i < n; ++i) {
                        const float bs = __shfl_sync(PR_FULL_MASK, es, i);
                        const int bd = __shfl_sync(PR_FULL_MASK, ed, i);
                        if (bs < 0.f || bd < 0 || bs < floor || !pr_beats(bs, bd, ks, kd)) {  // empty slot / missing entry /
                            done = true;                                                    // below the bound / sorted: the
                            break;                                                          // rest lose too
                        }
                        L.insert(bs, bd, lane);
                        L.kth(K, ks, kd);
                    }
                }
            }
        }
    }
}

// Fold the C per-item lists of a launch into the per-query running list (one warp per query), publish the new
// k-th score, and on the last launch write the output.  Also the merge behind pr_topk_merge (run_s == nullptr).
template <int E>
__global__ void __launch_bounds__(128) bm25_merge_kernel(
    const float *__restrict__ part_s, const int32_t *__restrict__ part_d, int C,
    int64_t stride_q, int64_t stride_c, float *run_s, int32_t *run_d, float *theta, float *const *peer_theta, int n_peers,
    int B, int K, int finalize, float *out_s, int32_t *out_d, int doc_id_base, int n_docs)
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
    const float floor = theta ? theta[q] : -1.f;
    merge_lists<E>(L, ks, kd, part_s + q * stride_q, part_d + q * stride_q, stride_c, 0, 1, C, K, floor, lane);
    if (run_s) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = e * 32 + lane;
            if (i < K) {
                run_s[(size_t)q * K + i] = L.s[e];
                run_d[(size_t)q * K + i] = L.d[e];
            }
        }
        if (ks > 0.f) prw::raise_theta(theta, peer_theta, n_peers, q, ks, lane);
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
        if (n_docs >= 0 && nvalid < K && lane == 0)
            fill_zero_tail(out_s + (size_t)q * K, out_d + (size_t)q * K, K, nvalid, doc_id_base, n_docs);
    }
}

// The same fold for SMALL batches, where one launch leaves thousands of per-item lists per query (a single query over
// 21M documents: 10,262): one CTA of 32 warps per query, every warp folds a strided share of the lists into its own
// register list (the one-warp kernel above needed 100 us of dependent loads for it), the 32 lists meet in shared memory
// and warp 0 folds them into the running list.
constexpr int kWideWarps = 32;
template <int E>
__global__ void __launch_bounds__(kWideWarps * 32) bm25_merge_wide_kernel(
    const float *__restrict__ part_s, const int32_t *__restrict__ part_d, int C, int64_t stride_q, int64_t stride_c,
    float *run_s, int32_t *run_d, float *theta, float *const *peer_theta, int n_peers, int K, int finalize, float *out_s,
    int32_t *out_d, int doc_id_base, int n_docs)
{
    extern __shared__ __align__(16) unsigned char merge_smem[];
    float *sh_s = reinterpret_cast<float *>(merge_smem);                    // [kWideWarps][K]
    int32_t *sh_d = reinterpret_cast<int32_t *>(sh_s + kWideWarps * K);
    __shared__ float s_floor;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x;
    const float *ps0 = part_s + q * stride_q;
    const int32_t *pd0 = part_d + q * stride_q;
    WarpTopK<E> L;
    float ks = PR_SENT_SCORE;
    int kd = PR_SENT_DOC;
    // Phase A: a bound from the list HEADS alone.  The heads are distinct documents, so the K-th largest head score is a
    // lower bound of the final K-th score -- far stronger than theta[q] here, which is only the best K-th score of a
    // single short item (a single query over 21M documents: items of 6,144 documents).  With it phase B walks about K
    // lists in total instead of a sizeable share of the thousands.  The K-th largest of the C head scores is found by a
    // block-wide RADIX SELECT on the score bits (scores >= 0 order like their bit patterns): four passes of 8 bits, a
    // 256-bin histogram in shared memory each.  (Folding the heads through per-warp register lists and then the 32
    // lists through warp 0 was ~150 dependent insertions: 20 of the 27 us this kernel took for a single query.)
    __shared__ int s_hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int s_remaining;
    constexpr int kHeadRegs = 4;   // heads kept in registers (C <= 4096: every small batch on one GPU); the rest is re-read
    constexpr uint32_t kNoKey = 0xffffffffu;
    auto head_key = [&](int c) -> uint32_t {
        const float hs = ps0[c * stride_c];
        const int hd = pd0[c * stride_c];
        return (hs >= 0.f && hd >= 0) ? __float_as_uint(hs) : kNoKey;
    };
    uint32_t hk[kHeadRegs];
#pragma unroll
    for (int r = 0; r < kHeadRegs; ++r) {
        const int c = (int)threadIdx.x + r * kWideWarps * 32;
        hk[r] = c < C ? head_key(c) : kNoKey;
    }
    uint32_t prefix = 0u;
    int remaining = K;
#pragma unroll 1
    for (int shift = 24; shift >= 0 && remaining > 0; shift -= 8) {
        if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t hi_mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
        // one shared-memory atomic per (warp, bin): the scores of a query share their exponent, so in the first passes
        // every key falls into the same one or two bins (thousands of same-address atomics otherwise)
        auto vote = [&](uint32_t key) {
            const uint32_t bin = (key != kNoKey && (key & hi_mask) == prefix) ? ((key >> shift) & 255u) : 256u;
            const unsigned same = __match_any_sync(PR_FULL_MASK, bin);
            if (bin < 256u && lane == __ffs(same) - 1) atomicAdd(&s_hist[bin], __popc(same));
        };
#pragma unroll
        for (int r = 0; r < kHeadRegs; ++r) vote(hk[r]);
        for (int c0 = kHeadRegs * kWideWarps * 32; c0 < C; c0 += kWideWarps * 32) {   // (warp-uniform trip count)
            const int c = c0 + (int)threadIdx.x;
            vote(c < C ? head_key(c) : kNoKey);
        }
        __syncthreads();
        if (warp == 0) {   // lane l holds bins 255 - 8 l ... 248 - 8 l: the bin in which the running count from the top reaches `remaining`
            int cnt[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                cnt[j] = s_hist[255 - (lane * 8 + j)];
                sum += cnt[j];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(PR_FULL_MASK, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(PR_FULL_MASK, incl, 31);
            if (total < remaining) {   // fewer than K heads (only possible in the first pass): no bound from the heads
                if (lane == 0) s_remaining = -1;
            } else if (incl - sum < remaining && remaining <= incl) {
                int cum = incl - sum;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (cum + cnt[j] >= remaining) {
                        s_prefix = prefix | ((uint32_t)(255 - (lane * 8 + j)) << shift);
                        s_remaining = remaining - cum;
                        break;
                    }
                    cum += cnt[j];
                }
            }
        }
        __syncthreads();
        prefix = s_prefix;
        remaining = s_remaining;
    }
    if (threadIdx.x == 0) s_floor = remaining > 0 ? fmaxf(theta[q], __uint_as_float(prefix)) : theta[q];
    __syncthreads();
    const float floor = s_floor;
    // Phase B: every warp folds its share of the lists, entries below the bound never looked at
    L.reset();
    ks = PR_SENT_SCORE;
    kd = PR_SENT_DOC;
    merge_lists<E>(L, ks, kd, ps0, pd0, stride_c, warp, kWideWarps, C, K, floor, lane);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        if (i < K) {
            sh_s[warp * K + i] = L.s[e];
            sh_d[warp * K + i] = L.d[e];
        }
    }
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        L.s[e] = i < K ? run_s[(size_t)q * K + i] : PR_SENT_SCORE;
        L.d[e] = i < K ? run_d[(size_t)q * K + i] : PR_SENT_DOC;
    }
    L.kth(K, ks, kd);
    merge_lists<E>(L, ks, kd, sh_s, sh_d, K, 0, 1, kWideWarps, K, floor, lane);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        if (i < K) {
            run_s[(size_t)q * K + i] = L.s[e];
            run_d[(size_t)q * K + i] = L.d[e];
        }
    }
    if (ks > 0.f) prw::raise_theta(theta, peer_theta, n_peers, q, ks, lane);
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
        if (nvalid < K && lane == 0)
            fill_zero_tail(out_s + (size_t)q * K, out_d + (size_t)q * K, K, nvalid, doc_id_base, n_docs);
    }
}

// K-th largest of the union of n_lists descending score lists per query (pr_bm25_raise_union_bound): one warp per
// query folds them through a register list; positions stand in for document ids (the lists hold distinct documents).
template <int E>
__global__ void __launch_bounds__(128) bm25_union_bound_kernel(const float *__restrict__ gs, int n_lists, int B, int K, float *theta)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= B) return;
    WarpTopK<E> L;
    L.reset();
    float ks = PR_SENT_SCORE;
    int kd = PR_SENT_DOC;
    for (int g = 0; g < n_lists; ++g) {
        const float *ps = gs + ((size_t)g * B + q) * K;
        bool done = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i0 = e * 32;
            if (done || i0 >= K) break;
            const float es = i0 + lane < K ? ps[i0 + lane] : -1.f;
            const int n = min(32, K - i0);
            for (int i = 0; i < n; ++i) {
                const float bs = __shfl_sync(PR_FULL_MASK, es, i);
                const int bd = g * K + i0 + i;
                if (bs < 0.f || !pr_beats(bs, bd, ks, kd)) {   // empty slot / sorted: the rest of this list loses too
                    done = true;
                    break;
                }
                L.insert(bs, bd, lane);
                L.kth(K, ks, kd);
            }
        }
    }
    if (ks > 0.f) prw::raise_theta(theta, nullptr, 0, q, ks, lane);
}

// Workspace initialisation.  (The query CSR is validated by the scoring warps, item by item: bm25_lean.cuh.)
__global__ void bm25_init_kernel(float *run_s, int32_t *run_d, float *theta, int64_t n_run, int B, int32_t *counters,
                                 int n_counters, int32_t *status)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_run) {
        run_s[i] = PR_SENT_SCORE;
        run_d[i] = PR_SENT_DOC;
    }
    // (peer-shared thresholds are cleared by the host side, one call ahead: theta == nullptr here)
    if (i < B && theta) theta[i] = PR_SENT_SCORE;
    if (i < n_counters) counters[i] = 0;
    if (i == 0) *status = 0;
}

__global__ void bm25_fill_kernel(float *p, int64_t n, float v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
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
    uint32_t wmin = 0xffffffffu, wmax = 0u;   // non-negative floats order like their bit patterns
    for (int64_t p = i0; p < nnz; p += stride) {
        const int32_t d = doc_ids[p];
        const float w = weights[p];
        if (d < 0 || d >= n_docs) atomicOr(bad, 2);
        if (!(w >= 0.f) || w > 3.0e38f) atomicOr(bad, 8);
        else {
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
constexpr int kNwChoices = 3;  // warps per CTA: 4, 8, 12
constexpr int kEChoices = 3;   // E = 1, 2, 4

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
    // boundary table (pr_index_build_aux), in caller-owned memory
    const int32_t *heavy_row;
    const uint32_t *tp;
    int32_t n_rows, n_sub;
    int64_t heavy_min_df;
    // hot posting stream (bm25_hot.cuh), also in the aux buffer
    const int32_t *hot_of_row;
    const uint32_t *hot_off;
    int32_t n_hot;
    int64_t hot_min_df, hot_stream_bytes;
    // cold stream: every CSR posting as an (offset, weight) pair, in the aux buffer in front of the hot stream
    const unsigned char *cold_stream;
    uint32_t hot_base_g;
    // per-kernel launch configuration, resolved once (cudaFuncSetAttribute + occupancy query are host latency that a
    // single-query call would pay every time): occupancy by (warps-per-CTA choice, E choice), 0 = not resolved yet
    int occ[kNwChoices][kEChoices][3];
    float max_weight;  // largest weight of the index
    bool scale_ok;     // weights in [2^-30, 2^8]: large batches use the four-epoch kernel variant
    // thresholds shared with the other GPUs of a doc-sharded corpus (pr_index_set_peer_thetas): our array of
    // 2 x peer_capacity floats (one half per call parity), the device table of the peers' arrays, the call counter
    float *peer_local;
    float *const *peer_table;
    int32_t n_peers;
    int64_t peer_capacity;
    uint64_t peer_calls;
    int union_lists;   // shards behind the last union bound of the running call (1 = none): see the variant choice per launch
};

namespace {

struct Layout {
    int n_chunks, C, L;
    int G;  // sub-tiles per work item actually used: tuning.subs_per_item, halved for small batches
    std::vector<int> launch_chunk0, launch_chunks;  // launch li covers chunks [chunk0, chunk0 + chunks)
    size_t off_status, off_counters, off_theta, off_run_s, off_run_d, off_part_s, off_part_d, off_cursors, total;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNominalWarpsPerSm = 24;  // 3 CTAs x 8 warps

Layout make_layout(const pr_index *ix, int32_t B, int32_t K, int32_t n_sub_tiles = -1)
{
    const pr_bm25_tuning_t &t = ix->tuning;
    const int32_t n_sub = n_sub_tiles >= 0 ? n_sub_tiles : ix->n_sub;
    Layout l;
    l.G = t.subs_per_item;
    // A small batch (the reference calls retrieve() with ONE query) has too few (query, chunk) items to fill
    // 148 SMs x 24 warps evenly: cut the items shorter until every resident warp gets about items_per_warp of them.
    const int64_t target = (int64_t)ix->num_sms * kNominalWarpsPerSm * t.items_per_warp;
    while (l.G > 1 && (int64_t)B * ((n_sub + l.G - 1) / l.G) < target) l.G = (l.G + 1) / 2;
    l.n_chunks = (n_sub + l.G - 1) / l.G;
    // chunks per launch: docs_per_launch bounds the per-item list storage of a big batch ([B][C][K] pairs); a small
    // batch gets launches of at least min_items work items -- usually ONE launch: thresholds travel between the
    // warps of a launch (bm25_lean.cuh), so launches are not needed to make them known.
    const int64_t chunk_docs = (int64_t)l.G * prw::kSub;
    int64_t c = (t.docs_per_launch + chunk_docs - 1) / chunk_docs;
    const int64_t c_items = B > 0 ? ((int64_t)t.min_items + B - 1) / B : 1;
    if (c_items > c) c = c_items;
    if (c < 1) c = 1;
    if (c > l.n_chunks) c = l.n_chunks > 0 ? l.n_chunks : 1;
    l.C = (int)c;
    // Launch plan.  Thresholds travel between the warps of a launch, but only as the k-th scores of single items; the
    // merge after a launch publishes the k-th score of everything scored so far, which is much stronger while few
    // documents have been seen (a sub-tile is scanned when ANY of its 2048 documents reaches the bound).  So a large
    // batch starts with a one-chunk launch and doubles -- c0, c0, 2 c0, 4 c0, ... up to C chunks: on a short shard
    // (8 GPUs: 2.6M documents each) the first full-size launch would be 15% of the range.  Between launches is also
    // where doc-range shards exchange their bounds (pr_bm25_topk_range).
    {
        int64_t c0 = B > 0 ? (4096 + (int64_t)B - 1) / B : l.C;
        if (c0 < 1) c0 = 1;
        if (c0 > l.C || l.C >= l.n_chunks) c0 = l.C;  // (a batch that fits ONE launch keeps it: no merges, no tails)
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
    l.off_run_s = o;    o = align_up(o + (size_t)B * K * 4, 256);
    l.off_run_d = o;    o = align_up(o + (size_t)B * K * 4, 256);
    l.off_part_s = o;   o = align_up(o + (size_t)B * l.C * K * 4, 256);
    l.off_part_d = o;   o = align_up(o + (size_t)B * l.C * K * 4, 256);
    // posting cursors of long queries, kCursorCap per resident warp (at most 32 warps per SM)
    l.off_cursors = o;  o = align_up(o + (size_t)ix->num_sms * 32 * prl::kCursorCap * prl::kCursorWords * 4, 256);
    l.total = o;
    return l;
}

using prk::pick_lean_fn;
using prk::score_fn_t;

inline int e_of(int k) { return k <= 32 ? 1 : (k <= 64 ? 2 : 4); }
inline int e_idx(int E) { return E == 1 ? 0 : E == 2 ? 1 : 2; }
inline int nw_idx(int nw) { return nw == 4 ? 0 : nw == 8 ? 1 : 2; }

const int kMergeWideMinLists = 128;  // lists per query and launch from which the 32-warp merge pays

template <int E>
int launch_merge_e(bool wide, dim3 grid, cudaStream_t st, const float *ps, const int32_t *pd, int C,
                   int64_t sq, int64_t sc, float *rs, int32_t *rd, float *th, float *const *pt, int np, int B, int K, int fin,
                   float *os, int32_t *od, int base, int n_docs)
{
    if (wide) {
        const size_t smem = (size_t)kWideWarps * K * 8;  // <= 32 KB
        bm25_merge_wide_kernel<E><<<B, kWideWarps * 32, smem, st>>>(ps, pd, C, sq, sc, rs, rd, th, pt, np, K, fin, os, od, base, n_docs);
    } else {
        bm25_merge_kernel<E><<<grid, 128, 0, st>>>(ps, pd, C, sq, sc, rs, rd, th, pt, np, B, K, fin, os, od, base, n_docs);
    }
    PR_CUDA_CHECK(cudaGetLastError());
    return PR_OK;
}

int launch_merge(int E, cudaStream_t st, const float *ps, const int32_t *pd, int C, int64_t sq, int64_t sc,
                 float *rs, int32_t *rd, float *th, float *const *pt, int np, int B, int K, int fin, float *os, int32_t *od,
                 int base, int n_docs)
{
    const bool wide = rs && C >= kMergeWideMinLists;
    const dim3 grid((unsigned)((B + 3) / 4));
    if (E == 1) return launch_merge_e<1>(wide, grid, st, ps, pd, C, sq, sc, rs, rd, th, pt, np, B, K, fin, os, od, base, n_docs);
    if (E == 2) return launch_merge_e<2>(wide, grid, st, ps, pd, C, sq, sc, rs, rd, th, pt, np, B, K, fin, os, od, base, n_docs);
    return launch_merge_e<4>(wide, grid, st, ps, pd, C, sq, sc, rs, rd, th, pt, np, B, K, fin, os, od, base, n_docs);
}

void default_tuning(pr_bm25_tuning_t *t)
{
    t->subs_per_item = 24;
    t->warps_per_cta = 8;
    t->docs_per_launch = 393216;
    t->min_items = 32768;
    t->items_per_warp = 1;
    t->tile_epochs = 4;
    t->batch_variant = 3;
}

int check_tuning(const pr_bm25_tuning_t &t)
{
    if (t.subs_per_item < 1 || t.docs_per_launch < 1 || t.min_items < 1 || t.items_per_warp < 1 ||
        (t.tile_epochs != 2 && t.tile_epochs != 4) || t.batch_variant < 1 || t.batch_variant > 3 || (t.warps_per_cta != 4 && t.warps_per_cta != 8 && t.warps_per_cta != 12)) {
        pr_set_error("bad tuning (subs_per_item=%d docs_per_launch=%d min_items=%d items_per_warp=%d warps_per_cta=%d; "
                     "warps_per_cta is 4, 8 or 12)",
                     t.subs_per_item, t.docs_per_launch, t.min_items, t.items_per_warp, t.warps_per_cta);
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
    if ((uint64_t)nnz + 64 >= ((uint64_t)1 << 32)) {
        pr_set_error("pr_index_create: %lld postings on one device; the posting streams are addressed with 32-bit granule "
                     "indices (< 4.29e9 postings): cut the corpus into doc-range shards", (long long)nnz);
        return PR_EUNSUPPORTED;
    }
    PR_CUDA_CHECK(cudaSetDevice(device));
    int32_t *bad = nullptr;
    int32_t h_bad3[3] = {0, -1, 0};  // flags, smallest weight (bits, starts at 0xffffffff), largest weight (bits)
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
    memset(ix->occ, 0, sizeof(ix->occ));
    ix->peer_local = nullptr;
    ix->peer_table = nullptr;
    ix->n_peers = 0;
    ix->peer_capacity = 0;
    ix->peer_calls = 0;
    ix->union_lists = 1;
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
    ix->n_hot = 0;
    ix->hot_min_df = 0;
    ix->hot_stream_bytes = 0;
    ix->cold_stream = nullptr;
    ix->hot_base_g = 0;
    {   // four tile epochs (bm25_lean.cuh, kernel variant 2) need every weight in [2^-30, 2^8]
        float wmin = 0.f, wmax = 0.f;
        if (nnz > 0) {
            memcpy(&wmin, &h_bad3[1], 4);
            memcpy(&wmax, &h_bad3[2], 4);
        }
        ix->max_weight = wmax;
        ix->scale_ok = nnz > 0 && wmin >= 9.313225746154785e-10f && wmax <= 256.f;
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
    // heavy_row[n_terms] + block counts, the cold stream (8 bytes per posting), then row_term/tp and the hot stream
    // within the budget
    const size_t fixed = align_up((size_t)index->n_terms * 4, 256) + align_up(((size_t)index->n_terms / 1024 + 2) * 4, 256) + 256;
    return fixed + align_up((size_t)index->nnz * 8 + 256, 256) + align_up(table_budget_bytes, 256);
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
    int32_t *block_cnt = (int32_t *)(p + o);  o = align_up(o + ((size_t)nt / 1024 + 2) * 4, 256);
    int32_t *count = (int32_t *)(p + o);      o += 256;
    unsigned char *cold = p + o;              o += align_up((size_t)index->nnz * 8 + 256, 256);
    if (o > aux_bytes) {
        pr_set_error("pr_index_build_aux: aux buffer of %zu bytes, need at least %zu (tables + 8 bytes per posting)", aux_bytes, o);
        return PR_EWORKSPACE;
    }
    // the boundary table gets a fifth of what is left (at least 16 MB of it), the hot stream the rest
    const size_t budget = aux_bytes - o;
    size_t table_bytes = budget / 5;
    if (table_bytes < ((size_t)16 << 20)) table_bytes = budget < ((size_t)16 << 20) ? budget : ((size_t)16 << 20);
    const size_t per_row = ((size_t)index->n_sub + 1) * 4 + 4;  // tp row + row_term entry
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
    uint32_t *tp = (uint32_t *)(p + o);       o = align_up(o + (size_t)rows * ((size_t)index->n_sub + 1) * 4, 256);
    if (nt > 0) {
        prw::heavy_block_count_kernel<<<n_blocks, 1024, 0, st>>>(index->indptr, nt, min_df, block_cnt);
        prw::heavy_block_scan_kernel<<<1, 32, 0, st>>>(block_cnt, n_blocks);
        prw::heavy_assign_kernel<<<n_blocks, 1024, 0, st>>>(index->indptr, nt, min_df, block_cnt, heavy_row, row_term);
        if (rows > 0)
            prw::tp_fill_kernel<<<2368, 256, 0, st>>>(index->indptr, index->doc_ids, row_term, rows, index->n_sub, tp);
    }
    if (index->nnz > 0)
        prl::cold_fill_kernel<<<2368, 256, 0, st>>>(index->doc_ids, index->weights, index->nnz, (uint2 *)cold);
    PR_CUDA_CHECK(cudaGetLastError());
    PR_CUDA_CHECK(cudaStreamSynchronize(st));
    index->heavy_row = heavy_row;
    index->tp = tp;
    index->n_rows = rows;
    index->heavy_min_df = min_df;
    index->hot_of_row = nullptr;
    index->hot_off = nullptr;
    index->n_hot = 0;
    index->hot_min_df = 0;
    index->hot_stream_bytes = 0;
    index->cold_stream = cold;
    index->hot_base_g = 0;

    // ---- hot posting stream (bm25_hot.cuh) for the tabulated terms with >= kHotMinSeg postings per sub-tile, as many
    // of them as the rest of the buffer AND the 32-bit granule space hold (the threshold doubles until both fit)
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
        const uint64_t g0 = (uint64_t)(stream_dev - cold) >> 3;  // one granule space: cold stream, tables, hot stream
        if (g0 + (uint64_t)total * 32 + 64 >= ((uint64_t)1 << 32)) continue;
        prh::hot_offsets_kernel<<<(unsigned)nb, 256, 0, st>>>(tp, hot_rows, n_sub, n_seg, block_sum, hot_off);
        prh::hot_fill_kernel<<<2368, 256, 0, st>>>(index->indptr, index->doc_ids, index->weights, row_term, tp, hot_rows, n_sub,
                                                  n_seg, hot_off, stream_dev);
        PR_CUDA_CHECK(cudaGetLastError());
        PR_CUDA_CHECK(cudaStreamSynchronize(st));
        index->hot_of_row = hot_of_row;
        index->hot_off = hot_off;
        index->n_hot = H;
        index->hot_min_df = hot_df;
        index->hot_stream_bytes = (int64_t)total * prh::kUnitBytes;
        index->hot_base_g = (uint32_t)g0;
        return PR_OK;
    }
}

extern "C" int pr_index_aux_info(const pr_index_t *index, pr_index_aux_info_t *info)
{
    if (!index || !info) {
        pr_set_error("pr_index_aux_info: null argument");
        return PR_EINVAL;
    }
    info->table_rows = index->n_rows;
    info->table_min_df = index->heavy_min_df;
    info->hot_rows = index->n_hot;
    info->hot_min_df = index->hot_min_df;
    info->hot_stream_bytes = index->hot_stream_bytes;
    info->cold_stream_bytes = index->cold_stream ? index->nnz * 8 : 0;
    info->n_sub_tiles = index->n_sub;
    return PR_OK;
}

extern "C" int pr_index_set_peer_thetas(pr_index_t *index, float *local_dev, int64_t capacity, int32_t n_peers,
                                        float *const *peer_bases_host, void *peer_table_dev)
{
    if (!index) {
        pr_set_error("pr_index_set_peer_thetas: null index");
        return PR_EINVAL;
    }
    if (!local_dev) {  // back to thresholds private to this GPU
        index->peer_local = nullptr;
        index->peer_table = nullptr;
        index->n_peers = 0;
        index->peer_capacity = 0;
        return PR_OK;
    }
    if (capacity < 1 || n_peers < 0 || n_peers > PR_MAX_PEERS || (n_peers > 0 && (!peer_bases_host || !peer_table_dev))) {
        pr_set_error("pr_index_set_peer_thetas: bad argument (capacity=%lld n_peers=%d, at most %d peers)", (long long)capacity,
                     n_peers, PR_MAX_PEERS);
        return PR_EINVAL;
    }
    float *tab[2 * PR_MAX_PEERS];
    for (int par = 0; par < 2; ++par)
        for (int p = 0; p < PR_MAX_PEERS; ++p)
            tab[par * PR_MAX_PEERS + p] = p < n_peers ? peer_bases_host[p] + (size_t)par * capacity : nullptr;
    if (n_peers > 0) PR_CUDA_CHECK(cudaMemcpy(peer_table_dev, tab, sizeof(tab), cudaMemcpyHostToDevice));
    bm25_fill_kernel<<<64, 256>>>(local_dev, 2 * capacity, PR_SENT_SCORE);
    PR_CUDA_CHECK(cudaGetLastError());
    PR_CUDA_CHECK(cudaDeviceSynchronize());
    index->peer_local = local_dev;
    index->peer_table = (float *const *)peer_table_dev;
    index->n_peers = n_peers;
    index->peer_capacity = capacity;
    index->peer_calls = 0;
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
    if (tuning->subs_per_item) t.subs_per_item = tuning->subs_per_item;
    if (tuning->warps_per_cta) t.warps_per_cta = tuning->warps_per_cta;
    if (tuning->docs_per_launch) t.docs_per_launch = tuning->docs_per_launch;
    if (tuning->min_items) t.min_items = tuning->min_items;
    if (tuning->items_per_warp) t.items_per_warp = tuning->items_per_warp;
    if (tuning->tile_epochs) t.tile_epochs = tuning->tile_epochs;
    if (tuning->batch_variant) t.batch_variant = tuning->batch_variant;
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

extern "C" int32_t pr_bm25_num_launches(const pr_index_t *index, int32_t n_queries, int32_t k, int64_t n_docs)
{
    if (!index || n_queries < 0 || k < 1 || k > PR_MAX_K || n_docs > 0x7fffffffLL) return -1;
    const int L = make_layout(index, n_queries, k, n_docs < 0 ? -1 : (int32_t)((n_docs + prw::kSub - 1) >> prw::kSubShift)).L;
    return L > 0 ? L : 1;  // a shard without documents still has its (empty) finalisation
}

extern "C" size_t pr_bm25_theta_offset(const pr_index_t *index, int32_t n_queries, int32_t k)
{
    if (!index || n_queries < 0 || k < 1 || k > PR_MAX_K) return 0;
    return make_layout(index, n_queries, k).off_theta;
}

extern "C" size_t pr_bm25_running_scores_offset(const pr_index_t *index, int32_t n_queries, int32_t k)
{
    if (!index || n_queries < 0 || k < 1 || k > PR_MAX_K) return 0;
    return make_layout(index, n_queries, k).off_run_s;
}

extern "C" int pr_bm25_raise_union_bound(pr_index_t *index, int32_t n_queries, int32_t k, const float *gathered_scores_dev,
                                         int32_t n_lists, void *workspace_dev, size_t workspace_bytes, pr_stream_t stream)
{
    if (!index || n_queries < 0 || k < 1 || k > PR_MAX_K || n_lists < 1 || !gathered_scores_dev || !workspace_dev) {
        pr_set_error("pr_bm25_raise_union_bound: bad argument");
        return PR_EINVAL;
    }
    const Layout l = make_layout(index, n_queries, k);
    if (workspace_bytes < l.total) {
        pr_set_error("pr_bm25_raise_union_bound: workspace of %zu bytes, need %zu", workspace_bytes, l.total);
        return PR_EWORKSPACE;
    }
    if (n_queries == 0) return PR_OK;
    float *theta = (float *)((unsigned char *)workspace_dev + l.off_theta);
    if (index->peer_local) {   // the half of the peer-shared array the running call uses
        if (n_queries > index->peer_capacity) {
            pr_set_error("pr_bm25_raise_union_bound: %d queries, the peer threshold arrays hold %lld", n_queries, (long long)index->peer_capacity);
            return PR_EINVAL;
        }
        theta = index->peer_local + (size_t)(index->peer_calls & 1) * index->peer_capacity;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n_queries + 3) / 4);
    switch (e_of(k)) {
    case 1: bm25_union_bound_kernel<1><<<grid, 128, 0, st>>>(gathered_scores_dev, n_lists, n_queries, k, theta); break;
    case 2: bm25_union_bound_kernel<2><<<grid, 128, 0, st>>>(gathered_scores_dev, n_lists, n_queries, k, theta); break;
    default: bm25_union_bound_kernel<4><<<grid, 128, 0, st>>>(gathered_scores_dev, n_lists, n_queries, k, theta); break;
    }
    PR_CUDA_CHECK(cudaGetLastError());
    index->last_launches += 1;
    index->union_lists = n_lists;
    return PR_OK;
}

extern "C" int64_t pr_bm25_last_launches(const pr_index_t *index) { return index ? index->last_launches : 0; }

extern "C" int pr_bm25_topk_range(pr_index_t *index, int32_t n_queries, const int64_t *q_indptr_dev,
                                  const int32_t *q_terms_dev, int64_t n_q_terms, int32_t k, float *out_scores_dev,
                                  int32_t *out_doc_ids_dev, void *workspace_dev, size_t workspace_bytes,
                                  int32_t launch_begin, int32_t launch_end, pr_stream_t stream)
{
    if (!index || n_queries < 0 || !q_indptr_dev || !out_scores_dev || !out_doc_ids_dev || !workspace_dev || n_q_terms < 0 ||
        (n_q_terms > 0 && !q_terms_dev)) {
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
    if (!index->heavy_row || !index->cold_stream) {
        pr_set_error("pr_bm25_topk: the index has no posting streams: call pr_index_build_aux first");
        return PR_EINVAL;
    }
    const Layout l = make_layout(index, n_queries, k);
    const int L = l.L > 0 ? l.L : 1;
    if (launch_begin < 0 || launch_end > L || launch_begin >= launch_end) {
        pr_set_error("pr_bm25_topk_range: launches [%d, %d) outside [0, %d)", launch_begin, launch_end, L);
        return PR_EINVAL;
    }
    if (workspace_bytes < l.total) {
        pr_set_error("pr_bm25_topk: workspace of %zu bytes, need %zu", workspace_bytes, l.total);
        return PR_EWORKSPACE;
    }
    if (launch_begin == 0) {
        index->last_launches = 0;
        index->ev_used = 0;
        index->union_lists = 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *ws = (unsigned char *)workspace_dev;
    int32_t *status = (int32_t *)(ws + l.off_status);
    if (n_queries == 0) {
        if (launch_begin == 0) PR_CUDA_CHECK(cudaMemsetAsync(status, 0, 4, st));
        return PR_OK;
    }
    const pr_bm25_tuning_t &t = index->tuning;
    const int E = e_of(k);
    int32_t *counters = (int32_t *)(ws + l.off_counters);
    float *theta = (float *)(ws + l.off_theta);
    float *theta_next = nullptr;
    float *const *peer_tab = nullptr;
    if (index->peer_local) {
        // thresholds shared with the peer GPUs live outside the workspace, one half of the array per call parity
        if (n_queries > index->peer_capacity) {
            pr_set_error("pr_bm25_topk: %d queries, the peer threshold arrays hold %lld", n_queries, (long long)index->peer_capacity);
            return PR_EINVAL;
        }
        if (launch_begin == 0) index->peer_calls++;
        const int parity = (int)(index->peer_calls & 1);
        theta = index->peer_local + (size_t)parity * index->peer_capacity;
        theta_next = index->peer_local + (size_t)(parity ^ 1) * index->peer_capacity;
        peer_tab = index->peer_table + parity * PR_MAX_PEERS;
    }
    const int n_peers = index->peer_local ? index->n_peers : 0;
    float *run_s = (float *)(ws + l.off_run_s);
    int32_t *run_d = (int32_t *)(ws + l.off_run_d);
    float *part_s = (float *)(ws + l.off_part_s);
    int32_t *part_d = (int32_t *)(ws + l.off_part_d);

    if (launch_begin == 0) {
        const int64_t n_run = (int64_t)n_queries * k;
        int64_t n_init = n_run > l.L + 1 ? n_run : l.L + 1;
        if (n_init < n_queries) n_init = n_queries;
        if (theta_next) bm25_fill_kernel<<<64, 256, 0, st>>>(theta_next, index->peer_capacity, PR_SENT_SCORE);
        bm25_init_kernel<<<(unsigned)((n_init + 255) / 256), 256, 0, st>>>(run_s, run_d, theta_next ? nullptr : theta, n_run, n_queries,
                                                                           counters, l.L + 1, status);
        PR_CUDA_CHECK(cudaGetLastError());
        index->last_launches += 1;
    }

    if (l.L == 0)  // shard without documents: only the (empty) finalisation
        return launch_merge(E, st, part_s, part_d, 0, 0, 0, run_s, run_d, theta, peer_tab, n_peers, n_queries, k, 1,
                            out_scores_dev, out_doc_ids_dev, index->doc_id_base, index->n_docs);

    const int nw = t.warps_per_cta;
    const int threads = nw * 32;
    const size_t smem = prl::lean_smem_bytes(nw);
    // items of one query run side by side when the batch is smaller than the resident warps: those warps re-read the
    // query's bound in front of tile scans (REFRESH variant of the kernel)
    const bool refresh = t.batch_variant == 1 || (t.batch_variant != 2 && (int64_t)n_queries < 2 * (int64_t)index->num_sms * kNominalWarpsPerSm);
    // Large batches: four tile epochs pay where tile scans are rare.  While bounds are still weak (the ramp launches
    // and the first full-size one) most sub-tiles are scanned, the epoch never advances and the larger kernel only
    // costs (measured on a 2.6M-document shard: +3%), so those launches run the two-epoch variant.
    const int var_late = refresh ? 1 : (index->scale_ok && t.tile_epochs != 2 ? 2 : 0);
    const int var_early = refresh ? 1 : 0;
    int first_full = 0;
    while (first_full < l.L && l.launch_chunks[first_full] < l.C) ++first_full;
    score_fn_t fns[2] = {pick_lean_fn(nw, E, var_early), pick_lean_fn(nw, E, var_late)};
    int occs[2] = {0, 0};
    for (int i = 0; i < 2; ++i) {
        int &occ = index->occ[nw_idx(nw)][e_idx(E)][i == 0 ? var_early : var_late];
        if (occ == 0) {
            PR_CUDA_CHECK(cudaFuncSetAttribute((const void *)fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PR_CUDA_CHECK(cudaFuncSetAttribute((const void *)fns[i], cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fns[i], threads, smem));
            if (occ < 1) {
                occ = 0;
                pr_set_error("pr_bm25_topk: kernel does not fit (threads=%d smem=%zu)", threads, smem);
                return PR_EINVAL;
            }
        }
        occs[i] = occ;
    }

    prw::ScoreArgs w;
    w.indptr = index->indptr;
    w.doc_ids = index->doc_ids;
    w.weights = index->weights;
    w.heavy_row = index->heavy_row;
    w.tp = index->tp;
    w.hot_of_row = index->hot_of_row;
    w.hot_off = index->hot_off;
    w.stream_base = index->cold_stream;
    w.hot_base_g = index->hot_base_g;
    w.cursors = (uint32_t *)(ws + l.off_cursors);
    w.q_indptr = q_indptr_dev;
    w.q_terms = q_terms_dev;
    w.theta = theta;
    w.peer_theta = peer_tab;
    w.n_peers = n_peers;
    w.part_s = part_s;
    w.part_d = part_d;
    w.status = status;
    w.n_q_terms = n_q_terms;
    w.n_docs = index->n_docs;
    w.n_terms = index->n_terms;
    w.doc_id_base = index->doc_id_base;
    w.n_queries = n_queries;
    w.K = k;
    w.n_sub = index->n_sub;
    w.subs_per_item = l.G;
    w.max_weight = index->max_weight;

    for (int li = launch_begin; li < launch_end; ++li) {
        const int chunk0 = l.launch_chunk0[li], Cl = l.launch_chunks[li];
        const int64_t items = (int64_t)n_queries * Cl;
        // ... or, on a doc-range shard, once the union bound (pr_bm25_raise_union_bound) stands for enough documents
        // over all shards that scans are rare (a sub-tile is scanned with probability ~2048 k / documents seen)
        const int64_t seen = (int64_t)chunk0 * l.G * prw::kSub * index->union_lists;
        const int late = l.L > 1 && (li > first_full || (index->union_lists > 1 && seen >= (int64_t)81920 * k)) ? 1 : 0;
        const score_fn_t fn = fns[late];
        int64_t grid = (int64_t)occs[late] * index->num_sms;
        const int64_t need = (items + nw - 1) / nw;
        if (grid > need) grid = need;
        if (index->profiling) {
            while ((int)index->ev.size() < index->ev_used + 2) {
                cudaEvent_t ev;
                PR_CUDA_CHECK(cudaEventCreate(&ev));
                index->ev.push_back(ev);
            }
            PR_CUDA_CHECK(cudaEventRecord(index->ev[index->ev_used], st));
        }
        w.chunk0 = chunk0;
        w.n_chunks_launch = Cl;
        w.counter = counters + li;
        fn<<<(unsigned)grid, threads, smem, st>>>(w);
        PR_CUDA_CHECK(cudaGetLastError());
        if (index->profiling) {
            PR_CUDA_CHECK(cudaEventRecord(index->ev[index->ev_used + 1], st));
            index->ev_used += 2;
        }
        const int rc = launch_merge(E, st, part_s, part_d, Cl, (int64_t)Cl * k, k, run_s, run_d, theta, peer_tab, n_peers,
                                    n_queries, k, li == l.L - 1, out_scores_dev, out_doc_ids_dev, index->doc_id_base,
                                    index->n_docs);
        if (rc != PR_OK) return rc;
        index->last_launches += 2;
    }
    return PR_OK;
}

extern "C" int pr_bm25_topk(pr_index_t *index, int32_t n_queries, const int64_t *q_indptr_dev,
                            const int32_t *q_terms_dev, int64_t n_q_terms, int32_t k, float *out_scores_dev,
                            int32_t *out_doc_ids_dev, void *workspace_dev, size_t workspace_bytes, pr_stream_t stream)
{
    const int32_t L = pr_bm25_num_launches(index, n_queries, k, -1);
    return pr_bm25_topk_range(index, n_queries, q_indptr_dev, q_terms_dev, n_q_terms, k, out_scores_dev, out_doc_ids_dev,
                              workspace_dev, workspace_bytes, 0, L > 0 ? L : 1, stream);
}

extern "C" int pr_bm25_status(const void *workspace_dev, pr_stream_t stream)
{
    if (!workspace_dev) {
        pr_set_error("pr_bm25_status: null workspace");
        return PR_EINVAL;
    }
    int32_t h = 0;
    PR_CUDA_CHECK(cudaMemcpyAsync(&h, workspace_dev, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PR_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    if (h & 2) {
        pr_set_error("query CSR inconsistent: q_indptr must start at 0, be non-decreasing and end within q_terms");
        return PR_EINVAL;
    }
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
    // lists are [n_lists, n_queries, k]: query stride k, list stride n_queries*k
    return launch_merge(e_of(k), (cudaStream_t)stream, scores_dev, ids_dev, n_lists, k, (int64_t)n_queries * k, nullptr,
                        nullptr, nullptr, nullptr, 0, n_queries, k, 1, out_scores_dev, out_ids_dev, 0, -1);
}
