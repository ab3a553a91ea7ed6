// Index-side tables of the BM25 scoring kernel (bm25_lean.cuh) and the argument block of its launches.
//
// The kernel cuts the document range into 2048-document SUB-TILES (one fp32 accumulator tile of 8 KB in shared
// memory per warp).  For every term with df above a threshold the index keeps `tp[row][s]` = number of the term's
// postings with doc < s * 2048, so the posting range of a (term, sub-tile) pair is two table reads instead of a
// search; rarer terms are located once per work item by a warp-collective 32-ary search and then followed with a
// cursor.  The tables are built once per index (pr_index_build_aux) into caller-owned memory.
#pragma once

#include "common.cuh"

#ifndef PR_SUB_SHIFT
#define PR_SUB_SHIFT 11
#endif

namespace prw {

// theta[q]: lower bound of query q's final k-th score, shared by everything that scores the query -- the warps of
// this GPU and, for a doc-sharded corpus, the other GPUs of the box.  Scores are > 0 and "nothing known" is -1.0f, so
// the bit patterns order like the floats for every raise.  Raising is an atomicMax on our copy plus one system-scope
// atomicMax per peer copy, issued by lanes 1..n_peers over NVLink (fire-and-forget reductions; a bound is advisory
// and monotone, so relaxed ordering is enough).  Reads are of our own copy only.
__device__ __forceinline__ float ld_theta(const float *p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void raise_theta(float *local, float *const *peer_theta, int n_peers, int q, float v, int lane)
{
    if (lane == 0) {
        atomicMax(reinterpret_cast<int *>(local + q), __float_as_int(v));
    } else if (lane <= n_peers) {
        int *dst = reinterpret_cast<int *>(peer_theta[lane - 1] + q);
        asm volatile("red.relaxed.sys.global.max.s32 [%0], %1;" ::"l"(dst), "r"(__float_as_int(v)) : "memory");
    }
}

constexpr int kSubShift = PR_SUB_SHIFT;
constexpr int kSub = 1 << kSubShift;  // documents per warp sub-tile (8 KB of fp32)
constexpr int kLightDf = 128;         // smallest df threshold of the boundary table

struct ScoreArgs {
    const int64_t *__restrict__ indptr;
    const int32_t *__restrict__ doc_ids;
    const float *__restrict__ weights;
    const int32_t *__restrict__ heavy_row;  // [n_terms] row of `tp`, or -1
    const uint32_t *__restrict__ tp;        // [n_rows][n_sub+1] postings with doc < s*kSub
    const int32_t *__restrict__ hot_of_row; // [n_rows] row of `hot_off`, or -1 (null: no hot stream)
    const uint32_t *__restrict__ hot_off;   // [n_hot][n_sub+1] 256-byte units of the hot stream before (row, sub-tile)
    const unsigned char *__restrict__ stream_base;  // cold stream (8-byte granule p = CSR posting p), hot stream behind it
    uint32_t hot_base_g;                    // granule index of the hot stream's first byte relative to stream_base
    uint32_t *cursors;                      // per-warp scratch of kCursorCap * kCursorWords words (long queries)
    const int64_t *__restrict__ q_indptr;
    const int32_t *__restrict__ q_terms;
    float *theta;                           // [n_queries] best known lower bound of each query's final k-th score:
                                            // read AND raised (atomicMax) by the scoring warps, see bm25_lean.cuh
    float *const *peer_theta;               // device table of n_peers pointers: the same array on the other GPUs of a
    int32_t n_peers;                        // doc-sharded corpus (peer memory over NVLink), raised together with ours
    float *part_s;                          // [n_queries][n_chunks_launch][K] per-item ranked lists
    int32_t *part_d;
    int32_t *counter;                       // work-item counter of this launch
    int32_t *status;
    int64_t n_q_terms;
    int32_t n_docs, n_terms, doc_id_base, n_queries, K;
    int32_t n_sub, subs_per_item, chunk0, n_chunks_launch;
    float max_weight;                       // largest weight of the index (bounds a query's scores: n_terms * max_weight)
};

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

}  // namespace prw
