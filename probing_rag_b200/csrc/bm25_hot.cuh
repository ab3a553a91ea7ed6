// Hot posting stream: a private, kernel-friendly copy of the posting lists of the frequent terms.
//
// ncu on the scoring kernels (profiles/r01) shows the scatter into the shared-memory score tile
// as the real bound: ~0.44 shared-memory wavefronts per posting (3.5-way bank conflicts on
// random document offsets) and ~2 warp instructions per posting, most of them masks, offset
// arithmetic and per-segment control for lists that hold a handful of postings per 2048-document
// sub-tile.  99% of the postings a Zipf query batch touches belong to the few thousand terms
// with >= 8 postings per sub-tile, so for those terms the index keeps, next to the CSR, every
// (term, sub-tile) segment re-laid-out as a sequence of mask-free STEPS:
//
//   wide step   768 B: [32 lanes x 4 tile byte-offsets u16][32 lanes x 4 weights f32]  (a tile is 8 KB + pads: 16 bits)
//   narrow step 256 B: 32 lanes x (tile byte-offset u32, weight f32) pairs, interleaved
//
//   * offsets are pre-scaled byte offsets into the warp's tile (no masking / shifting per slot);
//   * segments are padded to whole steps with (dummy word, +0.0f) slots -- one dummy word per 32-slot
//     group, in a bank the group's postings leave free -- so there are no
//     validity masks and no alignment heads; a remainder of more than 32 postings takes one
//     padded wide step rather than up to three narrow ones (fewer steps beat fewer slots: +10%);
//   * inside a segment the postings are placed bank-aware (bank = doc % 32): most 32-slot groups are CLEAN -- one
//     posting per bank, never a conflict -- and the few last groups absorb what the banks hold beyond that
//     (hot_fill_kernel).
//
// The order of a term's postings inside a sub-tile is irrelevant for the sums (a document occurs
// once per term), so scores stay bit-identical.  Everything is built on the device by
// pr_index_build_aux into caller-owned memory.
#pragma once

#include "bm25_tables.cuh"

namespace prh {

using prw::kSub;
using prw::kSubShift;

#ifndef PR_HOT_MIN_SEG
#define PR_HOT_MIN_SEG 8
#endif
constexpr int kHotMinSeg = PR_HOT_MIN_SEG;  // a term is hot when it averages >= this many postings per sub-tile
constexpr int kUnitBytes = 256;  // narrow step; a wide step is kWideUnits units
constexpr int kWideUnits = 3;    // 256 B of 16-bit offsets + 512 B of weights
#ifndef PR_HOT_WIDE_REM
#define PR_HOT_WIDE_REM 32
#endif
constexpr int kWideRem = PR_HOT_WIDE_REM;  // a remainder above this many postings takes one more (padded) wide step
static_assert(kWideRem <= 64, "a segment's unit count must split uniquely into wide steps (3 units) and at most 2 narrow ones");

// 256-byte units a segment of n postings occupies: full wide steps, then the rest as narrow
// steps (or one more wide step when the rest is > kWideRem)
__host__ __device__ __forceinline__ int seg_units(int n)
{
    int w = n >> 7, r = n & 127;
    if (r > kWideRem) {
        ++w;
        r = 0;
    }
    return kWideUnits * w + ((r + 31) >> 5);
}

// hot_of_row[r] = exclusive rank of row r among the rows with df >= min_df, or -1; one warp.
static __global__ void hot_assign_kernel(const int64_t *indptr, const int32_t *row_term, int n_rows, int64_t min_df,
                                  int32_t *hot_of_row, int32_t *hot_rows, int32_t *n_hot)
{
    const int lane = threadIdx.x & 31;
    int acc = 0;
    for (int r0 = 0; r0 < n_rows; r0 += 32) {
        const int r = r0 + lane;
        bool f = false;
        if (r < n_rows) {
            const int t = row_term[r];
            f = (indptr[t + 1] - indptr[t]) >= min_df;
        }
        const unsigned m = __ballot_sync(PR_FULL_MASK, f);
        const int h = acc + __popc(m & ((1u << lane) - 1u));
        if (r < n_rows) {
            hot_of_row[r] = f ? h : -1;
            if (f && hot_rows) hot_rows[h] = r;
        }
        acc += __popc(m);
    }
    if (lane == 0) *n_hot = acc;
}

__device__ __forceinline__ int hot_seg_len(const uint32_t *tp, const int32_t *hot_rows, int n_sub, int64_t s)
{
    const int h = (int)(s / n_sub), g = (int)(s % n_sub);
    const uint32_t *r = tp + (size_t)hot_rows[h] * ((size_t)n_sub + 1) + g;
    return (int)(r[1] - r[0]);
}

constexpr int kScanChunk = 2048;  // segments per block of the offset scan (256 threads x 8)

// block sums of seg_units over chunks of the flat (hot row, sub-tile) segment array
static __global__ void __launch_bounds__(256) hot_units_kernel(const uint32_t *tp, const int32_t *hot_rows, int n_sub, int64_t n_seg,
                                                        uint32_t *block_sum)
{
    __shared__ uint32_t ws[8];
    const int64_t s0 = (int64_t)blockIdx.x * kScanChunk + threadIdx.x * 8;
    uint32_t local = 0;
    for (int i = 0; i < 8; ++i)
        if (s0 + i < n_seg) local += (uint32_t)seg_units(hot_seg_len(tp, hot_rows, n_sub, s0 + i));
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(PR_FULL_MASK, local, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += ws[i];
        block_sum[blockIdx.x] = t;
    }
}

static __global__ void hot_block_scan_kernel(uint32_t *block_sum, int n_blocks, uint32_t *total)  // one thread: a few thousand blocks
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < n_blocks; ++i) {
            const uint32_t v = block_sum[i];
            block_sum[i] = acc;
            acc += v;
        }
        *total = acc;
    }
}

// hot_off[h][g] = units before segment (h, g); hot_off[h][n_sub] = units before row h+1
static __global__ void __launch_bounds__(256) hot_offsets_kernel(const uint32_t *tp, const int32_t *hot_rows, int n_sub, int64_t n_seg,
                                                          const uint32_t *block_off, uint32_t *hot_off)
{
    __shared__ uint32_t ws[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t s0 = (int64_t)blockIdx.x * kScanChunk + threadIdx.x * 8;
    uint32_t u[8], local = 0;
    for (int i = 0; i < 8; ++i) {
        u[i] = s0 + i < n_seg ? (uint32_t)seg_units(hot_seg_len(tp, hot_rows, n_sub, s0 + i)) : 0u;
        local += u[i];
    }
    uint32_t incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(PR_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    uint32_t pre = block_off[blockIdx.x] + incl - local;
    for (int i = 0; i < w; ++i) pre += ws[i];
    for (int i = 0; i < 8; ++i) {
        const int64_t s = s0 + i;
        if (s < n_seg) {
            const int64_t h = s / n_sub;
            const int g = (int)(s % n_sub);
            hot_off[h * ((int64_t)n_sub + 1) + g] = pre;
            if (g == n_sub - 1) hot_off[h * ((int64_t)n_sub + 1) + n_sub] = pre + u[i];
        }
        pre += u[i];
    }
}

// one warp per (hot row, sub-tile) segment: the bank-aware deal of the postings, then the pads
constexpr int kMaxGroups = kSub / 32 + 4;  // 32-slot groups of the longest possible segment

static __global__ void __launch_bounds__(256) hot_fill_kernel(const int64_t *indptr, const int32_t *doc_ids, const float *weights,
                                                       const int32_t *row_term, const uint32_t *tp, const int32_t *hot_rows,
                                                       int n_sub, int64_t n_seg, const uint32_t *hot_off, unsigned char *stream)
{
    __shared__ int s_fill[8][32];
    __shared__ int s_dirty0[8][32];
    __shared__ unsigned s_banks[8][kMaxGroups];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int *fill = s_fill[w];
    unsigned *banks = s_banks[w];  // per 32-slot group: banks its postings occupy
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    for (int64_t s = (int64_t)blockIdx.x * 8 + w; s < n_seg; s += n_warps) {
        const int h = (int)(s / n_sub), g = (int)(s % n_sub);
        const int row = hot_rows[h];
        const uint32_t *r = tp + (size_t)row * ((size_t)n_sub + 1) + g;
        const int sb = (int)r[0], n = (int)(r[1] - r[0]);
        if (n == 0) continue;
        const int units = seg_units(n);
        const int n_wide = units / kWideUnits, n_narrow = units % kWideUnits, groups = 4 * n_wide + n_narrow;
        uint32_t *out = reinterpret_cast<uint32_t *>(stream + (size_t)hot_off[(size_t)h * ((size_t)n_sub + 1) + g] * kUnitBytes);
        uint16_t *out16 = reinterpret_cast<uint16_t *>(out);
        // slot (group, lane) <- (tile byte offset, weight bits): wide steps hold [32 x 4 u16 offsets][32 x float4 weights],
        // narrow steps 32 (u32 offset, weight) pairs
        auto put = [&](int grp, int ln, uint32_t off, uint32_t wbits) {
            if (grp < 4 * n_wide) {
                const int base = (grp >> 2) * (kWideUnits * 64);  // words in front of this wide step
                out16[2 * base + ln * 4 + (grp & 3)] = (uint16_t)off;
                out[base + 64 + ln * 4 + (grp & 3)] = wbits;
            } else {
                const int ow = n_wide * (kWideUnits * 64) + (grp - 4 * n_wide) * 64 + 2 * ln;
                out[ow] = off;
                out[ow + 1] = wbits;
            }
        };
        // ---- bank histogram (bank = tile word index mod 32 = doc % 32)
        const int64_t p0 = indptr[row_term[row]] + sb;
        fill[lane] = 0;
        for (int i = lane; i < groups; i += 32) banks[i] = 0u;
        __syncwarp();
        for (int i = lane; i < n; i += 32) atomicAdd(&fill[doc_ids[p0 + i] & 31], 1);
        __syncwarp();
        const int c = fill[lane];  // postings in bank `lane`
        // ---- CLEAN and DIRTY groups.  A group (the 32 slots one LDS/STS of the scoring kernel touches) costs as many
        // shared-memory wavefronts as its fullest bank holds postings.  n ~ 32 x groups leaves no slack: half of the banks
        // hold more postings than there are groups, and dealing the postings round-robin put one of those extras into
        // nearly every group (measured: 1.75 wavefronts per wide-step LDS/STS).  Instead the first C groups are CLEAN --
        // group g takes the g-th posting of every bank that has one, into lane = bank, so it never conflicts -- and what
        // the banks have beyond C postings goes round-robin into the last D = groups - C DIRTY groups, which absorb all
        // the imbalance.  C = the largest count for which the leftovers fit: 32 D >= sum_b max(0, c_b - C).
        int C = groups, R = 0;
#pragma unroll 1
        for (; C >= 0; --C) {
            int r = max(0, c - C);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(PR_FULL_MASK, r, o);
            R = r;
            if (32 * (groups - C) >= r) break;
        }
        const int D = groups - C;
        const int left = max(0, c - C);  // this bank's postings in the dirty groups
        int incl = left;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(PR_FULL_MASK, incl, o);
            if (lane >= o) incl += v;
        }
        __syncwarp();
        fill[lane] = 0;                      // running count per bank
        s_dirty0[w][lane] = incl - left;     // first dirty rank of this bank
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const int d = doc_ids[p0 + i];
            const float wt = weights[p0 + i];
            const int b = d & 31;
            const int k = atomicAdd(&fill[b], 1);
            int grp, ln;
            if (k < C) {
                grp = k;
                ln = b;
            } else {
                const int j = s_dirty0[w][b] + (k - C);
                grp = C + j % D;
                ln = j / D;
                atomicOr(&banks[grp], 1u << b);
            }
            put(grp, ln, (uint32_t)(d & (kSub - 1)) * 4u, __float_as_uint(wt));
        }
        __syncwarp();
        // ---- pads: every unused slot adds +0.0f to a dummy word behind the tile.  Clean group g: lane = bank, unused
        // where the bank has no g-th posting -- its dummy word sits in that very bank.  Dirty group x holds the ranks
        // x, x + D, ... in its first lanes; the other lanes share ONE dummy word (same address: a broadcast) in a bank
        // none of the group's postings uses.  Padding never costs a wavefront.
        for (int grp = 0; grp < groups; ++grp) {
            if (grp < C) {
                if (c <= grp) put(grp, lane, (uint32_t)(kSub + lane) * 4u, 0u);
            } else {
                const int x = grp - C;
                const int cnt = R > x ? (R - x + D - 1) / D : 0;
                if (lane >= cnt) {
                    const int free_bank = __ffs(~banks[grp]) - 1;  // cnt < 32 postings: a bank is free
                    put(grp, lane, (uint32_t)(kSub + free_bank) * 4u, 0u);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace prh
