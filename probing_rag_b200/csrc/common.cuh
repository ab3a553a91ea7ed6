// Shared device/host helpers of libprobingrag.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/probing_rag.h"

#define PR_FULL_MASK 0xffffffffu
#define PR_SENT_SCORE (-1.0f)        // internal "empty slot": every real score is > 0
#define PR_SENT_DOC 0x7fffffff
#define PR_DENORM_MIN 1.401298464e-45f

void pr_set_error(const char *fmt, ...);

#define PR_CUDA_CHECK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            pr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                         __LINE__);                                                           \
            return PR_ECUDA;                                                                  \
        }                                                                                     \
    } while (0)

// Canonical total order of ranked lists: score descending, doc id ascending.
__host__ __device__ __forceinline__ bool pr_beats(float s, int d, float s2, int d2)
{
    return s > s2 || (s == s2 && d < d2);
}

// Streaming 128-bit loads of postings: read once per (query, tile), keep them out of L1.
__device__ __forceinline__ int4 pr_ldg_stream_i4(const int32_t *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float4 pr_ldg_stream_f4(const float *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// Sorted list of the best 32*E (score, doc) pairs held by one warp in registers, best first:
// entry i lives in lane i%32, slot i/32.  All methods are warp-collective and must be called
// by all 32 lanes with warp-uniform arguments.
template <int E>
struct WarpTopK {
    float s[E];
    int d[E];

    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            s[e] = PR_SENT_SCORE;
            d[e] = PR_SENT_DOC;
        }
    }

    // (score, doc) of entry K-1, the one a newcomer has to beat.
    __device__ __forceinline__ void kth(int K, float &ks, int &kd) const
    {
        const int i = K - 1, slot = i >> 5, src = i & 31;
        float vs = s[0];
        int vd = d[0];
#pragma unroll
        for (int e = 1; e < E; ++e)
            if (slot == e) {
                vs = s[e];
                vd = d[e];
            }
        ks = __shfl_sync(PR_FULL_MASK, vs, src);
        kd = __shfl_sync(PR_FULL_MASK, vd, src);
    }

    // Insert (ns, nd) at its rank, shifting worse entries down by one (the last falls off).
    __device__ __forceinline__ void insert(float ns, int nd, int lane)
    {
        int p = 0;
#pragma unroll
        for (int e = 0; e < E; ++e)
            p += __popc(__ballot_sync(PR_FULL_MASK, pr_beats(s[e], d[e], ns, nd)));
#pragma unroll
        for (int e = E - 1; e >= 0; --e) {
            float us = __shfl_up_sync(PR_FULL_MASK, s[e], 1);
            int ud = __shfl_up_sync(PR_FULL_MASK, d[e], 1);
            if (e > 0) {
                const float ws = __shfl_sync(PR_FULL_MASK, s[e - 1], 31);
                const int wd = __shfl_sync(PR_FULL_MASK, d[e - 1], 31);
                if (lane == 0) {
                    us = ws;
                    ud = wd;
                }
            }
            const int i = e * 32 + lane;
            if (i == p) {
                s[e] = ns;
                d[e] = nd;
            } else if (i > p) {
                s[e] = us;
                d[e] = ud;
            }
        }
    }
};

// Smallest idx in [lo, hi] with a[idx] >= target (hi if none); a ascending.  Warp-collective
// 32-ary search: each round the 32 lanes probe the last element of 32 equal blocks, so a
// 6M-entry posting list resolves in 5 dependent loads instead of 23.
__device__ __forceinline__ int64_t pr_lower_bound_warp(const int32_t *__restrict__ a, int64_t lo,
                                                       int64_t hi, int32_t target, int lane)
{
    while (true) {
        const int64_t n = hi - lo;
        if (n <= 0) return lo;
        if (n <= 32) {
            const int32_t v = (lane < n) ? __ldg(a + lo + lane) : 0x7fffffff;
            return lo + __popc(__ballot_sync(PR_FULL_MASK, v < target));
        }
        const int64_t stride = (n + 31) >> 5;
        int64_t last = (int64_t)(lane + 1) * stride;
        if (last > n) last = n;
        const int32_t v = __ldg(a + lo + last - 1);
        const int c = __popc(__ballot_sync(PR_FULL_MASK, v < target));
        if (c == 32) return hi;
        const int64_t nlo = lo + (int64_t)c * stride;
        int64_t nhi = nlo + stride;
        if (nhi > hi) nhi = hi;
        lo = nlo;
        hi = nhi;
    }
}
