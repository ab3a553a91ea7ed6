// On-device pooling of the probed layers' hidden states: replaces the forward hooks of
// /root/reference/exp_rag.py:317-321 -- which append `activations.detach().cpu()` to a Python list on
// every forward call, 6 device->host syncs per decode step -- and the concat / H2D / sum over tokens of
// exp_rag.py:385-386.  A hook now adds its activations straight into the prober's input matrix
// X[n_rows, n_probers, d_model] on the device; nothing leaves HBM until the gate has run.
//
// HBM-bound elementwise work: one thread owns 4 consecutive features of one row, walks the tokens of this
// forward call in order (fp32 accumulate, like torch.sum over the token axis) and does one 128-bit
// read-modify-write of the accumulator.  Algorithmic bytes: n_rows * n_tokens * d_model * sizeof(act)
// read + 8 * n_rows * d_model accumulator traffic.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ float4 load4(const T *p);

template <>
__device__ __forceinline__ float4 load4<float>(const float *p)
{
    return *reinterpret_cast<const float4 *>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p)
{
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&r.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&r.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <>
__device__ __forceinline__ float4 load4<__half>(const __half *p)
{
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    const __half2 a = *reinterpret_cast<const __half2 *>(&r.x), b = *reinterpret_cast<const __half2 *>(&r.y);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// grid: (n_rows, d_model / 4 / 128 rounded up); block 128
template <typename T>
__global__ void __launch_bounds__(128) pool_accumulate_kernel(float *__restrict__ acc, int n_acc_rows, int n_probers, int slot, int d_model,
                                                              const T *__restrict__ act, int n_tokens, int64_t row_stride,
                                                              int64_t tok_stride, const int32_t *__restrict__ row_map)
{
    const int c4 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    if (c4 >= d_model) return;
    const int r = blockIdx.x;
    const int dst = row_map ? row_map[r] : r;
    if (dst < 0 || dst >= n_acc_rows) return;
    const T *p = act + (int64_t)r * row_stride + c4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < n_tokens; ++t, p += tok_stride) {
        const float4 v = load4<T>(p);
        s.x += v.x;
        s.y += v.y;
        s.z += v.z;
        s.w += v.w;
    }
    float4 *a = reinterpret_cast<float4 *>(acc + ((int64_t)dst * n_probers + slot) * d_model + c4);
    float4 o = *a;
    o.x += s.x;
    o.y += s.y;
    o.z += s.z;
    o.w += s.w;
    *a = o;
}

}  // namespace

extern "C" int pr_pool_accumulate(float *acc_dev, int32_t n_acc_rows, int32_t n_probers, int32_t slot, int32_t d_model,
                                  const void *act_dev, int32_t act_dtype, int32_t n_rows, int32_t n_tokens,
                                  int64_t row_stride, int64_t tok_stride, const int32_t *row_map_dev, pr_stream_t stream)
{
    if (!acc_dev || (!act_dev && n_rows > 0 && n_tokens > 0) || n_acc_rows < 0 || n_rows < 0 || n_tokens < 0) {
        pr_set_error("pr_pool_accumulate: bad argument");
        return PR_EINVAL;
    }
    if (n_probers < 1 || slot < 0 || slot >= n_probers || d_model < 4 || d_model % 4 != 0) {
        pr_set_error("pr_pool_accumulate: slot %d of %d probers, d_model %d (must be a multiple of 4)", slot, n_probers, d_model);
        return PR_EINVAL;
    }
    if (!row_map_dev && n_rows > n_acc_rows) {
        pr_set_error("pr_pool_accumulate: %d activation rows for %d accumulator rows", n_rows, n_acc_rows);
        return PR_EINVAL;
    }
    const size_t esz = act_dtype == 0 ? 4 : 2;
    if (act_dtype < 0 || act_dtype > 2 || ((uintptr_t)act_dev % (4 * esz)) || row_stride % 4 || tok_stride % 4 ||
        ((uintptr_t)acc_dev % 16)) {
        pr_set_error("pr_pool_accumulate: dtype %d (0 f32, 1 bf16, 2 f16), pointers and strides must allow 4-element vector access", act_dtype);
        return PR_EINVAL;
    }
    if (n_rows == 0 || n_tokens == 0) return PR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)n_rows, (unsigned)((d_model / 4 + 127) / 128));
    if (act_dtype == 0)
        pool_accumulate_kernel<float><<<grid, 128, 0, st>>>(acc_dev, n_acc_rows, n_probers, slot, d_model, (const float *)act_dev, n_tokens,
                                                            row_stride, tok_stride, row_map_dev);
    else if (act_dtype == 1)
        pool_accumulate_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(acc_dev, n_acc_rows, n_probers, slot, d_model, (const __nv_bfloat16 *)act_dev,
                                                                    n_tokens, row_stride, tok_stride, row_map_dev);
    else
        pool_accumulate_kernel<__half><<<grid, 128, 0, st>>>(acc_dev, n_acc_rows, n_probers, slot, d_model, (const __half *)act_dev, n_tokens,
                                                             row_stride, tok_stride, row_map_dev);
    PR_CUDA_CHECK(cudaGetLastError());
    return PR_OK;
}
