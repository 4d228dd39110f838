// nn_probe.cu -- measures the denominator of the FP32 roofline on the device it runs on: the
// sustained issue rate of NON-FUSED single-precision adds and multiplies (the only arithmetic the
// search kernels are allowed to use), in lane-operations per second.  bench.py reports the
// search kernels against this measured figure next to the nominal SMs x 128 lanes x clock.
#include "../../include/nn_b200.h"

#include <cuda_runtime.h>

namespace
{
constexpr int CHAINS = 8;

template <bool PACKED>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *out, int iters, float a, float b)
{
    // Half of the chains only multiply, the other half only add: a multiply feeding an add is what
    // ptxas contracts to FFMA2 for the packed forms (even with .rn), and this probe must measure
    // the non-fused instructions the search kernels use.
    float2 x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
        x[i] = make_float2(__int2float_rn(threadIdx.x + i), __int2float_rn(blockIdx.x + 2 * i));
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i)
            {
                if (PACKED)
                    x[i] = (i & 1) ? __fmul2_rn(x[i], a2) : __fadd2_rn(x[i], b2);
                else
                {
                    x[i].x = (i & 1) ? __fmul_rn(x[i].x, a) : __fadd_rn(x[i].x, b);
                    x[i].y = (i & 1) ? __fmul_rn(x[i].y, a) : __fadd_rn(x[i].y, b);
                }
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
        s = __fadd_rn(s, __fadd_rn(x[i].x, x[i].y));
    if (s == 12345.678f)
        out[0] = s; // never true in practice; keeps the chains alive
}
} // namespace

extern "C" int nn_b200_probe_fp32(int packed, int iters, double *lane_ops_per_s)
{
    if (!lane_ops_per_s || iters < 1)
        return NN_B200_EINVAL;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return NN_B200_ECUDA;
    float *out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess)
        return NN_B200_ECUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int ctas = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep)
    {
        cudaEventRecord(e0);
        if (packed)
            fp32_probe_kernel<true><<<ctas, 256>>>(out, iters, 0.999f, 1e-3f);
        else
            fp32_probe_kernel<false><<<ctas, 256>>>(out, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess)
            break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)ctas * 256.0 * CHAINS * 4.0 * (double)iters;
        if (rep > 0 && ms > 0.f)
            best = ops / (ms * 1e-3) > best ? ops / (ms * 1e-3) : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess || best == 0.0)
        return NN_B200_ECUDA;
    *lane_ops_per_s = best;
    return NN_B200_OK;
}
