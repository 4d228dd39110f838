// nn_probe.cu -- measures the denominator of the FP32 roofline on the device it runs on: the
// sustained issue rate of NON-FUSED single-precision adds and multiplies (the only arithmetic the
// search kernels are allowed to use), in lane-operations per second.  bench.py reports the
// search kernels against this measured figure next to the nominal SMs x 128 lanes x clock.
#include "../../include/nn_b200.h"

#include <cuda_runtime.h>

namespace
{
constexpr int CHAINS = 8;

// MODE 0: scalar FMUL/FADD only           MODE 1: packed FMUL2/FADD2 only
// MODE 2: packed and scalar alternating   MODE 3: blocks of four packed, four scalar
// MODE 4: packed + one FMNMX per packed   MODE 5: scalar + one FMNMX per two scalar
// Every chain is independent; half of the chains only multiply, the other half only add: a
// multiply feeding an add is what ptxas contracts to FFMA2 for the packed forms (even with .rn),
// and this probe must measure the non-fused instructions the search kernels use.
template <int MODE>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *out, int iters, float a, float b)
{
    float2 x[CHAINS], y[CHAINS];
    float mn[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
    {
        x[i] = make_float2(__int2float_rn(threadIdx.x + i), __int2float_rn(blockIdx.x + 2 * i));
        y[i] = make_float2(__int2float_rn(threadIdx.x + 3 * i), __int2float_rn(blockIdx.x + 5 * i));
        mn[i] = __int2float_rn(i);
    }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    auto P = [&](int i) { x[i] = (i & 1) ? __fmul2_rn(x[i], a2) : __fadd2_rn(x[i], b2); };
    auto S = [&](int i) {
        y[i].x = (i & 1) ? __fmul_rn(y[i].x, a) : __fadd_rn(y[i].x, b);
        y[i].y = (i & 1) ? __fmul_rn(y[i].y, a) : __fadd_rn(y[i].y, b);
    };
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep)
        {
            if (MODE == 0)
            {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i)
                    S(i);
            }
            else if (MODE == 1)
            {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i)
                    P(i);
            }
            else if (MODE == 2)
            {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i)
                {
                    P(i);
                    y[i].x = (i & 1) ? __fmul_rn(y[i].x, a) : __fadd_rn(y[i].x, b);
                }
            }
            else if (MODE == 3)
            {
#pragma unroll
                for (int i = 0; i < CHAINS; i += 4)
                {
                    P(i), P(i + 1), P(i + 2), P(i + 3);
                    S(i), S(i + 1);
                }
            }
            else if (MODE == 4)
            {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i)
                {
                    P(i);
                    mn[i] = fminf(mn[i], x[(i + 3) % CHAINS].x);
                }
            }
            else
            {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i)
                {
                    S(i);
                    mn[i] = fminf(mn[i], y[(i + 3) % CHAINS].x);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
        s = __fadd_rn(s, __fadd_rn(__fadd_rn(x[i].x, x[i].y), __fadd_rn(__fadd_rn(y[i].x, y[i].y), mn[i])));
    if (s == 12345.678f)
        out[0] = s; // never true in practice; keeps the chains alive
}

// FP32 lane-operations per thread and loop iteration of each mode
constexpr int kLaneOps[6] = {32, 32, 48, 48, 32, 32};
} // namespace

extern "C" int nn_b200_probe_fp32(int mode, int iters, double *lane_ops_per_s)
{
    if (!lane_ops_per_s || iters < 1 || mode < 0 || mode > 5)
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
        switch (mode)
        {
        case 0:
            fp32_probe_kernel<0><<<ctas, 256>>>(out, iters, 0.999f, 1e-3f);
            break;
        case 1:
            fp32_probe_kernel<1><<<ctas, 256>>>(out, iters, 0.999f, 1e-3f);
            break;
        case 2:
            fp32_probe_kernel<2><<<ctas, 256>>>(out, iters, 0.999f, 1e-3f);
            break;
        case 3:
            fp32_probe_kernel<3><<<ctas, 256>>>(out, iters, 0.999f, 1e-3f);
            break;
        case 4:
            fp32_probe_kernel<4><<<ctas, 256>>>(out, iters, 0.999f, 1e-3f);
            break;
        default:
            fp32_probe_kernel<5><<<ctas, 256>>>(out, iters, 0.999f, 1e-3f);
            break;
        }
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess)
            break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)ctas * 256.0 * (double)kLaneOps[mode] * (double)iters;
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
