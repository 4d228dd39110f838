// nn_kernels_k.cu -- instantiates the search kernels for ONE dimension count NN_K (3..16).
// The build compiles this file once per k (-DNN_K=<k>) so the fourteen variants build in parallel
// and every inner loop over the dimensions is fully unrolled.
#include "nn_kernels.cuh"

#include <type_traits>

#ifndef NN_K
#error "compile with -DNN_K=<3..16>"
#endif

namespace nnb200
{

// Queries per thread of the wide query-register tile: as many as fit beside one reference group
// in a 128-register budget.
template <int K>
struct QregDefault
{
    static constexpr int BUDGET = (96 - Geo<K>::G * K) / K;
    static constexpr int Q = BUDGET >= 8 ? 8 : (BUDGET >= 4 ? 4 : (BUDGET >= 2 ? 2 : 1));
};

template <int K, int Q, int NT, bool PACKED>
static cudaError_t launch_qreg_one(const QregArgs &a, uint32_t qtiles, cudaStream_t st)
{
    auto kern = nn_qreg_kernel<K, Q, NT, PACKED>;
    static bool configured = false;
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QregCfg<K>::SMEM);
        if (e != cudaSuccess)
            return e;
        configured = true;
    }
    kern<<<qtiles * a.splits, NT, QregCfg<K>::SMEM, st>>>(a);
    return cudaGetLastError();
}

template <int K, int Q, int NT, bool PACKED>
static cudaError_t query_qreg_one(LaunchInfo *info, int *tile_queries, int *tile_refs)
{
    auto kern = nn_qreg_kernel<K, Q, NT, PACKED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QregCfg<K>::SMEM);
    if (e != cudaSuccess)
        return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess)
        return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, QregCfg<K>::SMEM);
    if (e != cudaSuccess)
        return e;
    info->regs = fa.numRegs;
    info->smem = (int)QregCfg<K>::SMEM;
    info->occ = occ;
    *tile_queries = NT * Q;
    *tile_refs = QregCfg<K>::TR;
    return cudaSuccess;
}

// q_sel: 0 = wide default tile, otherwise the requested queries/thread (1, 2, 4, 8; only tiles up
// to the default width are built for a given k -- wider ones would spill).
// nt_sel: bit 0 clear = packed f32x2 math, set = scalar math (kept for A/B measurements).
template <class F>
static cudaError_t qreg_dispatch(int q_sel, int nt_sel, F f)
{
    constexpr int QD = QregDefault<NN_K>::Q;
    const bool scalar = (nt_sel & 1) != 0;
    const int qq = q_sel == 0 ? QD : q_sel;
    auto go = [&](auto qc) -> cudaError_t {
        constexpr int QV = decltype(qc)::value;
        if constexpr (QV <= QD)
            return scalar ? f(qc, std::false_type{}) : f(qc, std::true_type{});
        else
            return cudaErrorInvalidValue;
    };
    switch (qq)
    {
    case 1:
        return go(std::integral_constant<int, 1>{});
    case 2:
        return go(std::integral_constant<int, 2>{});
    case 4:
        return go(std::integral_constant<int, 4>{});
    case 8:
        return go(std::integral_constant<int, 8>{});
    default:
        return cudaErrorInvalidValue;
    }
}

template <>
cudaError_t launch_qreg<NN_K>(int q_sel, int nt_sel, const QregArgs &a, uint32_t qtiles, cudaStream_t st)
{
    return qreg_dispatch(q_sel, nt_sel, [&](auto qc, auto pk) {
        return launch_qreg_one<NN_K, decltype(qc)::value, 128, decltype(pk)::value>(a, qtiles, st);
    });
}

template <>
cudaError_t query_qreg<NN_K>(int q_sel, int nt_sel, LaunchInfo *info, int *tile_queries, int *tile_refs)
{
    return qreg_dispatch(q_sel, nt_sel, [&](auto qc, auto pk) {
        return query_qreg_one<NN_K, decltype(qc)::value, 128, decltype(pk)::value>(info, tile_queries, tile_refs);
    });
}

// ---- reference-register kernel ------------------------------------------------------------------
template <int K>
struct RregCfg
{
    static constexpr int NT = 128;
    // groups per thread per batch: ~32 reference floats in flight per buffer
    static constexpr int PG = (32 / (Geo<K>::G * K)) >= 1 ? (32 / (Geo<K>::G * K)) : 1;
    static constexpr int PG_SOA = (32 / K) >= 2 ? (32 / K) : 2;
    // CTAs per SM the register budget is compiled for: two reference buffers + per-query state
    static constexpr int minb(int mq, bool soa)
    {
        const int refs = 2 * (soa ? PG_SOA : PG * Geo<K>::G) * K;
        return (refs + 3 * mq + K + 28 <= 128) ? 4 : ((refs + 3 * mq + K + 28 <= 168) ? 3 : 2);
    }
};

template <int K, int MQ, bool SOA>
static cudaError_t launch_rreg_one(const RregArgs &a, dim3 grid, cudaStream_t st)
{
    constexpr int PG = SOA ? RregCfg<K>::PG_SOA : RregCfg<K>::PG;
    if (a.mq_total != (int)grid.y * MQ)
        return cudaErrorInvalidValue;
    nn_rreg_kernel<K, MQ, PG, RregCfg<K>::NT, SOA, RregCfg<K>::minb(MQ, SOA)><<<grid, RregCfg<K>::NT, 0, st>>>(a);
    return cudaGetLastError();
}

template <int K, int MQ, bool SOA>
static cudaError_t query_rreg_one(LaunchInfo *info, int *refs_per_batch)
{
    constexpr int PG = SOA ? RregCfg<K>::PG_SOA : RregCfg<K>::PG;
    auto kern = nn_rreg_kernel<K, MQ, PG, RregCfg<K>::NT, SOA, RregCfg<K>::minb(MQ, SOA)>;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess)
        return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RregCfg<K>::NT, 0);
    if (e != cudaSuccess)
        return e;
    info->regs = fa.numRegs;
    info->smem = (int)fa.sharedSizeBytes;
    info->occ = occ;
    *refs_per_batch = RregCfg<K>::NT * PG * (SOA ? 1 : Geo<K>::G);
    return cudaSuccess;
}

#define NN_RREG_DISPATCH(FN, ...)                                                                                      \
    switch (mq)                                                                                                        \
    {                                                                                                                  \
    case 1:                                                                                                            \
        return soa ? FN<NN_K, 1, true>(__VA_ARGS__) : FN<NN_K, 1, false>(__VA_ARGS__);                                 \
    case 2:                                                                                                            \
        return soa ? FN<NN_K, 2, true>(__VA_ARGS__) : FN<NN_K, 2, false>(__VA_ARGS__);                                 \
    case 4:                                                                                                            \
        return soa ? FN<NN_K, 4, true>(__VA_ARGS__) : FN<NN_K, 4, false>(__VA_ARGS__);                                 \
    case 8:                                                                                                            \
        return soa ? FN<NN_K, 8, true>(__VA_ARGS__) : FN<NN_K, 8, false>(__VA_ARGS__);                                 \
    default:                                                                                                           \
        return cudaErrorInvalidValue;                                                                                  \
    }

template <>
cudaError_t launch_rreg<NN_K>(int mq, bool soa, const RregArgs &a, dim3 grid, cudaStream_t st)
{
    NN_RREG_DISPATCH(launch_rreg_one, a, grid, st)
}

template <>
cudaError_t query_rreg<NN_K>(int mq, bool soa, LaunchInfo *info, int *refs_per_batch)
{
    NN_RREG_DISPATCH(query_rreg_one, info, refs_per_batch)
}

template <>
cudaError_t launch_plain<NN_K>(const float *S, const float *R, int m, uint32_t n, uint32_t index_base, uint32_t splits,
                               unsigned long long *keys, cudaStream_t st)
{
    if (splits < 1)
        splits = 1;
    const uint32_t per = (n + splits - 1) / splits;
    dim3 grid((m + 127) / 128, splits);
    nn_plain_kernel<NN_K><<<grid, 128, 0, st>>>(S, R, m, n, index_base, per, keys);
    return cudaGetLastError();
}

template <>
cudaError_t launch_repack_soa<NN_K>(const float *in, float *out, uint32_t n, int num_sms, cudaStream_t st)
{
    using C = RepackCfg<NN_K>;
    const uint32_t ntiles = (n + C::TN - 1) / C::TN;
    if (ntiles == 0)
        return cudaSuccess;
    const uint32_t cap = (uint32_t)num_sms * 4u;
    nn_repack_soa_kernel<NN_K><<<ntiles < cap ? ntiles : cap, C::NT, 0, st>>>(in, out, n);
    return cudaGetLastError();
}

} // namespace nnb200
