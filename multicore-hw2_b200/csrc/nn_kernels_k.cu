// nn_kernels_k.cu -- instantiates the search kernels for ONE dimension count NN_K (3..16).
// The build compiles this file once per k (-DNN_K=<k>) so the fourteen variants build in parallel
// and every inner loop over the dimensions is fully unrolled.
#include "nn_kernels.cuh"

#include <atomic>
#include <type_traits>

#ifndef NN_K
#error "compile with -DNN_K=<3..16>"
#endif

namespace nnb200
{

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only, so the
// "already opted in" flag of a kernel is one bit per device ordinal.
struct PerDeviceOnce
{
    std::atomic<uint64_t> done{0};
    template <class F>
    cudaError_t operator()(F set)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess)
            return e;
        const uint64_t bit = 1ull << (dev & 63);
        if (done.load(std::memory_order_acquire) & bit)
            return cudaSuccess;
        e = set();
        if (e == cudaSuccess)
            done.fetch_or(bit, std::memory_order_release);
        return e;
    }
};

// SUPER: the super-chunk form of the tile loop (a.super_chunks > 1; built for the pair-packed tiles of the
// k for which it measured faster, qreg_super_chunks(k) > 0)
template <int K, int Q, int NT, int MATH, bool SUPER = false>
static cudaError_t launch_qreg_one(const QregArgs &a, uint32_t qtiles, cudaStream_t st)
{
    if constexpr (!SUPER && MATH == 2 && qreg_super_chunks(K) > 0)
        if (a.super_chunks > 1u)
            return launch_qreg_one<K, Q, NT, MATH, true>(a, qtiles, st);
    auto kern = nn_qreg_kernel<K, Q, NT, MATH, SUPER>;
    static PerDeviceOnce once;
    const cudaError_t e = once([&] {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QregCfg<K>::SMEM);
    });
    if (e != cudaSuccess)
        return e;
#if NN_CTA_ORDER == 2
    const uint32_t qg = a.qgroup < 1 ? 1u : (a.qgroup > qtiles ? qtiles : a.qgroup);
    QregArgs b = a;
    b.qgroup = qg;
    kern<<<dim3(qg * a.splits, (qtiles + qg - 1) / qg), NT, QregCfg<K>::SMEM, st>>>(b);
#else
    kern<<<qtiles * a.splits, NT, QregCfg<K>::SMEM, st>>>(a);
#endif
    return cudaGetLastError();
}

template <int K, int Q, int NT, int MATH>
static cudaError_t query_qreg_one(LaunchInfo *info, int *tile_queries, int *tile_refs)
{
    auto kern = nn_qreg_kernel<K, Q, NT, MATH>;
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
// math: 2 = f32x2 over query pairs (default; needs Q >= 2, Q = 1 uses f32x2 over dimension pairs);
//       1 = f32x2 over dimension pairs, 0 = scalar: only in builds with -DNN_AB_MATH (A/B measurements).
template <class F>
static cudaError_t qreg_dispatch(int q_sel, int math, F f)
{
    constexpr int QD = QregDefault<NN_K>::Q;
    const int qq = q_sel == 0 ? QD : q_sel;
    auto go = [&](auto qc) -> cudaError_t {
        constexpr int QV = decltype(qc)::value;
        if constexpr (QV <= QD)
        {
#ifdef NN_AB_MATH // A/B builds only: the shipped library holds the pair-packed kernels (and mode 1 for Q = 1)
            if (math == 0)
                return f(qc, std::integral_constant<int, 0>{});
            if (math == 1 || QV == 1)
                return f(qc, std::integral_constant<int, 1>{});
#else
            if (math != 2)
                return cudaErrorInvalidValue;
#endif
            return f(qc, std::integral_constant<int, (QV >= 2 ? 2 : 1)>{});
        }
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
cudaError_t launch_qreg<NN_K>(int q_sel, int math, const QregArgs &a, uint32_t qtiles, cudaStream_t st)
{
    return qreg_dispatch(q_sel, math, [&](auto qc, auto pk) {
        return launch_qreg_one<NN_K, decltype(qc)::value, 128, decltype(pk)::value>(a, qtiles, st);
    });
}

template <>
cudaError_t query_qreg<NN_K>(int q_sel, int math, LaunchInfo *info, int *tile_queries, int *tile_refs)
{
    return qreg_dispatch(q_sel, math, [&](auto qc, auto pk) {
        return query_qreg_one<NN_K, decltype(qc)::value, 128, decltype(pk)::value>(info, tile_queries, tile_refs);
    });
}

// ---- phased query-register kernel ---------------------------------------------------------------
template <int K, int Q>
static cudaError_t launch_qflex_one(const QflexArgs &a, uint32_t qtiles, cudaStream_t st)
{
    auto kern = nn_qflex_kernel<K, Q, 128>;
    static PerDeviceOnce once;
    const cudaError_t e = once([&] {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlexMaxSmem);
    });
    if (e != cudaSuccess)
        return e;
    const size_t smem = (size_t)kFlexRingOffset + (size_t)a.stages * a.stage_floats * 4;
    if (a.stages < 2 || a.stages > (uint32_t)kFlexMaxStages || smem > (size_t)kFlexMaxSmem ||
        (size_t)a.stages * a.stage_floats * 4 < (size_t)128 * Q * 8 || (size_t)a.tile_groups * Geo<K>::G * K > a.stage_floats)
        return cudaErrorInvalidValue;
#if NN_CTA_ORDER == 2
    const uint32_t qg = a.qgroup < 1 ? 1u : (a.qgroup > qtiles ? qtiles : a.qgroup);
    QflexArgs b = a;
    b.qgroup = qg;
    kern<<<dim3(qg * a.splits, (qtiles + qg - 1) / qg), 128, smem, st>>>(b);
#else
    kern<<<qtiles * a.splits, 128, smem, st>>>(a);
#endif
    return cudaGetLastError();
}

template <int K, int Q>
static cudaError_t query_qflex_one(int smem_bytes, FlexInfo *info)
{
    auto kern = nn_qflex_kernel<K, Q, 128>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlexMaxSmem);
    if (e != cudaSuccess)
        return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess)
        return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, (size_t)smem_bytes);
    if (e != cudaSuccess)
        return e;
    info->regs = fa.numRegs;
    info->smem = smem_bytes;
    info->occ = occ;
    info->ch = QregCfg<K>::CH;
    info->g = Geo<K>::G;
    info->tr = QregCfg<K>::TR;
    return cudaSuccess;
}

template <class F>
static cudaError_t qflex_dispatch(int q, F f)
{
    constexpr int QD = QregDefault<NN_K>::Q;
    if (q == 2)
        return f(std::integral_constant<int, 2>{});
    if (q == 4)
    {
        if constexpr (QD >= 4)
            return f(std::integral_constant<int, 4>{});
    }
    if (q == 8)
    {
        if constexpr (QD >= 8)
            return f(std::integral_constant<int, 8>{});
    }
    return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_qflex<NN_K>(int q, const QflexArgs &a, uint32_t qtiles, cudaStream_t st)
{
    return qflex_dispatch(q, [&](auto qc) { return launch_qflex_one<NN_K, decltype(qc)::value>(a, qtiles, st); });
}

template <>
cudaError_t query_qflex<NN_K>(int q, int smem_bytes, FlexInfo *info)
{
    return qflex_dispatch(q, [&](auto qc) { return query_qflex_one<NN_K, decltype(qc)::value>(smem_bytes, info); });
}

// ---- reference-register kernel ------------------------------------------------------------------
template <int K>
struct RregCfg
{
    static constexpr int NT = 128;
    static constexpr int NS = NN_RREG_SLOTS;
    // groups per thread per ring slot: ~NN_RREG_SLOT_FLOATS reference floats
    static constexpr int PS = (NN_RREG_SLOT_FLOATS / (Geo<K>::G * K)) >= 1 ? (NN_RREG_SLOT_FLOATS / (Geo<K>::G * K)) : 1;
    static constexpr int PS_SOA = (NN_RREG_SLOT_FLOATS / K) >= 2 ? (NN_RREG_SLOT_FLOATS / K) : 2;
    // CTAs per SM the register budget is compiled for: the reference ring + per-query state
    static constexpr int minb(int mq, bool soa)
    {
        const int refs = NS * (soa ? PS_SOA : PS * Geo<K>::G) * K;
        const int need = refs + 3 * mq + 2 * K + 30; // ring, best/bref/rm, one query pair
#ifdef NN_RREG_FORCE_MINB
        return NN_RREG_FORCE_MINB;
#endif
        // measured on B200 (k = 8, m = 8): 4 CTAs/SM at a 128-register cap beat 3 CTAs at 148
        return need <= 136 ? 4 : (need <= 168 ? 3 : 2);
    }
};

template <int K, int MQ, bool SOA>
static cudaError_t launch_rreg_one(const RregArgs &a, dim3 grid, cudaStream_t st)
{
    constexpr int PS = SOA ? RregCfg<K>::PS_SOA : RregCfg<K>::PS;
    if ((a.mq_total + MQ - 1) / MQ != (int)grid.y || a.mq_total < 1)
        return cudaErrorInvalidValue;
    nn_rreg_kernel<K, MQ, PS, RregCfg<K>::NS, RregCfg<K>::NT, SOA, RregCfg<K>::minb(MQ, SOA)>
        <<<grid, RregCfg<K>::NT, 0, st>>>(a);
    return cudaGetLastError();
}

template <int K, int MQ, bool SOA>
static cudaError_t query_rreg_one(LaunchInfo *info, int *refs_per_batch)
{
    constexpr int PS = SOA ? RregCfg<K>::PS_SOA : RregCfg<K>::PS;
    auto kern = nn_rreg_kernel<K, MQ, PS, RregCfg<K>::NS, RregCfg<K>::NT, SOA, RregCfg<K>::minb(MQ, SOA)>;
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
    *refs_per_batch = RregCfg<K>::NT * PS * (SOA ? 1 : Geo<K>::G);
    return cudaSuccess;
}

#define NN_RREG_DISPATCH(FN, ...)                                                                                      \
    switch (mq)                                                                                                        \
    {                                                                                                                  \
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

// ---- reference-stream kernel (TMA ring) ----------------------------------------------------------
// One CTA per SM: 12 consumer warps + 1 producer warp, ring of up to 4 tiles in ~200 KB of shared
// memory (measured at k = 8, m = 8, n = 2^26 on B200: 0.41 ms, against 0.43-0.44 ms for two CTAs of
// 8+1 warps per SM and 0.42 ms for 16+1 warps).
#ifndef NN_RTMA_NW
#define NN_RTMA_NW 12 // consumer warps per CTA (one more warp produces)
#endif
#ifndef NN_RTMA_STAGES
#define NN_RTMA_STAGES 4 // at most; fewer when a tile is large (odd k: 4-point groups)
#endif
#ifndef NN_RTMA_MINB
#define NN_RTMA_MINB 1 // CTAs per SM the register budget is compiled for
#endif
#ifndef NN_RTMA_SLOT_FLOATS
#define NN_RTMA_SLOT_FLOATS 32 // reference floats per thread per tile
#endif
template <int K>
struct RtmaSel
{
    static constexpr int NW = NN_RTMA_NW, MINB = NN_RTMA_MINB;
    static constexpr int PT = (NN_RTMA_SLOT_FLOATS / (Geo<K>::G * K)) >= 1 ? (NN_RTMA_SLOT_FLOATS / (Geo<K>::G * K)) : 1;
    static constexpr int TILE_BYTES = NW * 32 * PT * Geo<K>::G * K * 4;
    static constexpr int FIT = (200 * 1024 / MINB) / TILE_BYTES;
    static constexpr int STAGES = FIT >= NN_RTMA_STAGES ? NN_RTMA_STAGES : (FIT >= 2 ? FIT : 2);
    using Cfg = RtmaCfg<K, PT, NW, STAGES>;
};

template <int K, int MQ>
static cudaError_t launch_rtma_one(const RregArgs &a, dim3 grid, cudaStream_t st)
{
    using S = RtmaSel<K>;
    if ((a.mq_total + MQ - 1) / MQ != (int)grid.y || a.mq_total < 1)
        return cudaErrorInvalidValue;
    auto kern = nn_rtma_kernel<K, MQ, S::PT, S::NW, S::STAGES, S::MINB>;
    static PerDeviceOnce once;
    const cudaError_t e = once([&] {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::Cfg::smem(MQ));
    });
    if (e != cudaSuccess)
        return e;
    kern<<<grid, (S::NW + 1) * 32, S::Cfg::smem(MQ), st>>>(a);
    return cudaGetLastError();
}

template <int K, int MQ>
static cudaError_t query_rtma_one(LaunchInfo *info, int *tile_refs)
{
    using S = RtmaSel<K>;
    auto kern = nn_rtma_kernel<K, MQ, S::PT, S::NW, S::STAGES, S::MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::Cfg::smem(MQ));
    if (e != cudaSuccess)
        return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess)
        return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (S::NW + 1) * 32, S::Cfg::smem(MQ));
    if (e != cudaSuccess)
        return e;
    info->regs = fa.numRegs;
    info->smem = (int)S::Cfg::smem(MQ);
    info->occ = occ;
    *tile_refs = S::Cfg::TILE_REFS;
    return cudaSuccess;
}

template <>
cudaError_t launch_rtma<NN_K>(int mq, const RregArgs &a, dim3 grid, cudaStream_t st)
{
    switch (mq)
    {
    case 2:
        return launch_rtma_one<NN_K, 2>(a, grid, st);
    case 4:
        return launch_rtma_one<NN_K, 4>(a, grid, st);
    case 8:
        return launch_rtma_one<NN_K, 8>(a, grid, st);
    default:
        return cudaErrorInvalidValue;
    }
}

template <>
cudaError_t query_rtma<NN_K>(int mq, LaunchInfo *info, int *tile_refs)
{
    switch (mq)
    {
    case 2:
        return query_rtma_one<NN_K, 2>(info, tile_refs);
    case 4:
        return query_rtma_one<NN_K, 4>(info, tile_refs);
    case 8:
        return query_rtma_one<NN_K, 8>(info, tile_refs);
    default:
        return cudaErrorInvalidValue;
    }
}

template <>
cudaError_t launch_plain<NN_K>(const float *S, const float *R, int m, uint32_t n, uint32_t index_base, uint32_t splits,
                               unsigned long long *keys, int peer_keys, cudaStream_t st)
{
    if (splits < 1)
        splits = 1;
    const uint32_t per = (n + splits - 1) / splits;
    dim3 grid((m + 127) / 128, splits);
    nn_plain_kernel<NN_K><<<grid, 128, 0, st>>>(S, R, m, n, index_base, per, keys, peer_keys);
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
