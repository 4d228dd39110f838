// nn_kernels.cuh -- sm_100a kernels of the brute-force 1-NN path.
//
// What is replaced (all in /root/reference/sources/src/core.cu):
//   cudaCallbackKernel<1024>   808-855 (v7 copy 662-709)   -> nn_qreg_kernel / nn_rreg_kernel
//   mat_inv_kernel             792-807                      -> nn_repack_soa_kernel (nn_repack.cu);
//                                                              the search kernels read AoS directly
//   host second-level reduce   765-787, 936-957             -> 64-bit atomicMin on packed keys
//
// Arithmetic contract (v0, core.cu:44-54): d2 = ((d0*d0 + d1*d1) + d2*d2) + ... with
// d_i = q_i - r_i, every operation IEEE round-to-nearest, never fused.  The packed
// FADD2/FMUL2 (f32x2) forms used here are element-wise IEEE operations, so they give the same
// bits as the scalar ones; the adds stay scalar and sequential.  `0 + d0*d0` is elided: it is
// exact for every d0*d0 (which is never -0).
//
// Tie rule (strict `>` over ascending nInd, core.cu:50-54): the winner is the LOWEST index
// among the references at minimum distance; NaN distances never win; a query that nothing
// beats keeps (INFINITY, index 0).  Every kernel reduces (distance, index) as the packed key
// (float_bits(d2) << 32) | index, whose unsigned order is exactly that rule (d2 >= +0).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "nn_launch.h"

namespace nnb200
{

constexpr unsigned long long KEY_INIT = 0x7F80000000000000ull;
constexpr uint32_t NO_REF = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned long long pack_key(float d2, uint32_t idx)
{
    return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)idx;
}

// References come in 16-byte-aligned groups of G points (G*K floats = F4 float4) so that every
// k in 3..16 can be moved with 128-bit accesses from the native AoS layout.
template <int K>
struct Geo
{
    static constexpr int G = (K % 4 == 0) ? 1 : ((K % 2 == 0) ? 2 : 4);
    static constexpr int F4 = G * K / 4;
};

// ---- squared distance, v0 arithmetic ---------------------------------------------------------
// PAR = parity of r's offset inside a register array that starts on an even register: pairs
// (d, d+1) are formed where r's pair is register-aligned, the odd dimension out is scalar.
template <int K, int PAR, bool PACKED>
__device__ __forceinline__ float sqdist(const float (&q)[K], const float *r)
{
    float p[K];
    if (PACKED)
    {
        if (PAR)
        {
            const float d = __fsub_rn(q[0], r[0]);
            p[0] = __fmul_rn(d, d);
        }
#pragma unroll
        for (int i = PAR; i + 1 < K; i += 2)
        {
            const float2 d = __fadd2_rn(make_float2(q[i], q[i + 1]), make_float2(-r[i], -r[i + 1]));
            const float2 s = __fmul2_rn(d, d);
            p[i] = s.x;
            p[i + 1] = s.y;
        }
        if ((K - PAR) & 1)
        {
            const float d = __fsub_rn(q[K - 1], r[K - 1]);
            p[K - 1] = __fmul_rn(d, d);
        }
    }
    else
    {
#pragma unroll
        for (int i = 0; i < K; ++i)
        {
            const float d = __fsub_rn(q[i], r[i]);
            p[i] = __fmul_rn(d, d);
        }
    }
    float acc = p[0];
#pragma unroll
    for (int i = 1; i < K; ++i)
        acc = __fadd_rn(acc, p[i]);
    return acc;
}

// `par` is a constant after unrolling; the untaken side folds away.
template <int K, bool PACKED>
__device__ __forceinline__ float sqdist_par(const float (&q)[K], const float *r, const int par)
{
    return par ? sqdist<K, 1, PACKED>(q, r) : sqdist<K, 0, PACKED>(q, r);
}

// Scalar distance straight from global memory (index resolution, plain kernel).
template <int K>
__device__ __forceinline__ float sqdist_gmem(const float (&q)[K], const float *__restrict__ r)
{
    float rr[K];
#pragma unroll
    for (int i = 0; i < K; ++i)
        rr[i] = __ldg(r + i);
    return sqdist<K, 0, false>(q, rr);
}

// ---- mbarrier / bulk-copy (TMA 1-D) primitives ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// =============================================================================================
// Kernel A -- "query-register" kernel, for many queries (FP32-pipe bound).
//
// A CTA of NT threads owns a tile of NT*Q queries; each thread keeps Q queries (Q*K floats) in
// registers.  The CTA walks a contiguous range of reference tiles (TR points, native AoS) that the
// TMA unit streams into a STAGES-deep shared-memory ring (cp.async.bulk + mbarrier).  Every thread
// reads the SAME reference from shared memory (128-bit broadcast loads), so one LDS.128 feeds
// Q*4 dimensions of work.  Per chunk of CH references a thread keeps only the chunk minimum
// (FMNMX3) and a strict-less select on (best, chunk start); the exact index is resolved once per
// query at the end by re-scanning the winning chunk -- no per-pair index bookkeeping.
// Work item = (query tile, reference split); results are folded with atomicMin on packed keys.
// =============================================================================================
template <int K>
struct QregCfg
{
    static constexpr int CH = 4;                                // references per chunk
    static constexpr int TR = ((2048 / K) / 8) * 8;             // references per tile (~8 KB)
    static constexpr int STAGES = 3;
    static constexpr int TILE_FLOATS = TR * K;
    static constexpr uint32_t TILE_BYTES = TILE_FLOATS * 4u;
    static constexpr size_t SMEM = (size_t)STAGES * TILE_BYTES + 64;
};

template <int K, int Q, bool PACKED>
__device__ __forceinline__ void qreg_chunk(const float *__restrict__ sm, const float (&q)[Q][K], float (&cm)[Q])
{
    constexpr int G = Geo<K>::G, F4 = Geo<K>::F4, CH = QregCfg<K>::CH;
    float hold[Q];
#pragma unroll
    for (int g0 = 0; g0 < CH; g0 += G)
    {
        float grp[G * K];
        const float4 *p4 = reinterpret_cast<const float4 *>(sm + g0 * K);
#pragma unroll
        for (int i = 0; i < F4; ++i)
        {
            const float4 v = p4[i];
            grp[4 * i + 0] = v.x;
            grp[4 * i + 1] = v.y;
            grp[4 * i + 2] = v.z;
            grp[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
        {
            const int c = g0 + g;
#pragma unroll
            for (int j = 0; j < Q; ++j)
            {
                const float d = sqdist_par<K, PACKED>(q[j], &grp[g * K], (g * K) & 1);
                if ((c & 1) == 0)
                    hold[j] = d;
                else if (c == 1)
                    cm[j] = fminf(hold[j], d);
                else
                    cm[j] = fminf(fminf(cm[j], hold[j]), d);
            }
        }
    }
}

template <int K, int Q, int NT, bool PACKED>
__global__ void __launch_bounds__(NT, 512 / NT) nn_qreg_kernel(const QregArgs a)
{
    using C = QregCfg<K>;
    constexpr int CH = C::CH, TR = C::TR, STAGES = C::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * C::TILE_BYTES);

    const int tid = threadIdx.x;
    const uint32_t split = blockIdx.x % a.splits;
    const uint32_t qtile = blockIdx.x / a.splits;
    const uint32_t full_tiles = a.n / TR;
    const uint32_t rem = a.n - full_tiles * TR;
    const uint32_t t0 = min(split * a.tiles_per_split, full_tiles);
    const uint32_t t1 = min(t0 + a.tiles_per_split, full_tiles);
    const bool tail = (rem != 0) && (split == a.splits - 1);
    if (t0 >= t1 && !tail)
        return;

    // this thread's queries (clamped so that out-of-range slots compute on a valid row)
    float q[Q][K];
    float best[Q];
    uint32_t bref[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j)
    {
        int qi = (int)(qtile * (NT * Q)) + j * NT + tid;
        qi = qi < a.m ? qi : a.m - 1;
        const float *src = a.S + (size_t)qi * K;
#pragma unroll
        for (int i = 0; i < K; ++i)
            q[j][i] = __ldg(src + i);
        best[j] = __int_as_float(0x7f800000);
        bref[j] = NO_REF;
    }

    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if (t0 + s < t1)
            {
                mbar_expect_tx(&full[s], C::TILE_BYTES);
                bulk_g2s(tiles + (size_t)s * C::TILE_FLOATS, a.R + (size_t)(t0 + s) * C::TILE_FLOATS, C::TILE_BYTES,
                         &full[s]);
            }
    }

    uint32_t stage = 0, parity = 0;
    for (uint32_t t = t0; t < t1; ++t)
    {
        __syncthreads(); // everyone is done with tile t-1: its stage may be refilled
        if (tid == 0)
        {
            const uint32_t tn = t + STAGES - 1;
            if (tn < t1)
            {
                const uint32_t sn = (stage + STAGES - 1) % STAGES;
                mbar_expect_tx(&full[sn], C::TILE_BYTES);
                bulk_g2s(tiles + (size_t)sn * C::TILE_FLOATS, a.R + (size_t)tn * C::TILE_FLOATS, C::TILE_BYTES,
                         &full[sn]);
            }
        }
        mbar_wait(&full[stage], parity);
        const float *sm = tiles + (size_t)stage * C::TILE_FLOATS;
        const uint32_t ref0 = t * TR;
#pragma unroll 2
        for (int c = 0; c < TR; c += CH)
        {
            float cm[Q];
            qreg_chunk<K, Q, PACKED>(sm + c * K, q, cm);
#pragma unroll
            for (int j = 0; j < Q; ++j)
                if (cm[j] < best[j])
                {
                    best[j] = cm[j];
                    bref[j] = ref0 + c;
                }
        }
        if (++stage == STAGES)
        {
            stage = 0;
            parity ^= 1;
        }
    }

    if (tail)
    {
        // last, partial tile: plain cooperative loads; slots past n are NaN (a NaN distance never wins)
        __syncthreads();
        const uint32_t ref0 = full_tiles * TR;
        const uint32_t padded = ((rem + CH - 1) / CH) * CH;
        const float *src = a.R + (size_t)ref0 * K;
        for (uint32_t i = tid; i < padded * K; i += NT)
            tiles[i] = (i < rem * K) ? __ldg(src + i) : __int_as_float(0x7fffffff);
        __syncthreads();
        for (uint32_t c = 0; c < padded; c += CH)
        {
            float cm[Q];
            qreg_chunk<K, Q, PACKED>(tiles + c * K, q, cm);
#pragma unroll
            for (int j = 0; j < Q; ++j)
                if (cm[j] < best[j])
                {
                    best[j] = cm[j];
                    bref[j] = ref0 + c;
                }
        }
    }

    // resolve the exact (lowest) index inside the winning chunk, then fold into the global keys
#pragma unroll
    for (int j = 0; j < Q; ++j)
    {
        const int qi = (int)(qtile * (NT * Q)) + j * NT + tid;
        if (bref[j] != NO_REF && qi < a.m)
        {
            uint32_t idx = bref[j];
#pragma unroll
            for (int c = CH - 1; c >= 0; --c)
            {
                const uint32_t r = bref[j] + c;
                if (r < a.n)
                {
                    const float d = sqdist_gmem<K>(q[j], a.R + (size_t)r * K);
                    if (d == best[j])
                        idx = r;
                }
            }
            atomicMin(a.keys + qi, pack_key(best[j], a.index_base + idx));
        }
    }
}

// =============================================================================================
// Kernel B -- "reference-register" kernel, for few queries (HBM-streaming / SM-fill bound).
//
// The roles are swapped: every thread streams its OWN references from HBM straight into registers
// (128-bit loads from the native AoS layout, next batch prefetched while the current one is
// computed) and all MQ queries of the pass are broadcast from shared memory.  A persistent grid
// (SMs x occupancy CTAs) strides over the reference set, so all SMs are busy even for m = 1.
// Per thread and query: batch minimum (FMNMX3) + strict-less select on (best, batch start); the
// exact index is resolved at the end, keys are reduced across the warp with __shfl_xor and folded
// with one 64-bit atomicMin per warp and query.
// SOA = true reads references from the repacked [k][n] layout instead (coalesced 32-bit loads).
// =============================================================================================
template <int K, int MQ, int PG, int NT, bool SOA, int MINB>
__global__ void __launch_bounds__(NT, MINB) nn_rreg_kernel(const RregArgs a)
{
    constexpr int G = SOA ? 1 : Geo<K>::G;
    constexpr int F4 = G * K / 4; // (AoS only) float4 per group
    constexpr int P = G * PG; // references per thread per batch
    __shared__ __align__(16) float sq[MQ * K];

    const int tid = threadIdx.x;
    const int pass = blockIdx.y;
    const float *S = a.S + (size_t)pass * MQ * K;
    unsigned long long *keys = a.keys + (size_t)pass * MQ;
    for (int i = tid; i < MQ * K; i += NT)
        sq[i] = __ldg(S + i);
    __syncthreads();

    float best[MQ];
    uint32_t bref[MQ];
#pragma unroll
    for (int j = 0; j < MQ; ++j)
    {
        best[j] = __int_as_float(0x7f800000);
        bref[j] = NO_REF;
    }

    // Batch b covers groups [b*NT*PG, (b+1)*NT*PG); thread t owns groups b*NT*PG + i*NT + t.
    const uint32_t ngroups = (a.n + G - 1) / G;
    const uint32_t groups_per_batch = NT * PG;
    const uint32_t nbatches = (ngroups + groups_per_batch - 1) / groups_per_batch;

    float cur[P * K], nxt[P * K];

    auto load_batch = [&](uint32_t b, float(&dst)[P * K]) {
#pragma unroll
        for (int i = 0; i < PG; ++i)
        {
            const uint32_t grp = b * groups_per_batch + i * NT + tid;
            const uint32_t r0 = grp * G;
            if constexpr (SOA)
            {
#pragma unroll
                for (int d = 0; d < K; ++d)
                    dst[i * K + d] = (r0 < a.n) ? __ldg(a.R + (size_t)d * a.n + r0) : __int_as_float(0x7fffffff);
            }
            else if (r0 + G <= a.n)
            {
                const float4 *p4 = reinterpret_cast<const float4 *>(a.R + (size_t)r0 * K);
#pragma unroll
                for (int f = 0; f < F4; ++f)
                {
                    const float4 v = __ldg(p4 + f);
                    dst[i * G * K + 4 * f + 0] = v.x;
                    dst[i * G * K + 4 * f + 1] = v.y;
                    dst[i * G * K + 4 * f + 2] = v.z;
                    dst[i * G * K + 4 * f + 3] = v.w;
                }
            }
            else
            {
#pragma unroll
                for (int e = 0; e < G * K; ++e)
                    dst[i * G * K + e] =
                        ((size_t)r0 * K + e < (size_t)a.n * K) ? __ldg(a.R + (size_t)r0 * K + e) : __int_as_float(0x7fffffff);
            }
        }
    };

    auto compute_batch = [&](uint32_t b, const float(&ref)[P * K]) {
#pragma unroll
        for (int j = 0; j < MQ; ++j)
        {
            float qv[K];
            if (K % 4 == 0)
            {
                const float4 *q4 = reinterpret_cast<const float4 *>(sq + j * K);
#pragma unroll
                for (int f = 0; f < K / 4; ++f)
                {
                    const float4 v = q4[f];
                    qv[4 * f + 0] = v.x;
                    qv[4 * f + 1] = v.y;
                    qv[4 * f + 2] = v.z;
                    qv[4 * f + 3] = v.w;
                }
            }
            else
            {
#pragma unroll
                for (int d = 0; d < K; ++d)
                    qv[d] = sq[j * K + d];
            }
            float bm = 0.f, hold = 0.f;
#pragma unroll
            for (int p = 0; p < P; ++p)
            {
                const float d = sqdist_par<K, true>(qv, &ref[p * K], (p * K) & 1);
                if (p == 0)
                    bm = d;
                else if ((p & 1) == 1 && p + 1 < P)
                    hold = d;
                else if ((p & 1) == 0)
                    bm = fminf(fminf(bm, hold), d);
                else
                    bm = fminf(bm, d);
            }
            if (bm < best[j])
            {
                best[j] = bm;
                bref[j] = b;
            }
        }
    };

    // software pipeline over this CTA's batches: the loads of the next batch are in flight while
    // the current one is computed; two register buffers alternate (no copies)
    uint32_t b = blockIdx.x;
    if (b < nbatches)
    {
        load_batch(b, cur);
        for (;;)
        {
            uint32_t bn = b + gridDim.x;
            if (bn < nbatches)
                load_batch(bn, nxt);
            compute_batch(b, cur);
            if (bn >= nbatches)
                break;
            b = bn;
            bn = b + gridDim.x;
            if (bn < nbatches)
                load_batch(bn, cur);
            compute_batch(b, nxt);
            if (bn >= nbatches)
                break;
            b = bn;
        }
    }

    // resolve: lowest index within this thread's winning batch (its groups ascend with i)
    const int lane = tid & 31;
#pragma unroll
    for (int j = 0; j < MQ; ++j)
    {
        unsigned long long key = KEY_INIT | NO_REF;
        if (bref[j] != NO_REF)
        {
            float qv[K];
#pragma unroll
            for (int d = 0; d < K; ++d)
                qv[d] = sq[j * K + d];
            uint32_t idx = 0;
#pragma unroll
            for (int i = PG - 1; i >= 0; --i)
            {
#pragma unroll
                for (int g = G - 1; g >= 0; --g)
                {
                    const uint32_t r = (bref[j] * groups_per_batch + i * NT + tid) * G + g;
                    if (r < a.n)
                    {
                        float rr[K];
#pragma unroll
                        for (int d = 0; d < K; ++d)
                            rr[d] = SOA ? __ldg(a.R + (size_t)d * a.n + r) : __ldg(a.R + (size_t)r * K + d);
                        const float d2 = sqdist<K, 0, false>(qv, rr);
                        if (d2 == best[j])
                            idx = r;
                    }
                }
            }
            key = pack_key(best[j], a.index_base + idx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if (lane == 0 && key < (KEY_INIT | NO_REF))
            atomicMin(keys + j, key);
    }
}

// =============================================================================================
// Plain kernel -- one thread per query, references read from global memory, per-pair strict-less
// update.  An independent, deliberately simple formulation kept as an on-device cross-check of
// the two tuned kernels at sizes the CPU oracle cannot reach.
// =============================================================================================
template <int K>
__global__ void __launch_bounds__(128) nn_plain_kernel(const float *__restrict__ S, const float *__restrict__ R, int m,
                                                      uint32_t n, uint32_t index_base, uint32_t refs_per_split,
                                                      unsigned long long *keys)
{
    const int qi = blockIdx.x * 128 + threadIdx.x;
    if (qi >= m)
        return;
    float q[K];
#pragma unroll
    for (int i = 0; i < K; ++i)
        q[i] = S[(size_t)qi * K + i];
    const uint32_t r0 = blockIdx.y * refs_per_split;
    const uint32_t r1 = min(n, r0 + refs_per_split);
    float best = __int_as_float(0x7f800000);
    uint32_t bidx = NO_REF;
    for (uint32_t r = r0; r < r1; ++r)
    {
        const float d = sqdist_gmem<K>(q, R + (size_t)r * K);
        if (d < best)
        {
            best = d;
            bidx = r;
        }
    }
    if (bidx != NO_REF)
        atomicMin(keys + qi, pack_key(best, index_base + bidx));
}

// =============================================================================================
// AoS [n][k] -> SoA [k][n] repack: out[d*n + j] = in[j*k + d]      (mat_inv_kernel, core.cu:792-807)
//
// The reference reads one 4-byte element per thread at stride k (1/3 .. 1/16 sector efficiency) and
// idles 16..29 of every 32 thread rows.  Here a CTA stages a tile of TN points through shared
// memory: 128-bit coalesced global loads of the contiguous AoS tile, scatter into a padded [k][TN+4]
// shared tile, then 128-bit shared loads and 128-bit coalesced global stores per dimension row.
// HBM traffic is the algorithmic 2*n*k*4 bytes.  Persistent grid-stride loop over tiles.
// =============================================================================================
template <int K>
struct RepackCfg
{
    static constexpr int NT = 256;
    static constexpr int TN = ((6144 / K) / 4) * 4; // points per tile (~24 KB)
    static constexpr int TNP = TN + 4;              // padded row: keeps 16-byte alignment, spreads banks
    static constexpr int V = TN * K / 4;            // float4 per full tile
    static constexpr int VPT = (V + NT - 1) / NT;   // float4 per thread
};

template <int K>
__global__ void __launch_bounds__(RepackCfg<K>::NT) nn_repack_soa_kernel(const float *__restrict__ in,
                                                                          float *__restrict__ out, uint32_t n)
{
    using C = RepackCfg<K>;
    constexpr int NT = C::NT, TN = C::TN, TNP = C::TNP, V = C::V, VPT = C::VPT;
    __shared__ __align__(16) float sm[K * TNP];
    const int tid = threadIdx.x;
    const uint32_t ntiles = (n + TN - 1) / TN;
    const bool vec_out = (n % 4u) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        const uint32_t r0 = tile * TN;
        const uint32_t cnt = min((uint32_t)TN, n - r0);
        const uint32_t nflt = cnt * K;
        const float *src = in + (size_t)r0 * K;
        const float4 *src4 = reinterpret_cast<const float4 *>(src);
        float4 x[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i)
        {
            const uint32_t v = i * NT + tid;
            const uint32_t e = 4 * v;
            if (v < V && e + 3 < nflt)
                x[i] = __ldg(src4 + v);
            else
            {
                x[i].x = (v < V && e + 0 < nflt) ? __ldg(src + e + 0) : 0.f;
                x[i].y = (v < V && e + 1 < nflt) ? __ldg(src + e + 1) : 0.f;
                x[i].z = (v < V && e + 2 < nflt) ? __ldg(src + e + 2) : 0.f;
                x[i].w = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < VPT; ++i)
        {
            const uint32_t v = i * NT + tid;
            if (v < V)
            {
                const uint32_t e = 4 * v;
                const float c[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
                {
                    const uint32_t j = (e + u) / K, d = (e + u) % K;
                    sm[d * TNP + j] = c[u];
                }
            }
        }
        __syncthreads();
        for (int w = tid; w < K * (TN / 4); w += NT)
        {
            const uint32_t d = w / (TN / 4), j = 4 * (w % (TN / 4));
            if (j < cnt)
            {
                const float4 v = *reinterpret_cast<const float4 *>(&sm[d * TNP + j]);
                float *dst = out + (size_t)d * n + r0 + j;
                if (vec_out && j + 3 < cnt)
                    *reinterpret_cast<float4 *>(dst) = v;
                else
                {
                    dst[0] = v.x;
                    if (j + 1 < cnt)
                        dst[1] = v.y;
                    if (j + 2 < cnt)
                        dst[2] = v.z;
                    if (j + 3 < cnt)
                        dst[3] = v.w;
                }
            }
        }
        __syncthreads();
    }
}

} // namespace nnb200
