// nn_kernels.cuh -- sm_100a kernels of the brute-force 1-NN path.
//
// What is replaced (all in /root/reference/sources/src/core.cu):
//   cudaCallbackKernel<1024>   808-855 (v7 copy 662-709)   -> nn_qreg_kernel (many queries), nn_qflex_kernel
//                                                              (25 .. a few hundred), nn_rtma_kernel (5..24),
//                                                              nn_rreg_kernel (1..4)
//   mat_inv_kernel             792-807                      -> nn_repack_soa_kernel; the search kernels read
//                                                              AoS directly
//   host second-level reduce   765-787, 936-957             -> 64-bit atomicMin on packed keys; finish_group:
//                                                              the last CTA per query tile stores the indices
//                                                              (one launch per search); peer_push_group: the
//                                                              same across processes over NVLink peer memory
//
// Arithmetic contract (v0, core.cu:44-54): d2 = ((d0*d0 + d1*d1) + d2*d2) + ... with
// d_i = q_i - r_i, every operation IEEE round-to-nearest, never fused.  The packed
// FADD2/FMUL2 (f32x2) forms used here are element-wise IEEE operations, so they give the same
// bits as the scalar ones; the sum over the dimensions stays sequential.  `0 + d0*d0` is elided:
// it is exact for every d0*d0 (which is never -0).
//
// Math modes (MATH template parameter):
//   0  scalar FADD/FMUL/FADD                                   (kept for A/B measurements)
//   1  f32x2 across dimension pairs, scalar sequential adds    (kept for A/B measurements)
//   2  f32x2 across a PAIR OF QUERIES: (qa_d, qb_d) - bcast(r_d), squared, accumulated -- every
//      FP instruction of the hot loop is packed.  Measured on B200: streams that interleave packed
//      and scalar FP32 instructions lose ~20% of the FMA pipe, all-packed streams reach 99%.
// ptxas 12.9 contracts mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 although `.rn` forbids it, so
// mode 2 writes the square as fma(d, d, -0.0) with the -0.0 passed in at run time (exact, see
// sqdist_pair).  tests/test_abi_and_host.py asserts that the shipped SASS holds no scalar FFMA and
// exactly one FFMA2 per (2k-1)/k FADD2 in the pair kernels (a contracted add would change the
// ratio), and the `twins` parity family fails on any fused or re-ordered sum.
//
// Tie rule (strict `>` over ascending nInd, core.cu:50-54): the winner is the LOWEST index
// among the references at minimum distance; NaN distances never win; a query that nothing
// beats keeps (INFINITY, index 0).  Every kernel reduces (distance, index) as the packed key
// (float_bits(d2) << 32) | index, whose unsigned order is exactly that rule (d2 >= +0).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "nn_launch.h"

// build-time tuning knobs (overridable with -D for A/B builds; defaults are the measured best)
#ifndef NN_QREG_UNROLL
#define NN_QREG_UNROLL 1 // chunks of CH references unrolled in the tile loop
#endif
#ifndef NN_RREG_SLOTS
#define NN_RREG_SLOTS 2 // register ring depth of the reference-register kernel
#endif
#ifndef NN_RREG_SLOT_FLOATS
#define NN_RREG_SLOT_FLOATS 32 // reference floats per thread per ring slot
#endif
#ifndef NN_RREG_LDG256
#define NN_RREG_LDG256 0 // 1: 256-bit global loads where a reference group is a multiple of 32 bytes (no gain measured)
#endif
#ifndef NN_RREG_L2PF
#define NN_RREG_L2PF 0 // slot-batches ahead that a CTA asks the L2 to prefetch (0 = off; no gain measured)
#endif
#ifndef NN_RTMA_ROTATE
#define NN_RTMA_ROTATE 1 // reference-stream kernel, k = 16: bank-conflict-free rotated chunk reads (m=8, n=2^25: 0.521 -> 0.444 ms)
#endif
#ifndef NN_RTMA_ROTATE2
#define NN_RTMA_ROTATE2 0 // same idea for k = 8 (2-way conflict): measured 3% SLOWER at k=8, m=8 (0.422 vs 0.409 ms)
#endif
#ifndef NN_RTMA_QUERY_REGS
#define NN_RTMA_QUERY_REGS 0 // reference-stream kernel: keep the query pairs in registers (A/B)
#endif
#ifndef NN_RTMA_PIPELINE
#define NN_RTMA_PIPELINE 1 // reference-stream kernel: software-pipeline the shared loads over half tiles
#endif
#ifndef NN_QREG_UNROLL_NARROW
#define NN_QREG_UNROLL_NARROW 1 // unroll the chunk loop 2x (Q = 4) / 4x (Q <= 2): -10% at Q = 1 and at k = 16, m = 1024; +-1% elsewhere
#endif
#ifndef NN_QREG_CH8
#define NN_QREG_CH8 1 // 8-point chunks at k = 3 (fewer compares and selects per pair): +1..4% there, -0.4% at k = 4
#endif
#ifndef NN_QREG_PREFETCH
#define NN_QREG_PREFETCH 1 // query-register kernel: load the next reference group while computing the current one
#endif
#ifndef NN_QFLEX_MINB5
#define NN_QFLEX_MINB5 0 // phased kernel: compile the small-k instantiations for 5 CTAs per SM (A/B)
#endif
#ifndef NN_QFLEX_UNROLL_Q8
#define NN_QFLEX_UNROLL_Q8 1 // phased query-register kernel: chunks unrolled in the tile loop, by queries per thread
#endif
#ifndef NN_QFLEX_UNROLL_Q4
#define NN_QFLEX_UNROLL_Q4 2
#endif
#ifndef NN_QFLEX_UNROLL_Q2
#define NN_QFLEX_UNROLL_Q2 4
#endif
#ifndef NN_CTA_ORDER
#define NN_CTA_ORDER 2 // CTA index -> (split, query tile): 2 grouped on a 2-D grid (default); A/B: 0 split fastest, 1 grouped, 1-D
#endif
#ifndef NN_QREG_REGCAP_LOW
#define NN_QREG_REGCAP_LOW 0 // 1: always compile the query-register kernel for 4 CTAs/SM (128 regs)
#endif

namespace nnb200
{

constexpr int kQregUnroll = NN_QREG_UNROLL;
constexpr unsigned long long KEY_INIT = 0x7F80000000000000ull;
constexpr uint32_t NO_REF = 0xFFFFFFFFu;
constexpr unsigned long long KEY_NONE = 0xFFFFFFFFFFFFFFFFull; // "no candidate": larger than every real key

__device__ __forceinline__ unsigned long long pack_key(float d2, uint32_t idx)
{
    return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)idx;
}

// Fold one candidate into a key array.  `peer` != 0: the array lives in ANOTHER GPU's memory (mapped
// over NVLink by cudaDeviceEnablePeerAccess): the minimum is then a system-scope atomic executed at
// the owning GPU's L2, which is how the shards of a multi-GPU search merge inside the search kernel
// itself (replaces the reference's host merge, core.cu:925-957, without a separate collective).
__device__ __forceinline__ void fold_key(unsigned long long *slot, unsigned long long key, int peer)
{
    if (peer)
        atomicMin_system(slot, key);
    else
        atomicMin(slot, key);
}

// Last-CTA-finishes protocol (struct Finish, nn_launch.h).  Called by every thread of the CTA after
// its folds.  The CTAs of a ticket group count themselves on tickets[group]; the one that draws the
// last ticket knows that every other CTA's folds are visible (each fenced before drawing), reads the
// group's final keys at the L2, emits results / keys_out and restores the start state of keys and
// ticket, so the workspace is ready for the next launch without an init kernel.
// BAR_ID / BAR_THREADS: the CTA barrier the protocol synchronises on -- barrier 0 with all threads
// (__syncthreads) or a named barrier over the first BAR_THREADS threads (kernels whose producer warp has
// already left).  Ordering: every thread's folds happen-before the barrier; thread 0 then draws the
// ticket with an acq_rel atomic at GPU scope (release: cumulative over what it observed through the
// barrier; acquire: the last drawer sees every earlier CTA's folds), which is the CUTLASS semaphore
// pattern and costs ONE fence per CTA instead of one per thread.
template <int BAR_ID, int BAR_THREADS>
__device__ __forceinline__ void cta_bar()
{
    if constexpr (BAR_ID == 0)
        __syncthreads();
    else
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(BAR_THREADS) : "memory");
}

// ---- multi-process merge (struct PeerSync) ----------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spins (one thread) until *flag >= target; gives up after about a minute and raises the error flag, so
// that a rank that never shows up cannot hang the GPU for good (ranks may well be seconds apart: context
// creation, lazy code loading, host-side work between searches).
__device__ __forceinline__ void peer_spin(const unsigned int *flag, unsigned int target, unsigned int *error)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < target)
    {
        __nanosleep(64);
        if (clock64() - t0 > 120000000000ll)
        {
            if (error)
                atomicExch(error, 1u);
            break;
        }
    }
}

// The last CTA of a ticket group in multi-process mode (all its threads): push the group's final keys
// from this rank's workspace into rank 0's buffer, restore the workspace, count the group; the last
// group counts the rank in, and on rank 0 waits for every rank and finishes the search.
// `keys` = this rank's workspace keys (already offset like the launch's queries), `q_abs` = index of
// the group's first query in the whole query set.
template <int BAR_ID, int BAR_THREADS>
__device__ __forceinline__ void peer_push_group(const Finish &f, unsigned long long *keys, int q_begin, int q_count,
                                                int q_abs)
{
    const PeerSync &p = f.peer;
    __shared__ uint32_t s_role; // 0: nothing more to do; 1: rank 0, last group of the call
    const int nthr = BAR_ID == 0 ? (int)blockDim.x : BAR_THREADS;
    const unsigned int buf = p.step & 1u;
    // the buffer was last used by search step-2, which rank 0 must have consumed
    if (p.step >= 2u && threadIdx.x == 0)
        peer_spin(p.done_local, p.step - 1u, p.error);
    cta_bar<BAR_ID, BAR_THREADS>();
    for (int i = threadIdx.x; i < q_count; i += nthr)
    {
        const unsigned long long key = __ldcg(keys + q_begin + i);
        if (key != KEY_INIT)
            atomicMin_system(p.keys + q_abs + i, key);
        __stcg(keys + q_begin + i, KEY_INIT);
    }
    cta_bar<BAR_ID, BAR_THREADS>();
    if (threadIdx.x == 0)
    {
        __threadfence_system(); // the pushes (observed through the barrier) are out before the group is counted
        const bool last = atomicAdd(p.groups_done, 1u) == p.num_groups - 1u;
        if (last)
        {
            *p.groups_done = 0u; // every group of the call has pushed: ready for the next call
            __threadfence_system();
            atomicAdd_system(p.arrive + buf, 1u);
        }
        s_role = (last && p.rank == 0u) ? 1u : 0u;
    }
    cta_bar<BAR_ID, BAR_THREADS>();
    if (s_role == 0u)
        return;
    if (threadIdx.x == 0)
        peer_spin(p.arrive + buf, p.world, p.error); // every rank's shard is in
    cta_bar<BAR_ID, BAR_THREADS>();
    for (int i = threadIdx.x; i < p.m; i += nthr)
    {
        const unsigned long long key = __ldcg(p.keys + i);
        if (f.results)
            f.results[i] = (int)(unsigned int)(key & 0xffffffffull);
        if (f.keys_out)
            f.keys_out[i] = key;
        __stcg(p.keys + i, KEY_INIT);
    }
    cta_bar<BAR_ID, BAR_THREADS>();
    if (threadIdx.x == 0)
    {
        p.arrive[buf] = 0u;
        __threadfence_system(); // the restored keys and counter precede the announcement
        for (unsigned int r = 0; r + 1u < p.world; ++r)
            st_release_sys(p.done_peers[r], p.step + 1u);
        st_release_sys(p.done_local, p.step + 1u);
    }
}

template <int BAR_ID = 0, int BAR_THREADS = 0>
__device__ __forceinline__ void finish_group(const Finish &f, unsigned long long *keys, uint32_t group,
                                             uint32_t expected, int q_begin, int q_count, int q_abs = -1)
{
    if (f.tickets == nullptr) // (uniform over the grid)
        return;
    __shared__ uint32_t s_last;
    const int nthr = BAR_ID == 0 ? (int)blockDim.x : BAR_THREADS;
    cta_bar<BAR_ID, BAR_THREADS>();
    if (threadIdx.x == 0)
    {
        uint32_t t;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(f.tickets + group) : "memory");
        s_last = (t == expected - 1u) ? 1u : 0u;
        if (s_last)
            f.tickets[group] = 0u;
    }
    cta_bar<BAR_ID, BAR_THREADS>();
    if (!s_last)
        return;
    if (f.peer.arrive != nullptr)
    { // multi-process mode (uniform over the grid)
        peer_push_group<BAR_ID, BAR_THREADS>(f, keys, q_begin, q_count, q_abs < 0 ? q_begin : q_abs);
        return;
    }
    for (int i = threadIdx.x; i < q_count; i += nthr)
    {
        const unsigned long long key = __ldcg(keys + q_begin + i);
        if (f.results)
            f.results[q_begin + i] = (int)(unsigned int)(key & 0xffffffffull);
        if (f.keys_out)
            f.keys_out[q_begin + i] = key;
        __stcg(keys + q_begin + i, KEY_INIT);
    }
}

// Debug builds (-DNN_QREG_TIMELINE): thread 0 of every CTA stamps the phases of the query-register
// kernel with the global nanosecond timer; nn_bench --timeline prints where a small search's time goes.
#ifdef NN_QREG_TIMELINE
__device__ __forceinline__ unsigned long long tl_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define NN_TL(a, i)                                                                                         \
    do                                                                                                      \
    {                                                                                                       \
        if ((a).timeline && threadIdx.x == 0)                                                               \
        {                                                                                                   \
            unsigned long long *tl_ = (a).timeline + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16;    \
            tl_[(i)] = tl_now();                                                                            \
            tl_[8 + (i)] = (unsigned long long)clock64();                                                   \
        }                                                                                                   \
    } while (0)
#else
#define NN_TL(a, i) ((void)0)
#endif

// CTA index -> (reference split, query tile).  The query tiles are taken in groups of `qgroup`; inside
// a group the tile index runs fastest, the split next, and the groups follow one another.  qgroup = 1 is
// "split fastest" (consecutive CTAs = the splits of one tile), qgroup >= #tiles is "tile fastest".
__device__ __forceinline__ void cta_to_work(uint32_t splits, uint32_t qgroup, uint32_t &split, uint32_t &qtile)
{
    if (qgroup <= 1u)
    {
        split = blockIdx.x % splits;
        qtile = blockIdx.x / splits;
        return;
    }
    const uint32_t qtiles = gridDim.x / splits;
    const uint32_t per_group = qgroup * splits;
    const uint32_t grp = blockIdx.x / per_group, within = blockIdx.x - grp * per_group;
    const uint32_t in_grp = min(qgroup, qtiles - grp * qgroup); // the last group may be smaller
    split = within / in_grp;
    qtile = grp * qgroup + within % in_grp;
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

// Two queries against one reference, all packed: qp[i] = (qa_i, qb_i), r broadcast to both lanes
// (SASS: FADD2 Rd, Rq.F32x2.HI_LO, -Rr.F32).  Returns (d2(qa, r), d2(qb, r)).
// The square is written fma(d, d, nz) with nz = (-0.0f, -0.0f) supplied at run time through the
// kernel arguments: x*x + (-0) is x*x rounded once, bit for bit the IEEE product in every case
// (+0 + -0 = +0 under round-to-nearest; Inf and NaN propagate), and because it already IS an fma
// ptxas cannot contract it with the following add.  (A plain mul.rn.f32x2 feeding add.rn.f32x2 is
// contracted to FFMA2 by ptxas 12.9 even though `.rn` forbids it; that would break parity.)
template <int K>
__device__ __forceinline__ float2 sqdist_pair(const float2 (&qp)[K], const float *r, const float2 nz)
{
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < K; ++i)
    {
        const float2 d = __fadd2_rn(qp[i], make_float2(-r[i], -r[i]));
        const float2 p = __ffma2_rn(d, d, nz);
        acc = (i == 0) ? p : __fadd2_rn(acc, p);
    }
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

// 256-bit read-only global load (sm_100: LDG.E.256): one instruction and one L1 lookup per 32-byte
// sector instead of two 128-bit loads that each touch the sector.  `p` must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float *p, float *dst)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]), "=f"(dst[4]), "=f"(dst[5]), "=f"(dst[6]),
                   "=f"(dst[7])
                 : "l"(p));
}

// Asks the L2 to fetch `bytes` (multiple of 16) starting at `src` from HBM; no destination, no
// completion tracking.  One instruction per 16 KB moves the HBM latency out of the register loads.
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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
// Queries per thread of the WIDE query-register tile: as many as fit beside one reference group in
// a 128-register budget.  Narrower tiles (fewer queries per thread) serve small query counts.
template <int K>
struct QregDefault
{
    static constexpr int BUDGET = (96 - Geo<K>::G * K) / K;
    static constexpr int Q = BUDGET >= 8 ? 8 : (BUDGET >= 4 ? 4 : (BUDGET >= 2 ? 2 : 1));
};

template <int K>
struct QregCfg
{
    static constexpr int CH = (NN_QREG_CH8 && K == 3) ? 8 : 4;  // references per chunk
    static constexpr int TR = ((2048 / K) / 8) * 8;             // references per tile (~8 KB)
    static constexpr int STAGES = 3;
    static constexpr int TILE_FLOATS = TR * K;
    static constexpr uint32_t TILE_BYTES = TILE_FLOATS * 4u;
    // ring + barriers + slack: the software-pipelined chunk loop reads one reference group (at most
    // 240 bytes) past the end of the tile it is working on; the value is never used
    static constexpr size_t SMEM = (size_t)STAGES * TILE_BYTES + 64 + 256;
};

// Software pipelining of the shared-memory loads pays in the narrow tiles (Q <= 4) that small
// query counts use: few warps per scheduler, short chunks, so the load latency at the head of every
// chunk is exposed (ncu at k=3, m=1024, n=65536: FMA pipe 66% -> 75% of active cycles).  The wide
// tiles keep their registers for queries (measured: -1% at Q = 8).  It needs one more reference
// group of registers under the 128-register cap.
template <int K, int Q, int MATH>
struct QregPrefetch
{
    static constexpr bool value =
        NN_QREG_PREFETCH && Q <= 4 && (Q * K + 2 * Geo<K>::G * K + (MATH == 2 ? 52 : 36) <= 128);
};

// One chunk of CH references against the thread's Q queries; cm[] receives the chunk minima
// (ACC: is folded with them -- the caller starts it at +INF and reads it after several chunks).
// PF: `nxt` holds the chunk's first reference group on entry (loaded while the previous chunk was
// computed) and the first group of the FOLLOWING chunk on exit, so no shared-memory latency sits
// between two chunks.
template <int K, int Q, int MATH, bool PF, bool ACC = false>
__device__ __forceinline__ void qreg_chunk(const float *__restrict__ sm, const float (&q)[Q][K], float (&cm)[Q],
                                           const float2 nz, float4 (&nxt)[Geo<K>::F4])
{
    constexpr int G = Geo<K>::G, F4 = Geo<K>::F4, CH = QregCfg<K>::CH;
    float hold[Q];
#pragma unroll
    for (int g0 = 0; g0 < CH; g0 += G)
    {
        float grp[G * K];
        if constexpr (PF)
        {
#pragma unroll
            for (int i = 0; i < F4; ++i)
            {
                grp[4 * i + 0] = nxt[i].x;
                grp[4 * i + 1] = nxt[i].y;
                grp[4 * i + 2] = nxt[i].z;
                grp[4 * i + 3] = nxt[i].w;
            }
            const float4 *n4 = reinterpret_cast<const float4 *>(sm + (g0 + G) * K);
#pragma unroll
            for (int i = 0; i < F4; ++i)
                nxt[i] = n4[i];
        }
        else
        {
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
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
        {
            const int c = g0 + g;
            float dist[Q];
            if constexpr (MATH == 2)
            {
                static_assert(MATH != 2 || Q % 2 == 0, "pair-packed math needs an even number of queries per thread");
#pragma unroll
                for (int j = 0; j < Q; j += 2)
                {
                    float2 qp[K];
#pragma unroll
                    for (int i = 0; i < K; ++i)
                        qp[i] = make_float2(q[j][i], q[j + 1][i]);
                    const float2 d2 = sqdist_pair<K>(qp, &grp[g * K], nz);
                    dist[j] = d2.x;
                    dist[j + 1] = d2.y;
                }
            }
            else
            {
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    dist[j] = sqdist_par<K, MATH == 1>(q[j], &grp[g * K], (g * K) & 1);
            }
#pragma unroll
            for (int j = 0; j < Q; ++j)
            {
                if ((c & 1) == 0)
                    hold[j] = dist[j];
                else if (c == 1 && !ACC)
                    cm[j] = fminf(hold[j], dist[j]);
                else
                    cm[j] = fminf(fminf(cm[j], hold[j]), dist[j]);
            }
        }
    }
}

// CTAs per SM the register allocation is capped for: 4 (128 registers) when the query tile and one
// reference group fit comfortably, else 3 (168 registers) -- spilling costs more than occupancy here.
template <int K, int Q, int MATH>
constexpr int qreg_minb()
{
    return (NN_QREG_REGCAP_LOW || Q * K + Geo<K>::G * K + (MATH == 2 ? 52 : 36) <= 128) ? 4 : 3;
}

// SUPER: the (best, where) update runs once per super-chunk of a.super_chunks chunks instead of once per
// chunk (see the tile loop).  Two instantiations rather than a run-time branch: with both loop forms in
// one kernel ptxas took 160 registers at k = 16, Q = 4 (3 CTAs per SM instead of 4).
template <int K, int Q, int NT, int MATH, bool SUPER = false>
__global__ void __launch_bounds__(NT, qreg_minb<K, Q, MATH>()) nn_qreg_kernel(const QregArgs a)
{
    using C = QregCfg<K>;
    constexpr int CH = C::CH, TR = C::TR, STAGES = C::STAGES;
    constexpr bool PF = QregPrefetch<K, Q, MATH>::value;
    // Narrow tiles run with little work per CTA, where the query loads of the prologue and the
    // re-read of the winning chunks in the epilogue weigh (ncu at k=16, m=1024, n=65536: lg/mio
    // throttle and long-scoreboard stalls from 64..256 scalar loads per thread, each touching 32
    // sectors per warp): they use 128-bit loads there.  The widest tile keeps the scalar forms, whose
    // register footprint lets it run 4 CTAs per SM.
    constexpr bool VEC = Q < QregDefault<K>::Q;
    // chunks unrolled in the tile loop: the narrow tiles have short chunks (Q/2 * (3K-1) * CH packed
    // instructions), so the loop branch and the compare/select chain at the end of every chunk weigh
    // more; unrolling lets the next chunk's arithmetic cover them
    constexpr int UNR = (NN_QREG_UNROLL_NARROW && Q < QregDefault<K>::Q) ? (Q >= 4 ? 2 : 4) : kQregUnroll;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * C::TILE_BYTES);

    const int tid = threadIdx.x;
    NN_TL(a, 0);
    // CTA order.  Launch order is x fastest, then y: x = (reference split, query tile within its group),
    // y = group of a.qgroup query tiles.  The CTAs that are resident together then cover MANY query tiles
    // of a FEW reference splits, so a split is fetched from HBM once and served to the other tiles by the
    // L2.  With the splits of one tile consecutive instead (round 1), BASELINE config 4 -- 128 tiles x 37
    // splits = 8 waves -- re-streamed the set in every wave and read 55 GB from DRAM for a 1.07 GB
    // reference set (L2 hit rate 46%); grouped by 128 tiles: 1.6 GB (97%).  Same speed either way (the
    // kernel is FP32-bound) -- but HOW the two indices are computed matters: the same order written as one
    // division/modulo chain on a 1-D grid (NN_CTA_ORDER=1) ran 1.6% slower at configs 4 and 2 whatever
    // the group size (1496 vs 1473 ms), a pure code-generation effect on the unrolled loop that follows.
#if NN_CTA_ORDER == 0 // the original: split fastest, no grouping
    const uint32_t split = blockIdx.x % a.splits;
    const uint32_t qtile = blockIdx.x / a.splits;
#elif NN_CTA_ORDER == 2 // 2-D grid: x = (split, tile within its group), y = group of a.qgroup tiles
    const uint32_t split = blockIdx.x / a.qgroup;
    const uint32_t qtile = blockIdx.y * a.qgroup + blockIdx.x % a.qgroup;
    if (qtile * (uint32_t)(NT * Q) >= (uint32_t)a.m)
        return; // (the last group may hold fewer tiles)
#else
    uint32_t split, qtile;
    cta_to_work(a.splits, a.qgroup, split, qtile);
#endif
    // This CTA's references: [r0, r1).  r0 is a multiple of 8 points (a whole chunk, 16-byte aligned for every
    // k); whole chunks [r0, r1c) stream through the TMA ring in tiles of up to TR points, and the
    // ragged end of the reference set (< CH points, last split only) is handled after the loop.
    const uint32_t r0 = min(split * a.refs_per_split, a.n);
    const uint32_t r1 = (split == a.splits - 1) ? a.n : min(r0 + a.refs_per_split, a.n);
    const uint32_t r1c = r0 + ((r1 - r0) / CH) * CH;
    const uint32_t ntiles = (r1c - r0 + TR - 1) / TR;
    const int qt_begin = (int)(qtile * (NT * Q));
    if (r1 <= r0)
    { // (never planned: every split receives references) -- still counted by the finish protocol
        finish_group(a.fin, a.keys, qtile, a.splits, qt_begin, min(NT * Q, a.m - qt_begin));
        return;
    }

    auto issue = [&](uint32_t t, uint32_t stage) { // tile t of this CTA -> ring stage
        const uint32_t first = r0 + t * TR;
        const uint32_t bytes = min((uint32_t)TR, r1c - first) * (uint32_t)(K * 4);
        mbar_expect_tx(&full[stage], bytes);
        bulk_g2s(tiles + (size_t)stage * C::TILE_FLOATS, a.R + (size_t)first * K, bytes, &full[stage]);
    };
    // The first reference tiles are requested before anything else so that their HBM latency
    // overlaps the query loads (only thread 0 touches the barriers before the CTA-wide sync).
    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            mbar_init(&full[s], 1);
        mbar_fence_init();
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if ((uint32_t)s < ntiles)
                issue(s, s);
    }

    // this thread's queries (clamped so that out-of-range slots compute on a valid row)
    const bool s_align16 = (reinterpret_cast<uintptr_t>(a.S) & 15) == 0;
    const bool s_align8 = (reinterpret_cast<uintptr_t>(a.S) & 7) == 0;
    float q[Q][K];
    float best[Q];
    uint32_t bref[Q];
    const float2 nz = make_float2(a.neg_zero, a.neg_zero);
#pragma unroll
    for (int j = 0; j < Q; ++j)
    {
        int qi = (int)(qtile * (NT * Q)) + j * NT + tid;
        qi = qi < a.m ? qi : a.m - 1;
        const float *src = a.S + (size_t)qi * K;
        // rows of 4n (2n) floats are 16 (8) byte aligned when the query array is: vector loads cut the
        // load instructions -- each touches 32 different sectors per warp -- by 4 (2)
        if (VEC && K % 4 == 0 && s_align16)
        {
#pragma unroll
            for (int i = 0; i < K / 4; ++i)
            {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
                q[j][4 * i + 0] = v.x;
                q[j][4 * i + 1] = v.y;
                q[j][4 * i + 2] = v.z;
                q[j][4 * i + 3] = v.w;
            }
        }
        else if (VEC && K % 2 == 0 && s_align8)
        {
#pragma unroll
            for (int i = 0; i < K / 2; ++i)
            {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(src) + i);
                q[j][2 * i + 0] = v.x;
                q[j][2 * i + 1] = v.y;
            }
        }
        else
        {
#pragma unroll
            for (int i = 0; i < K; ++i)
                q[j][i] = __ldg(src + i);
        }
        best[j] = __int_as_float(0x7f800000);
        bref[j] = NO_REF;
    }
    __syncthreads();
    NN_TL(a, 1);

    const int sref = SUPER ? (int)max(a.super_chunks, 1u) * CH : CH; // references per (best, where) update
    uint32_t stage = 0, parity = 0;
    for (uint32_t t = 0; t < ntiles; ++t)
    {
        __syncthreads(); // everyone is done with tile t-1: its stage may be refilled
        if (tid == 0 && t + STAGES - 1 < ntiles)
            issue(t + STAGES - 1, (stage + STAGES - 1) % STAGES);
        mbar_wait(&full[stage], parity);
#ifdef NN_QREG_TIMELINE
        if (t == 0)
            NN_TL(a, 2);
#endif
        const float *sm = tiles + (size_t)stage * C::TILE_FLOATS;
        const uint32_t ref0 = r0 + t * TR;
        const int cnt = (int)min((uint32_t)TR, r1c - ref0);
        float4 nxt[Geo<K>::F4];
        if constexpr (PF)
        {
#pragma unroll
            for (int i = 0; i < Geo<K>::F4; ++i)
                nxt[i] = reinterpret_cast<const float4 *>(sm)[i];
        }
        // The strict-less update of (best, where) costs three instructions per query.  Long splits do it
        // once per SUPER-CHUNK of `sref` references (a.super_chunks chunks), across which the minimum
        // simply keeps accumulating.  It matters because the packed FP32 instructions hold the issue port
        // for two cycles, so EVERY other instruction costs the FMA pipe a cycle (ncu on all kernels of
        // this file: fma-pipe-active % + other-instructions-issued % = 100): at k = 16, Q = 4 a chunk is
        // 376 packed + 41 other instructions (94.8% pipe), 12 of them this update.  Short splits keep the
        // per-chunk update: their winner re-scan (one super-chunk per query) would cost more than it saves.
        if constexpr (SUPER)
        {
            for (int c0 = 0; c0 < cnt; c0 += sref)
            {
                const int c1 = min(cnt, c0 + sref);
                float cm[Q];
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    cm[j] = __int_as_float(0x7f800000);
#pragma unroll UNR
                for (int c = c0; c < c1; c += CH)
                    qreg_chunk<K, Q, MATH, PF, true>(sm + c * K, q, cm, nz, nxt);
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    if (cm[j] < best[j])
                    {
                        best[j] = cm[j];
                        bref[j] = ref0 + c0;
                    }
            }
        }
        else
        {
#pragma unroll UNR
            for (int c = 0; c < cnt; c += CH)
            {
                float cm[Q];
                qreg_chunk<K, Q, MATH, PF>(sm + c * K, q, cm, nz, nxt);
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    if (cm[j] < best[j])
                    {
                        best[j] = cm[j];
                        bref[j] = ref0 + c;
                    }
            }
        }
        if (++stage == STAGES)
        {
            stage = 0;
            parity ^= 1;
        }
    }

    NN_TL(a, 3);
    if (r1c < r1)
    {
        // ragged end of the reference set (1..CH-1 points): plain loads; the chunk is padded with NaN
        // (a NaN distance never wins)
        __syncthreads();
        const float *src = a.R + (size_t)r1c * K;
        const uint32_t have = (r1 - r1c) * K;
        for (uint32_t i = tid; i < (uint32_t)(CH * K); i += NT)
            tiles[i] = (i < have) ? __ldg(src + i) : __int_as_float(0x7fffffff);
        __syncthreads();
        float cm[Q];
        float4 unused[Geo<K>::F4];
        qreg_chunk<K, Q, MATH, false>(tiles, q, cm, nz, unused);
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (cm[j] < best[j])
            {
                best[j] = cm[j];
                bref[j] = r1c;
            }
    }

    // resolve the exact (lowest) index inside the winning chunk, then fold into the global keys.
    // A CTA whose whole range went through the ring without re-using a stage (small problems: BASELINE
    // config 1 is 448 references per CTA) still holds every chunk in shared memory and re-reads the
    // winner there; otherwise it comes from global memory (L2), whose latency -- two dependent round
    // trips at Q = 2 -- was 12% of the warps' time at config 1 (ncu source view, r01_cfg1_qreg).
    const bool in_ring = ntiles <= (uint32_t)STAGES && r1c == r1;
#pragma unroll
    for (int j = 0; j < Q; ++j)
    {
        const int qi = (int)(qtile * (NT * Q)) + j * NT + tid;
        if (bref[j] != NO_REF && qi < a.m)
        {
            uint32_t idx = bref[j];
            if constexpr (SUPER)
            {
                // The winner is the lowest index at distance best[j] inside the super-chunk that first reached
                // it: [bref, bref + sref), cut at the end of this CTA's range (the window may run into later
                // chunks of the range -- their references are no closer, and a tie there has a higher index).
                // Scanned from the top down so that the lowest index is the last one written.
                const uint32_t wn = min(bref[j] + (uint32_t)sref, r1) - bref[j];
                constexpr int G = Geo<K>::G, F4 = Geo<K>::F4;
                if (in_ring || (VEC && wn % CH == 0))
                {
                    // whole chunks: they start on a 16-byte boundary (chunk starts are multiples of 4 points),
                    // so they are re-read group by group with 128-bit loads -- from the ring when the CTA's
                    // whole range is still there, else from global memory (L2)
                    const float4 *p4 = in_ring ? reinterpret_cast<const float4 *>(tiles + (size_t)(bref[j] - r0) * K)
                                               : reinterpret_cast<const float4 *>(a.R + (size_t)bref[j] * K);
                    for (int cc = (int)wn - CH; cc >= 0; cc -= CH)
                    {
#pragma unroll
                        for (int g0 = CH - G; g0 >= 0; g0 -= G)
                        {
                            float grp[G * K];
#pragma unroll
                            for (int i = 0; i < F4; ++i)
                            {
                                const float4 *src = p4 + ((cc + g0) / G) * F4 + i;
                                const float4 v = in_ring ? *src : __ldg(src);
                                grp[4 * i + 0] = v.x;
                                grp[4 * i + 1] = v.y;
                                grp[4 * i + 2] = v.z;
                                grp[4 * i + 3] = v.w;
                            }
#pragma unroll
                            for (int g = G - 1; g >= 0; --g)
                            {
                                const float d = sqdist<K, 0, false>(q[j], &grp[g * K]);
                                if (d == best[j])
                                    idx = bref[j] + cc + g0 + g;
                            }
                        }
                    }
                }
                else
                {
                    for (int c = (int)wn - 1; c >= 0; --c)
                    {
                        const uint32_t r = bref[j] + c;
                        const float d = sqdist_gmem<K>(q[j], a.R + (size_t)r * K);
                        if (d == best[j])
                            idx = r;
                    }
                }
            }
            else
            {
                if (in_ring)
                {
                    constexpr int G = Geo<K>::G, F4 = Geo<K>::F4;
                    const float4 *p4 = reinterpret_cast<const float4 *>(tiles + (size_t)(bref[j] - r0) * K);
#pragma unroll
                    for (int g0 = CH - G; g0 >= 0; g0 -= G)
                    {
                        float grp[G * K];
#pragma unroll
                        for (int i = 0; i < F4; ++i)
                        {
                            const float4 v = p4[(g0 / G) * F4 + i];
                            grp[4 * i + 0] = v.x;
                            grp[4 * i + 1] = v.y;
                            grp[4 * i + 2] = v.z;
                            grp[4 * i + 3] = v.w;
                        }
#pragma unroll
                        for (int g = G - 1; g >= 0; --g)
                        {
                            const float d = sqdist<K, 0, false>(q[j], &grp[g * K]);
                            if (d == best[j])
                                idx = bref[j] + g0 + g;
                        }
                    }
                }
                else if (VEC && bref[j] + CH <= a.n)
                {
                    // whole chunk inside the set: it starts on a 16-byte boundary (chunk starts are
                    // multiples of 4 points), so it is re-read group by group with 128-bit loads -- a
                    // quarter of the load instructions, each of which touches 32 sectors per warp
                    constexpr int G = Geo<K>::G, F4 = Geo<K>::F4;
                    const float4 *p4 = reinterpret_cast<const float4 *>(a.R + (size_t)bref[j] * K);
#pragma unroll
                    for (int g0 = CH - G; g0 >= 0; g0 -= G)
                    {
                        float grp[G * K];
#pragma unroll
                        for (int i = 0; i < F4; ++i)
                        {
                            const float4 v = __ldg(p4 + (g0 / G) * F4 + i);
                            grp[4 * i + 0] = v.x;
                            grp[4 * i + 1] = v.y;
                            grp[4 * i + 2] = v.z;
                            grp[4 * i + 3] = v.w;
                        }
#pragma unroll
                        for (int g = G - 1; g >= 0; --g)
                        {
                            const float d = sqdist<K, 0, false>(q[j], &grp[g * K]);
                            if (d == best[j])
                                idx = bref[j] + g0 + g;
                        }
                    }
                }
                else
                {
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
                }
            }
            fold_key(a.keys + qi, pack_key(best[j], a.index_base + idx), a.peer_keys);
        }
    }
    NN_TL(a, 4);
    finish_group(a.fin, a.keys, qtile, a.splits, qt_begin, min(NT * Q, a.m - qt_begin));
    NN_TL(a, 5);
}

// =============================================================================================
// Kernel A' -- "phased query-register" kernel, for query counts that do not fill 128-query tiles
// (9 .. a few hundred queries).
//
// Kernel A gives every thread Q queries and lets all 128 threads walk the SAME references, so a
// search of 100 queries either computes on a padded 128-query tile (and, at one query per thread,
// without the pair-packed math) or -- in the reference-stream kernel C -- re-streams the reference
// set once per 8 queries.  Here the 128 threads of a CTA are laid out as NG query groups x NP
// PHASES (NG * NP <= 128, both chosen by the host for the query count): thread (g, p) keeps the Q
// queries of group g in registers, exactly as in kernel A, but only visits the reference groups
// p, p + NP, p + 2 NP, ... of each tile.  100 queries are then 25 groups of 4 queries x 5 phases = 125
// busy lanes of 128 with the all-packed math of kernel A and ONE pass over the reference set.
// Same TMA ring, same chunk-minimum / late index resolution as kernel A; a thread's chunk is CH/G
// groups NP groups apart (consecutive phases read consecutive 16-byte-aligned groups, which keeps
// the lanes of a warp that belong to different phases out of each other's shared-memory banks).
// The phases of a query meet in shared memory at the end (the ring is reused as a [Q][128] key
// array), so a CTA still issues one atomicMin per query.
// =============================================================================================
constexpr int kFlexMaxStages = 8;
constexpr int kFlexRingOffset = 128; // barriers first (a fixed place whatever the ring), then the ring
constexpr int kFlexMaxSmem = 56 * 1024; // per CTA: four CTAs per SM stay resident

template <int K, int Q>
__device__ __forceinline__ void qflex_chunk(const float *__restrict__ sm, const uint32_t gstride,
                                            const float (&q)[Q][K], float (&cm)[Q], const float2 nz,
                                            float4 (&nxt)[Geo<K>::F4], const bool more)
{
    constexpr int G = Geo<K>::G, F4 = Geo<K>::F4, CHG = QregCfg<K>::CH / Geo<K>::G;
    static_assert(Q % 2 == 0, "pair-packed math needs an even number of queries per thread");
    float hold[Q];
#pragma unroll
    for (int u = 0; u < CHG; ++u)
    {
        float grp[G * K];
#pragma unroll
        for (int i = 0; i < F4; ++i)
        {
            grp[4 * i + 0] = nxt[i].x;
            grp[4 * i + 1] = nxt[i].y;
            grp[4 * i + 2] = nxt[i].z;
            grp[4 * i + 3] = nxt[i].w;
        }
        // the thread's next group is loaded while this one is computed; `more` is false only for the
        // last chunk of a tile, whose successor would lie past the tile
        if (u + 1 < CHG || more)
        {
            const float4 *n4 = reinterpret_cast<const float4 *>(sm + (size_t)(u + 1) * gstride);
#pragma unroll
            for (int i = 0; i < F4; ++i)
                nxt[i] = n4[i];
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
        {
            const int c = u * G + g;
            float dist[Q];
#pragma unroll
            for (int j = 0; j < Q; j += 2)
            {
                float2 qp[K];
#pragma unroll
                for (int i = 0; i < K; ++i)
                    qp[i] = make_float2(q[j][i], q[j + 1][i]);
                const float2 d2 = sqdist_pair<K>(qp, &grp[g * K], nz);
                dist[j] = d2.x;
                dist[j + 1] = d2.y;
            }
#pragma unroll
            for (int j = 0; j < Q; ++j)
            {
                if ((c & 1) == 0)
                    hold[j] = dist[j];
                else if (c == 1)
                    cm[j] = fminf(hold[j], dist[j]);
                else
                    cm[j] = fminf(fminf(cm[j], hold[j]), dist[j]);
            }
        }
    }
}

// CTAs per SM the phased kernel is compiled for: 5 where the thread's queries and two reference groups
// leave room under 102 registers (small k), else what the query-register kernel uses.
template <int K, int Q>
constexpr int qflex_minb()
{
    return (NN_QFLEX_MINB5 && Q * K + 2 * Geo<K>::G * K + 52 <= 102) ? 5 : qreg_minb<K, Q, 2>();
}

template <int K, int Q, int NT>
__global__ void __launch_bounds__(NT, qflex_minb<K, Q>()) nn_qflex_kernel(const QflexArgs a)
{
    using C = QregCfg<K>;
    constexpr int G = Geo<K>::G, F4 = Geo<K>::F4, CH = C::CH, CHG = CH / G;
    // The ring geometry is a launch parameter: layouts with many phases give every thread only one or two
    // chunks per 8 KB tile, so the per-tile barrier and the TMA latency (two tiles in flight) showed
    // (time = compute + ~0.6 x HBM time); they get more and larger stages (host: flex_ring).
    const uint32_t STAGES = a.stages;
    // (Q = 4: not unrolled for k = 5..8 -- 25-30 fewer registers, k = 5: 0.74 -> 0.80 of the roofline at
    // m = 100, k = 8: 0.805 -> 0.818; k <= 4 and k >= 9 measured 0.5-1% better with two chunks per trip)
    constexpr int UNR = Q >= 8 ? NN_QFLEX_UNROLL_Q8
                               : (Q >= 4 ? ((NN_QFLEX_UNROLL_Q4 == 2 && K >= 5 && K <= 8) ? 1 : NN_QFLEX_UNROLL_Q4)
                                         : NN_QFLEX_UNROLL_Q2);
    static_assert(CH % 2 == 0 && CH % G == 0, "chunks are whole groups and an even number of points");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);         // up to kFlexMaxStages barriers
    float *tiles = reinterpret_cast<float *>(smem_raw + kFlexRingOffset); // stages x stage_floats (>= NT*Q*8 bytes in all)

    const int tid = threadIdx.x;
    // CTA order: see nn_qreg_kernel
#if NN_CTA_ORDER == 0 // the original: split fastest, no grouping
    const uint32_t split = blockIdx.x % a.splits;
    const uint32_t qtile = blockIdx.x / a.splits;
#elif NN_CTA_ORDER == 2 // 2-D grid: x = (split, tile within its group), y = group of a.qgroup tiles
    const uint32_t split = blockIdx.x / a.qgroup;
    const uint32_t qtile = blockIdx.y * a.qgroup + blockIdx.x % a.qgroup;
    if (qtile * a.tile_queries >= (uint32_t)a.m)
        return; // (the last group may hold fewer tiles)
#else
    uint32_t split, qtile;
    cta_to_work(a.splits, a.qgroup, split, qtile);
#endif
    const uint32_t ng = a.ng, np = a.np;
    // lanes beyond NG*NP shadow the last phase of group 0 (same addresses: broadcast) and publish nothing
    const bool live = (uint32_t)tid < ng * np;
    const uint32_t g = live ? (uint32_t)tid % ng : 0u;
    const uint32_t p = live ? (uint32_t)tid / ng : np - 1u;
    const uint32_t round = np * CH;                  // references one step of all phases covers
    const uint32_t gstride = np * (uint32_t)(G * K); // floats between two groups of one thread
    const uint32_t tile_refs = a.tile_groups * G;    // per ring stage; a whole number of rounds

    const uint32_t r0 = min(split * a.refs_per_split, a.n);
    const uint32_t r1 = (split == a.splits - 1) ? a.n : min(r0 + a.refs_per_split, a.n);
    const uint32_t r1c = r0 + ((r1 - r0) / round) * round; // whole rounds stream through the ring
    const uint32_t ntiles = (r1c - r0 + tile_refs - 1) / tile_refs;
    const int qt_begin = (int)(qtile * a.tile_queries);
    const int qt_count = min((int)a.tile_queries, a.m - qt_begin);
    if (r1 <= r0)
    {
        finish_group(a.fin, a.keys, qtile, a.splits, qt_begin, qt_count);
        return;
    }

    auto issue = [&](uint32_t t, uint32_t stage) {
        const uint32_t first = r0 + t * tile_refs;
        const uint32_t bytes = min(tile_refs, r1c - first) * (uint32_t)(K * 4);
        mbar_expect_tx(&full[stage], bytes);
        bulk_g2s(tiles + (size_t)stage * a.stage_floats, a.R + (size_t)first * K, bytes, &full[stage]);
    };
    if (tid == 0)
    {
        for (uint32_t s = 0; s < STAGES; ++s)
            mbar_init(&full[s], 1);
        mbar_fence_init();
        for (uint32_t s = 0; s + 1 < STAGES; ++s)
            if (s < ntiles)
                issue(s, s);
    }

    // this thread's queries: local query j*NG + g of the tile (clamped; surplus slots publish nothing)
    float q[Q][K];
    float best[Q];
    uint32_t bref[Q];
    const float2 nz = make_float2(a.neg_zero, a.neg_zero);
#pragma unroll
    for (int j = 0; j < Q; ++j)
    {
        const int lq = min(j * (int)ng + (int)g, qt_count - 1);
        const float *src = a.S + (size_t)(qt_begin + lq) * K;
#pragma unroll
        for (int i = 0; i < K; ++i)
            q[j][i] = __ldg(src + i);
        best[j] = __int_as_float(0x7f800000);
        bref[j] = NO_REF;
    }
    __syncthreads();

    uint32_t stage = 0, parity = 0;
    for (uint32_t t = 0; t < ntiles; ++t)
    {
        __syncthreads(); // everyone is done with tile t-1: its stage may be refilled
        if (tid == 0 && t + STAGES - 1 < ntiles)
            issue(t + STAGES - 1, (stage + STAGES - 1) % STAGES);
        mbar_wait(&full[stage], parity);
        const uint32_t ref0 = r0 + t * tile_refs;
        const uint32_t nch = min(tile_refs, r1c - ref0) / round; // chunks per thread in this tile
        const float *sm = tiles + (size_t)stage * a.stage_floats + (size_t)p * (G * K);
        float4 nxt[F4];
#pragma unroll
        for (int i = 0; i < F4; ++i)
            nxt[i] = reinterpret_cast<const float4 *>(sm)[i];
#pragma unroll UNR
        for (uint32_t c = 0; c < nch; ++c)
        {
            float cm[Q];
            qflex_chunk<K, Q>(sm + (size_t)c * CHG * gstride, gstride, q, cm, nz, nxt, c + 1 < nch);
#pragma unroll
            for (int j = 0; j < Q; ++j)
                if (cm[j] < best[j])
                {
                    best[j] = cm[j];
                    bref[j] = ref0 + (p + np * c * CHG) * G;
                }
        }
        if (++stage == STAGES)
        {
            stage = 0;
            parity ^= 1;
        }
    }

    if (r1c < r1)
    {
        // ragged end of the range (less than one round): plain loads, padded with NaN up to a whole
        // round (a NaN distance never wins)
        __syncthreads();
        const float *src = a.R + (size_t)r1c * K;
        const uint32_t have = (r1 - r1c) * K;
        for (uint32_t i = tid; i < round * K; i += NT)
            tiles[i] = (i < have) ? __ldg(src + i) : __int_as_float(0x7fffffff);
        __syncthreads();
        const float *sm = tiles + (size_t)p * (G * K);
        float4 nxt[F4];
#pragma unroll
        for (int i = 0; i < F4; ++i)
            nxt[i] = reinterpret_cast<const float4 *>(sm)[i];
        float cm[Q];
        qflex_chunk<K, Q>(sm, gstride, q, cm, nz, nxt, false);
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (cm[j] < best[j])
            {
                best[j] = cm[j];
                bref[j] = r1c + p * G;
            }
    }

    // The phases of every query meet in shared memory (the ring becomes best[Q][NT], chunk[Q][NT]).
    // One thread per query then takes the minimum over the phases and resolves the exact (lowest)
    // index ONLY in the winning chunk(s): groups chunk + u*NP*G, points g, re-read from global memory
    // with v0's arithmetic.  (Every thread resolving its own Q candidates first cost Q dependent L2
    // round trips per thread -- 8 at Q = 8 -- for candidates of which all but one per query lose.)
    __syncthreads();
    float *xb = tiles;
    uint32_t *xr = reinterpret_cast<uint32_t *>(xb + Q * NT);
#pragma unroll
    for (int j = 0; j < Q; ++j)
    {
        xb[j * NT + tid] = best[j];
        xr[j * NT + tid] = live ? bref[j] : NO_REF;
    }
    __syncthreads();
    for (uint32_t s = tid; s < ng * Q; s += NT)
    {
        if ((int)s >= qt_count) // local query s = j*NG + gg
            continue;
        const uint32_t j = s / ng, gg = s % ng;
        float bmin = __int_as_float(0x7f800000);
        bool any = false;
        for (uint32_t pp = 0; pp < np; ++pp)
            if (xr[j * NT + pp * ng + gg] != NO_REF)
            {
                bmin = fminf(bmin, xb[j * NT + pp * ng + gg]);
                any = true;
            }
        if (!any)
            continue; // nothing beat v0's start state in this CTA's range
        float qv[K];
#pragma unroll
        for (int i = 0; i < K; ++i)
            qv[i] = __ldg(a.S + (size_t)(qt_begin + (int)s) * K + i);
        uint32_t idx = NO_REF;
        for (uint32_t pp = 0; pp < np; ++pp)
        {
            const uint32_t c0 = xr[j * NT + pp * ng + gg];
            if (c0 == NO_REF || xb[j * NT + pp * ng + gg] != bmin)
                continue;
#pragma unroll
            for (int u = 0; u < CHG; ++u)
            {
#pragma unroll
                for (int pt = 0; pt < G; ++pt)
                {
                    const uint32_t r = c0 + (uint32_t)u * np * G + pt;
                    if (r < a.n)
                    {
                        const float d = sqdist_gmem<K>(qv, a.R + (size_t)r * K);
                        if (d == bmin)
                            idx = min(idx, r);
                    }
                }
            }
        }
        fold_key(a.keys + qt_begin + (int)s, pack_key(bmin, a.index_base + idx), a.peer_keys);
    }
    finish_group(a.fin, a.keys, qtile, a.splits, qt_begin, qt_count);
}

// =============================================================================================
// Kernel B -- "reference-register" kernel, for few queries (HBM-streaming / SM-fill bound).
//
// The roles are swapped: every thread streams its OWN references from HBM straight into registers
// (128-bit loads from the native AoS layout) and the MQ queries of the pass are broadcast from
// shared memory, stored there as interleaved PAIRS (qa_d, qb_d) so that the whole distance
// computation is packed f32x2 with the thread's reference coordinate as the broadcast operand.
// A persistent grid (SMs x occupancy CTAs) strides over the reference set, so all SMs are busy
// even for m = 1.
//
// At k = 8, m = 8 the path needs the FP32 pipe AND the HBM stream near their peaks at the same
// time, so the loads must be in flight continuously.  The thread's reference registers form a RING
// of NS slots (PS groups each): as soon as a slot has been consumed its loads for NS slots ahead
// are issued, so NS-1 slots are always in flight (a two-buffer scheme only has one buffer in
// flight for part of the time; ncu showed long-scoreboard stalls dominating).  Measured dead ends,
// both slower than plain LDG on B200: a cp.async.bulk (TMA) ring per CTA (5.2 TB/s at m = 1) and a
// warp-private cp.async ring (3.3 TB/s at m = 1, any depth or occupancy) versus 6.2 TB/s here.
//
// Per thread and query: minimum over a ring round (FMNMX3) + strict-less select on (best, round);
// the exact index is resolved at the end, keys are reduced across the warp with __shfl_xor and
// folded with one 64-bit atomicMin per warp and query.
// SOA = true reads references from the repacked [k][n] layout instead (coalesced 32-bit loads).
// MQ is even; a pass over an odd tail of queries duplicates its last query (valid_q masks it).
// =============================================================================================
template <int K, int MQ, int PS, int NS, int NT, bool SOA, int MINB>
__global__ void __launch_bounds__(NT, MINB) nn_rreg_kernel(const RregArgs a)
{
    static_assert(MQ % 2 == 0, "queries are processed in pairs");
    constexpr int G = SOA ? 1 : Geo<K>::G;
    constexpr int F4 = G * K / 4; // (AoS only) float4 per group
    constexpr int P = G * PS;     // references per thread per slot
    constexpr int NP = MQ / 2;    // query pairs
    __shared__ __align__(16) float2 sq[NP * K]; // sq[pair*K + d] = (qa_d, qb_d)
    __shared__ unsigned long long wkeys[NT / 32][MQ]; // per-warp results, merged per CTA at the end

    const int tid = threadIdx.x;
    const int pass = blockIdx.y;
    const int q0 = pass * MQ;                     // first query of this pass (relative to a.S)
    const int valid_q = min(MQ, a.mq_total - q0); // >= 1
    for (int i = tid; i < NP * K; i += NT)
    {
        const int pr = i / K, d = i % K;
        const int qa = min(2 * pr, valid_q - 1), qb = min(2 * pr + 1, valid_q - 1);
        sq[i] = make_float2(__ldg(a.S + (size_t)(q0 + qa) * K + d), __ldg(a.S + (size_t)(q0 + qb) * K + d));
    }
    __syncthreads();

    float best[MQ];
    uint32_t bref[MQ];
    const float2 nz = make_float2(a.neg_zero, a.neg_zero);
#pragma unroll
    for (int j = 0; j < MQ; ++j)
    {
        best[j] = __int_as_float(0x7f800000);
        bref[j] = NO_REF;
    }

    // Slot-batch sb covers groups [sb*NT*PS, (sb+1)*NT*PS); thread t owns groups sb*NT*PS + i*NT + t.
    // This CTA's j-th slot-batch is blockIdx.x + j*gridDim.x and lives in ring slot j % NS.
    const uint32_t ngroups = (a.n + G - 1) / G;
    const uint32_t groups_per_sb = NT * PS;
    const uint32_t nsb = (ngroups + groups_per_sb - 1) / groups_per_sb;
    const uint32_t mine = blockIdx.x < nsb ? (nsb - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    float ring[NS][P * K];
    const bool wide = (reinterpret_cast<uintptr_t>(a.R) & 31) == 0; // 256-bit loads need 32-byte alignment

    auto load_slot = [&](uint32_t j, float(&dst)[P * K]) {
        const uint32_t sb = blockIdx.x + j * gridDim.x;
#pragma unroll
        for (int i = 0; i < PS; ++i)
        {
            const uint32_t grp = sb * groups_per_sb + i * NT + tid;
            const uint32_t r0 = grp * G;
            if constexpr (SOA)
            {
#pragma unroll
                for (int d = 0; d < K; ++d)
                    dst[i * K + d] = (r0 < a.n) ? __ldg(a.R + (size_t)d * a.n + r0) : __int_as_float(0x7fffffff);
            }
            else if (r0 + G <= a.n)
            {
                if (NN_RREG_LDG256 && (G * K) % 8 == 0 && wide)
                {
#pragma unroll
                    for (int f = 0; f < (G * K) / 8; ++f)
                        ldg256(a.R + (size_t)r0 * K + 8 * f, &dst[i * G * K + 8 * f]);
                }
                else
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
            }
            else
            { // ragged end of the reference set: slots past n are NaN (a NaN distance never wins)
#pragma unroll
                for (int e = 0; e < G * K; ++e)
                    dst[i * G * K + e] =
                        ((size_t)r0 * K + e < (size_t)a.n * K) ? __ldg(a.R + (size_t)r0 * K + e) : __int_as_float(0x7fffffff);
            }
        }
    };

    // fold one slot into the running round minima rm[]
    auto compute_slot = [&](const float(&ref)[P * K], float (&rm)[MQ]) {
#pragma unroll
        for (int pr = 0; pr < NP; ++pr)
        {
            float2 qp[K];
            if constexpr (K % 2 == 0)
            {
                const float4 *q4 = reinterpret_cast<const float4 *>(sq + pr * K);
#pragma unroll
                for (int f = 0; f < K / 2; ++f)
                {
                    const float4 v = q4[f];
                    qp[2 * f] = make_float2(v.x, v.y);
                    qp[2 * f + 1] = make_float2(v.z, v.w);
                }
            }
            else
            {
#pragma unroll
                for (int d = 0; d < K; ++d)
                    qp[d] = sq[pr * K + d];
            }
            float2 hold = make_float2(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < P; ++p)
            {
                const float2 d = sqdist_pair<K>(qp, &ref[p * K], nz);
                if ((p & 1) == 0 && p + 1 < P)
                    hold = d;
                else if ((p & 1) == 1)
                {
                    rm[2 * pr] = fminf(fminf(rm[2 * pr], hold.x), d.x);
                    rm[2 * pr + 1] = fminf(fminf(rm[2 * pr + 1], hold.y), d.y);
                }
                else
                {
                    rm[2 * pr] = fminf(rm[2 * pr], d.x);
                    rm[2 * pr + 1] = fminf(rm[2 * pr + 1], d.y);
                }
            }
        }
    };

#pragma unroll
    for (int s = 0; s < NS; ++s)
        if ((uint32_t)s < mine)
            load_slot(s, ring[s]);

    for (uint32_t base = 0; base < mine; base += NS)
    {
        float rm[MQ];
#pragma unroll
        for (int j = 0; j < MQ; ++j)
            rm[j] = __int_as_float(0x7f800000);
#pragma unroll
        for (int s = 0; s < NS; ++s)
        {
            const uint32_t j = base + s;
            if (j < mine)
            {
                if (NN_RREG_L2PF > 0 && !SOA && tid == 0 && j + NS + NN_RREG_L2PF < mine)
                { // HBM -> L2 for the slot-batch this CTA will load into registers L2PF steps from now
                    const uint32_t sbp = blockIdx.x + (j + NS + NN_RREG_L2PF) * gridDim.x;
                    const size_t f0 = (size_t)sbp * groups_per_sb * G * K;
                    if (f0 + (size_t)groups_per_sb * G * K <= (size_t)a.n * K)
                        bulk_prefetch_l2(a.R + f0, groups_per_sb * G * K * 4u);
                }
                compute_slot(ring[s], rm);
                if (j + NS < mine)
                    load_slot(j + NS, ring[s]); // refill at once: NS-1 slots stay in flight
            }
        }
#pragma unroll
        for (int j = 0; j < MQ; ++j)
            if (rm[j] < best[j])
            {
                best[j] = rm[j];
                bref[j] = base;
            }
    }

    // Warp-level merge.  Only lanes whose minimum equals the warp minimum can hold the winner (several
    // may, on ties), so only they resolve their exact index -- lowest index within the thread's winning
    // round (slot-batches ascend with s, groups with i) -- by re-reading that round; the others skip
    // the re-read entirely (it would otherwise cost ~7% extra HBM traffic at k = 8, m = 8).
    const int lane = tid & 31;
#pragma unroll
    for (int j = 0; j < MQ; ++j)
    {
        float wmin = best[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            wmin = fminf(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
        unsigned long long key = KEY_INIT | NO_REF;
        if (bref[j] != NO_REF && best[j] == wmin)
        {
            float qv[K];
#pragma unroll
            for (int d = 0; d < K; ++d)
                qv[d] = (j & 1) ? sq[(j / 2) * K + d].y : sq[(j / 2) * K + d].x;
            uint32_t idx = 0;
            for (int s = NS - 1; s >= 0; --s)
            {
                const uint32_t jj = bref[j] + s;
                if (jj >= mine)
                    continue;
                const uint32_t sb = blockIdx.x + jj * gridDim.x;
#pragma unroll
                for (int i = PS - 1; i >= 0; --i)
                {
#pragma unroll
                    for (int g = G - 1; g >= 0; --g)
                    {
                        const uint32_t r = (sb * groups_per_sb + i * NT + tid) * G + g;
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
            }
            key = pack_key(best[j], a.index_base + idx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if (lane == 0)
            wkeys[tid >> 5][j] = key;
    }
    __syncthreads(); // one atomicMin per CTA and query (see nn_rtma_kernel)
    if (tid < MQ && tid < valid_q)
    {
        unsigned long long kmin = wkeys[0][tid];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w)
            kmin = wkeys[w][tid] < kmin ? wkeys[w][tid] : kmin;
        if (kmin < (KEY_INIT | NO_REF))
            fold_key(a.keys + q0 + tid, kmin, a.peer_keys);
    }
    finish_group(a.fin, a.keys, (uint32_t)pass, gridDim.x, q0, valid_q, a.q_first + q0);
}

// =============================================================================================
// Kernel C -- "reference-stream" kernel: the few-query case with the HBM stream decoupled from the
// register file.
//
// Same roles as kernel B (thread <-> its own references, query pairs broadcast from shared memory,
// all-packed f32x2 math), but the references reach the SM through a STAGES-deep shared-memory ring
// filled by the TMA unit: a producer warp issues one cp.async.bulk per tile (TILE_BYTES contiguous
// bytes of the native AoS array) signalled on a `full` mbarrier; each of the NW consumer warps
// copies ITS references of the tile from shared memory into registers (128-bit LDS), releases the
// stage at once on the `empty` mbarrier and only then does the arithmetic.  So the bytes in flight
// per SM are (STAGES-1) x TILE_BYTES x CTAs/SM regardless of the register budget, HBM is read in
// large fully-used bursts, and no warp ever waits on a global load inside the math.
// (ncu on kernel B at k = 8, m = 8: 27% more DRAM sectors than the reference set holds, FMA pipe
// 63% active with long-scoreboard the top stall -- profiles/r01_ncu_summary.txt.)
// Measured alternatives that did not win at k = 8, m = 8, n = 2^26 (0.409 ms here): two CTAs of 8+1
// warps per SM (0.43-0.44 ms), 16+1 warps (0.42), query pairs held in registers (0.44, spills at
// 12 warps), and per-warp private rings without a producer warp (each warp fetching its own 4 KB
// slice: 0.407 ms at 16 warps, but 3-4% slower at k = 3 and k = 16), and walking the passes of a
// many-query launch inside the CTA without draining the ring (grid.y = 1): the extra loop level cost
// the hot loop 8% at one pass and gained nothing at 13.
// Round 2: 16 queries per pass (MQ = 16 fits 128 registers with 4-16 bytes of spills): 3-6% faster where
// the query count is a multiple of 16 (k=8, m=32, n=2^22: 133 -> 127 us), but 3-7% SLOWER with a padded
// tail (m = 9, 24), at small n (n=2^16: 15.9 -> 20.8 us) and 1% slower at n = 2^26 -- not kept.
// Tile t of the reference set goes to CTA t % gridDim.x (persistent grid); the ragged end of the set
// (n % TILE_REFS references) is read with plain loads by the CTA whose turn it is.
// =============================================================================================
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// One thread's reference group (F4 float4) out of a shared-memory tile.  When consecutive lanes'
// groups are 64 bytes apart modulo 128 (k = 16: one 64-byte point per lane) the four 16-byte chunks
// of lanes l and l+2 fall into the same banks: a 4-way conflict on every LDS.128, which made the
// shared-memory pipe as busy as the FMA pipe.  There each lane reads its chunks in an order rotated
// by (lane/2)%4, which spreads a quarter-warp over all eight 16-byte bank groups, and rotates the
// registers back with two select stages (ALU pipe, which has room).
template <int K>
__device__ __forceinline__ void lds_group(const float4 *__restrict__ p4, float *dst, const int lane)
{
    constexpr int F4 = Geo<K>::F4;
    constexpr bool ROT = NN_RTMA_ROTATE && (F4 == 4) && ((F4 * 16) % 128 == 64);
    if constexpr (ROT)
    {
        const int rot = (lane >> 1) & 3;
        float4 v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            v[c] = p4[(c + rot) & 3]; // v[c] = chunk (c + rot) & 3
        const bool r1 = rot & 1, r2 = rot & 2;
        float4 t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        { // undo the odd part: t[j] = chunk (j + (rot & 2)) & 3
            const float4 a = v[j], b = v[(j + 3) & 3];
            t[j] = make_float4(r1 ? b.x : a.x, r1 ? b.y : a.y, r1 ? b.z : a.z, r1 ? b.w : a.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
        { // undo the even part: chunk j
            const float4 a = t[j], b = t[(j + 2) & 3];
            dst[4 * j + 0] = r2 ? b.x : a.x;
            dst[4 * j + 1] = r2 ? b.y : a.y;
            dst[4 * j + 2] = r2 ? b.z : a.z;
            dst[4 * j + 3] = r2 ? b.w : a.w;
        }
    }
    else if constexpr (NN_RTMA_ROTATE2 && F4 == 2)
    { // k = 8: 32-byte points, lanes l and l+4 collide (2-way): lanes 4..7 of every 8 swap the halves
        const bool sw = (lane >> 2) & 1;
        const float4 a = p4[sw ? 1 : 0], b = p4[sw ? 0 : 1];
        dst[0] = sw ? b.x : a.x;
        dst[1] = sw ? b.y : a.y;
        dst[2] = sw ? b.z : a.z;
        dst[3] = sw ? b.w : a.w;
        dst[4] = sw ? a.x : b.x;
        dst[5] = sw ? a.y : b.y;
        dst[6] = sw ? a.z : b.z;
        dst[7] = sw ? a.w : b.w;
    }
    else
    {
#pragma unroll
        for (int f = 0; f < F4; ++f)
        {
            const float4 v = p4[f];
            dst[4 * f + 0] = v.x;
            dst[4 * f + 1] = v.y;
            dst[4 * f + 2] = v.z;
            dst[4 * f + 3] = v.w;
        }
    }
}

template <int K, int PT, int NW, int STAGES>
struct RtmaCfg
{
    static constexpr int G = Geo<K>::G;
    static constexpr int NTC = NW * 32;              // consumer threads
    static constexpr int TILE_REFS = NTC * PT * G;   // references per tile
    static constexpr int TILE_FLOATS = TILE_REFS * K;
    static constexpr uint32_t TILE_BYTES = TILE_FLOATS * 4u;
    static constexpr size_t smem(int mq) { return (size_t)STAGES * TILE_BYTES + (size_t)STAGES * 16 + (size_t)(mq / 2) * K * 8; }
};

template <int K, int MQ, int PT, int NW, int STAGES, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) nn_rtma_kernel(const RregArgs a)
{
    static_assert(MQ % 2 == 0, "queries are processed in pairs");
    using C = RtmaCfg<K, PT, NW, STAGES>;
    constexpr int G = C::G, F4 = Geo<K>::F4, P = G * PT, NP = MQ / 2, NTC = C::NTC;
    constexpr uint32_t TILE_REFS = C::TILE_REFS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * C::TILE_BYTES);
    uint64_t *empty = full + STAGES;
    float2 *sq = reinterpret_cast<float2 *>(empty + STAGES); // sq[pair*K + d] = (qa_d, qb_d)
    __shared__ unsigned long long wkeys[NW][MQ];             // per-warp results, merged per CTA at the end

    const int tid = threadIdx.x;
    const int pass = blockIdx.y;
    const int q0 = pass * MQ;
    const int valid_q = min(MQ, a.mq_total - q0); // >= 1
    for (int i = tid; i < NW * MQ; i += (NW + 1) * 32)
        (&wkeys[0][0])[i] = KEY_NONE;
    for (int i = tid; i < NP * K; i += (NW + 1) * 32)
    {
        const int pr = i / K, d = i % K;
        const int qa = min(2 * pr, valid_q - 1), qb = min(2 * pr + 1, valid_q - 1);
        sq[i] = make_float2(__ldg(a.S + (size_t)(q0 + qa) * K + d), __ldg(a.S + (size_t)(q0 + qb) * K + d));
    }
    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
        {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t nfull = a.n / TILE_REFS;
    const uint32_t rem = a.n - nfull * TILE_REFS;
    const uint32_t mine = blockIdx.x < nfull ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool tail = rem != 0 && blockIdx.x == nfull % gridDim.x;

    if (tid >= NTC)
    { // ---- producer warp: one elected lane keeps the ring full ----
        if (tid == NTC)
        {
            uint32_t s = 0, ph = 0;
            for (uint32_t j = 0; j < mine; ++j)
            {
                if (j >= (uint32_t)STAGES)
                    mbar_wait(&empty[s], ph ^ 1); // the stage's previous tile has been copied out by all warps
                const uint32_t tile = blockIdx.x + j * gridDim.x;
                mbar_expect_tx(&full[s], C::TILE_BYTES);
                bulk_g2s(ring + (size_t)s * C::TILE_FLOATS, a.R + (size_t)tile * C::TILE_FLOATS, C::TILE_BYTES, &full[s]);
                if (++s == (uint32_t)STAGES)
                {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        return; // (the consumers' closing barriers are named barriers over their NTC threads)
    }

    // ---- consumer warps ----
    float best[MQ];
    uint32_t bref[MQ];
    const float2 nz = make_float2(a.neg_zero, a.neg_zero);
#pragma unroll
    for (int j = 0; j < MQ; ++j)
    {
        best[j] = __int_as_float(0x7f800000);
        bref[j] = NO_REF;
    }
    const int lane = tid & 31;

    // minima of this thread's P references against the MQ queries, folded into (best, bref = tile)
    auto fold_tile = [&](const float(&ref)[P * K], uint32_t tile) {
#pragma unroll
        for (int pr = 0; pr < NP; ++pr)
        {
            float2 qp[K];
            if constexpr (K % 2 == 0)
            {
                const float4 *q4 = reinterpret_cast<const float4 *>(sq + pr * K);
#pragma unroll
                for (int f = 0; f < K / 2; ++f)
                {
                    const float4 v = q4[f];
                    qp[2 * f] = make_float2(v.x, v.y);
                    qp[2 * f + 1] = make_float2(v.z, v.w);
                }
            }
            else
            {
#pragma unroll
                for (int d = 0; d < K; ++d)
                    qp[d] = sq[pr * K + d];
            }
            float2 rm = make_float2(0.f, 0.f), hold = make_float2(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < P; ++p)
            {
                const float2 d = sqdist_pair<K>(qp, &ref[p * K], nz);
                if (p == 0 && P > 1)
                    hold = d;
                else if (p == 0)
                    rm = d;
                else if (p == 1)
                    rm = make_float2(fminf(hold.x, d.x), fminf(hold.y, d.y));
                else if ((p & 1) == 0 && p + 1 < P)
                    hold = d;
                else if ((p & 1) == 1)
                    rm = make_float2(fminf(fminf(rm.x, hold.x), d.x), fminf(fminf(rm.y, hold.y), d.y));
                else
                    rm = make_float2(fminf(rm.x, d.x), fminf(rm.y, d.y));
            }
            if (rm.x < best[2 * pr])
            {
                best[2 * pr] = rm.x;
                bref[2 * pr] = tile;
            }
            if (rm.y < best[2 * pr + 1])
            {
                best[2 * pr + 1] = rm.y;
                bref[2 * pr + 1] = tile;
            }
        }
    };

    if constexpr (PT % 2 == 0 && NN_RTMA_PIPELINE)
    {
        // Software pipeline over HALF tiles: while the arithmetic of one half runs, the 128-bit shared
        // loads of the next half (the second half of this tile, or the first half of the next tile
        // when it has already landed) are in flight, so neither the shared-memory latency nor the
        // wait for the TMA sits on the warp's critical path.
        constexpr int HT = PT / 2, HP = P / 2;
        float refA[HP * K], refB[HP * K];
        auto lds_half = [&](float(&dst)[HP * K], uint32_t st, int half) {
            const float4 *t4 = reinterpret_cast<const float4 *>(ring + (size_t)st * C::TILE_FLOATS);
#pragma unroll
            for (int i = 0; i < HT; ++i)
            {
                lds_group<K>(t4 + (size_t)((half * HT + i) * NTC + tid) * F4, &dst[i * G * K], lane);
            }
        };
        // query pairs kept in registers for the whole kernel where they fit (NP*K float2 = 2*NP*K
        // registers): saves 4 shared loads per pair and half tile
        constexpr bool QR = NN_RTMA_QUERY_REGS && (2 * NP * K <= 64);
        float2 qall[QR ? NP * K : 1];
        if constexpr (QR)
        {
#pragma unroll
            for (int i = 0; i < NP * K; ++i)
                qall[i] = sq[i];
        }
        // minima of HP references against the MQ queries, folded into rm[] (FIRST: rm is unset)
        auto fold_half = [&](const float(&ref)[HP * K], float2(&rm)[NP], const bool FIRST) {
#pragma unroll
            for (int pr = 0; pr < NP; ++pr)
            {
                float2 qp[K];
                if constexpr (QR)
                {
#pragma unroll
                    for (int d = 0; d < K; ++d)
                        qp[d] = qall[pr * K + d];
                }
                else if constexpr (K % 2 == 0)
                {
                    const float4 *q4 = reinterpret_cast<const float4 *>(sq + pr * K);
#pragma unroll
                    for (int f = 0; f < K / 2; ++f)
                    {
                        const float4 v = q4[f];
                        qp[2 * f] = make_float2(v.x, v.y);
                        qp[2 * f + 1] = make_float2(v.z, v.w);
                    }
                }
                else
                {
#pragma unroll
                    for (int d = 0; d < K; ++d)
                        qp[d] = sq[pr * K + d];
                }
                float2 hold = make_float2(0.f, 0.f);
                bool have = !FIRST;
#pragma unroll
                for (int p = 0; p < HP; ++p)
                {
                    const float2 d = sqdist_pair<K>(qp, &ref[p * K], nz);
                    if ((p & 1) == 0 && p + 1 < HP)
                        hold = d;
                    else if ((p & 1) == 1)
                    {
                        rm[pr] = have ? make_float2(fminf(fminf(rm[pr].x, hold.x), d.x), fminf(fminf(rm[pr].y, hold.y), d.y))
                                      : make_float2(fminf(hold.x, d.x), fminf(hold.y, d.y));
                        have = true;
                    }
                    else
                    {
                        rm[pr] = have ? make_float2(fminf(rm[pr].x, d.x), fminf(rm[pr].y, d.y)) : d;
                        have = true;
                    }
                }
            }
        };
        uint32_t s = 0, ph = 0;
        if (mine > 0)
        {
            mbar_wait(&full[0], 0);
            lds_half(refA, 0, 0);
        }
        for (uint32_t j = 0; j < mine; ++j)
        {
            lds_half(refB, s, 1);
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&empty[s]); // every shared load of this tile by this warp precedes the (release) arrive
            float2 rm[NP];
            fold_half(refA, rm, true);
            uint32_t s2 = s + 1, ph2 = ph;
            if (s2 == (uint32_t)STAGES)
            {
                s2 = 0;
                ph2 ^= 1;
            }
            const bool more = j + 1 < mine;
            bool ready = false;
            if (more)
            {
                ready = __all_sync(0xffffffffu, mbar_test(&full[s2], ph2));
                if (ready)
                    lds_half(refA, s2, 0);
            }
            fold_half(refB, rm, false);
            const uint32_t tile = blockIdx.x + j * gridDim.x;
#pragma unroll
            for (int pr = 0; pr < NP; ++pr)
            {
                if (rm[pr].x < best[2 * pr])
                {
                    best[2 * pr] = rm[pr].x;
                    bref[2 * pr] = tile;
                }
                if (rm[pr].y < best[2 * pr + 1])
                {
                    best[2 * pr + 1] = rm[pr].y;
                    bref[2 * pr + 1] = tile;
                }
            }
            if (more && !ready)
            {
                mbar_wait(&full[s2], ph2);
                lds_half(refA, s2, 0);
            }
            s = s2;
            ph = ph2;
        }
    }
    else
    {
        uint32_t s = 0, ph = 0;
        for (uint32_t j = 0; j < mine; ++j)
        {
            mbar_wait(&full[s], ph);
            float ref[P * K];
            const float4 *t4 = reinterpret_cast<const float4 *>(ring + (size_t)s * C::TILE_FLOATS);
#pragma unroll
            for (int i = 0; i < PT; ++i)
            {
                lds_group<K>(t4 + (size_t)(i * NTC + tid) * F4, &ref[i * G * K], lane);
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&empty[s]); // this warp's copy is in registers: the stage may be refilled
            fold_tile(ref, blockIdx.x + j * gridDim.x);
            if (++s == (uint32_t)STAGES)
            {
                s = 0;
                ph ^= 1;
            }
        }
    }
    if (tail)
    { // ragged end of the reference set: plain loads, slots past n are NaN (a NaN distance never wins)
        float ref[P * K];
        const size_t f0 = (size_t)nfull * C::TILE_FLOATS, fend = (size_t)a.n * K;
#pragma unroll
        for (int i = 0; i < PT; ++i)
#pragma unroll
            for (int e = 0; e < G * K; ++e)
            {
                const size_t f = f0 + (size_t)(i * NTC + tid) * (G * K) + e;
                ref[i * G * K + e] = f < fend ? __ldg(a.R + f) : __int_as_float(0x7fffffff);
            }
        fold_tile(ref, nfull);
    }

    // Warp-level merge.  For every query the lanes whose minimum equals the warp minimum are the
    // candidates (usually one; several on ties), and each candidate's exact answer is the lowest
    // index among ITS references of its winning tile at that distance.  The references are re-read
    // COOPERATIVELY -- the whole warp loads the candidate's P*K floats, one float per lane -- and the
    // loads of all MQ queries are issued before the first one is used: one memory round trip per
    // pass instead of one per query in divergent branches (that serial chain cost ~11 us per pass,
    // more than the arithmetic of a pass over 2^20 references).  Lane p < P then rebuilds reference
    // p from the lanes' registers with shuffles and evaluates it with v0's arithmetic.
    constexpr int F = P * K;             // floats of one thread's references in a tile
    constexpr int NV = (F + 31) / 32;    // registers per lane holding them
    static_assert(NV <= 2 && P <= 32, "index resolution shuffles v[0], v[1] only and lets lane p evaluate reference p: "
                                      "keep NN_RTMA_SLOT_FLOATS <= 64 and at most 32 references per thread and tile");
    const int warp_tid0 = tid - lane;    // first consumer thread of this warp
    auto load_cand = [&](uint32_t tile, int src, float(&v)[NV]) {
#pragma unroll
        for (int u = 0; u < NV; ++u)
        {
            const int e = lane + 32 * u; // float number e of the candidate's P*K
            const int i = e / (G * K), off = e % (G * K);
            const uint32_t r = tile * TILE_REFS + (uint32_t)(i * NTC + warp_tid0 + src) * G; // first point of the group
            const bool ok = e < F && r + (uint32_t)(off / K) < a.n;
            v[u] = ok ? __ldg(a.R + (size_t)r * K + off) : __int_as_float(0x7fffffff);
        }
    };
    // lowest index among the candidate's references whose distance to query j is exactly wmin
    auto eval_cand = [&](int j, uint32_t tile, int src, const float(&v)[NV], float wmin) -> uint32_t {
        const int pme = lane < P ? lane : P - 1; // lane p evaluates reference p
        float rr[K];
#pragma unroll
        for (int d = 0; d < K; ++d)
        {
            const int e = pme * K + d;
            float x = __shfl_sync(0xffffffffu, v[0], e & 31);
            if constexpr (NV > 1)
            {
                const float y = __shfl_sync(0xffffffffu, v[1], e & 31);
                x = (e >> 5) ? y : x;
            }
            rr[d] = x;
        }
        float qv[K];
#pragma unroll
        for (int d = 0; d < K; ++d)
            qv[d] = (j & 1) ? sq[(j / 2) * K + d].y : sq[(j / 2) * K + d].x;
        const float d2 = sqdist<K, 0, false>(qv, rr);
        const uint32_t r = tile * TILE_REFS + (uint32_t)((pme / G) * NTC + warp_tid0 + src) * G + (uint32_t)(pme % G);
        uint32_t idx = (lane < P && r < a.n && d2 == wmin) ? r : NO_REF;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            idx = min(idx, __shfl_xor_sync(0xffffffffu, idx, o));
        return idx;
    };

    float wmins[MQ];
    uint32_t masks[MQ], tiles[MQ];
    float vals[MQ][NV];
#pragma unroll
    for (int j = 0; j < MQ; ++j)
    {
        float wmin = best[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            wmin = fminf(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
        wmins[j] = wmin;
        masks[j] = __ballot_sync(0xffffffffu, bref[j] != NO_REF && best[j] == wmin);
        const int src = masks[j] ? __ffs(masks[j]) - 1 : 0;
        tiles[j] = __shfl_sync(0xffffffffu, bref[j], src);
        if (masks[j]) // (warp-uniform)
            load_cand(tiles[j], src, vals[j]);
    }
#pragma unroll
    for (int j = 0; j < MQ; ++j)
    {
        if (!masks[j])
            continue; // nothing beat the start state for this query in this warp (its slot stays KEY_NONE)
        uint32_t idx = eval_cand(j, tiles[j], __ffs(masks[j]) - 1, vals[j], wmins[j]);
        for (uint32_t rest = masks[j] & (masks[j] - 1); rest; rest &= rest - 1)
        { // further lanes tied at the warp minimum (rare): same procedure, one after the other
            const int src = __ffs(rest) - 1;
            const uint32_t tile = __shfl_sync(0xffffffffu, bref[j], src);
            float v[NV];
            load_cand(tile, src, v);
            idx = min(idx, eval_cand(j, tile, src, v, wmins[j]));
        }
        if (lane == 0)
            wkeys[tid >> 5][j] = pack_key(wmins[j], a.index_base + idx);
    }
    // One atomicMin per CTA and query: the warps' keys meet in shared memory first.  (Every warp of
    // every CTA folding on its own meant 148 x 12 atomics on each of the pass's 8 addresses, all at the
    // end of the kernel and serialised at one L2 slice: 20.5 -> 14.0 us for a one-tile-per-CTA launch.)
    cta_bar<1, NTC>();
    if (tid < MQ && tid < valid_q)
    {
        unsigned long long kmin = wkeys[0][tid];
#pragma unroll
        for (int w = 1; w < NW; ++w)
            kmin = wkeys[w][tid] < kmin ? wkeys[w][tid] : kmin;
        if (kmin != KEY_NONE)
            fold_key(a.keys + q0 + tid, kmin, a.peer_keys);
    }
    finish_group<1, NTC>(a.fin, a.keys, (uint32_t)pass, gridDim.x, q0, valid_q, a.q_first + q0);
}

// =============================================================================================
// Plain kernel -- one thread per query, references read from global memory, per-pair strict-less
// update.  An independent, deliberately simple formulation kept as an on-device cross-check of
// the two tuned kernels at sizes the CPU oracle cannot reach.
// =============================================================================================
template <int K>
__global__ void __launch_bounds__(128) nn_plain_kernel(const float *__restrict__ S, const float *__restrict__ R, int m,
                                                      uint32_t n, uint32_t index_base, uint32_t refs_per_split,
                                                      unsigned long long *keys, int peer_keys)
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
        fold_key(keys + qi, pack_key(best, index_base + bidx), peer_keys);
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
