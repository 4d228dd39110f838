// nn_launch.h -- host-visible launch interface of the per-k kernel translation units
// (nn_kernels_k.cu, one object per k in 3..16).  Plain C++: no device code in here.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nnb200
{

#ifdef NN_AB_MATH
constexpr bool kAbMath = true; // the scalar / dimension-pair math modes are compiled in (A/B builds)
#else
constexpr bool kAbMath = false;
#endif

// "Finish in the same launch": when `tickets` is non-null the CTAs of one ticket group (a query tile of
// the query-register kernels, a pass of the few-query kernels) count themselves on tickets[group]
// after folding their candidates; the LAST one to arrive reads the group's final keys, stores
// results[i] = index (and/or keys_out[i] = key) and puts keys and ticket back into the start
// state.  One launch then does what keys_init -> search -> keys_unpack did in three, and the key
// array it folds into (the workspace) is left ready for the next search.  Replaces
// `result[bx*m+by] = ind_s[0]` (core.cu:853-854) + the host reduce (core.cu:765-787).
// Multi-PROCESS merge over NVLink peer memory (one process per GPU, nn_b200_peer_*).  Every rank's
// search kernel first folds into the rank's OWN key workspace (GPU-scope atomics, as on one GPU); the
// last CTA of every ticket group then pushes the group's final keys into rank 0's key array (CUDA-IPC
// mapped) with ONE system-scope atomicMin per query, and the last group to do so counts the rank in
// on `arrive`.  Rank 0's last CTA waits for all ranks, stores the indices, restores the keys and tells
// every rank -- by a flag in THAT rank's memory -- how many searches are finished, which is what lets
// a fast rank push search s only after search s-2 (same key buffer: two alternate) has been consumed.
// No collective launch, no host round trip, m remote atomics per rank and search.
struct PeerSync
{
    unsigned long long *keys = nullptr;        // rank 0's memory: this search's key buffer (m keys; two alternate)
    unsigned int *arrive = nullptr;            // rank 0's memory: [2] ranks folded, per key buffer
    unsigned int *done_local = nullptr;        // this rank's memory: searches finished by rank 0
    unsigned int *const *done_peers = nullptr; // rank 0 only: the other ranks' flags (peer-mapped), world-1 of them
    unsigned int *error = nullptr;             // this rank's memory: set when a wait timed out
    unsigned int *groups_done = nullptr;       // this rank's memory: ticket groups of this call that have pushed
    unsigned int num_groups = 0;               // ticket groups of this call over all its launches
    unsigned int step = 0, world = 0, rank = 0;
    int m = 0;                                 // queries (rank 0 unpacks all of them)
};

struct Finish
{
    unsigned int *tickets = nullptr;        // one counter per ticket group, all zero between launches
    int *results = nullptr;                 // optional: int[m] nearest indices
    unsigned long long *keys_out = nullptr; // optional: final packed keys (multi-GPU merge input)
    PeerSync peer;                          // peer.arrive != nullptr: multi-process mode (results/keys_out: rank 0's)
};

// Super-chunk form of nn_qreg_kernel (the (best, where) update once per several chunks): chunks per update
// on long splits, by k; 0 = the form is not built for this k.  From the per-k A/B on B200
// (profiles/r02_super_chunks.txt: m = 65536, n = 2^21): +2.7% at k = 7, +3.6% at k = 15, +1.4% at k = 14,
// +0.4..1.1% at k = 4, 5, 6, 11, 16; nothing or a loss at k = 3, 8, 9, 10, 12, 13.
constexpr int qreg_super_chunks(int k)
{
    return (k == 4 || k == 5 || k == 6 || k == 7 || k == 15 || k == 16) ? 16 : ((k == 11 || k == 14) ? 8 : 0);
}
constexpr uint32_t kQregSuperMinRefs = 65536; // splits at least this long use it (its winner re-scan is per CTA)

struct QregArgs
{
    const float *S;
    const float *R;
    int m;
    uint32_t n;
    uint32_t index_base;
    uint32_t splits;          // reference splits per query tile
    uint32_t qgroup = 1;      // CTA order: query tiles per group (cta_to_work, nn_kernels.cuh)
    uint32_t refs_per_split;  // references per split (multiple of 4; the last split takes what is left)
    uint32_t super_chunks = 1; // chunks between two (best, where) updates, >= 1 (nn_qreg_kernel: super-chunks)
    unsigned long long *keys;
    float neg_zero;           // must be -0.0f: run-time addend of the exact fma(d, d, -0) square
    int peer_keys;            // keys live in another GPU's memory: fold with system-scope atomics
    Finish fin;               // ticket group = query tile (blockIdx.x / splits), `splits` CTAs each
#ifdef NN_QREG_TIMELINE
    unsigned long long *timeline = nullptr; // debug builds: 6 global-timer + 6 SM-clock stamps per CTA, 16 slots (nn_bench --timeline)
#endif
};

// Phased query-register kernel (nn_qflex_kernel): 128 threads = ng query groups x np phases.
struct QflexArgs
{
    const float *S;
    const float *R;
    int m;
    uint32_t n;
    uint32_t index_base;
    uint32_t splits;         // reference splits per query tile
    uint32_t qgroup = 1;     // CTA order: query tiles per group (cta_to_work, nn_kernels.cuh)
    uint32_t refs_per_split; // multiple of np * CH references (the last split takes what is left)
    uint32_t tile_queries;   // queries per query tile, <= ng * Q
    uint32_t ng, np;         // ng * np <= 128
    uint32_t tile_groups;    // reference groups per ring stage: a multiple of np * CH / G that fits a stage
    uint32_t stages;         // ring depth (2..8)
    uint32_t stage_floats;   // floats per ring stage (a multiple of 4; dynamic smem = 128 + stages * stage_floats * 4)
    unsigned long long *keys;
    float neg_zero;
    int peer_keys;
    Finish fin;              // ticket group = query tile, `splits` CTAs each
};

struct RregArgs
{
    const float *S;   // queries of this launch (already offset to its first query)
    const float *R;
    int mq_total;     // queries covered by this launch: gridDim.y = ceil(mq_total / MQ) passes
    int q_first = 0;  // index of the launch's first query in the whole query set (multi-process merge)
    uint32_t n;
    uint32_t index_base;
    unsigned long long *keys; // already offset to q0
    float neg_zero;           // must be -0.0f (see QregArgs)
    int peer_keys;            // see QregArgs
    Finish fin;               // ticket group = pass (blockIdx.y), gridDim.x CTAs each; results/keys_out offset like keys
};

// ---- per-K launchers (defined in nn_kernels_k.cu) -----------------------------------------
struct LaunchInfo
{
    int regs;
    int smem;
    int occ; // CTAs per SM
};

template <int K>
cudaError_t launch_qreg(int q_sel, int math, const QregArgs &a, uint32_t qtiles, cudaStream_t st);
template <int K>
cudaError_t query_qreg(int q_sel, int math, LaunchInfo *info, int *tile_queries, int *tile_refs);
// q = 2, 4 or 8 queries per thread (8 only where kernel A has it); info: chunk (CH) and group (G) sizes, TR
struct FlexInfo
{
    int regs, smem, occ;
    int ch, g, tr;
};
template <int K>
cudaError_t launch_qflex(int q, const QflexArgs &a, uint32_t qtiles, cudaStream_t st);
template <int K>
cudaError_t query_qflex(int q, int smem_bytes, FlexInfo *info); // occupancy for that much dynamic shared memory
template <int K>
cudaError_t launch_rreg(int mq, bool soa, const RregArgs &a, dim3 grid, cudaStream_t st);
template <int K>
cudaError_t query_rreg(int mq, bool soa, LaunchInfo *info, int *refs_per_batch);
template <int K>
cudaError_t launch_rtma(int mq, const RregArgs &a, dim3 grid, cudaStream_t st);
template <int K>
cudaError_t query_rtma(int mq, LaunchInfo *info, int *tile_refs);
template <int K>
cudaError_t launch_plain(const float *S, const float *R, int m, uint32_t n, uint32_t index_base, uint32_t splits,
                         unsigned long long *keys, int peer_keys, cudaStream_t st);


// AoS [n][k] -> SoA [k][n] repack (replaces mat_inv_kernel, core.cu:792-807).
template <int K>
cudaError_t launch_repack_soa(const float *in, float *out, uint32_t n, int num_sms, cudaStream_t st);

} // namespace nnb200
