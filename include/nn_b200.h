/*
 * nn_b200.h -- C ABI of the B200-native brute-force 1-nearest-neighbour path.
 *
 * Drop-in boundary for wu-kan/multicore-hw2's hot path.  Every entry point names the
 * reference interface it replaces (files under /root/reference/sources/src/).
 *
 * Conventions: plain pointers and sizes only, no C++/torch types.  Points are AoS float32,
 * queries S = [m][k], references R = [n][k], exactly as the reference's harness passes them
 * (generator.h:32-50).  Results are 0-based indices into R; ties go to the lowest index and the
 * distance arithmetic is v0's (core.cu:44-54): IEEE round-to-nearest sub/mul/add, no FMA,
 * summed over the dimensions in order -- results are bit-identical to v0::cudaCallback.
 * 3 <= k <= 16.  Functions returning int return NN_B200_OK (0) or a negative NN_B200_E* code;
 * nn_b200_last_error() gives the message.  There is NO CPU fallback: without a usable CUDA
 * device every compute entry fails (the reference falls back to v0, core.cu:869-870).
 */
#ifndef NN_B200_H
#define NN_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define NN_B200_API __attribute__((visibility("default")))
#else
#define NN_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define NN_B200_OK 0
#define NN_B200_EINVAL (-1)  /* bad k/m/n, null or misaligned pointer */
#define NN_B200_ECUDA (-2)   /* a CUDA runtime call or kernel launch failed */
#define NN_B200_ENCCL (-3)   /* NCCL could not be loaded / initialised / failed */
#define NN_B200_ENODEV (-4)  /* no CUDA device visible */

#define NN_B200_KMIN 3
#define NN_B200_KMAX 16

/* v0's start state (INFINITY, index 0), core.cu:39-40, as a packed key. */
#define NN_B200_KEY_INIT 0x7F80000000000000ull

/* ------------------------------------------------------------------------------------------
 * 1. The reference's entry point, host pointers in, malloc'ed indices out.
 * ------------------------------------------------------------------------------------------ */

/* Replaces the global `cudaCallback` (core.h:71, core.cu:1282-1297) and therefore
 * `v8::cudaCallback` (core.cu:856-958) which it forwards to.  Same contract: S and R are host
 * arrays owned by the caller and only read; *results is allocated with malloc(sizeof(int)*m)
 * and freed by the caller (core.cu:935; main.cu:98).  Synchronous.  Uses as many of the visible GPUs
 * as pay for themselves (nn_b200_plan_gpus -- the role of the small-n shortcut core.cu:871-872; at
 * most n, core.cu:867-868; NN_B200_GPUS=<count> caps it), sharding R contiguously
 * (core.cu:875-883) and merging per-query packed keys on the devices -- every GPU's search kernel
 * folds into GPU 0's key array with 64-bit system-scope atomicMin over NVLink peer memory, or, where
 * peer access is unavailable, one ncclAllReduce(min, u64) -- instead of the reference's host-side
 * second-level reduce (core.cu:936-957).  On failure it prints
 * "Error: ..." and exit(1)s like the reference's CHECK macro (core.h:77-87).
 * The library ALSO exports the C++ symbol `::cudaCallback(int,int,int,float*,float*,int**)`
 * (_Z12cudaCallbackiiiPfS_PPi) with the same body, so the reference's unmodified main.cu links. */
NN_B200_API void nn_b200_cudaCallback(int k, int m, int n, float *searchPoints, float *referencePoints,
                          int **results);

/* Same work, caller-provided result buffer (m ints, host) and an error code instead of exit(1).
 * num_gpus <= 0 means "all visible" (subject to NN_B200_GPUS). */
NN_B200_API int nn_b200_search_host(int k, int m, int n, const float *searchPoints,
                        const float *referencePoints, int *results, int num_gpus);

/* ------------------------------------------------------------------------------------------
 * 2. Device-resident building blocks (what v7/v8 run between their H2D and D2H copies).
 *    All pointers are DEVICE pointers on the current CUDA device; `stream` is a cudaStream_t
 *    (NULL = default stream).  Calls are asynchronous with respect to the host.
 * ------------------------------------------------------------------------------------------ */

/* keys[i] = NN_B200_KEY_INIT for i < m: v0's per-query start state (core.cu:39-40). */
NN_B200_API int nn_b200_keys_init(uint64_t *d_keys, int m, void *stream);

/* Fused squared-distance + argmin of m queries against n references, folded into d_keys with a
 * 64-bit atomicMin: keys[i] = min(keys[i], (float_bits(d2) << 32) | (index_base + j)).
 * Replaces `cudaCallbackKernel<1024>` (core.cu:808-855 / 662-709), the transpose it depends on
 * (`mat_inv_kernel`, core.cu:792-807 -- references are consumed in their native AoS layout) and
 * the host second-level reduction (core.cu:765-787, 936-957).  `index_base` is the global index
 * of d_R[0] (shard offset, core.cu:932-933).  Because the fold is a min, a reference set may be
 * fed in any number of calls (chunks, shards) in any order.  d_R must be 16-byte aligned. */
NN_B200_API int nn_b200_nearest_keys(int k, int m, int64_t n, const float *d_S, const float *d_R,
                         uint32_t index_base, uint64_t *d_keys, void *stream);

/* results[i] = (int)(keys[i] & 0xffffffff): the `result[...] = ind_s[0]` store of core.cu:853-854
 * after the merge. */
NN_B200_API int nn_b200_keys_unpack(const uint64_t *d_keys, int m, int *d_results, void *stream);

/* AoS [n][k] -> SoA [k][n], out[kInd*n + nInd] = in[nInd*k + kInd].  Replaces `mat_inv_kernel`
 * (core.cu:792-807; also 315-330, 533-548, 646-661) with a shared-memory staged repack whose
 * global loads and stores are both 128-bit and fully coalesced. */
NN_B200_API int nn_b200_repack_soa(int k, int64_t n, const float *d_in, float *d_out, void *stream);

/* Same fused search over a reference set already repacked to SoA [k][n] (the layout v4/v7/v8
 * search in, core.cu:830-835).  Same keys as nn_b200_nearest_keys. */
NN_B200_API int nn_b200_nearest_keys_soa(int k, int m, int64_t n, const float *d_S, const float *d_R_soa,
                             uint32_t index_base, uint64_t *d_keys, void *stream);

/* ------------------------------------------------------------------------------------------
 * 2a. The same search in ONE launch.
 *     v7's device section is transpose -> search kernel -> copy-out (core.cu:726-764); the building
 *     blocks above make it keys_init -> search -> keys_unpack.  For small problems (BASELINE config 1
 *     is 16 us of arithmetic) the launches themselves weigh, so the search kernels can also finish
 *     the job: the last CTA of every query tile to fold its candidates (a ticket counter decides)
 *     stores results[i] and puts keys and ticket back into the start state.  The state lives in a
 *     caller-owned WORKSPACE that is initialised once and is left initialised by every call.
 * ------------------------------------------------------------------------------------------ */

/* Bytes of a workspace for searches of up to m queries (ticket counters + m packed keys). */
NN_B200_API size_t nn_b200_workspace_bytes(int m);
/* Once after allocation (device memory, 16-byte aligned): tickets = 0, keys = NN_B200_KEY_INIT. */
NN_B200_API int nn_b200_workspace_init(void *d_ws, int m, void *stream);
/* The workspace's key array (pure pointer arithmetic): a reference set that arrives in pieces is
 * folded into it with nn_b200_nearest_keys, then nn_b200_workspace_finish emits the answer. */
NN_B200_API uint64_t *nn_b200_workspace_keys(void *d_ws);
/* results[i] = index of the nearest of the n references for each of the m queries (what
 * cudaCallbackKernel + the host reduce leave in `results`, core.cu:853-854, 765-787), and/or
 * keys_out[i] = the final packed key (input of a multi-GPU merge); either may be NULL, not both.
 * One kernel launch for every plan except the diagnostic plain kernel.  The workspace must be in its
 * initialised state and is left in it.  d_R must be 16-byte aligned. */
NN_B200_API int nn_b200_search_device(int k, int m, int64_t n, const float *d_S, const float *d_R,
                                      uint32_t index_base, void *d_ws, int *d_results, uint64_t *d_keys_out,
                                      void *stream);
/* Emits results / keys_out from the workspace's keys and restores its initialised state (one small
 * kernel): the closing step after folding several pieces with nn_b200_nearest_keys. */
NN_B200_API int nn_b200_workspace_finish(void *d_ws, int m, int *d_results, uint64_t *d_keys_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * 2b. Resident reference index: build once, query many times.
 *     The reference pays the host->device copy of the reference set on every call
 *     (core.cu:885-891); its only build-once/query-many variants are the KD-trees v9/v10
 *     (build core.cu:984-1009, ask 1010-1026, use 1037-1047; README.md:335-344 reports query and total time
 *     separately).  The same usage on the brute-force path: the shards stay in HBM (native AoS,
 *     one contiguous shard per GPU, core.cu:875-883), a search moves only the queries and the
 *     indices across PCIe.  Results are identical to nn_b200_search_host on the same data.
 * ------------------------------------------------------------------------------------------ */
typedef struct nn_b200_index nn_b200_index;

/* Uploads the n references (host, AoS [n][k]) to num_gpus devices (<= 0: all visible). */
NN_B200_API int nn_b200_index_create(int k, int n, const float *referencePoints, int num_gpus, nn_b200_index **index);
/* results[i] = index of the nearest resident reference of query i (m queries, host, AoS [m][k]).
 * On a single-GPU index, batches of up to 1 MiB of queries are launch-bound: their copy-in, init,
 * search, unpack and copy-out are captured once per batch size into a CUDA graph and replayed with
 * one launch (option "index_graph" = 0 disables). */
NN_B200_API int nn_b200_index_search(nn_b200_index *index, int m, const float *searchPoints, int *results);
NN_B200_API int nn_b200_index_info(const nn_b200_index *index, int *k, int64_t *n, int *gpus);
NN_B200_API void nn_b200_index_destroy(nn_b200_index *index);

/* ------------------------------------------------------------------------------------------
 * 2c. Multi-PROCESS merge over NVLink peer memory (one process per GPU of one node).
 *     v8 gathers the per-GPU candidates on the host and reduces them there (core.cu:925-957); with one
 *     process per GPU the obvious replacement is search -> ncclAllReduce(min, u64) -> unpack.  Here the
 *     exchange happens INSIDE the search kernels instead: rank 0 owns two alternating key arrays in
 *     device memory that every other rank maps through CUDA IPC; every rank's search kernel folds its
 *     shard's candidates into them with system-scope 64-bit atomicMin; the last CTA of a rank counts
 *     the rank in; rank 0's last CTA waits for all ranks, stores the indices, restores the keys and
 *     releases the buffer to the ranks (a flag in each rank's own memory).  One kernel launch per rank
 *     and search, no collective, no host round trip.
 *     Set-up: every rank calls _create, the 64-byte handles are exchanged by whatever transport the
 *     host program has (torch.distributed in multicore_hw2_b200.sharded.PeerMerge), every rank calls
 *     _attach with all of them in rank order.
 * ------------------------------------------------------------------------------------------ */
typedef struct nn_b200_peer_merge nn_b200_peer_merge;
#define NN_B200_PEER_HANDLE_BYTES 64
NN_B200_API int nn_b200_peer_create(int m, int rank, int world, nn_b200_peer_merge **pm);
NN_B200_API int nn_b200_peer_handle(const nn_b200_peer_merge *pm, void *handle, size_t len);
NN_B200_API int nn_b200_peer_attach(nn_b200_peer_merge *pm, const void *handles_in_rank_order, size_t len);
/* This rank's n references (device, AoS, global indices from index_base) against the m queries
 * (device; the same on every rank).  Rank 0's d_results (device) receives the merged indices when its
 * kernel completes; other ranks pass NULL.  Every rank calls it the same number of times.  Async. */
NN_B200_API int nn_b200_peer_search(nn_b200_peer_merge *pm, int k, int m, int64_t n, const float *d_S, const float *d_R,
                                    uint32_t index_base, int *d_results, void *stream);
/* 1 if a wait inside one of this rank's kernels timed out (a rank missing or out of step). Synchronises. */
NN_B200_API int nn_b200_peer_error(nn_b200_peer_merge *pm);
NN_B200_API void nn_b200_peer_destroy(nn_b200_peer_merge *pm);

/* ------------------------------------------------------------------------------------------
 * 3. Host-side helpers
 * ------------------------------------------------------------------------------------------ */

/* Contiguous reference shard of rank `shard` of `num_shards` (v8's partition, core.cu:875-883,
 * without its duplicated-last-point hack: an empty shard is allowed).  Shard starts are
 * multiples of 4 references so that every shard stays 16-byte aligned for any k. */
NN_B200_API int nn_b200_shard_range(int64_t n, int num_shards, int shard, int64_t *begin, int64_t *count);

/* Number of CUDA devices the host entry would use for n references (core.cu:865-868). */
NN_B200_API int nn_b200_device_count(int64_t n);

/* GPUs the host entry uses for a call when the caller names no count (pure arithmetic, no device
 * needed): the analogue of the reference's device-count clamp and small-n single-GPU shortcut
 * (core.cu:865-872).  Sharding divides the copy and the search but costs a host thread per GPU and the
 * merge, so small calls stay on one GPU.  `visible` = GPUs available (nn_b200_device_count). */
NN_B200_API int nn_b200_plan_gpus(int k, int m, int64_t n, int visible);
/* Introspection of the host entry's ingest pipeline (pure arithmetic, no device needed): the reference
 * set of a shard reaches the GPU in `nchunks` H2D chunks of chunk_refs[i] references (`gpus` GPUs of the
 * call share the host's copy bandwidth); the searches over them are launched in groups of consecutive chunks.  Writes, for every launch g, the index one
 * past its last chunk into group_ends[g] (room for nchunks entries) and returns the number of
 * launches.  When the search is the slower side (many queries) a launch takes every chunk that must
 * have landed by the time it starts; otherwise a fixed few (option "search_group"). */
NN_B200_API int nn_b200_plan_search_groups(int k, int m, int gpus, const int64_t *chunk_refs, int nchunks,
                                           int *group_ends);

/* GPUs the most recent nn_b200_cudaCallback / nn_b200_search_host call of this process used. */
NN_B200_API int nn_b200_last_gpus(void);

/* Loads every search kernel (all k, all tile shapes) on the current device so that no later call
 * pays the lazy code loading of a first use (about 1-3 ms per kernel).  The reference hides the same
 * cold start with its static WarmUP object (core.cu:1274).  The host entry points call it once per
 * device on first use unless the environment has NN_B200_WARMUP=0. */
NN_B200_API int nn_b200_warmup(void);

/* Kernels launched by this library in this process so far (bench.py's `gpu_launches`). */
NN_B200_API int64_t nn_b200_launch_count(void);

/* Message of the last error on the calling thread ("" if none). */
NN_B200_API const char *nn_b200_last_error(void);

/* Tuning knobs for benchmarking/sweeps; production code never needs them.  Known names:
 * "variant" (0 auto, 1 query-register kernel, 2 reference-register kernel, 3 plain kernel, 4 reference-stream kernel,
 * 5 phased query-register kernel),
 * "splits" (0 auto), "h2d_chunk_bytes", "auto_gpus" (1: nn_b200_plan_gpus picks the GPU count of a host call, 0: all visible), "p2p_merge" (multi-GPU host entry: 1 = the search kernels of
 * every GPU fold into GPU 0's key array with system-scope atomics over NVLink, 0 = NCCL all-reduce),
 * "qreg_super" (query-register kernel: chunks per update of the running (best, where) pair; -1 = by k
 * and split length, 0 = always per chunk).
 * Returns NN_B200_EINVAL for an unknown name. */
NN_B200_API int nn_b200_set_option(const char *name, int64_t value);

/* Measurement aid: sustained rate, in lane-operations per second, at which this device issues
 * NON-fused FP32 multiplies and adds -- the measured denominator of the FP32 roofline.
 * mode 0: scalar FMUL/FADD; 1: packed f32x2 FMUL2/FADD2; 2: packed and scalar alternating;
 * 3: blocks of packed then scalar; 4: packed + FMNMX; 5: scalar + FMNMX.  Synchronous. */
NN_B200_API int nn_b200_probe_fp32(int mode, int iters, double *lane_ops_per_s);

/* Human-readable description of the launch plan nn_b200_nearest_keys would use (variant, tile,
 * grid); written into buf (NUL-terminated, truncated to len). */
NN_B200_API int nn_b200_describe_plan(int k, int m, int64_t n, char *buf, size_t len);

/* Which kernel family nn_b200_nearest_keys picks for m queries x n references of k dimensions (pure
 * arithmetic, no device needed): 1 query-register, 2 reference-register, 4 reference-stream, 5 phased query-register;
 * NN_B200_EINVAL for a bad shape. */
NN_B200_API int nn_b200_plan_variant(int k, int m, int64_t n);

/* Layout of the phased query-register kernel for m queries (pure arithmetic): q queries per thread,
 * `groups` x `phases` <= 128 threads per CTA, `qtiles` query tiles of `tile_queries` queries each
 * (tile_queries <= groups * q, qtiles * tile_queries >= m). */
NN_B200_API int nn_b200_plan_flex(int k, int m, int *q, int *groups, int *phases, int *qtiles, int *tile_queries);

/* Introspection of the launch planner (pure arithmetic, no device needed): for the query-register
 * kernel with `q` queries per thread (tile = 128 q queries) at `occ` CTAs per SM on `sms` SMs, how
 * many reference splits per query tile a search of m queries x n references is launched with and
 * how many references each split covers (a multiple of 4). */
NN_B200_API int nn_b200_plan_splits(int k, int q, int occ, int sms, int64_t m, int64_t n, int64_t *splits,
                                    int64_t *refs_per_split);

#ifdef __cplusplus
} /* extern "C" */

/* The reference's own C++-linkage symbol (core.h:71). */
NN_B200_API void cudaCallback(int k, int m, int n, float *searchPoints, float *referencePoints, int **results);
#endif

#endif /* NN_B200_H */
