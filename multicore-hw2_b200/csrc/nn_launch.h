// nn_launch.h -- host-visible launch interface of the per-k kernel translation units
// (nn_kernels_k.cu, one object per k in 3..16).  Plain C++: no device code in here.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nnb200
{

struct QregArgs
{
    const float *S;
    const float *R;
    int m;
    uint32_t n;
    uint32_t index_base;
    uint32_t splits;          // reference splits per query tile
    uint32_t refs_per_split;  // references per split (multiple of 4; the last split takes what is left)
    unsigned long long *keys;
    float neg_zero;           // must be -0.0f: run-time addend of the exact fma(d, d, -0) square
    int peer_keys;            // keys live in another GPU's memory: fold with system-scope atomics
};

struct RregArgs
{
    const float *S;   // queries of this launch (already offset to its first query)
    const float *R;
    int mq_total;     // queries covered by this launch: gridDim.y = ceil(mq_total / MQ) passes
    uint32_t n;
    uint32_t index_base;
    unsigned long long *keys; // already offset to q0
    float neg_zero;           // must be -0.0f (see QregArgs)
    int peer_keys;            // see QregArgs
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
