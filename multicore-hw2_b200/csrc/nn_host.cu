// nn_host.cu -- host side of the B200 brute-force 1-NN path: launch planning, the C ABI of
// include/nn_b200.h and the drop-in `cudaCallback` (replaces v8::cudaCallback / v7::cudaCallback,
// /root/reference/sources/src/core.cu:856-958 and 710-788).
//
// Differences from the reference's orchestration, by design:
//   * no transpose pass and no thrust: references are searched in the AoS layout they arrive in;
//   * the host->device copy of the reference set is chunked on a copy stream and overlapped with
//     the search of the previous chunk (min-folding makes chunks independent), instead of one
//     synchronous thrust::device_vector construction (core.cu:885-891);
//   * per-GPU candidates are merged on the devices, not on the host (core.cu:925-957): by default every
//     GPU's search kernel folds its packed keys into GPU 0's key array with system-scope 64-bit
//     atomicMin over NVLink peer memory; without native peer atomics, ncclAllReduce(min, uint64) --
//     and the merge is correct for m > 1;
//   * a single-GPU, single-chunk call is ONE kernel launch: the search kernel's last CTA per query
//     tile stores the indices and restores the key workspace (struct Finish, nn_launch.h);
//   * device buffers, streams and communicators live in a lazily created context that survives
//     across calls (the reference re-allocates per call and hides a 30 ms cold start with its
//     static WarmUP object, core.cu:1274);
//   * no CPU fallback (core.cu:869-870 falls back to v0): without a GPU the call fails loudly.
#include "../../include/nn_b200.h"
#include "nn_kernels.cuh" // (templates only: finish_group for the arrive-only kernel of the multi-process merge)
#include "nn_launch.h"

#include <algorithm>
#include <cmath>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace nnb200;

// ---------------------------------------------------------------------------------------------
// errors, options, counters
// ---------------------------------------------------------------------------------------------
constexpr int kWsTickets = 8192; // ticket counters at the head of a search workspace (see nn_b200_workspace_bytes)
static thread_local std::string t_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
    return code;
}
#define CU(call)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
            return fail(NN_B200_ECUDA, "%s:%d: %s -> code %d, reason: %s", __FILE__, __LINE__, #call, (int)e__,        \
                        cudaGetErrorString(e__));                                                                      \
    } while (0)

struct Options
{
    std::atomic<int64_t> variant{0};         // 0 auto, 1 qreg, 2 rreg, 3 plain, 4 rtma, 5 qflex
    std::atomic<int64_t> splits{0};          // qreg reference splits per query tile (0 auto)
    std::atomic<int64_t> qreg_q{0};          // qreg queries per thread (0 auto)
    std::atomic<int64_t> math{2};            // 2 f32x2 over query pairs, 1 f32x2 over dims, 0 scalar (A/B only)
    std::atomic<int64_t> rreg_max_m{112};    // m <= this -> few-query kernels (rreg up to 4 queries, rtma above)
    std::atomic<int64_t> rreg_ctas_per_sm{0}; // 0: occupancy
    std::atomic<int64_t> h2d_chunk_bytes{16 << 20};
    std::atomic<int64_t> waves{8};           // qreg: most CTA waves considered
    std::atomic<int64_t> stage_threads{-1};  // pageable inputs: host threads staging into pinned buffers (-1 auto, 0 off)
    std::atomic<int64_t> stage_min_bytes{48 << 20}; // pageable reference sets from this size on go through the staging threads
    std::atomic<int64_t> flex_deep_ring{1};  // phased kernel ring: 1 = 4 x ~12 KB (default), 0 = 3 x ~8 KB, -1 = deep only beyond 8 phases
    std::atomic<int64_t> qgroup{128};        // CTA order of the query-register kernels: query tiles per group (1 = split fastest)
    std::atomic<int64_t> qreg_super{-1};     // query-register kernel: chunks per (best, where) update; -1 = by the split's length
    std::atomic<int64_t> search_group{8};    // host entry: at most this many landed H2D chunks are searched by one launch
    std::atomic<int64_t> stage_one_stream{1}; // staging threads push their copies on the shared copy stream (5% faster than a stream each)
    std::atomic<int64_t> index_graph{1};     // resident index on one GPU: replay a captured CUDA graph for small batches
    std::atomic<int64_t> p2p_merge{1};       // multi-GPU host entry: 1 fold into GPU 0's keys over NVLink, 0 NCCL all-reduce
    std::atomic<int64_t> auto_gpus{1};       // host entry without an explicit GPU count: 1 = plan_gpus decides, 0 = all visible
};
static Options g_opt;
static std::atomic<int> g_last_gpus{0};   // GPUs the most recent host-entry call used (nn_b200_last_gpus)
static std::atomic<int> g_active_gpus{1}; // GPUs driven by the host entry call in flight (sizes the staging threads)
static std::atomic<int64_t> g_opt_epoch{0}; // bumped by every set_option: cached plans of older epochs are stale

extern "C" int nn_b200_set_option(const char *name, int64_t value)
{
    if (!name)
        return fail(NN_B200_EINVAL, "null option name");
    const std::string s(name);
    if (s == "variant")
        g_opt.variant = value;
    else if (s == "splits")
        g_opt.splits = value;
    else if (s == "qreg_q")
        g_opt.qreg_q = value;
    else if (s == "math")
    {
        if (value != 2 && !kAbMath)
            return fail(NN_B200_EINVAL, "math mode %lld needs a library built with -DNN_AB_MATH", (long long)value);
        g_opt.math = value;
    }
    else if (s == "rreg_max_m")
        g_opt.rreg_max_m = value;
    else if (s == "rreg_ctas_per_sm")
        g_opt.rreg_ctas_per_sm = value;
    else if (s == "h2d_chunk_bytes")
        g_opt.h2d_chunk_bytes = value;
    else if (s == "waves")
        g_opt.waves = value;
    else if (s == "p2p_merge")
        g_opt.p2p_merge = value;
    else if (s == "auto_gpus")
        g_opt.auto_gpus = value;
    else if (s == "index_graph")
        g_opt.index_graph = value;
    else if (s == "stage_threads")
        g_opt.stage_threads = value;
    else if (s == "stage_min_bytes")
        g_opt.stage_min_bytes = value;
    else if (s == "search_group")
        g_opt.search_group = value;
    else if (s == "qgroup")
        g_opt.qgroup = value;
    else if (s == "qreg_super")
        g_opt.qreg_super = value;
    else if (s == "flex_deep_ring")
        g_opt.flex_deep_ring = value;
    else if (s == "stage_one_stream")
        g_opt.stage_one_stream = value;
    else
        return fail(NN_B200_EINVAL, "unknown option '%s'", name);
    g_opt_epoch++;
    return NN_B200_OK;
}

extern "C" const char *nn_b200_last_error(void) { return t_err.c_str(); }
extern "C" int64_t nn_b200_launch_count(void) { return g_launches.load(); }
extern "C" int nn_b200_last_gpus(void) { return g_last_gpus.load(); }

#ifdef NN_QREG_TIMELINE
// debug builds only (make EXTRA=-DNN_QREG_TIMELINE): device buffer of 8 time stamps per CTA that the
// query-register kernel fills (nn_bench --timeline); not part of the ABI of include/nn_b200.h
static unsigned long long *g_timeline = nullptr;
extern "C" __attribute__((visibility("default"))) void nn_b200_debug_timeline(unsigned long long *p) { g_timeline = p; }
#endif

// Chunks per (best, where) update of the query-register kernel (its super-chunk form, nn_kernels.cuh): three
// instructions per query saved on every chunk but the first of a super-chunk, against a winner re-scan of
// one super-chunk per query and CTA -- so only on long splits, and only for the k where it measured faster.
// Option qreg_super: -1 this rule, 0 or 1 off, n > 1 forced (for the k the form is built for).
static uint32_t qreg_super_for(int k, uint32_t refs_per_split)
{
    const int64_t sc = g_opt.qreg_super.load();
    if (qreg_super_chunks(k) == 0 || sc == 0)
        return 1u;
    if (sc > 0)
        return (uint32_t)std::min<int64_t>(sc, 64);
    return refs_per_split >= kQregSuperMinRefs ? (uint32_t)qreg_super_chunks(k) : 1u;
}

// ---------------------------------------------------------------------------------------------
// per-k dispatch
// ---------------------------------------------------------------------------------------------
#define NN_FOR_K(X) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

static cudaError_t k_launch_qreg(int k, int q, int nt, const QregArgs &a, uint32_t qtiles, cudaStream_t st)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return launch_qreg<KK>(q, nt, a, qtiles, st);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_query_qreg(int k, int q, int nt, LaunchInfo *li, int *tq, int *tr)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return query_qreg<KK>(q, nt, li, tq, tr);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_launch_qflex(int k, int q, const QflexArgs &a, uint32_t qtiles, cudaStream_t st)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return launch_qflex<KK>(q, a, qtiles, st);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_query_qflex(int k, int q, int smem_bytes, FlexInfo *fi)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return query_qflex<KK>(q, smem_bytes, fi);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_launch_rreg(int k, int mq, bool soa, const RregArgs &a, dim3 grid, cudaStream_t st)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return launch_rreg<KK>(mq, soa, a, grid, st);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_query_rreg(int k, int mq, bool soa, LaunchInfo *li, int *rpb)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return query_rreg<KK>(mq, soa, li, rpb);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_launch_rtma(int k, int mq, const RregArgs &a, dim3 grid, cudaStream_t st)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return launch_rtma<KK>(mq, a, grid, st);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_query_rtma(int k, int mq, LaunchInfo *li, int *tile_refs)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return query_rtma<KK>(mq, li, tile_refs);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_launch_plain(int k, const float *S, const float *R, int m, uint32_t n, uint32_t base,
                                  uint32_t splits, unsigned long long *keys, int peer, cudaStream_t st)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return launch_plain<KK>(S, R, m, n, base, splits, keys, peer, st);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
static cudaError_t k_launch_repack(int k, const float *in, float *out, uint32_t n, int sms, cudaStream_t st)
{
    switch (k)
    {
#define X(KK)                                                                                                          \
    case KK:                                                                                                           \
        return launch_repack_soa<KK>(in, out, n, sms, st);
        NN_FOR_K(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// small kernels: key init / unpack
// ---------------------------------------------------------------------------------------------
__global__ void nn_keys_init_kernel(unsigned long long *keys, int m)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m)
        keys[i] = NN_B200_KEY_INIT;
}
__global__ void nn_keys_unpack_kernel(const unsigned long long *__restrict__ keys, int m, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m)
        out[i] = (int)(unsigned int)(keys[i] & 0xffffffffull);
}

// unpack + restore the start state in one pass: what the last CTA of a ticket group does inside the
// search kernels (finish_group, nn_kernels.cuh), as a kernel of its own for the cases that fold a
// workspace in several launches (H2D chunks, the plain kernel, n = 0)
__global__ void nn_keys_finish_kernel(unsigned long long *__restrict__ keys, int m, int *__restrict__ out,
                                      unsigned long long *__restrict__ keys_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m)
    {
        const unsigned long long key = keys[i];
        if (out)
            out[i] = (int)(unsigned int)(key & 0xffffffffull);
        if (keys_out)
            keys_out[i] = key;
        keys[i] = NN_B200_KEY_INIT;
    }
}

// ---------------------------------------------------------------------------------------------
// device properties (cached per device)
// ---------------------------------------------------------------------------------------------
struct DevInfo
{
    bool ok = false;
    int sms = 0;
    int cc = 0;
};
static std::mutex g_dev_mu;
static DevInfo g_dev[64];

static int dev_info(int dev, DevInfo *out)
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (dev < 0 || dev >= 64)
        return fail(NN_B200_EINVAL, "device ordinal %d out of range", dev);
    if (!g_dev[dev].ok)
    {
        int sms = 0, major = 0, minor = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        CU(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
        g_dev[dev].sms = sms;
        g_dev[dev].cc = major * 10 + minor;
        g_dev[dev].ok = true;
    }
    *out = g_dev[dev];
    if (out->cc != 100)
        return fail(NN_B200_ENODEV, "device %d has compute capability %d.%d; this library is built for sm_100a only",
                    dev, out->cc / 10, out->cc % 10);
    return NN_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// launch planning
// ---------------------------------------------------------------------------------------------
static bool check_shape_quiet(int k, int m, int64_t n)
{
    return k < NN_B200_KMIN || k > NN_B200_KMAX || m < 0 || n < 0 || n > 0x7fffffffLL;
}

struct Plan
{
    int variant = 0; // 1 qreg, 2 rreg, 3 plain, 4 rtma, 5 qflex
    int tile_refs = 0; // rtma
    // qflex (shares q, occ, regs, qtiles, splits, refs_per_split with qreg)
    int ng = 0, np = 0;
    uint32_t tile_queries = 0, tile_groups = 0, stages = 0, stage_floats = 0;
    // qreg
    int q = 0, scalar = 0, tile_q = 0, tile_r = 0, occ = 0, regs = 0;
    uint32_t qtiles = 0, splits = 0, refs_per_split = 0;
    // rreg
    int rreg_ctas = 0;
    // plain
    uint32_t plain_splits = 1;
};

// Time model of the query-register kernel (pure arithmetic, SM clock cycles) for `splits` reference
// splits per query tile.  A CTA is four warps, one per SM sub-partition, so c co-resident CTAs put
// c warps on every FMA pipe; a warp needs (q/2)(3k-1) packed instructions of 2 pipe cycles per
// reference.  eff(c) is the measured fraction of the pipe that c warps keep busy (one warp alone
// exposes its shared-memory and dependent-issue latencies), ovh the per-CTA prologue (query loads,
// first TMA tile) and epilogue (index resolution, atomics).
static double qreg_model_cycles(int k, int q, int occ, int sms, int64_t m, int64_t n, int64_t splits, int64_t *rps_out)
{
    const int64_t qtiles = (m + 128 * (int64_t)q - 1) / (128 * (int64_t)q);
    const int64_t total = qtiles * splits;
    int64_t rps = (n + splits - 1) / splits;
    rps = (rps + 7) / 8 * 8; // whole chunks (4 points; 8 in the NN_QREG_CH8 build) and 16-byte aligned starts
    if (rps_out)
        *rps_out = rps;
    // f_q: measured pipe share of the tile shape itself (shared loads and compares per packed
    // instruction grow as the tile narrows; one query per thread cannot use the pair-packed math)
    double f_q = q >= 8 ? 1.0 : (q >= 4 ? 0.988 : (q >= 2 ? 0.965 : 1.0 / 1.3));
    // measured at m = 16384, n = 2^19 on B200 (fraction of the FP32 roofline) where 8 and 4 queries per
    // thread are both built: the 8-wide tile runs at 128 registers and loses to the 4-wide one at
    // k = 5, 7, 8
    static const double share8[9] = {0, 0, 0, 0.993, 0.985, 0.941, 0.954, 0.949, 0.924};
    static const double share4[9] = {0, 0, 0, 0.974, 0.975, 0.973, 0.943, 0.960, 0.954};
    if (k <= 8 && q >= 2)
        f_q = (q == 8 ? share8[k] : (q == 4 ? share4[k] : share4[k] * 0.975)) / 0.993;
    const double cpr = (q >= 2 ? (q / 2) * (3.0 * k - 1.0) * 2.0 : (3.0 * k - 1.0)) / f_q;
    auto eff = [](int64_t c) { return c <= 1 ? 0.58 : (c == 2 ? 0.78 : (c == 3 ? 0.86 : 0.92)); };
    // Per-CTA prologue (query loads) and epilogue (re-read of the Q winning 4-point chunks): every
    // load instruction of a warp touches 32 different sectors, so it costs ~32 LSU cycles; the CTAs
    // of an SM share the LSU.  Narrow tiles use vector loads there, the widest tile scalar ones
    // (nn_kernels.cuh, VEC); about half of it hides behind other CTAs' arithmetic.
    const int g = (k % 4 == 0) ? 1 : ((k % 2 == 0) ? 2 : 4);
    const int budget = (96 - g * k) / k;
    const int qwide = budget >= 8 ? 8 : (budget >= 4 ? 4 : (budget >= 2 ? 2 : 1));
    const bool vec = q < qwide;
    const double loads = vec ? (double)q * k / (k % 4 == 0 ? 4 : (k % 2 == 0 ? 2 : 1)) + (double)q * k
                             : (double)q * k + 4.0 * q * k;
    const double lsu = loads * 32.0 * 4.0 * 0.5;
    const double ovh = 5000.0;
    const int64_t slots = (int64_t)sms * occ;
    const int64_t full_waves = total / slots, rem = total % slots;
    double t = (double)full_waves * (ovh + (double)occ * (lsu + (double)rps * cpr / eff(occ)));
    if (rem)
    {
        const int64_t c = (rem + sms - 1) / sms;
        t += ovh + (double)c * (lsu + (double)rps * cpr / eff(c));
    }
    return t;
}

// Best split count for one tile shape: candidates are the counts that fill the SMs evenly -- c CTAs
// on every SM for c = 1..occ, or w whole waves of SMs x occ CTAs -- never fewer than 32 references
// per split.  Returns the modelled cycles.
static double plan_splits(int k, int q, int occ, int sms, int64_t m, int64_t n, int64_t wmax, int64_t forced,
                          int64_t *splits_out, int64_t *rps_out)
{
    const int64_t qtiles = (m + 128 * (int64_t)q - 1) / (128 * (int64_t)q);
    const int64_t smax = std::max<int64_t>(1, n / 32);
    std::vector<int64_t> cand;
    if (forced > 0)
        cand.push_back(std::min(forced, smax));
    else
    {
        cand.push_back(1);
        for (int64_t c = 1; c <= occ; ++c)
            cand.push_back(((int64_t)sms * c) / qtiles);
        for (int64_t w = 2; w <= std::max<int64_t>(1, wmax); ++w)
            cand.push_back(((int64_t)sms * occ * w) / qtiles);
    }
    double best = -1.0;
    for (int64_t sp : cand)
    {
        sp = std::max<int64_t>(1, std::min(sp, smax));
        int64_t rps = 0;
        const double t = qreg_model_cycles(k, q, occ, sms, m, n, sp, &rps);
        const int64_t used = (n + rps - 1) / rps; // splits that actually receive references
        if (best < 0 || t < best * 0.999)
        {
            best = t;
            *splits_out = std::max<int64_t>(1, used);
            *rps_out = rps;
        }
    }
    return best;
}

extern "C" int nn_b200_plan_splits(int k, int q, int occ, int sms, int64_t m, int64_t n, int64_t *splits,
                                   int64_t *refs_per_split)
{
    if (k < NN_B200_KMIN || k > NN_B200_KMAX || q < 1 || occ < 1 || sms < 1 || m < 1 || n < 1 || !splits || !refs_per_split)
        return fail(NN_B200_EINVAL, "bad plan_splits arguments");
    plan_splits(k, q, occ, sms, m, n, 8, 0, splits, refs_per_split);
    return NN_B200_OK;
}

// ---- phased query-register kernel: layout and split planning (pure arithmetic) ----------------
// Per-k constants of nn_kernels.cuh (Geo, QregCfg, QregDefault), restated so that the choice can be
// made -- and unit-tested -- without a device.
struct FlexGeo
{
    int g, ch, chg, trg, qmax;
};
static FlexGeo flex_geo(int k)
{
    FlexGeo f;
    f.g = (k % 4 == 0) ? 1 : ((k % 2 == 0) ? 2 : 4);
    f.ch = (k == 3) ? 8 : 4;
    f.chg = f.ch / f.g;
    f.trg = (((2048 / k) / 8) * 8) / f.g;
    const int budget = (96 - f.g * k) / k;
    f.qmax = budget >= 8 ? 8 : (budget >= 4 ? 4 : 2);
    return f;
}
struct FlexLayout
{
    int q = 0, ng = 0, np = 0;
    int64_t qtiles = 0, tile_queries = 0;
    double cost = 0; // FMA-pipe cycles per reference and SM sub-partition, up to a constant
};
// Best (queries per thread, query tiles) for m queries: minimise the lane-cycles spent per reference,
// qtiles * (q/2) / np, corrected by the measured pipe share of the tile width.  np = 1 means the
// layout IS kernel A's (one phase, all threads on the same reference).
static FlexLayout flex_layout(int k, int64_t m, int forced_q)
{
    const FlexGeo fg = flex_geo(k);
    FlexLayout best;
    for (int q : {8, 4, 2})
    {
        if (q > fg.qmax || (forced_q && q != forced_q))
            continue;
        const double f_q = q >= 8 ? 1.0 : (q >= 4 ? 0.988 : 0.965);
        const int64_t tmin = (m + 128 * (int64_t)q - 1) / (128 * (int64_t)q);
        for (int64_t T = tmin; T <= tmin + 7; ++T)
        {
            const int64_t tq = (m + T - 1) / T;
            const int ng = (int)((tq + q - 1) / q);
            if (ng < 1 || ng > 128)
                continue;
            const int np = std::min(128 / ng, fg.trg / fg.chg);
            if (np < 1)
                continue;
            // many phases per CTA mean short per-thread runs inside every tile and more distinct
            // shared-memory addresses per warp: measured +11% lane-cycles per doubling beyond 8 phases
            // (B200, k = 3/8/16, m = 16..64, n = 2^22: profiles/r02_fewquery_crossover.json)
            const double pen = np <= 8 ? 1.0 : 1.0 + 0.11 * std::log2((double)np / 8.0);
            const double cost = (double)T * q * pen / ((double)np * f_q);
            if (best.q == 0 || cost < best.cost * 0.999)
            {
                best.q = q;
                best.ng = ng;
                best.np = np;
                best.qtiles = T;
                best.tile_queries = tq;
                best.cost = cost;
            }
        }
    }
    return best;
}
// Ring geometry of the phased kernel for a layout: four stages of ~12 KB (48 KB per CTA: four CTAs per
// SM still fit).  With the query-register kernel's own 3 x ~8 KB a thread of a many-phase layout gets only
// one or two chunks per tile; the deeper ring measured 1-3% faster at every shape (B200, n = 2^22:
// k=16, m=25: 220 -> 213 us; k=3, m=64: 92.1 -> 90.1 us), never slower.
struct FlexRing
{
    int stages, stage_floats, tile_groups, smem_bytes;
};
static FlexRing flex_ring(int k, const FlexLayout &L, int64_t deep)
{
    const FlexGeo fg = flex_geo(k);
    const int group_floats = fg.g * k, unit = L.np * fg.chg; // groups in one round of all phases
    FlexRing r;
    const bool many = deep > 0 || (deep < 0 && L.np > 8);
    r.stages = many ? 4 : 3;
    const int cap_groups = many ? std::max(unit, (12 * 1024 / 4) / group_floats) : fg.trg;
    r.tile_groups = std::max(1, cap_groups / unit) * unit;
    r.stage_floats = (r.tile_groups * group_floats + 3) / 4 * 4;
    while (r.stages > 2 && 128 + (int64_t)r.stages * r.stage_floats * 4 > 56 * 1024)
        --r.stages;
    r.smem_bytes = 128 + r.stages * r.stage_floats * 4;
    return r;
}

// Split count for a phased layout: same candidates and time model as plan_splits, with every thread
// visiting 1/np of its CTA's references and splits that are whole rounds of np * CH references.
static double plan_flex_splits(int k, const FlexLayout &L, int occ, int sms, int64_t n, int64_t wmax, int64_t forced,
                               int64_t *splits_out, int64_t *rps_out)
{
    const FlexGeo fg = flex_geo(k);
    const int64_t round = (int64_t)L.np * fg.ch;
    const int64_t smax = std::max<int64_t>(1, n / std::max<int64_t>(round, 32));
    std::vector<int64_t> cand;
    if (forced > 0)
        cand.push_back(std::min(forced, smax));
    else
    {
        cand.push_back(1);
        for (int64_t c = 1; c <= occ; ++c)
            cand.push_back(((int64_t)sms * c) / L.qtiles);
        for (int64_t w = 2; w <= std::max<int64_t>(1, wmax); ++w)
            cand.push_back(((int64_t)sms * occ * w) / L.qtiles);
    }
    const double f_q = L.q >= 8 ? 1.0 : (L.q >= 4 ? 0.988 : 0.965);
    const double cpr = (L.q / 2) * (3.0 * k - 1.0) * 2.0 / f_q / (double)L.np;
    auto eff = [](int64_t c) { return c <= 1 ? 0.58 : (c == 2 ? 0.78 : (c == 3 ? 0.86 : 0.92)); };
    const double lsu = 5.0 * L.q * k * 32.0 * 4.0 * 0.5, ovh = 6000.0; // scalar query loads + chunk re-reads; phase merge
    double best = -1.0;
    for (int64_t sp : cand)
    {
        sp = std::max<int64_t>(1, std::min(sp, smax));
        int64_t rps = (n + sp - 1) / sp;
        rps = (rps + round - 1) / round * round;
        const int64_t used = (n + rps - 1) / rps;
        const int64_t total = used * L.qtiles, slots = (int64_t)sms * occ;
        const int64_t full_waves = total / slots, rem = total % slots;
        double t = (double)full_waves * (ovh + (double)occ * (lsu + (double)rps * cpr / eff(occ)));
        if (rem)
        {
            const int64_t c = (rem + sms - 1) / sms;
            t += ovh + (double)c * (lsu + (double)rps * cpr / eff(c));
        }
        if (best < 0 || t < best * 0.999)
        {
            best = t;
            *splits_out = std::max<int64_t>(1, used);
            *rps_out = rps;
        }
    }
    return best;
}

// Which kernel family searches m queries against n references (pure arithmetic).
//   m <= 4: purely HBM-bound, plain register loads stream fastest (6.9 TB/s)  -> 2, reference-register
//   above that, three families compete and the one with the smallest modelled time wins:
//   the reference-stream kernel (4) re-streams the set once per pass of 8 queries; the query-register
//   kernel (1) streams it once but computes on padded 128-query tiles; the phased query-register
//   kernel (5) lays the 128 threads of a CTA out as query groups x reference phases, so neither a
//   padded tile nor a re-streamed set is paid for, at a higher fixed cost.
static int auto_variant(int k, int m, int64_t n)
{
    if (m <= 4)
        return 2;
    // Modelled kernel time in microseconds of each family that can take the shape (fitted to the B200
    // sweep profiles/r02_fewquery_crossover.json: k = 3/8/16, m = 5..500, n = 2^16/2^20/2^22; the model
    // is within ~10% of every measured point that decides a pick):
    //   a lane-cost unit = one thread visiting every reference with one query pair; it costs
    //   3.6 (3k-1) us per 10^6 references (all-packed math at ~90% of the pipe, 4 CTAs per SM)
    const double mrefs = (double)n * 1e-6, x = mrefs * k;
    const double c = 3.6 * (3.0 * k - 1.0);
    const FlexGeo fg = flex_geo(k);
    // query-register kernel: best padded tile; one query per thread has no pair to pack (0.80)
    double tile_cost = -1;
    for (int q : {1, 2, 4, 8})
    {
        if (q > fg.qmax)
            continue;
        const double f_q = q >= 8 ? 1.0 : (q >= 4 ? 0.988 : (q >= 2 ? 0.965 : 0.80));
        const double u = (double)((m + 128 * q - 1) / (128 * q)) * q / f_q;
        if (tile_cost < 0 || u < tile_cost)
            tile_cost = u;
    }
    const double t_tile = 7.0 + tile_cost * c * mrefs;
    int best = 1;
    double t_best = t_tile;
    // reference-stream kernel: one pass over the set per 8 queries
    if (m <= 64)
    {
        const double passes = (double)((m + 7) / 8);
        const double t_stream = 6.0 + passes * (4.5 + 0.80 * x);
        if (t_stream < t_best)
        {
            best = 4;
            t_best = t_stream;
        }
    }
    // phased query-register kernel: ~7% over its lane cost (per-lane reference addresses, phase
    // merge), a larger fixed part, and its short per-thread runs leave part of the HBM stream exposed
    if (m >= 9)
    {
        const FlexLayout L = flex_layout(k, m, 0);
        if (L.np >= 2)
        {
            const double t_flex = 15.0 + 0.3 * L.np + 1.07 * L.cost * c * mrefs + 0.6 * (x * 4.0 / 6.5);
            if (t_flex < t_best)
            {
                best = 5;
                t_best = t_flex;
            }
        }
    }
    return best;
}

extern "C" int nn_b200_plan_flex(int k, int m, int *q, int *groups, int *phases, int *qtiles, int *tile_queries)
{
    if (check_shape_quiet(k, m, 1) || m < 1 || !q || !groups || !phases || !qtiles || !tile_queries)
        return fail(NN_B200_EINVAL, "bad plan_flex arguments");
    const FlexLayout L = flex_layout(k, m, 0);
    if (L.q == 0)
        return fail(NN_B200_EINVAL, "no layout");
    *q = L.q;
    *groups = L.ng;
    *phases = L.np;
    *qtiles = (int)L.qtiles;
    *tile_queries = (int)L.tile_queries;
    return NN_B200_OK;
}

extern "C" int nn_b200_plan_variant(int k, int m, int64_t n)
{
    if (check_shape_quiet(k, m, n))
        return NN_B200_EINVAL;
    return auto_variant(k, m, n);
}

static int make_plan_uncached(int k, int m, int64_t n, bool soa, const DevInfo &di, Plan *p)
{
    int variant = (int)g_opt.variant.load();
    if (soa)
        variant = 2;
    // Few queries: every thread streams its own references (roles swapped).  Up to 4 queries the
    // search is purely HBM-bound and plain register loads stream fastest (6.9 TB/s); from 5 queries
    // on the FP32 pipe matters as well and the TMA-ring kernel wins at every k (B200 sweeps in
    // profiles/); beyond ~112 queries the query-register kernel's 128-query tile is full enough.
    if (variant == 0)
        variant = auto_variant(k, m, n);
    p->variant = variant;
    if (variant == 1)
    {
        // Candidates: queries/thread Q (tile = 128*Q queries) x reference splits.  Each is scored with
        // a small efficiency model -- query padding x SM fill / wave quantisation x per-split
        // prologue amortisation x a measured per-tile-shape factor -- and the best one wins.
        const int math = (int)g_opt.math.load();
        const int forced_q = (int)g_opt.qreg_q.load();
        const int64_t forced_splits = g_opt.splits.load();
        const int cand[4] = {0, 4, 2, 1}; // 0 = wide default for this k
        double best_score = -1.0;
        for (int ci = 0; ci < 4; ++ci)
        {
            const int qs = forced_q ? forced_q : cand[ci];
            LaunchInfo li{};
            int tq = 0, tr = 0;
            cudaError_t e = k_query_qreg(k, qs, math, &li, &tq, &tr);
            if (e != cudaSuccess)
            {
                (void)cudaGetLastError();
                if (forced_q)
                    return fail(NN_B200_EINVAL, "qreg_q=%d is not built for k=%d", forced_q, k);
                continue;
            }
            const int q = tq / 128;
            const int occ = li.occ > 0 ? li.occ : 1;
            const int64_t qtiles = ((int64_t)m + tq - 1) / tq;
            int64_t spl = 1, rps = 0;
            const double cyc = plan_splits(k, q, occ, di.sms, m, n, g_opt.waves.load(), forced_splits, &spl, &rps);
            if (best_score < 0 || cyc < best_score)
            {
                best_score = cyc;
                p->q = qs;
                p->scalar = math;
                p->tile_q = tq;
                p->tile_r = tr;
                p->occ = occ;
                p->regs = li.regs;
                p->qtiles = (uint32_t)qtiles;
                p->refs_per_split = (uint32_t)rps;
                p->splits = (uint32_t)spl;
            }
            if (forced_q)
                break;
        }
        if (best_score < 0)
            return fail(NN_B200_ECUDA, "no query-register kernel available for k=%d", k);
    }
    else if (variant == 5)
    {
        const FlexLayout L = flex_layout(k, m, (int)g_opt.qreg_q.load());
        if (L.q == 0)
            return fail(NN_B200_EINVAL, "no phased query-register layout for k=%d m=%d (qreg_q=%d)", k, m,
                        (int)g_opt.qreg_q.load());
        const FlexRing ring = flex_ring(k, L, g_opt.flex_deep_ring.load());
        FlexInfo fi{};
        cudaError_t e = k_query_qflex(k, L.q, ring.smem_bytes, &fi);
        if (e != cudaSuccess)
            return fail(NN_B200_ECUDA, "phased query-register kernel query failed for k=%d q=%d: %s", k, L.q,
                        cudaGetErrorString(e));
        const FlexGeo fg = flex_geo(k);
        if (fg.ch != fi.ch || fg.g != fi.g || fg.trg != fi.tr / fi.g)
            return fail(NN_B200_ECUDA, "host and device disagree on the tile geometry for k=%d", k);
        int64_t spl = 1, rps = 0;
        const int occ = fi.occ > 0 ? fi.occ : 1;
        plan_flex_splits(k, L, occ, di.sms, n, g_opt.waves.load(), g_opt.splits.load(), &spl, &rps);
        p->q = L.q;
        p->ng = L.ng;
        p->np = L.np;
        p->occ = occ;
        p->regs = fi.regs;
        p->qtiles = (uint32_t)L.qtiles;
        p->tile_queries = (uint32_t)L.tile_queries;
        p->tile_groups = (uint32_t)ring.tile_groups;
        p->stages = (uint32_t)ring.stages;
        p->stage_floats = (uint32_t)ring.stage_floats;
        p->splits = (uint32_t)spl;
        p->refs_per_split = (uint32_t)rps;
    }
    else if (variant == 2)
    {
        LaunchInfo li{};
        int rpb = 0;
        cudaError_t e = k_query_rreg(k, 8, soa, &li, &rpb);
        if (e != cudaSuccess)
            return fail(NN_B200_ECUDA, "reference-register kernel query failed for k=%d: %s", k, cudaGetErrorString(e));
        int per_sm = (int)g_opt.rreg_ctas_per_sm.load();
        if (per_sm <= 0)
            per_sm = li.occ > 0 ? li.occ : 1;
        p->occ = per_sm;
        p->regs = li.regs;
        p->rreg_ctas = di.sms * per_sm;
    }
    else if (variant == 4)
    {
        LaunchInfo li{};
        int tr = 0;
        cudaError_t e = k_query_rtma(k, 8, &li, &tr);
        if (e != cudaSuccess)
            return fail(NN_B200_ECUDA, "reference-stream kernel query failed for k=%d: %s", k, cudaGetErrorString(e));
        int per_sm = (int)g_opt.rreg_ctas_per_sm.load();
        if (per_sm <= 0)
            per_sm = li.occ > 0 ? li.occ : 1;
        p->occ = per_sm;
        p->regs = li.regs;
        p->tile_refs = tr;
        // persistent grid, but never more CTAs than full tiles (+1 so that the ragged end has an owner)
        const int64_t tiles = n / tr + 1;
        p->rreg_ctas = (int)std::min<int64_t>((int64_t)di.sms * per_sm, tiles);
    }
    else if (variant == 3)
    {
        const int64_t qblocks = ((int64_t)m + 127) / 128;
        int64_t s = ((int64_t)di.sms * 16 + qblocks - 1) / qblocks;
        s = std::max<int64_t>(1, std::min<int64_t>(s, std::max<int64_t>(1, n / 256)));
        p->plain_splits = (uint32_t)s;
    }
    else
        return fail(NN_B200_EINVAL, "unknown variant %d", variant);
    return NN_B200_OK;
}

// Planning asks the runtime for register counts and occupancies of several candidate kernels; that
// costs tens of microseconds on the host, more than a small search takes on the device.  Plans are
// therefore remembered per (device, shape, option epoch).
struct PlanKey
{
    int dev, k, m, soa;
    int64_t n, epoch;
    bool operator==(const PlanKey &o) const
    {
        return dev == o.dev && k == o.k && m == o.m && soa == o.soa && n == o.n && epoch == o.epoch;
    }
};
static std::mutex g_plan_mu;
static std::vector<std::pair<PlanKey, Plan>> g_plans;

static int make_plan(int dev, int k, int m, int64_t n, bool soa, const DevInfo &di, Plan *p)
{
    const PlanKey key{dev, k, m, soa ? 1 : 0, n, g_opt_epoch.load()};
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        for (size_t i = 0; i < g_plans.size(); ++i)
            if (g_plans[i].first == key)
            {
                *p = g_plans[i].second;
                return NN_B200_OK;
            }
    }
    const int rc = make_plan_uncached(k, m, n, soa, di, p);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (g_plans.size() >= 64)
        g_plans.erase(g_plans.begin());
    g_plans.emplace_back(key, *p);
    return NN_B200_OK;
}

static int check_shape(int k, int m, int64_t n)
{
    if (k < NN_B200_KMIN || k > NN_B200_KMAX)
        return fail(NN_B200_EINVAL, "k=%d outside %d..%d", k, NN_B200_KMIN, NN_B200_KMAX);
    if (m < 0 || n < 0)
        return fail(NN_B200_EINVAL, "negative size m=%d n=%lld", m, (long long)n);
    if (n > 0x7fffffffLL)
        return fail(NN_B200_EINVAL, "n=%lld does not fit the reference's int interface", (long long)n);
    return NN_B200_OK;
}

// peer_override: -1 = look at where d_keys lives; 1 = other GPUs fold into the same array
// concurrently, so even its owner must use system-scope atomics.
// fin: non-null = finish inside the launch (ticket protocol, struct Finish); *fin_done tells the
// caller whether the kernels did it (false: the plan has no single finishing launch -- the plain
// kernel -- and the caller must run nn_keys_finish_kernel itself).
static int nearest_keys_impl(int k, int m, int64_t n, const float *d_S, const float *d_R, uint32_t index_base,
                             uint64_t *d_keys, void *stream, bool soa, int peer_override = -1,
                             const Finish *fin = nullptr, bool *fin_done = nullptr)
{
    if (fin_done)
        *fin_done = false;
    int rc = check_shape(k, m, n);
    if (rc)
        return rc;
    if (m == 0 || n == 0)
        return NN_B200_OK;
    if (!d_S || !d_R || !d_keys)
        return fail(NN_B200_EINVAL, "null device pointer");
    if ((reinterpret_cast<uintptr_t>(d_R) & 15) != 0)
        return fail(NN_B200_EINVAL, "reference pointer must be 16-byte aligned");
    if ((uint64_t)index_base + (uint64_t)n > 0x100000000ull)
        return fail(NN_B200_EINVAL, "index_base + n exceeds 32 bits");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevInfo di;
    rc = dev_info(dev, &di);
    if (rc)
        return rc;
    Plan p;
    rc = make_plan(dev, k, m, n, soa, di, &p);
    if (rc)
        return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(d_keys);
    // keys in another GPU's memory (peer-mapped): the kernels fold with system-scope atomics
    int peer = peer_override > 0 ? 1 : 0;
    if (peer_override < 0)
    {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, d_keys) == cudaSuccess)
            peer = (pa.type == cudaMemoryTypeDevice && pa.device != dev) ? 1 : 0;
        else
            (void)cudaGetLastError();
    }
    // multi-process merge (struct PeerSync): the usual per-group finish, whose last CTA pushes the group's
    // keys to rank 0 instead of unpacking them; the call's groups are counted to know when the rank is done
    Finish peer_fin;
    const bool peer_mode = fin && fin->peer.arrive != nullptr;
    if (peer_mode)
    {
        unsigned int groups = 0;
        if (p.variant == 1 || p.variant == 5)
            groups = p.qtiles;
        else if (p.variant == 2 || p.variant == 4)
            groups = (unsigned int)(m / 8 + (m % 8 ? 1 : 0));
        else
            return fail(NN_B200_EINVAL, "the plain kernel has no in-kernel finish (multi-process merge)");
        if (groups > (unsigned int)kWsTickets)
            return fail(NN_B200_EINVAL, "m=%d needs %u ticket groups, more than the %d of a workspace", m, groups, kWsTickets);
        peer_fin = *fin;
        peer_fin.peer.num_groups = groups;
        fin = &peer_fin;
        peer = 0; // folds go to this rank's own workspace
    }
    if (p.variant == 1)
    {
        QregArgs a;
        a.S = d_S;
        a.R = d_R;
        a.m = m;
        a.n = (uint32_t)n;
        a.index_base = index_base;
        a.splits = p.splits;
        a.qgroup = (uint32_t)std::max<int64_t>(1, g_opt.qgroup.load());
        a.refs_per_split = p.refs_per_split;
        a.super_chunks = qreg_super_for(k, p.refs_per_split);
        a.keys = keys;
        a.neg_zero = -0.0f;
        a.peer_keys = peer;
        if (fin && !peer && p.qtiles <= (uint32_t)kWsTickets)
        {
            a.fin = *fin;
            if (fin_done)
                *fin_done = true;
        }
#ifdef NN_QREG_TIMELINE
        a.timeline = g_timeline;
#endif
        CU(k_launch_qreg(k, p.q, p.scalar, a, p.qtiles, st));
        g_launches++;
    }
    else if (p.variant == 5)
    {
        QflexArgs a;
        a.S = d_S;
        a.R = d_R;
        a.m = m;
        a.n = (uint32_t)n;
        a.index_base = index_base;
        a.splits = p.splits;
        a.qgroup = (uint32_t)std::max<int64_t>(1, g_opt.qgroup.load());
        a.refs_per_split = p.refs_per_split;
        a.tile_queries = p.tile_queries;
        a.ng = (uint32_t)p.ng;
        a.np = (uint32_t)p.np;
        a.tile_groups = p.tile_groups;
        a.stages = p.stages;
        a.stage_floats = p.stage_floats;
        a.keys = keys;
        a.neg_zero = -0.0f;
        a.peer_keys = peer;
        if (fin && !peer && p.qtiles <= (uint32_t)kWsTickets)
        {
            a.fin = *fin;
            if (fin_done)
                *fin_done = true;
        }
        CU(k_launch_qflex(k, p.q, a, p.qtiles, st));
        g_launches++;
    }
    else if (p.variant == 2 || p.variant == 4)
    {
        // full passes of 8 queries, then one pass with the smallest even width covering the tail
        const bool rtma = p.variant == 4;
        const bool use_fin = fin && !peer && (m + 7) / 8 + 1 <= kWsTickets;
        int ticket_off = 0; // one ticket per pass, numbered across the launches of this call
        auto launch = [&](int q0, int count, int mq) -> int {
            int done = 0;
            const int passes = (count + mq - 1) / mq;
            while (done < passes)
            {
                const int py = std::min(passes - done, 65535); // gridDim.y limit
                RregArgs a;
                a.S = d_S + (size_t)(q0 + done * mq) * k;
                a.R = d_R;
                a.mq_total = std::min(count - done * mq, py * mq);
                a.n = (uint32_t)n;
                a.index_base = index_base;
                a.keys = keys + q0 + done * mq;
                a.neg_zero = -0.0f;
                a.peer_keys = peer;
                a.q_first = q0 + done * mq;
                if (use_fin)
                {
                    a.fin = *fin;
                    a.fin.tickets = fin->tickets + ticket_off;
                    if (!peer_mode)
                    { // (multi-process mode: results belong to rank 0's final unpack of ALL queries)
                        a.fin.results = fin->results ? fin->results + q0 + done * mq : nullptr;
                        a.fin.keys_out = fin->keys_out ? fin->keys_out + q0 + done * mq : nullptr;
                    }
                    ticket_off += py;
                }
                if (rtma)
                    CU(k_launch_rtma(k, mq, a, dim3((unsigned)p.rreg_ctas, (unsigned)py), st));
                else
                    CU(k_launch_rreg(k, mq, soa, a, dim3((unsigned)p.rreg_ctas, (unsigned)py), st));
                g_launches++;
                done += py;
            }
            return NN_B200_OK;
        };
        const int full = (m / 8) * 8, tail = m - full;
        if (full)
        {
            rc = launch(0, full, 8);
            if (rc)
                return rc;
        }
        if (tail)
        {
            rc = launch(full, tail, tail > 4 ? 8 : (tail > 2 ? 4 : 2));
            if (rc)
                return rc;
        }
        if (use_fin && fin_done)
            *fin_done = true;
    }
    else
    {
        CU(k_launch_plain(k, d_S, d_R, m, (uint32_t)n, index_base, p.plain_splits, keys, peer, st));
        g_launches++;
    }
    return NN_B200_OK;
}

extern "C" int nn_b200_nearest_keys(int k, int m, int64_t n, const float *d_S, const float *d_R, uint32_t index_base,
                                    uint64_t *d_keys, void *stream)
{
    return nearest_keys_impl(k, m, n, d_S, d_R, index_base, d_keys, stream, false);
}

extern "C" int nn_b200_nearest_keys_soa(int k, int m, int64_t n, const float *d_S, const float *d_R_soa,
                                        uint32_t index_base, uint64_t *d_keys, void *stream)
{
    return nearest_keys_impl(k, m, n, d_S, d_R_soa, index_base, d_keys, stream, true);
}

extern "C" int nn_b200_describe_plan(int k, int m, int64_t n, char *buf, size_t len)
{
    int rc = check_shape(k, m, n);
    if (rc)
        return rc;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevInfo di;
    rc = dev_info(dev, &di);
    if (rc)
        return rc;
    Plan p;
    rc = make_plan(dev, k, m, n, false, di, &p);
    if (rc)
        return rc;
    if (p.variant == 1)
        snprintf(buf, len,
                 "qreg k=%d Q=%d %s tile=%dq x %dr regs=%d occ=%d qtiles=%u splits=%u refs/split=%u ctas=%u sms=%d super=%u", k,
                 p.q, p.scalar == 0 ? "scalar" : (p.scalar == 1 ? "f32x2-dims" : "f32x2-pairs"), p.tile_q, p.tile_r, p.regs, p.occ, p.qtiles, p.splits,
                 p.refs_per_split, p.qtiles * p.splits, di.sms, p.scalar == 2 ? qreg_super_for(k, p.refs_per_split) : 1u);
    else if (p.variant == 5)
        snprintf(buf, len,
                 "qflex k=%d Q=%d groups=%d phases=%d tile=%uq x %ug ring=%ux%uK regs=%d occ=%d qtiles=%u splits=%u refs/split=%u "
                 "ctas=%u sms=%d",
                 k, p.q, p.ng, p.np, p.tile_queries, p.tile_groups, p.stages, p.stage_floats * 4 / 1024, p.regs, p.occ, p.qtiles, p.splits, p.refs_per_split,
                 p.qtiles * p.splits, di.sms);
    else if (p.variant == 2)
        snprintf(buf, len, "rreg k=%d f32x2-pairs regs=%d ctas/sm=%d ctas=%d passes8=%d tail=%d sms=%d", k, p.regs,
                 p.occ, p.rreg_ctas, m / 8, m % 8, di.sms);
    else if (p.variant == 4)
        snprintf(buf, len, "rtma k=%d f32x2-pairs regs=%d ctas/sm=%d ctas=%d tile=%dr passes8=%d tail=%d sms=%d", k,
                 p.regs, p.occ, p.rreg_ctas, p.tile_refs, m / 8, m % 8, di.sms);
    else
        snprintf(buf, len, "plain k=%d splits=%u sms=%d", k, p.plain_splits, di.sms);
    return NN_B200_OK;
}

extern "C" int nn_b200_keys_init(uint64_t *d_keys, int m, void *stream)
{
    if (m < 0)
        return fail(NN_B200_EINVAL, "negative m");
    if (m == 0)
        return NN_B200_OK;
    if (!d_keys)
        return fail(NN_B200_EINVAL, "null keys pointer");
    nn_keys_init_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long *>(d_keys),
                                                                           m);
    CU(cudaGetLastError());
    g_launches++;
    return NN_B200_OK;
}

extern "C" int nn_b200_keys_unpack(const uint64_t *d_keys, int m, int *d_results, void *stream)
{
    if (m < 0)
        return fail(NN_B200_EINVAL, "negative m");
    if (m == 0)
        return NN_B200_OK;
    if (!d_keys || !d_results)
        return fail(NN_B200_EINVAL, "null pointer");
    nn_keys_unpack_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long *>(d_keys), m, d_results);
    CU(cudaGetLastError());
    g_launches++;
    return NN_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// one-launch search: workspace = m packed keys in the start state + ticket counters (all zero)
// ---------------------------------------------------------------------------------------------
// Layout: [kWsTickets ticket counters][m packed keys].  The ticket region has a fixed size, so a
// workspace initialised for m queries serves any search of up to m queries.
static unsigned int *ws_tickets(void *d_ws) { return reinterpret_cast<unsigned int *>(d_ws); }
static uint64_t *ws_keys(void *d_ws)
{
    return reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(d_ws) + (size_t)kWsTickets * 4);
}

extern "C" size_t nn_b200_workspace_bytes(int m)
{
    if (m < 0)
        return 0;
    return (size_t)kWsTickets * 4 + (size_t)m * 8;
}

extern "C" uint64_t *nn_b200_workspace_keys(void *d_ws) { return d_ws ? ws_keys(d_ws) : nullptr; }

extern "C" int nn_b200_workspace_init(void *d_ws, int m, void *stream)
{
    if (m < 0)
        return fail(NN_B200_EINVAL, "negative m");
    if (!d_ws || (reinterpret_cast<uintptr_t>(d_ws) & 15) != 0)
        return fail(NN_B200_EINVAL, "workspace pointer null or not 16-byte aligned");
    CU(cudaMemsetAsync(ws_tickets(d_ws), 0, (size_t)kWsTickets * 4, (cudaStream_t)stream));
    return nn_b200_keys_init(ws_keys(d_ws), m, stream);
}

extern "C" int nn_b200_workspace_finish(void *d_ws, int m, int *d_results, uint64_t *d_keys_out, void *stream)
{
    if (m < 0)
        return fail(NN_B200_EINVAL, "negative m");
    if (m == 0)
        return NN_B200_OK;
    if (!d_ws)
        return fail(NN_B200_EINVAL, "null workspace");
    nn_keys_finish_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<unsigned long long *>(ws_keys(d_ws)), m, d_results,
        reinterpret_cast<unsigned long long *>(d_keys_out));
    CU(cudaGetLastError());
    g_launches++;
    return NN_B200_OK;
}

static int search_device_impl(int k, int m, int64_t n, const float *d_S, const float *d_R, uint32_t index_base,
                              void *d_ws, int *d_results, uint64_t *d_keys_out, void *stream)
{
    int rc = check_shape(k, m, n);
    if (rc)
        return rc;
    if (m == 0)
        return NN_B200_OK;
    if (!d_ws || (reinterpret_cast<uintptr_t>(d_ws) & 15) != 0)
        return fail(NN_B200_EINVAL, "workspace pointer null or not 16-byte aligned");
    if (!d_results && !d_keys_out)
        return fail(NN_B200_EINVAL, "neither results nor keys_out given");
    Finish fin;
    fin.tickets = ws_tickets(d_ws);
    fin.results = d_results;
    fin.keys_out = reinterpret_cast<unsigned long long *>(d_keys_out);
    bool done = false;
    rc = nearest_keys_impl(k, m, n, d_S, d_R, index_base, ws_keys(d_ws), stream, false, 0, &fin, &done);
    if (rc)
        return rc;
    if (!done) // n = 0 (every query keeps v0's start state, index 0) or a plan without an in-kernel finish
        return nn_b200_workspace_finish(d_ws, m, d_results, d_keys_out, stream);
    return NN_B200_OK;
}

extern "C" int nn_b200_search_device(int k, int m, int64_t n, const float *d_S, const float *d_R, uint32_t index_base,
                                     void *d_ws, int *d_results, uint64_t *d_keys_out, void *stream)
{
    return search_device_impl(k, m, n, d_S, d_R, index_base, d_ws, d_results, d_keys_out, stream);
}

extern "C" int nn_b200_repack_soa(int k, int64_t n, const float *d_in, float *d_out, void *stream)
{
    int rc = check_shape(k, 0, n);
    if (rc)
        return rc;
    if (n == 0)
        return NN_B200_OK;
    if (!d_in || !d_out)
        return fail(NN_B200_EINVAL, "null pointer");
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) != 0)
        return fail(NN_B200_EINVAL, "input pointer must be 16-byte aligned");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevInfo di;
    rc = dev_info(dev, &di);
    if (rc)
        return rc;
    CU(k_launch_repack(k, d_in, d_out, (uint32_t)n, di.sms, (cudaStream_t)stream));
    g_launches++;
    return NN_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// warm-up: the reference hides its cold start behind a static WarmUP object that runs every version
// once before main() (core.cu:1274, README.md:222).  Here the cost of a first use is the lazy
// loading of each kernel's code; nn_b200_warmup() loads every search kernel of the current device
// (all k, all tile shapes) by querying its attributes, so that no later call pays for it.
// ---------------------------------------------------------------------------------------------
extern "C" int nn_b200_warmup(void)
{
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevInfo di;
    int rc = dev_info(dev, &di);
    if (rc)
        return rc;
    for (int k = NN_B200_KMIN; k <= NN_B200_KMAX; ++k)
    {
        LaunchInfo li{};
        int a = 0, b = 0;
        for (int qs : {0, 4, 2, 1})
            if (k_query_qreg(k, qs, 2, &li, &a, &b) != cudaSuccess)
                (void)cudaGetLastError(); // tile width not built for this k
        for (int mq : {2, 4, 8})
        {
            if (k_query_rreg(k, mq, false, &li, &a) != cudaSuccess || k_query_rtma(k, mq, &li, &a) != cudaSuccess)
                (void)cudaGetLastError();
        }
        FlexInfo fi{};
        for (int q : {2, 4, 8})
            if (k_query_qflex(k, q, 32 * 1024, &fi) != cudaSuccess)
                (void)cudaGetLastError(); // 8 queries per thread: k <= 8 only
    }
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, nn_keys_init_kernel));
    CU(cudaFuncGetAttributes(&fa, nn_keys_unpack_kernel));
    return NN_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// sharding helpers
// ---------------------------------------------------------------------------------------------
extern "C" int nn_b200_shard_range(int64_t n, int num_shards, int shard, int64_t *begin, int64_t *count)
{
    if (n < 0 || num_shards < 1 || shard < 0 || shard >= num_shards || !begin || !count)
        return fail(NN_B200_EINVAL, "bad shard arguments");
    // ceil(n / shards) like core.cu:875, rounded up to 4 points so every shard start is 16-byte
    // aligned for any k; the last shards take the remainder and may be empty.
    int64_t per = (n + num_shards - 1) / num_shards;
    per = (per + 3) / 4 * 4;
    const int64_t b = std::min<int64_t>(n, per * shard);
    const int64_t e = std::min<int64_t>(n, b + per);
    *begin = b;
    *count = e - b;
    return NN_B200_OK;
}

static int visible_devices()
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess)
    {
        (void)cudaGetLastError();
        return 0;
    }
    return cnt;
}

extern "C" int nn_b200_device_count(int64_t n)
{
    int cnt = visible_devices();
    const char *cap = getenv("NN_B200_GPUS");
    if (cap && atoi(cap) > 0)
        cnt = std::min(cnt, atoi(cap));
    if ((int64_t)cnt > n) // core.cu:867-868
        cnt = (int)std::max<int64_t>(n, 1);
    return cnt;
}

// How many GPUs a host-entry call should use (pure arithmetic; SURVEY a6, the analogue of the
// reference's small-n single-GPU shortcut `n <= min(1<<18, m<<10) -> v7`, core.cu:871-872, which
// overflows for m >= 2^21 and ignores k).  Sharding divides the copy of the reference set (every GPU
// pulls its shard over its own PCIe link, staged by its own host thread) and the search, but costs a
// host thread per GPU, cross-device event waits and the merge; below a few hundred microseconds of
// single-GPU work that overhead is the larger part.  Model, in microseconds:
//   t(G) = [G > 1] (kMultiBase + kMultiPerGpu G) + bytes(R)/G / copy_rate + max(3k m n / fp32, 4k n / hbm) / G
// with the constants measured on the 8x B200 box (profiles/r02_gpu_count.json); the smallest G within
// 3% of the best modelled time wins.
static const double kMultiBaseUs = 45.0, kMultiPerGpuUs = 6.0;
static int plan_gpus(int k, int m, int64_t n, int visible, bool pinned)
{
    if (visible <= 1 || n <= 1 || m <= 0)
        return visible < 1 ? visible : 1;
    const int64_t cap = std::min<int64_t>(visible, n);
    const double bytes_r = (double)n * k * 4.0;
    const double copy_us_per_byte = 1e6 / (pinned ? 52e9 : (bytes_r >= (double)(16 << 20) ? 40e9 : 11e9));
    const double fp32_us = 3.0 * k * (double)m * (double)n / (0.85 * 37.2e12) * 1e6;
    const double hbm_us = bytes_r / 6.0e12 * 1e6;
    const double work = std::max(fp32_us, hbm_us);
    double best = -1;
    std::vector<double> t((size_t)cap + 1, 0.0);
    for (int64_t g = 1; g <= cap; ++g)
    {
        t[g] = (g > 1 ? kMultiBaseUs + kMultiPerGpuUs * (double)g : 0.0) + (bytes_r * copy_us_per_byte + work) / (double)g;
        if (best < 0 || t[g] < best)
            best = t[g];
    }
    for (int64_t g = 1; g <= cap; ++g)
        if (t[g] <= best * 1.03)
            return (int)g;
    return (int)cap;
}

// Search launches over one shard's H2D chunk list (pure arithmetic): ends[g] = one past the last chunk of
// launch g.  The copies stay small (early start, double buffers), but once the pipeline runs several chunks
// are searched by ONE launch -- every launch of a big search has its own wave tail, and config 4's 1 GiB in
// 4 MiB chunks meant 260 launches (e2e 1538 ms for a 1473 ms search; 16 MiB pinned chunks: 1487 ms; a fixed
// 2 chunks per launch: 1476 ms).  The first four chunks go one by one; then
//  * a fixed few (`search_group`, at most 1/24 of the list) when the copy is the slower side, so that the
//    search of the last group, which nothing overlaps, stays short;
//  * when the search of a reference takes much longer than its copy (many queries) the copies run far
//    ahead of the searches, and a launch takes every chunk that must have landed by the time it starts:
//    `ahead` references copied per reference searched, from the FP32 bound of the search and a pessimistic
//    8 GB/s of copy shared by the `gpus` GPUs of the call, halved.  Config 4 on one GPU (64 chunks):
//    34 launches -> 7, each long enough for the query-register kernel's super-chunk form (e2e 1476.5 ->
//    1464-1467 ms for a 1463-1464 ms search); on 8 GPUs the shards keep the fixed groups.
static std::vector<size_t> plan_search_groups(int k, int m, const std::vector<std::pair<int64_t, int64_t>> &chunks,
                                              int64_t search_group, int gpus)
{
    const size_t nchunks = chunks.size();
    const size_t group_max = (size_t)std::max<int64_t>(1, std::min<int64_t>(search_group, (int64_t)nchunks / 24));
    const double copy_rate = 8e9 / (double)std::max(1, gpus); // bytes/s per GPU
    const double ahead =
        search_group > 1 ? 0.5 * (3.0 * k * (double)m / (0.9 * 37.2e12)) / ((double)k * sizeof(float) / copy_rate) : 0.0;
    std::vector<size_t> ends;
    for (size_t ci = 0; ci < nchunks;)
    {
        size_t ce = std::min(nchunks, ci + (ci < 4 ? (size_t)1 : group_max));
        if (ci >= 4 && ahead >= 4.0)
        {
            const double landed = (double)chunks[ci].first * ahead; // references searched so far x ahead
            while (ce < nchunks && (double)(chunks[ce].first + chunks[ce].second) <= landed)
                ++ce;
        }
        ends.push_back(ce);
        ci = ce;
    }
    return ends;
}

extern "C" int nn_b200_plan_search_groups(int k, int m, int gpus, const int64_t *chunk_refs, int nchunks, int *group_ends)
{
    if (k < 3 || k > 16 || m < 0 || gpus < 1 || nchunks < 0 || (nchunks > 0 && (!chunk_refs || !group_ends)))
        return fail(NN_B200_EINVAL, "bad arguments to nn_b200_plan_search_groups");
    std::vector<std::pair<int64_t, int64_t>> chunks;
    int64_t off = 0;
    for (int i = 0; i < nchunks; ++i)
    {
        if (chunk_refs[i] <= 0)
            return fail(NN_B200_EINVAL, "chunk %d holds %lld references", i, (long long)chunk_refs[i]);
        chunks.emplace_back(off, chunk_refs[i]);
        off += chunk_refs[i];
    }
    const std::vector<size_t> ends = plan_search_groups(k, m, chunks, g_opt.search_group.load(), gpus);
    for (size_t g = 0; g < ends.size(); ++g)
        group_ends[g] = (int)ends[g];
    return (int)ends.size();
}

extern "C" int nn_b200_plan_gpus(int k, int m, int64_t n, int visible)
{
    if (check_shape_quiet(k, m, n) || visible < 0)
        return NN_B200_EINVAL;
    return plan_gpus(k, m, n, visible, false);
}

// ---------------------------------------------------------------------------------------------
// NCCL, bound lazily with dlopen so that single-GPU use never loads it
// ---------------------------------------------------------------------------------------------
namespace
{
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
// values fixed by nccl.h's public enums (NCCL 2.x): ncclUint64 = 5, ncclMin = 3
constexpr int kNcclUint64 = 5;
constexpr int kNcclMin = 3;
struct Nccl
{
    void *h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl g_nccl;

int load_nccl()
{
    if (g_nccl.h)
        return NN_B200_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names)
    {
        g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h)
            break;
    }
    if (!g_nccl.h)
        return fail(NN_B200_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                                               \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                                                                 \
    if (!g_nccl.field)                                                                                                 \
        return fail(NN_B200_ENCCL, "libnccl lacks %s", name);
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return NN_B200_OK;
}
} // namespace

// ---------------------------------------------------------------------------------------------
// persistent host worker threads.  A host-entry call uses one thread per GPU (as v8 does with OpenMP,
// core.cu:873) and, for pageable inputs, several staging threads per GPU; creating them per call cost
// 50-100 us each -- a large part of a call that takes a millisecond.  The pools grow on demand, keep
// their threads parked on a condition variable and are never destroyed (no static-destruction order
// problems with threads that still wait).
// ---------------------------------------------------------------------------------------------
namespace
{
class TaskGroup
{
    std::mutex mu_;
    std::condition_variable cv_;
    int pending_ = 0;

  public:
    void add()
    {
        std::lock_guard<std::mutex> lk(mu_);
        ++pending_;
    }
    void done()
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0)
            cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return pending_ == 0; });
    }
    ~TaskGroup() { wait(); }
};

class WorkerPool
{
    struct State
    {
        std::mutex mu;
        std::condition_variable cv;
        std::vector<std::function<void()>> queue;
        int idle = 0, threads = 0;
    };
    State *st_ = new State; // leaked on purpose

  public:
    // runs f on a pool thread; g.wait() (or g's destructor) returns after it has finished
    void submit(TaskGroup &g, std::function<void()> f)
    {
        g.add();
        std::function<void()> task = [f = std::move(f), &g]() {
            f();
            g.done();
        };
        State *st = st_;
        std::unique_lock<std::mutex> lk(st->mu);
        st->queue.push_back(std::move(task));
        if ((int)st->queue.size() > st->idle) // more queued tasks than parked threads: one more thread
        {
            ++st->threads;
            lk.unlock();
            std::thread([st]() {
                for (;;)
                {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> l2(st->mu);
                        ++st->idle;
                        st->cv.wait(l2, [&] { return !st->queue.empty(); });
                        --st->idle;
                        job = std::move(st->queue.front());
                        st->queue.erase(st->queue.begin());
                    }
                    job();
                }
            }).detach();
        }
        else
        {
            lk.unlock();
            st->cv.notify_one();
        }
    }
};
WorkerPool g_gpu_workers;   // one task per GPU of a call
WorkerPool g_feed_workers;  // staging threads (submitted from within the GPU tasks: a pool of their own)
} // namespace

// ---------------------------------------------------------------------------------------------
// lazily created per-process context for the host entry point
// ---------------------------------------------------------------------------------------------
namespace
{
struct DevCtx
{
    int dev = -1;
    cudaStream_t compute = nullptr, copy = nullptr;
    std::vector<cudaEvent_t> events;
    float *dS = nullptr, *dR = nullptr;
    void *dWs = nullptr;                 // search workspace (tickets + keys), kept in its start state between calls
    unsigned long long *dKeys = nullptr; // = the workspace's key array
    int *dOut = nullptr;
    bool ws_dirty = false;  // a call is (or died) between folding into the workspace and finishing it
    bool finished = false;  // this call's search kernel already stored dOut and restored the workspace
    size_t capS = 0, capR = 0, capM = 0;
    int *hOut = nullptr; // pinned
    size_t capH = 0;
    cudaEvent_t done = nullptr; // end of this device's searches (P2P merge)
    cudaEvent_t keys_ready = nullptr; // device 0: its key array is initialised (P2P merge)
    int peer_to_0 = -1;         // -1 unknown, 0 no peer access to device 0, 1 enabled
    // pageable-input staging: per feeder thread one stream and two pinned chunk buffers
    static constexpr int kMaxFeeders = 8;
    cudaStream_t fstream[kMaxFeeders] = {};
    float *stage[kMaxFeeders][2] = {};
    cudaEvent_t stage_ev[kMaxFeeders][2] = {};
    size_t stage_cap[kMaxFeeders][2] = {};
    // small-batch graph path of the resident index: pinned copy of the queries, and a generation
    // that changes whenever one of the buffers a captured graph points at is re-allocated
    float *hS = nullptr;
    size_t capHS = 0;
    uint64_t generation = 0;
};
struct HostCtx
{
    std::mutex mu;
    std::vector<DevCtx> devs;
    std::vector<ncclComm_t> comms; // for the current device count
    int comm_gpus = 0;
};
HostCtx g_ctx;

int ensure_dev(DevCtx &c, int dev, size_t bytesS, size_t bytesR, size_t m, size_t nevents)
{
    CU(cudaSetDevice(dev));
    if (c.dev != dev)
    {
        if (!c.compute)
            CU(cudaStreamCreateWithFlags(&c.compute, cudaStreamNonBlocking));
        if (!c.copy)
            CU(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
        // first use of this device: load every kernel now (31 ms on B200) rather than a few
        // milliseconds at the first call of every new shape, as the reference's WarmUP does
        const char *w = getenv("NN_B200_WARMUP");
        if (!w || atoi(w) != 0)
        {
            const int rc = nn_b200_warmup();
            if (rc)
                return rc;
        }
        c.dev = dev; // only now: a context whose set-up failed half-way is set up again by the next call
    }
    // Buffers grow geometrically from generous floors (64 MiB of references, 4 MiB of queries, 64 Ki
    // results): a device or pinned allocation costs 1-3 ms, more than most small searches, and the
    // reference's harness walks through growing shapes (main.cu:28-39).
    auto grown = [](size_t want, size_t have, size_t floor_) { return std::max(std::max(want, floor_), have * 2); };
    if (bytesS > c.capS)
        bytesS = grown(bytesS, c.capS, (size_t)4 << 20);
    const size_t exactR = bytesR;
    if (bytesR > c.capR)
        bytesR = grown(bytesR, c.capR, (size_t)64 << 20);
    if (m > c.capM || m > c.capH)
        m = grown(m, std::max(c.capM, c.capH), (size_t)1 << 16);
    while (c.events.size() < nevents)
    {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c.events.push_back(e);
    }
    if (bytesS > c.capS)
    {
        if (c.dS)
            CU(cudaFree(c.dS));
        c.dS = nullptr;
        c.capS = 0;
        CU(cudaMalloc(&c.dS, bytesS));
        c.capS = bytesS;
        ++c.generation;
    }
    if (bytesR > c.capR)
    {
        if (c.dR)
            CU(cudaFree(c.dR));
        c.dR = nullptr;
        c.capR = 0;
        if (cudaMalloc(&c.dR, bytesR) != cudaSuccess)
        { // no room for the head-room: take exactly what this call needs
            (void)cudaGetLastError();
            bytesR = exactR;
            CU(cudaMalloc(&c.dR, bytesR));
        }
        c.capR = bytesR;
    }
    if (m > c.capM)
    {
        if (c.dWs)
            CU(cudaFree(c.dWs));
        if (c.dOut)
            CU(cudaFree(c.dOut));
        c.dWs = nullptr;
        c.dKeys = nullptr;
        c.dOut = nullptr;
        c.capM = 0;
        CU(cudaMalloc(&c.dWs, nn_b200_workspace_bytes((int)std::min<size_t>(m, 0x7fffffff))));
        CU(cudaMalloc(&c.dOut, m * sizeof(int)));
        c.dKeys = reinterpret_cast<unsigned long long *>(nn_b200_workspace_keys(c.dWs));
        c.capM = m;
        c.ws_dirty = true;
        ++c.generation;
    }
    if (c.ws_dirty)
    { // fresh allocation, or an earlier call failed half-way: (re)establish the start state
        const int rc = nn_b200_workspace_init(c.dWs, (int)std::min<size_t>(c.capM, 0x7fffffff), c.compute);
        if (rc)
            return rc;
        c.ws_dirty = false;
    }
    if (m > c.capH)
    {
        if (c.hOut)
            CU(cudaFreeHost(c.hOut));
        c.hOut = nullptr;
        c.capH = 0;
        CU(cudaMallocHost(&c.hOut, m * sizeof(int)));
        c.capH = m;
        ++c.generation;
    }
    return NN_B200_OK;
}

// Enqueue one device's share: queries, its contiguous reference shard in chunks (copy stream)
// and one search per chunk (compute stream).  Returns without synchronising.
// Is this host pointer ordinary pageable memory (malloc, as the reference's harness passes it,
// generator.h:37/44)?  Pinned or registered memory is copied by the DMA engines directly.
bool is_pageable(const void *p)
{
    cudaPointerAttributes pa{};
    if (cudaPointerGetAttributes(&pa, p) != cudaSuccess)
    {
        (void)cudaGetLastError();
        return true;
    }
    return pa.type == cudaMemoryTypeUnregistered;
}

int ensure_staging(DevCtx &c, int feeders, size_t chunk_bytes)
{
    for (int t = 0; t < feeders; ++t)
    {
        if (!c.fstream[t])
            CU(cudaStreamCreateWithFlags(&c.fstream[t], cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b)
        {
            if (!c.stage_ev[t][b])
                CU(cudaEventCreateWithFlags(&c.stage_ev[t][b], cudaEventDisableTiming));
            if (c.stage_cap[t][b] < chunk_bytes)
            {
                if (c.stage[t][b])
                    CU(cudaFreeHost(c.stage[t][b]));
                c.stage[t][b] = nullptr;
                c.stage_cap[t][b] = 0;
                CU(cudaMallocHost(&c.stage[t][b], chunk_bytes));
                c.stage_cap[t][b] = chunk_bytes;
            }
        }
    }
    return NN_B200_OK;
}

// `keys0`: non-null = fold into that (peer) key array, already initialised, once `keys_ready` has
// fired; null = this device's own key array, initialised here.
// NN_B200_TRACE=1: phase times of the staged ingest on stderr (diagnostics only)
static bool trace_on()
{
    static const bool on = [] {
        const char *e = getenv("NN_B200_TRACE");
        return e && atoi(e) != 0;
    }();
    return on;
}
static double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// `lone`: this device is the only one of the call (nothing to merge with).
int enqueue_device(DevCtx &c, int dev, int k, int m, const float *S, const float *R, int64_t begin, int64_t count,
                   unsigned long long *keys0, cudaEvent_t keys_ready, bool lone)
{
    const size_t bytesS = (size_t)m * k * sizeof(float);
    const size_t bytesR = (size_t)count * k * sizeof(float);
    // Pageable source (what the reference's harness passes): a cudaMemcpyAsync from it is staged by
    // the driver on one thread at ~11 GB/s.  From 48 MiB on, `feeders` host threads instead copy
    // alternate 4 MiB chunks into their own pinned double buffers and push them on their own
    // streams, so host copy, DMA and the search of earlier chunks all overlap.  Measured on the B200
    // box (16 cores), 2 GiB reference set: 193 ms (driver) -> 44 ms with 8 threads; pinned: 39 ms.
    // Threshold: 48 MiB (round 1: 128 MiB).  Measured on the B200 box, one GPU, ms per call, staged / driver
    // path / pinned: 4 MiB 1.03 / 0.35 / 0.33; 16 MiB 2.0 / 1.6 / 0.72; 32 MiB 3.0 / 2.2 / 1.2;
    // 64 MiB (BASELINE config 2) 6.7 / 7.0 / 6.2; 256 MiB 7.4 / 23.7 / 5.4.  Below ~48 MiB the driver's
    // own staging wins: a staged call carries ~1 ms that NN_B200_TRACE=1 shows is spent on the device
    // after everything has been enqueued (open item), and the chunks cost extra launches.
    int64_t want_feeders = g_opt.stage_threads.load();
    if (want_feeders < 0)
        // up to 8 threads; 4 below 96 MiB, where more threads only contend for the driver (64 MiB = BASELINE
        // config 2: 6.48 ms with 4, 7.05 ms with 8, 6.24 ms pinned; 256 MiB: 11.1 / 7.3 / 5.2 ms)
        want_feeders = bytesR >= (size_t)g_opt.stage_min_bytes.load()
                           ? std::max(2u, std::min(bytesR >= ((size_t)96 << 20) ? 8u : 4u,
                                                   std::thread::hardware_concurrency() / (unsigned)std::max(1, g_active_gpus.load())))
                           : 0;
    const bool staged = want_feeders > 0 && count > 0 && is_pageable(R);
    int64_t chunk_bytes = g_opt.h2d_chunk_bytes.load();
    // Staged path: 4 MiB chunks, ramping up -- unless the whole shard is only a few chunks: then every
    // staging thread gets ONE equal share (one thread copies at ~10 GB/s, so an 8 MiB set cut into
    // 0.25 .. 4 MiB chunks had its last 4 MiB chunk alone take 0.4 ms; many small chunks cost more in
    // launches and driver calls than they overlap: measured 0.83 / 1.38 ms for 8 MiB with 6 / 16 chunks).
    bool staged_even = false;
    if (staged)
    {
        chunk_bytes = std::min<int64_t>(chunk_bytes, 4 << 20);
        if ((int64_t)bytesR < 2 * want_feeders * chunk_bytes)
        {
            staged_even = true;
            chunk_bytes = ((int64_t)bytesR + want_feeders - 1) / want_feeders;
        }
    }
    int64_t chunk_refs = std::max<int64_t>(4096, (chunk_bytes + (int64_t)(k * sizeof(float)) - 1) / (int64_t)(k * sizeof(float)));
    chunk_refs = (chunk_refs + (staged_even ? 4095 : 0)) / 4096 * 4096; // keeps every chunk start 16-byte aligned and tile aligned
    // Chunk list.  Direct (pinned) path: the first chunks ramp up geometrically from 1/8 of the chunk
    // size, so the search starts after a short first copy instead of a whole chunk's.
    std::vector<std::pair<int64_t, int64_t>> chunks; // (first reference, count) relative to the shard
    {
        // (staged path too: the first 4 MiB chunk alone is ~0.4 ms of single-thread memcpy before the
        // GPU has anything to do; a ramp from 1/16 of the chunk size starts the search after ~30 us and
        // the first, small chunks are staged by different threads in parallel)
        int64_t step = staged_even ? chunk_refs : std::max<int64_t>(4096, chunk_refs / (staged ? 16 : 8) / 4096 * 4096);
        for (int64_t off = 0; off < count;)
        {
            const int64_t cnt = std::min<int64_t>(step, count - off);
            chunks.emplace_back(off, cnt);
            off += cnt;
            step = std::min<int64_t>(chunk_refs, step * 2);
        }
    }
    const size_t nchunks = chunks.size();
    int rc = ensure_dev(c, dev, std::max<size_t>(bytesS, 16), std::max<size_t>(bytesR, 16), (size_t)std::max(m, 1),
                        nchunks + 1);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(c.dS, S, bytesS, cudaMemcpyHostToDevice, c.copy));
    CU(cudaEventRecord(c.events[0], c.copy));
    // No init launch: the key workspace is in its start state between calls (whoever folds into it
    // restores it when finishing).  keys0: GPU 0's workspace, usable once `keys_ready` has fired.
    unsigned long long *keys = keys0 ? keys0 : c.dKeys;
    if (keys0)
        CU(cudaStreamWaitEvent(c.compute, keys_ready, 0));
    if (!keys0 || dev == 0)
        c.ws_dirty = true;
    c.finished = false;
    CU(cudaStreamWaitEvent(c.compute, c.events[0], 0));

    const int feeders = staged ? (int)std::min<int64_t>(std::min<int64_t>(want_feeders, DevCtx::kMaxFeeders), (int64_t)nchunks) : 0;
    std::mutex fmu;
    std::condition_variable fcv;
    std::vector<char> recorded(nchunks, 0);
    int frc = NN_B200_OK;
    std::string ferr;
    const double t_enq = now_us();
    const bool one_stream = g_opt.stage_one_stream.load() != 0;
    TaskGroup fgroup; // declared after everything the staging tasks touch: its destructor waits for them
    if (feeders > 0)
    {
        rc = ensure_staging(c, feeders, (size_t)chunk_refs * k * sizeof(float));
        if (rc)
            return rc;
        for (int t = 0; t < feeders; ++t)
            g_feed_workers.submit(fgroup, [&, t]() {
                auto bail = [&](cudaError_t e, size_t ci) {
                    std::lock_guard<std::mutex> lk(fmu);
                    if (frc == NN_B200_OK)
                    {
                        frc = NN_B200_ECUDA;
                        ferr = std::string("staged copy of chunk ") + std::to_string(ci) + ": " + cudaGetErrorString(e);
                    }
                    for (auto &r : recorded)
                        r = 1;
                    fcv.notify_all();
                };
                const double t_start = now_us();
                cudaError_t e = cudaSetDevice(dev);
                if (e != cudaSuccess)
                    return bail(e, 0);
                size_t use = 0;
                for (size_t ci = t; ci < nchunks; ci += feeders, ++use)
                {
                    const int b = (int)(use & 1);
                    const int64_t off = chunks[ci].first, cnt = chunks[ci].second;
                    const size_t bytes = (size_t)cnt * k * sizeof(float);
                    if (use >= 2 && (e = cudaEventSynchronize(c.stage_ev[t][b])) != cudaSuccess)
                        return bail(e, ci);
                    const double t_m0 = now_us();
                    memcpy(c.stage[t][b], R + (size_t)(begin + off) * k, bytes);
                    const double t_m1 = now_us();
                    cudaStream_t fs = one_stream ? c.copy : c.fstream[t];
                    if ((e = cudaMemcpyAsync(c.dR + (size_t)off * k, c.stage[t][b], bytes, cudaMemcpyHostToDevice, fs)) !=
                            cudaSuccess ||
                        (e = cudaEventRecord(c.events[ci + 1], fs)) != cudaSuccess ||
                        (e = cudaEventRecord(c.stage_ev[t][b], fs)) != cudaSuccess)
                        return bail(e, ci);
                    if (trace_on())
                        fprintf(stderr, "[nn_b200] feeder %d chunk %zu: start +%.0f us, memcpy %zu KiB %.0f us, api %.0f us\n", t, ci,
                                t_start - t_enq, bytes >> 10, t_m1 - t_m0, now_us() - t_m1);
                    {
                        std::lock_guard<std::mutex> lk(fmu);
                        recorded[ci] = 1;
                    }
                    fcv.notify_all();
                }
            });
    }

    // Copy granularity and search granularity are decoupled (plan_search_groups).
    const std::vector<size_t> group_ends = plan_search_groups(k, m, chunks, g_opt.search_group.load(), g_active_gpus.load());
    size_t gi = 0;
    for (size_t ci = 0; ci < nchunks;)
    {
        const size_t ce = group_ends[gi++];
        const int64_t off = chunks[ci].first;
        int64_t cnt = 0;
        for (size_t cj = ci; cj < ce; ++cj)
        {
            if (feeders > 0)
            {
                std::unique_lock<std::mutex> lk(fmu);
                fcv.wait(lk, [&] { return recorded[cj] != 0; });
                if (frc != NN_B200_OK)
                    return fail(frc, "%s", ferr.c_str());
            }
            else
            {
                CU(cudaMemcpyAsync(c.dR + (size_t)chunks[cj].first * k, R + (size_t)(begin + chunks[cj].first) * k,
                                   (size_t)chunks[cj].second * k * sizeof(float), cudaMemcpyHostToDevice, c.copy));
                CU(cudaEventRecord(c.events[cj + 1], c.copy));
            }
            if (trace_on() && feeders > 0)
                fprintf(stderr, "[nn_b200] main: chunk %zu ready at +%.0f us\n", cj, now_us() - t_enq);
            CU(cudaStreamWaitEvent(c.compute, c.events[cj + 1], 0));
            cnt += chunks[cj].second;
        }
        ci = ce;
        if (lone && nchunks == 1)
        { // the whole search in one launch: its last CTAs store dOut and restore the workspace
            rc = search_device_impl(k, m, cnt, c.dS, c.dR, (uint32_t)begin, c.dWs, c.dOut, nullptr, c.compute);
            if (rc)
                return rc;
            c.finished = true;
            continue;
        }
        rc = nearest_keys_impl(k, m, cnt, c.dS, c.dR + (size_t)off * k, (uint32_t)(begin + off),
                               reinterpret_cast<uint64_t *>(keys), c.compute, false, keys0 ? 1 : 0);
        if (rc)
            return rc;
    }
    if (!c.done)
        CU(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
    CU(cudaEventRecord(c.done, c.compute));
    return NN_B200_OK;
}

// Peer access from every used device to device 0 (whose key array all shards fold into).
bool ensure_peer_to_0(int gpus)
{
    bool all = true;
    for (int g = 1; g < gpus; ++g)
    {
        DevCtx &c = g_ctx.devs[g];
        if (c.peer_to_0 < 0)
        {
            int can = 0;
            c.peer_to_0 = 0;
            // the merge is a 64-bit atomic MIN executed at GPU 0's L2: that needs NATIVE peer atomics
            // (NVLink); PCIe peer access only has fetch-add/swap/CAS and would merge wrongly
            int native = 0;
            if (cudaDeviceCanAccessPeer(&can, g, 0) == cudaSuccess && can &&
                cudaDeviceGetP2PAttribute(&native, cudaDevP2PAttrNativeAtomicSupported, g, 0) == cudaSuccess && native &&
                cudaSetDevice(g) == cudaSuccess)
            {
                const cudaError_t e = cudaDeviceEnablePeerAccess(0, 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled)
                    c.peer_to_0 = 1;
            }
            (void)cudaGetLastError();
        }
        all = all && c.peer_to_0 == 1;
    }
    return all;
}

int ensure_comms(int gpus)
{
    if (g_ctx.comm_gpus == gpus)
        return NN_B200_OK;
    int rc = load_nccl();
    if (rc)
        return rc;
    for (ncclComm_t c : g_ctx.comms)
        g_nccl.CommDestroy(c);
    g_ctx.comms.assign(gpus, nullptr);
    g_ctx.comm_gpus = 0;
    std::vector<int> devs(gpus);
    for (int i = 0; i < gpus; ++i)
        devs[i] = i;
    ncclResult_t r = g_nccl.CommInitAll(g_ctx.comms.data(), gpus, devs.data());
    if (r != 0)
        return fail(NN_B200_ENCCL, "ncclCommInitAll(%d) failed: %s", gpus, g_nccl.GetErrorString(r));
    g_ctx.comm_gpus = gpus;
    return NN_B200_OK;
}
} // namespace

// Common driver of a sharded search on `gpus` devices: sets up the merge, runs enqueue(g, keys0,
// keys_ready) for every device (one host thread each when there are several), merges, unpacks on
// device 0 and copies the m indices to `results`.  Caller holds g_ctx.mu.
template <class Enqueue>
static int run_sharded_body(int m, int gpus, Enqueue enqueue, int *results, int prev_dev);

template <class Enqueue>
static int run_sharded(int m, int gpus, Enqueue enqueue, int *results)
{
    g_active_gpus = gpus;
    int prev_dev = 0;
    CU(cudaGetDevice(&prev_dev));
    if ((int)g_ctx.devs.size() < gpus)
        g_ctx.devs.resize(gpus);
    const int rc = run_sharded_body(m, gpus, enqueue, results, prev_dev);
    if (rc)
    { // whatever was enqueued before the failure must not outlive the call: the caller frees its
      // buffers as soon as we return (main.cu:76-77); the workspaces stay marked dirty
        const std::string keep = t_err;
        for (int g = 0; g < gpus; ++g)
            if (g_ctx.devs[g].compute && cudaSetDevice(g) == cudaSuccess)
                cudaDeviceSynchronize();
        (void)cudaGetLastError();
        t_err = keep;
    }
    cudaSetDevice(prev_dev);
    return rc;
}

template <class Enqueue>
static int run_sharded_body(int m, int gpus, Enqueue enqueue, int *results, int prev_dev)
{
    int rc = NN_B200_OK;
    const double t_call = now_us();

    // Merge of the per-GPU candidates.  Default: every GPU's search kernels fold straight into GPU 0's
    // key array with system-scope 64-bit atomicMin over NVLink (the exchange step happens inside the
    // search kernel, tile by tile; GPU 0 only waits for the other GPUs' completion events).  Without
    // peer access (or with option p2p_merge = 0): one in-place ncclAllReduce(min, uint64).
    const bool p2p = gpus > 1 && g_opt.p2p_merge.load() != 0 && ensure_peer_to_0(gpus);
    cudaEvent_t keys_ready = nullptr;
    if (p2p)
    {
        DevCtx &c0 = g_ctx.devs[0];
        // (its workspace is, or is being put, in the start state.  It is marked dirty -- "a call is folding
        // into it" -- by GPU 0's own enqueue, AFTER that enqueue's ensure_dev: marking it here would make
        // that ensure_dev re-initialise the keys while the other GPUs are already folding into them)
        rc = ensure_dev(c0, 0, 16, 16, (size_t)std::max(m, 1), 1);
        if (rc)
            return rc;
        if (!c0.keys_ready)
            CU(cudaEventCreateWithFlags(&c0.keys_ready, cudaEventDisableTiming));
        CU(cudaEventRecord(c0.keys_ready, c0.compute));
        keys_ready = c0.keys_ready;
    }
    unsigned long long *keys0 = p2p ? g_ctx.devs[0].dKeys : nullptr;

    std::vector<int> rcs(gpus, 0);
    std::vector<std::string> errs(gpus);
    auto work = [&](int g) {
        rcs[g] = enqueue(g, keys0, keys_ready);
        if (rcs[g])
            errs[g] = t_err;
    };
    if (gpus == 1)
        work(0);
    else
    {
        // one host thread per GPU, as v8 does with OpenMP (core.cu:873): pageable copies block the
        // issuing thread, so the shards are pushed in parallel (persistent pool threads; GPU 0's share
        // runs on the calling thread)
        TaskGroup grp;
        for (int g = 1; g < gpus; ++g)
            g_gpu_workers.submit(grp, [&work, g]() { work(g); });
        work(0);
        grp.wait();
    }
    for (int g = 0; g < gpus; ++g)
        if (rcs[g])
        {
            t_err = errs[g];
            return rcs[g];
        }
    if (trace_on())
        fprintf(stderr, "[nn_b200] call: everything enqueued at +%.0f us\n", now_us() - t_call);

    if (p2p)
    {
        for (int g = 1; g < gpus; ++g)
            CU(cudaStreamWaitEvent(g_ctx.devs[0].compute, g_ctx.devs[g].done, 0));
    }
    else if (gpus > 1)
    {
        rc = ensure_comms(gpus);
        if (rc)
            return rc;
        g_nccl.GroupStart();
        for (int g = 0; g < gpus; ++g)
        {
            DevCtx &c = g_ctx.devs[g];
            ncclResult_t r =
                g_nccl.AllReduce(c.dKeys, c.dKeys, (size_t)m, kNcclUint64, kNcclMin, g_ctx.comms[g], c.compute);
            if (r != 0)
            {
                g_nccl.GroupEnd();
                return fail(NN_B200_ENCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
            }
        }
        ncclResult_t r = g_nccl.GroupEnd();
        if (r != 0)
            return fail(NN_B200_ENCCL, "ncclGroupEnd failed: %s", g_nccl.GetErrorString(r));
    }

    // Finish: indices out of GPU 0's merged keys, every folded-into workspace back to its start state
    // (a single-GPU single-chunk search has already done both inside its kernel).
    DevCtx &c0 = g_ctx.devs[0];
    for (int g = (p2p ? 0 : gpus - 1); g >= 0; --g)
    {
        DevCtx &c = g_ctx.devs[g];
        if (c.finished)
            continue;
        CU(cudaSetDevice(g));
        rc = nn_b200_workspace_finish(c.dWs, m, g == 0 ? c.dOut : nullptr, nullptr, c.compute);
        if (rc)
            return rc;
    }
    CU(cudaSetDevice(0));
    CU(cudaMemcpyAsync(c0.hOut, c0.dOut, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c0.compute));
    for (int g = gpus - 1; g >= 0; --g)
    {
        CU(cudaSetDevice(g));
        CU(cudaStreamSynchronize(g_ctx.devs[g].copy));
        CU(cudaStreamSynchronize(g_ctx.devs[g].compute));
        g_ctx.devs[g].ws_dirty = false;
    }
    memcpy(results, c0.hOut, (size_t)m * sizeof(int));
    if (trace_on())
        fprintf(stderr, "[nn_b200] call: synchronised at +%.0f us\n", now_us() - t_call);
    (void)prev_dev;
    return NN_B200_OK;
}

extern "C" int nn_b200_search_host(int k, int m, int n, const float *S, const float *R, int *results, int num_gpus)
{
    int rc = check_shape(k, m, n);
    if (rc)
        return rc;
    if (m == 0)
        return NN_B200_OK;
    if (!S || !results || (n > 0 && !R))
        return fail(NN_B200_EINVAL, "null host pointer");
    int gpus = nn_b200_device_count(n > 0 ? n : 1);
    if (gpus < 1)
        return fail(NN_B200_ENODEV, "no CUDA device visible (this build has no CPU fallback)");
    if (num_gpus > 0)
        gpus = std::min(gpus, num_gpus); // the caller's choice
    else if (gpus > 1 && g_opt.auto_gpus.load() != 0)
        gpus = plan_gpus(k, m, n, gpus, n > 0 && !is_pageable(R)); // as many as pay for themselves
    g_last_gpus = gpus;

    std::lock_guard<std::mutex> lk(g_ctx.mu);
    return run_sharded(m, gpus,
                       [&](int g, unsigned long long *keys0, cudaEvent_t keys_ready) {
                           int64_t b = 0, cnt = 0;
                           nn_b200_shard_range(n, gpus, g, &b, &cnt);
                           return enqueue_device(g_ctx.devs[g], g, k, m, S, R, b, cnt, keys0, keys_ready, gpus == 1);
                       },
                       results);
}

// ---------------------------------------------------------------------------------------------
// Resident reference index: build once, query many times
// ---------------------------------------------------------------------------------------------
struct nn_b200_index
{
    int k = 0;
    int64_t n = 0;
    int gpus = 0;
    std::vector<float *> dR;          // per device: its contiguous shard, native AoS
    std::vector<int64_t> begin, count; // shard ranges
    struct Graph
    {
        int m;
        uint64_t generation; // of the device context the graph was captured against
        int64_t epoch;       // of the options
        cudaGraphExec_t exec;
    };
    std::vector<Graph> graphs; // one per batch size seen (single-GPU indexes only)
};

extern "C" void nn_b200_index_destroy(nn_b200_index *ix)
{
    if (!ix)
        return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto &gr : ix->graphs)
        cudaGraphExecDestroy(gr.exec);
    for (int g = 0; g < (int)ix->dR.size(); ++g)
        if (ix->dR[g])
        {
            cudaSetDevice(g);
            cudaFree(ix->dR[g]);
        }
    cudaSetDevice(prev);
    (void)cudaGetLastError();
    delete ix;
}

extern "C" int nn_b200_index_create(int k, int n, const float *R, int num_gpus, nn_b200_index **out)
{
    int rc = check_shape(k, 0, n);
    if (rc)
        return rc;
    if (!out || (n > 0 && !R))
        return fail(NN_B200_EINVAL, "null pointer");
    *out = nullptr;
    int gpus = nn_b200_device_count(n > 0 ? n : 1);
    if (gpus < 1)
        return fail(NN_B200_ENODEV, "no CUDA device visible (this build has no CPU fallback)");
    if (num_gpus > 0)
        gpus = std::min(gpus, num_gpus);
    int prev_dev = 0;
    CU(cudaGetDevice(&prev_dev));
    nn_b200_index *ix = new nn_b200_index;
    ix->k = k;
    ix->n = n;
    ix->gpus = gpus;
    ix->dR.assign(gpus, nullptr);
    ix->begin.assign(gpus, 0);
    ix->count.assign(gpus, 0);
    std::vector<int> rcs(gpus, 0);
    std::vector<std::string> errs(gpus);
    auto load = [&](int g) {
        auto body = [&]() -> int {
            nn_b200_shard_range(n, gpus, g, &ix->begin[g], &ix->count[g]);
            DevInfo di;
            int r = dev_info(g, &di); // refuses anything but sm_100
            if (r)
                return r;
            CU(cudaSetDevice(g));
            const size_t bytes = (size_t)ix->count[g] * k * sizeof(float);
            CU(cudaMalloc(&ix->dR[g], std::max<size_t>(bytes, 16)));
            if (bytes)
                CU(cudaMemcpy(ix->dR[g], R + (size_t)ix->begin[g] * k, bytes, cudaMemcpyHostToDevice));
            return NN_B200_OK;
        };
        rcs[g] = body();
        if (rcs[g])
            errs[g] = t_err;
    };
    if (gpus == 1)
        load(0);
    else
    {
        TaskGroup grp;
        for (int g = 1; g < gpus; ++g)
            g_gpu_workers.submit(grp, [&load, g]() { load(g); });
        load(0);
        grp.wait();
    }
    cudaSetDevice(prev_dev);
    for (int g = 0; g < gpus; ++g)
        if (rcs[g])
        {
            t_err = errs[g];
            const int r = rcs[g];
            nn_b200_index_destroy(ix);
            return r;
        }
    *out = ix;
    return NN_B200_OK;
}

extern "C" int nn_b200_index_info(const nn_b200_index *ix, int *k, int64_t *n, int *gpus)
{
    if (!ix)
        return fail(NN_B200_EINVAL, "null index");
    if (k)
        *k = ix->k;
    if (n)
        *n = ix->n;
    if (gpus)
        *gpus = ix->gpus;
    return NN_B200_OK;
}

// Caller holds g_ctx.mu.  Single-GPU index, small batch.
static int index_search_graph(nn_b200_index *ix, int m, const float *S, int *results)
{
    const int k = ix->k;
    const size_t bytesS = (size_t)m * k * sizeof(float);
    int prev_dev = 0;
    CU(cudaGetDevice(&prev_dev));
    if (g_ctx.devs.empty())
        g_ctx.devs.resize(1);
    DevCtx &c = g_ctx.devs[0];
    int rc = ensure_dev(c, 0, std::max<size_t>(bytesS, 16), 16, (size_t)m, 1);
    if (rc)
        return rc;
    if (c.capHS < bytesS)
    {
        if (c.hS)
            CU(cudaFreeHost(c.hS));
        c.hS = nullptr;
        c.capHS = 0;
        CU(cudaMallocHost(&c.hS, (size_t)1 << 20));
        c.capHS = (size_t)1 << 20;
        ++c.generation;
    }
    const int64_t epoch = g_opt_epoch.load();
    cudaGraphExec_t exec = nullptr;
    for (auto it = ix->graphs.begin(); it != ix->graphs.end();)
    {
        if (it->generation != c.generation || it->epoch != epoch)
        { // captured against buffers or plans that no longer exist
            cudaGraphExecDestroy(it->exec);
            it = ix->graphs.erase(it);
            continue;
        }
        if (it->m == m)
            exec = it->exec;
        ++it;
    }
    if (!exec)
    {
        // plan outside the capture (planning queries the runtime), then record copy-in, search, copy-out
        DevInfo di;
        rc = dev_info(0, &di);
        if (rc)
            return rc;
        Plan p;
        rc = make_plan(0, k, m, ix->count[0], false, di, &p);
        if (rc)
            return rc;
        CU(cudaStreamBeginCapture(c.compute, cudaStreamCaptureModeThreadLocal));
        int r = NN_B200_OK;
        cudaError_t ce = cudaMemcpyAsync(c.dS, c.hS, bytesS, cudaMemcpyHostToDevice, c.compute);
        if (ce == cudaSuccess)
        {
            r = search_device_impl(k, m, ix->count[0], c.dS, ix->dR[0], (uint32_t)ix->begin[0], c.dWs, c.dOut, nullptr,
                                   c.compute);
            if (!r)
                ce = cudaMemcpyAsync(c.hOut, c.dOut, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c.compute);
        }
        cudaGraph_t graph = nullptr;
        const cudaError_t ee = cudaStreamEndCapture(c.compute, &graph);
        if (r)
        {
            if (graph)
                cudaGraphDestroy(graph);
            return r;
        }
        if (ce != cudaSuccess || ee != cudaSuccess)
        {
            if (graph)
                cudaGraphDestroy(graph);
            (void)cudaGetLastError();
            return fail(NN_B200_ECUDA, "graph capture of the index search failed: %s",
                        cudaGetErrorString(ce != cudaSuccess ? ce : ee));
        }
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess)
            return fail(NN_B200_ECUDA, "cudaGraphInstantiate of the index search failed: %s", cudaGetErrorString(ie));
        if (ix->graphs.size() >= 16)
        {
            cudaGraphExecDestroy(ix->graphs.front().exec);
            ix->graphs.erase(ix->graphs.begin());
        }
        ix->graphs.push_back({m, c.generation, epoch, exec});
    }
    memcpy(c.hS, S, bytesS);
    c.ws_dirty = true;
    CU(cudaGraphLaunch(exec, c.compute));
    g_launches += 1; // the one-launch search replayed
    CU(cudaStreamSynchronize(c.compute));
    c.ws_dirty = false;
    memcpy(results, c.hOut, (size_t)m * sizeof(int));
    CU(cudaSetDevice(prev_dev));
    return NN_B200_OK;
}

extern "C" int nn_b200_index_search(nn_b200_index *ix, int m, const float *S, int *results)
{
    if (!ix)
        return fail(NN_B200_EINVAL, "null index");
    int rc = check_shape(ix->k, m, ix->n);
    if (rc)
        return rc;
    if (m == 0)
        return NN_B200_OK;
    if (!S || !results)
        return fail(NN_B200_EINVAL, "null host pointer");
    const int k = ix->k;
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    // Small batches on one GPU are launch-bound (copy in, init, search, unpack, copy out: five
    // enqueues for a few tens of microseconds of device work): the sequence is captured once per
    // batch size into a CUDA graph and replayed with a single launch.
    // (measured on B200: k=3, m=1024, n=65536 63 -> 50 us per call; k=8, m=8, n=2^22 72 -> 59 us;
    // nothing to gain once the search itself takes milliseconds, hence the bound on m*n*k)
    if (ix->gpus == 1 && ix->count[0] > 0 && g_opt.index_graph.load() != 0 &&
        (size_t)m * k * sizeof(float) <= ((size_t)1 << 20) && (double)m * (double)ix->n * k <= 2147483648.0)
        return index_search_graph(ix, m, S, results);
    return run_sharded(m, ix->gpus,
                       [&](int g, unsigned long long *keys0, cudaEvent_t keys_ready) -> int {
                           DevCtx &c = g_ctx.devs[g];
                           const size_t bytesS = (size_t)m * k * sizeof(float);
                           int r = ensure_dev(c, g, std::max<size_t>(bytesS, 16), 16, (size_t)m, 1);
                           if (r)
                               return r;
                           CU(cudaMemcpyAsync(c.dS, S, bytesS, cudaMemcpyHostToDevice, c.copy));
                           CU(cudaEventRecord(c.events[0], c.copy));
                           unsigned long long *keys = keys0 ? keys0 : c.dKeys;
                           if (keys0)
                               CU(cudaStreamWaitEvent(c.compute, keys_ready, 0));
                           if (!keys0 || g == 0)
                               c.ws_dirty = true;
                           c.finished = false;
                           CU(cudaStreamWaitEvent(c.compute, c.events[0], 0));
                           if (ix->gpus == 1)
                           { // one launch: search, merge, index store
                               r = search_device_impl(k, m, ix->count[g], c.dS, ix->dR[g], (uint32_t)ix->begin[g], c.dWs,
                                                      c.dOut, nullptr, c.compute);
                               if (r)
                                   return r;
                               c.finished = true;
                           }
                           else if (ix->count[g] > 0)
                           {
                               r = nearest_keys_impl(k, m, ix->count[g], c.dS, ix->dR[g], (uint32_t)ix->begin[g],
                                                     reinterpret_cast<uint64_t *>(keys), c.compute, false, keys0 ? 1 : 0);
                               if (r)
                                   return r;
                           }
                           if (!c.done)
                               CU(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
                           CU(cudaEventRecord(c.done, c.compute));
                           return NN_B200_OK;
                       },
                       results);
}

// ---------------------------------------------------------------------------------------------
// Multi-process merge over NVLink peer memory: one process per GPU (torchrun ranks), rank 0 owns the
// keys.  The exchange step of the sharded search (the reference's host-side gather + reduce,
// core.cu:925-957; one ncclAllReduce(min, u64) in the NCCL driver) happens INSIDE the search kernels:
// see struct PeerSync (nn_launch.h) and peer_finish (nn_kernels.cuh).
// Block layout, identical on every rank (only rank 0's keys / arrive are used):
//   [0] done  [64] groups_done  [128] error  [192] arrive[2]  [256] keys[2][m]  [..] this rank's search workspace
// ---------------------------------------------------------------------------------------------
struct nn_b200_peer_merge
{
    int m = 0, rank = 0, world = 0, dev = 0;
    unsigned char *block = nullptr;          // this rank's block (cudaMalloc)
    cudaIpcMemHandle_t handle{};
    std::vector<unsigned char *> mapped;     // by rank: peer blocks opened here (rank 0: all; others: rank 0's)
    unsigned int **d_done_peers = nullptr;   // rank 0: device array of the other ranks' done flags
    unsigned int step = 0;
    bool attached = false;
};
static size_t peer_ws_offset(int m) { return (256 + (size_t)2 * (size_t)std::max(m, 1) * 8 + 255) / 256 * 256; }
static size_t peer_block_bytes(int m) { return peer_ws_offset(m) + nn_b200_workspace_bytes(std::max(m, 1)); }

__global__ void nn_peer_arrive_kernel(const Finish f)
{ // a rank whose shard is empty still has to be counted (one CTA, nothing folded)
    finish_group(f, nullptr, 0, 1, 0, 0);
}

extern "C" void nn_b200_peer_destroy(nn_b200_peer_merge *pm)
{
    if (!pm)
        return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(pm->dev);
    cudaDeviceSynchronize();
    for (unsigned char *p : pm->mapped)
        if (p)
            cudaIpcCloseMemHandle(p);
    if (pm->d_done_peers)
        cudaFree(pm->d_done_peers);
    if (pm->block)
        cudaFree(pm->block);
    cudaSetDevice(prev);
    (void)cudaGetLastError();
    delete pm;
}

extern "C" int nn_b200_peer_create(int m, int rank, int world, nn_b200_peer_merge **out)
{
    if (m < 0 || world < 1 || rank < 0 || rank >= world || !out)
        return fail(NN_B200_EINVAL, "bad peer_create arguments");
    *out = nullptr;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevInfo di;
    int rc = dev_info(dev, &di);
    if (rc)
        return rc;
    nn_b200_peer_merge *pm = new nn_b200_peer_merge;
    pm->m = m;
    pm->rank = rank;
    pm->world = world;
    pm->dev = dev;
    pm->mapped.assign(world, nullptr);
    auto bail = [&](int code) {
        nn_b200_peer_destroy(pm);
        return code;
    };
    const size_t bytes = peer_block_bytes(m);
    if (cudaMalloc(&pm->block, bytes) != cudaSuccess)
        return bail(fail(NN_B200_ECUDA, "cudaMalloc of the peer block failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (cudaMemset(pm->block, 0, 256) != cudaSuccess)
        return bail(fail(NN_B200_ECUDA, "cudaMemset failed"));
    rc = nn_b200_keys_init(reinterpret_cast<uint64_t *>(pm->block + 256), 2 * std::max(m, 1), nullptr);
    if (!rc)
        rc = nn_b200_workspace_init(pm->block + peer_ws_offset(m), std::max(m, 1), nullptr);
    if (rc)
        return bail(rc);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaIpcGetMemHandle(&pm->handle, pm->block) != cudaSuccess)
        return bail(fail(NN_B200_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError())));
    *out = pm;
    return NN_B200_OK;
}

extern "C" int nn_b200_peer_handle(const nn_b200_peer_merge *pm, void *handle, size_t len)
{
    if (!pm || !handle || len < sizeof(cudaIpcMemHandle_t))
        return fail(NN_B200_EINVAL, "peer_handle needs a %zu-byte buffer", sizeof(cudaIpcMemHandle_t));
    memcpy(handle, &pm->handle, sizeof(cudaIpcMemHandle_t));
    return NN_B200_OK;
}

// handles: world x 64 bytes, the result of nn_b200_peer_handle on every rank, in rank order
extern "C" int nn_b200_peer_attach(nn_b200_peer_merge *pm, const void *handles, size_t len)
{
    if (!pm || !handles || len < (size_t)pm->world * sizeof(cudaIpcMemHandle_t))
        return fail(NN_B200_EINVAL, "peer_attach needs world x %zu bytes", sizeof(cudaIpcMemHandle_t));
    if (pm->attached)
        return fail(NN_B200_EINVAL, "already attached");
    CU(cudaSetDevice(pm->dev));
    const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(handles);
    auto open = [&](int r) -> int {
        void *p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess));
        pm->mapped[r] = reinterpret_cast<unsigned char *>(p);
        return NN_B200_OK;
    };
    if (pm->rank == 0)
    {
        std::vector<unsigned int *> flags;
        for (int r = 1; r < pm->world; ++r)
        {
            const int rc = open(r);
            if (rc)
                return rc;
            flags.push_back(reinterpret_cast<unsigned int *>(pm->mapped[r])); // [0] = that rank's done flag
        }
        if (!flags.empty())
        {
            CU(cudaMalloc(&pm->d_done_peers, flags.size() * sizeof(unsigned int *)));
            CU(cudaMemcpy(pm->d_done_peers, flags.data(), flags.size() * sizeof(unsigned int *), cudaMemcpyHostToDevice));
        }
    }
    else
    {
        const int rc = open(0);
        if (rc)
            return rc;
    }
    pm->attached = true;
    return NN_B200_OK;
}

// One sharded search: this rank's n references (global indices from index_base) against the m queries;
// rank 0's d_results receives the merged indices (other ranks pass NULL).  Every rank must call it the
// same number of times.  Asynchronous on `stream`.
extern "C" int nn_b200_peer_search(nn_b200_peer_merge *pm, int k, int m, int64_t n, const float *d_S, const float *d_R,
                                   uint32_t index_base, int *d_results, void *stream)
{
    if (!pm || !pm->attached)
        return fail(NN_B200_EINVAL, "peer merge not attached");
    int rc = check_shape(k, m, n);
    if (rc)
        return rc;
    if (m < 1 || m > pm->m)
        return fail(NN_B200_EINVAL, "m=%d outside 1..%d of this peer merge", m, pm->m);
    if (pm->rank == 0 && !d_results)
        return fail(NN_B200_EINVAL, "rank 0 needs a result buffer");
    unsigned char *b0 = pm->rank == 0 ? pm->block : pm->mapped[0];
    void *ws = pm->block + peer_ws_offset(pm->m); // this rank's own workspace: the folds stay on this GPU
    Finish fin;
    fin.tickets = ws_tickets(ws);
    fin.results = pm->rank == 0 ? d_results : nullptr;
    fin.peer.groups_done = reinterpret_cast<unsigned int *>(pm->block + 64);
    fin.peer.keys = reinterpret_cast<unsigned long long *>(b0 + 256) + (size_t)(pm->step & 1u) * (size_t)pm->m;
    fin.peer.arrive = reinterpret_cast<unsigned int *>(b0 + 192);
    fin.peer.done_local = reinterpret_cast<unsigned int *>(pm->block);
    fin.peer.done_peers = pm->d_done_peers;
    fin.peer.error = reinterpret_cast<unsigned int *>(pm->block + 128);
    fin.peer.step = pm->step;
    fin.peer.world = (unsigned int)pm->world;
    fin.peer.rank = (unsigned int)pm->rank;
    fin.peer.m = m;
    bool done = false;
    if (n > 0)
    {
        rc = nearest_keys_impl(k, m, n, d_S, d_R, index_base, ws_keys(ws), stream, false, 0, &fin, &done);
        if (rc)
            return rc;
        if (!done)
            return fail(NN_B200_EINVAL, "this plan has no in-kernel finish");
    }
    if (!done)
    { // empty shard: be counted all the same (one group, nothing to push)
        fin.peer.num_groups = 1;
        nn_peer_arrive_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(fin);
        CU(cudaGetLastError());
        g_launches++;
    }
    pm->step++;
    return NN_B200_OK;
}

// 1 if a wait inside a kernel of this rank timed out (a rank missing or out of step), else 0.  Synchronises.
extern "C" int nn_b200_peer_error(nn_b200_peer_merge *pm)
{
    if (!pm)
        return fail(NN_B200_EINVAL, "null peer merge");
    unsigned int e = 0;
    CU(cudaMemcpy(&e, pm->block + 128, sizeof e, cudaMemcpyDeviceToHost));
    return (int)e;
}

extern "C" void nn_b200_cudaCallback(int k, int m, int n, float *searchPoints, float *referencePoints, int **results)
{
    int *tmp = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1)); // caller frees (core.cu:935; main.cu:98)
    if (!tmp)
    {
        printf("Error: %s:%d, malloc of %d results failed\n", __FILE__, __LINE__, m);
        exit(1);
    }
    const int rc = nn_b200_search_host(k, m, n, searchPoints, referencePoints, tmp, 0);
    if (rc != NN_B200_OK)
    {
        // same behaviour as the reference's CHECK macro (core.h:77-87)
        printf("Error: %s:%d, code:%d, reason: %s \n", __FILE__, __LINE__, rc, nn_b200_last_error());
        exit(1);
    }
    *results = tmp;
}

// Opt-in warm-up at library load, the counterpart of the reference's `static WarmUP warm_up(1,1,1<<20)`
// (core.cu:1274, README.md:222), which runs every version once before main() so that no timed call
// pays for context creation and code loading.  With NN_B200_STATIC_WARMUP=<g> in the environment the
// first g GPUs (all for a value < 0) get their context, streams, kernels and floor-sized buffers
// here; without it the first cudaCallback pays that cost (0.7-2.5 s under the reference's harness).
namespace
{
struct StaticWarmup
{
    StaticWarmup()
    {
        const char *e = getenv("NN_B200_STATIC_WARMUP");
        if (!e || atoi(e) == 0)
            return;
        int want = atoi(e);
        const int vis = visible_devices();
        if (want < 0 || want > vis)
            want = vis;
        std::lock_guard<std::mutex> lk(g_ctx.mu);
        if ((int)g_ctx.devs.size() < want)
            g_ctx.devs.resize(want);
        for (int g = 0; g < want; ++g)
            if (ensure_dev(g_ctx.devs[g], g, 16, 16, 1, 2) == NN_B200_OK)
                cudaStreamSynchronize(g_ctx.devs[g].compute);
        if (want > 1)
            ensure_peer_to_0(want);
        if (want > 0)
            cudaSetDevice(0);
        (void)cudaGetLastError();
    }
};
StaticWarmup g_static_warmup; // (last in this file: everything it touches is constructed before it)
} // namespace

// The reference's C++-linkage entry point (core.h:71, core.cu:1282-1297).  Build with
// -DNN_B200_NO_CXX_ENTRY when the host program keeps its own ::cudaCallback that forwards to
// nn_b200_cudaCallback (variant b).
#ifndef NN_B200_NO_CXX_ENTRY
void cudaCallback(int k, int m, int n, float *searchPoints, float *referencePoints, int **results)
{
    nn_b200_cudaCallback(k, m, n, searchPoints, referencePoints, results);
}
#endif
