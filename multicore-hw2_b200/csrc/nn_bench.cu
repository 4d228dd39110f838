// nn_bench.cu -- standalone sweep tool for the device-resident search (links libnn_b200.so through
// the C ABI only).  Prints one JSON line per measured configuration.  Used to pick tile shapes and
// to produce the ncu captures under profiles/; bench.py is the contract benchmark.
//
//   nn_bench --k 16 --m 4096 --n 1048576 [--variant 1] [--q 4] [--scalar 1] [--splits S] [--waves W]
//            [--iters 10] [--warmup 3] [--check 1] [--soa 1] [--repack 1] [--rreg_ctas C]
//   nn_bench --sweep <name>     (preset lists: math, cfgs, small, repack)
#include "../../include/nn_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef NN_QREG_TIMELINE
extern "C" void nn_b200_debug_timeline(unsigned long long *p); // debug builds of the library only
#endif

#define CK(call)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
        {                                                                                                              \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);                  \
            exit(2);                                                                                                   \
        }                                                                                                              \
    } while (0)
#define NN(call)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        int r__ = (call);                                                                                              \
        if (r__ != 0)                                                                                                  \
        {                                                                                                              \
            fprintf(stderr, "nn_b200 error %d (%s) at %s:%d\n", r__, nn_b200_last_error(), __FILE__, __LINE__);        \
            exit(3);                                                                                                   \
        }                                                                                                              \
    } while (0)

// counter-based uniform [0,1) floats with 24 random bits (splitmix64 finaliser)
__global__ void fill_uniform(float *out, size_t count, uint64_t seed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride)
    {
        uint64_t z = seed + (uint64_t)i * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        out[i] = (float)(z >> 40) * (1.0f / 16777216.0f);
    }
}
// coarse grid: plenty of exact ties
__global__ void quantize(float *x, size_t count, float levels)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride)
        x[i] = floorf(x[i] * levels) / levels;
}

struct Cfg
{
    int k = 16, m = 4096;
    long long n = 1 << 20;
    int variant = 0, q = 0, scalar = 2, splits = 0, waves = 8, iters = 10, warmup = 3, check = 0, soa = 0, repack = 0,
        rreg_ctas = 0, quant = 0, ldg = 0, fused = 0, timeline = 0;
    std::string tag;
};

static double g_peak_ops = 0, g_clock_ghz = 0;
static int g_sms = 0;

static void run(const Cfg &c)
{
    const size_t ns = (size_t)c.m * c.k, nr = (size_t)c.n * c.k;
    float *dS, *dR, *dRs = nullptr;
    uint64_t *dK, *dK2;
    CK(cudaMalloc(&dS, std::max<size_t>(ns, 4) * 4));
    CK(cudaMalloc(&dR, std::max<size_t>(nr, 4) * 4));
    CK(cudaMalloc(&dK, std::max(c.m, 1) * 8));
    CK(cudaMalloc(&dK2, std::max(c.m, 1) * 8));
    fill_uniform<<<1024, 256>>>(dS, ns, 1000);
    fill_uniform<<<4096, 256>>>(dR, nr, 2000);
    if (c.quant)
    {
        quantize<<<1024, 256>>>(dS, ns, (float)c.quant);
        quantize<<<4096, 256>>>(dR, nr, (float)c.quant);
    }
    CK(cudaDeviceSynchronize());

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<float> ms;

    if (c.repack)
    {
        CK(cudaMalloc(&dRs, std::max<size_t>(nr, 4) * 4));
        for (int it = 0; it < c.warmup + c.iters; ++it)
        {
            CK(cudaEventRecord(e0));
            NN(nn_b200_repack_soa(c.k, c.n, dR, dRs, nullptr));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (it >= c.warmup)
                ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        const double med = ms[ms.size() / 2], best = ms[0];
        const double bytes = 2.0 * (double)nr * 4.0;
        printf("{\"op\":\"repack_soa\",\"tag\":\"%s\",\"k\":%d,\"n\":%lld,\"ms_med\":%.4f,\"ms_best\":%.4f,"
               "\"GBps_med\":%.1f,\"GBps_best\":%.1f}\n",
               c.tag.c_str(), c.k, c.n, med, best, bytes / med / 1e6, bytes / best / 1e6);
        fflush(stdout);
        CK(cudaFree(dRs));
        CK(cudaFree(dS));
        CK(cudaFree(dR));
        CK(cudaFree(dK));
        CK(cudaFree(dK2));
        return;
    }
    if (c.soa)
    {
        CK(cudaMalloc(&dRs, std::max<size_t>(nr, 4) * 4));
        NN(nn_b200_repack_soa(c.k, c.n, dR, dRs, nullptr));
    }

    NN(nn_b200_set_option("variant", c.variant));
    NN(nn_b200_set_option("qreg_q", c.q));
    NN(nn_b200_set_option("math", c.scalar));
    NN(nn_b200_set_option("splits", c.splits));
    NN(nn_b200_set_option("waves", c.waves));
    NN(nn_b200_set_option("rreg_ctas_per_sm", c.rreg_ctas));
    char plan[256] = "";
    if (!c.soa)
        NN(nn_b200_describe_plan(c.k, c.m, c.n, plan, sizeof plan));

    // --fused 1: the one-launch search (workspace + nn_b200_search_device): no init, no unpack
    void *dWs = nullptr;
    int *dOut = nullptr;
    if (c.fused)
    {
        CK(cudaMalloc(&dWs, nn_b200_workspace_bytes(c.m)));
        CK(cudaMalloc(&dOut, std::max(c.m, 1) * sizeof(int)));
        NN(nn_b200_workspace_init(dWs, c.m, nullptr));
    }
    for (int it = 0; it < c.warmup + c.iters; ++it)
    {
        if (c.fused)
        {
            CK(cudaEventRecord(e0));
            NN(nn_b200_search_device(c.k, c.m, c.n, dS, dR, 0, dWs, dOut, c.fused == 2 ? dK : nullptr, nullptr));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (it >= c.warmup)
                ms.push_back(t);
            continue;
        }
        NN(nn_b200_keys_init(dK, c.m, nullptr));
        CK(cudaEventRecord(e0));
        if (c.soa)
            NN(nn_b200_nearest_keys_soa(c.k, c.m, c.n, dS, dRs, 0, dK, nullptr));
        else
            NN(nn_b200_nearest_keys(c.k, c.m, c.n, dS, dR, 0, dK, nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float t;
        CK(cudaEventElapsedTime(&t, e0, e1));
        if (it >= c.warmup)
            ms.push_back(t);
    }
    CK(cudaGetLastError());
#ifdef NN_QREG_TIMELINE
    if (c.timeline && c.fused)
    { // debug build: where does a small query-register search spend its time? (thread 0 of every CTA)
        const size_t max_ctas = 1u << 16;
        unsigned long long *dT = nullptr;
        CK(cudaMalloc(&dT, max_ctas * 16 * 8));
        nn_b200_debug_timeline(dT);
        std::vector<unsigned long long> h(max_ctas * 16);
        for (int rep = 0; rep < 3; ++rep)
        {
            CK(cudaMemset(dT, 0, max_ctas * 16 * 8));
            CK(cudaDeviceSynchronize());
            NN(nn_b200_search_device(c.k, c.m, c.n, dS, dR, 0, dWs, dOut, nullptr, nullptr));
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h.data(), dT, max_ctas * 16 * 8, cudaMemcpyDeviceToHost));
        nn_b200_debug_timeline(nullptr);
        CK(cudaFree(dT));
        std::vector<size_t> ctas;
        unsigned long long t0 = ~0ull;
        for (size_t i = 0; i < max_ctas; ++i)
            if (h[i * 16] && h[i * 16 + 5])
            {
                ctas.push_back(i);
                t0 = std::min(t0, h[i * 16]);
            }
        auto stat = [&](std::vector<double> v, const char *name) {
            std::sort(v.begin(), v.end());
            double sum = 0;
            for (double x : v)
                sum += x;
            printf("  %-28s min %8.2f  p10 %8.2f  med %8.2f  p90 %8.2f  max %8.2f  mean %8.2f\n", name, v.front(),
                   v[v.size() / 10], v[v.size() / 2], v[v.size() * 9 / 10], v.back(), sum / v.size());
        };
        printf("timeline k=%d m=%d n=%lld: %zu CTAs stamped (us; global timer relative to the first CTA's entry)\n", c.k,
               c.m, c.n, ctas.size());
        const char *names[6] = {"entry", "queries loaded", "first tile landed", "tile loop done", "folded", "finished"};
        for (int s = 0; s < 6; ++s)
        {
            std::vector<double> v;
            for (size_t i : ctas)
                v.push_back((double)(h[i * 16 + s] - t0) * 1e-3);
            stat(v, names[s]);
        }
        printf("phase lengths by the SM clock (us at 1.965 GHz):\n");
        for (int s = 1; s < 6; ++s)
        {
            std::vector<double> v;
            for (size_t i : ctas)
                v.push_back((double)(long long)(h[i * 16 + 8 + s] - h[i * 16 + 8 + s - 1]) / 1965.0);
            stat(v, names[s]);
        }
        fflush(stdout);
    }
#endif
    std::sort(ms.begin(), ms.end());
    const double med = ms[ms.size() / 2], best = ms[0];
    const double pairs = (double)c.m * (double)c.n;
    const double ops = 3.0 * c.k * pairs;
    const double bytes = (double)nr * 4.0;

    long long mismatches = -1;
    if (c.check)
    {
        NN(nn_b200_set_option("variant", 3));
        NN(nn_b200_keys_init(dK2, c.m, nullptr));
        NN(nn_b200_nearest_keys(c.k, c.m, c.n, dS, dR, 0, dK2, nullptr));
        CK(cudaDeviceSynchronize());
        std::vector<uint64_t> a(c.m), b(c.m);
        CK(cudaMemcpy(a.data(), dK, (size_t)c.m * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), dK2, (size_t)c.m * 8, cudaMemcpyDeviceToHost));
        mismatches = 0;
        if (c.fused)
        { // indices of the one-launch search against the plain kernel's; fused == 2 also returned the keys
            std::vector<int> r(c.m);
            CK(cudaMemcpy(r.data(), dOut, (size_t)c.m * sizeof(int), cudaMemcpyDeviceToHost));
            for (int i = 0; i < c.m; ++i)
            {
                if ((uint32_t)r[i] != (uint32_t)(b[i] & 0xffffffffu))
                    ++mismatches;
                if (c.fused != 2)
                    a[i] = b[i];
            }
            // the workspace must be back in its start state: a second search gives the same answer
            NN(nn_b200_set_option("variant", c.variant));
            NN(nn_b200_search_device(c.k, c.m, c.n, dS, dR, 0, dWs, dOut, nullptr, nullptr));
            CK(cudaDeviceSynchronize());
            std::vector<int> r2(c.m);
            CK(cudaMemcpy(r2.data(), dOut, (size_t)c.m * sizeof(int), cudaMemcpyDeviceToHost));
            for (int i = 0; i < c.m; ++i)
                if (r2[i] != r[i])
                    ++mismatches;
        }
        for (int i = 0; i < c.m; ++i)
            if (a[i] != b[i])
            {
                if (mismatches < 4)
                    fprintf(stderr, "  mismatch q=%d got %016llx want %016llx\n", i, (unsigned long long)a[i],
                            (unsigned long long)b[i]);
                ++mismatches;
            }
        NN(nn_b200_set_option("variant", c.variant));
    }

    printf("{\"op\":\"%s\",\"tag\":\"%s\",\"k\":%d,\"m\":%d,\"n\":%lld,\"soa\":%d,\"quant\":%d,\"ms_med\":%.4f,"
           "\"ms_best\":%.4f,\"pairs_per_s\":%.4e,\"fp32_frac_maxclk\":%.4f,\"GBps\":%.1f,\"mismatch_vs_plain\":%lld,"
           "\"plan\":\"%s\"}\n",
           c.fused ? "search_device" : "nearest_keys", c.tag.c_str(), c.k, c.m, c.n, c.soa, c.quant, med, best,
           pairs / (med * 1e-3), ops / (med * 1e-3) / g_peak_ops,
           bytes / med / 1e6, mismatches, plan);
    fflush(stdout);
    if (dRs)
        CK(cudaFree(dRs));
    if (dWs)
        CK(cudaFree(dWs));
    if (dOut)
        CK(cudaFree(dOut));
    CK(cudaFree(dS));
    CK(cudaFree(dR));
    CK(cudaFree(dK));
    CK(cudaFree(dK2));
}

int main(int argc, char **argv)
{
    Cfg c;
    std::string sweep;
    for (int i = 1; i < argc; ++i)
    {
        const std::string a = argv[i];
        auto val = [&]() -> const char * { return (i + 1 < argc) ? argv[++i] : "0"; };
        if (a == "--k")
            c.k = atoi(val());
        else if (a == "--m")
            c.m = atoi(val());
        else if (a == "--n")
            c.n = atoll(val());
        else if (a == "--variant")
            c.variant = atoi(val());
        else if (a == "--q")
            c.q = atoi(val());
        else if (a == "--math")
            c.scalar = atoi(val());
        else if (a == "--splits")
            c.splits = atoi(val());
        else if (a == "--waves")
            c.waves = atoi(val());
        else if (a == "--iters")
            c.iters = atoi(val());
        else if (a == "--warmup")
            c.warmup = atoi(val());
        else if (a == "--check")
            c.check = atoi(val());
        else if (a == "--soa")
            c.soa = atoi(val());
        else if (a == "--repack")
            c.repack = atoi(val());
        else if (a == "--rreg_ctas")
            c.rreg_ctas = atoi(val());
        else if (a == "--ldg")
            c.ldg = atoi(val());
        else if (a == "--quant")
            c.quant = atoi(val());
        else if (a == "--opt")
        { // --opt name=value: any nn_b200_set_option knob
            std::string kv = val();
            const size_t eq = kv.find('=');
            if (eq == std::string::npos || nn_b200_set_option(kv.substr(0, eq).c_str(), atoll(kv.c_str() + eq + 1)) != 0)
            {
                fprintf(stderr, "bad --opt %s\n", kv.c_str());
                return 1;
            }
        }
        else if (a == "--fused")
            c.fused = atoi(val());
        else if (a == "--timeline") // debug builds (-DNN_QREG_TIMELINE) only
            c.timeline = atoi(val());
        else if (a == "--tag")
            c.tag = val();
        else if (a == "--sweep")
            sweep = val();
        else
        {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 1;
        }
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    g_sms = prop.multiProcessorCount;
    g_clock_ghz = clock_khz / 1e6;
    g_peak_ops = (double)g_sms * 128.0 * clock_khz * 1e3;
    printf("{\"device\":\"%s\",\"sms\":%d,\"max_clock_ghz\":%.3f,\"fp32_peak_ops\":%.4e}\n", prop.name, g_sms,
           g_clock_ghz, g_peak_ops);

    if (sweep == "warmup")
    {
        CK(cudaFree(0));
        for (int i = 0; i < 2; ++i)
        {
            cudaEvent_t a;
            CK(cudaEventCreate(&a));
            const auto t0 = std::chrono::steady_clock::now();
            NN(nn_b200_warmup());
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            printf("{\"op\":\"warmup\",\"call\":%d,\"ms\":%.3f}\n", i, ms);
        }
        return 0;
    }
    if (sweep == "probe")
    {
        const char *names[6] = {"scalar", "packed", "packed/scalar alternating", "4 packed + 4 scalar blocks",
                                "packed + FMNMX", "scalar + FMNMX"};
        for (int mode = 0; mode < 6; ++mode)
        {
            double v = 0;
            NN(nn_b200_probe_fp32(mode, 20000, &v));
            printf("{\"op\":\"probe_fp32\",\"mode\":%d,\"name\":\"%s\",\"lane_ops_per_s\":%.4e,\"frac_nominal\":%.4f}\n", mode,
                   names[mode], v, v / g_peak_ops);
        }
        return 0;
    }
    if (sweep.empty())
    {
        if (c.tag.empty())
            c.tag = "single";
        run(c);
        return 0;
    }
    std::vector<Cfg> list;
    auto add = [&](const char *tag, int k, int m, long long n, int variant, int q, int scalar, int waves, int check,
                   int soa = 0, int repack = 0, int rreg_ctas = 0, int splits = 0) {
        Cfg x;
        x.tag = tag;
        x.k = k;
        x.m = m;
        x.n = n;
        x.variant = variant;
        x.q = q;
        x.scalar = scalar;
        x.waves = waves;
        x.check = check;
        x.soa = soa;
        x.repack = repack;
        x.rreg_ctas = rreg_ctas;
        x.splits = splits;
        x.iters = c.iters;
        x.warmup = c.warmup;
        list.push_back(x);
    };
    if (sweep == "math")
    {
        // math mode (2 f32x2 over query pairs, 1 f32x2 over dims, 0 scalar) x queries per thread
        for (int math : {2, 1, 0})
        {
            for (int q : {2, 4})
                add("k16", 16, 4096, 1 << 20, 1, q, math, 4, 0);
            for (int q : {2, 4, 8})
                add("k8", 8, 4096, 1 << 20, 1, q, math, 4, 0);
            for (int q : {2, 4, 8})
                add("k3", 3, 16384, 1 << 20, 1, q, math, 4, 0);
        }
    }
    else if (sweep == "cfgs")
    {
        add("cfg1", 3, 1024, 65536, 0, 0, 2, 4, 1);
        add("cfg2", 16, 4096, 1 << 20, 0, 0, 2, 4, 1);
        add("cfg3", 8, 8, 1 << 26, 0, 0, 2, 4, 1);
        add("cfg5-1/16", 3, 1 << 16, 1 << 20, 0, 0, 2, 4, 0);
        add("cfg4-1/64", 16, 65536, 1 << 18, 0, 0, 2, 4, 0);
    }
    else if (sweep == "small")
    {
        for (int ctas : {0, 1, 2, 3, 4})
            add("cfg3-rreg", 8, 8, 1 << 26, 2, 0, 2, 4, 0, 0, 0, ctas);
        add("cfg3-rreg-soa", 8, 8, 1 << 26, 2, 0, 2, 4, 0, 1);

        for (int m : {1, 2, 4, 8, 16, 32, 64})
            add("k8-m", 8, m, 1 << 24, 2, 0, 2, 4, 0);
        for (int m : {16, 32, 64, 128, 256, 512})
            add("k8-m-qreg", 8, m, 1 << 24, 1, 0, 2, 4, 0);
        for (int k : {3, 16})
            for (int m : {1, 8})
                add("k-m", k, m, 1 << 24, 2, 0, 2, 4, 0);
    }
    else if (sweep == "repack")
    {
        for (int k : {3, 8, 16})
            add("repack", k, 1, 1 << 26, 0, 0, 2, 4, 0, 0, 1);
        add("repack", 16, 1, (1 << 24) + 3, 0, 0, 2, 4, 0, 0, 1);
    }
    else if (sweep == "cross")
    {
        // kernel-family crossover over the few-query band: every family on every shape, plus the
        // planner's own pick (tag "auto"); scripts/make_crossover.py turns the lines into
        // profiles/r02_fewquery_crossover.json, which the CPU test of nn_b200_plan_variant reads
        for (int k : {3, 8, 16})
            for (long long n : {1LL << 16, 1LL << 20, 1LL << 22})
                for (int m : {5, 8, 9, 16, 24, 25, 32, 48, 64, 100, 112, 128, 200, 256, 500})
                {
                    add("auto", k, m, n, 0, 0, 2, 8, 0);
                    add("qreg", k, m, n, 1, 0, 2, 8, 0);
                    if (m <= 64)
                        add("rtma", k, m, n, 4, 0, 2, 8, 0);
                    add("qflex", k, m, n, 5, 0, 2, 8, 0);
                }
    }
    else if (sweep == "fuzz")
    {
        // random shapes, planner in charge, every result against the plain kernel: odd sizes, every k, query
        // counts on both sides of every family boundary, tie-heavy and plain data, one-launch and building-block paths
        uint64_t z = 0x243F6A8885A308D3ull;
        auto rnd = [&](uint64_t lim) {
            z ^= z << 13;
            z ^= z >> 7;
            z ^= z << 17;
            return (long long)(z % lim);
        };
        for (int i = 0; i < 240; ++i)
        {
            Cfg x;
            x.tag = "fuzz";
            x.k = 3 + (int)rnd(14);
            const int cls = (int)rnd(6);
            x.m = cls == 0 ? 1 + (int)rnd(8) : (cls == 1 ? 5 + (int)rnd(30) : (cls == 2 ? 20 + (int)rnd(120) : (cls == 3 ? 100 + (int)rnd(500) : 1 + (int)rnd(3000))));
            const int ncls = (int)rnd(5);
            x.n = ncls == 0 ? 1 + rnd(40) : (ncls == 1 ? 1 + rnd(3000) : (ncls == 2 ? 1000 + rnd(60000) : 20000 + rnd(400000)));
            x.variant = 0;
            x.check = 1;
            x.quant = (i % 3 == 0) ? 8 : 0;
            x.fused = i % 2;
            x.iters = 1;
            x.warmup = 0;
            list.push_back(x);
        }
    }
    else if (sweep == "check")
    {
        // correctness against the plain kernel on tie-heavy data, all k, awkward sizes
        for (int k = 3; k <= 16; ++k)
        {
            Cfg x;
            x.tag = "check";
            x.k = k;
            x.m = 1000 + k;
            x.n = 200000 + 37 * k;
            x.variant = 1;
            x.check = 1;
            x.quant = 8;
            x.iters = 1;
            x.warmup = 0;
            list.push_back(x);
            x.variant = 2;
            x.m = 15;
            list.push_back(x);
            x.m = 1;
            list.push_back(x);
            x.variant = 4;
            x.m = 15;
            list.push_back(x);
            x.variant = 5;
            x.m = 100 + k;
            list.push_back(x);
            x.m = 15;
            x.variant = 2;
            x.soa = 1;
            list.push_back(x);
        }
    }
    else
    {
        fprintf(stderr, "unknown sweep %s\n", sweep.c_str());
        return 1;
    }
    for (const Cfg &x : list)
        run(x);
    return 0;
}
