// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/nn_oracle.c header).
//
// C-ABI shim around the REFERENCE's own serial implementation `v0::cudaCallback`
// (/root/reference/sources/src/core.cu:25-63).  The reference source is never copied into
// this repository: oracle/Makefile extracts the `namespace v0` block from core.cu where it
// lies into the git-ignored oracle/_ref/ directory at build time and compiles it together
// with this shim.  The shim adds nothing to the arithmetic; it only adapts the calling
// convention (the reference mallocs its result, core.cu:35/59) and lets several host
// threads each run the unmodified v0 on a slice of the queries (v0 is re-entrant and its
// queries are independent, core.cu:37).
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#ifdef _OPENMP
#include <omp.h>
#endif

// Host threads this process may run on.  Deliberately NOT omp_get_max_threads(): launchers such as
// torchrun export OMP_NUM_THREADS=1, which would silently turn the "all host cores" baseline into a
// single-thread one on multi-GPU runs.
static int host_threads()
{
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0)
    {
        const int c = CPU_COUNT(&set);
        if (c > 0)
            return c;
    }
    return 1;
}

namespace v0
{
    extern void cudaCallback(int k, int m, int n, float *searchPoints, float *referencePoints, int **results);
}

extern "C" void ref_v0(int k, int m, int n, float *S, float *R, int *out)
{
    int *res = nullptr;
    v0::cudaCallback(k, m, n, S, R, &res);
    memcpy(out, res, sizeof(int) * (size_t)m);
    free(res);
}

// Every host thread runs the reference's v0 on a contiguous slice of the queries.
// Returns the number of threads used.
extern "C" int ref_v0_mt(int k, int m, int n, float *S, float *R, int *out, int threads)
{
    int used = 1;
#ifdef _OPENMP
    if (threads < 1)
        threads = host_threads();
    if (threads > m)
        threads = m > 0 ? m : 1;
    used = threads;
    // slices of <= 16 queries handed out dynamically, so a busy core does not set the time
    const int slice = m / (threads * 8) > 16 ? 16 : (m / (threads * 8) > 0 ? m / (threads * 8) : 1);
    const int nslices = (m + slice - 1) / slice;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int s = 0; s < nslices; ++s)
    {
        const int lo = s * slice, cnt = (lo + slice <= m) ? slice : m - lo;
        int *res = nullptr;
        v0::cudaCallback(k, cnt, n, S + (size_t)k * lo, R, &res);
        memcpy(out + lo, res, sizeof(int) * (size_t)cnt);
        free(res);
    }
#else
    ref_v0(k, m, n, S, R, out);
#endif
    return used;
}
