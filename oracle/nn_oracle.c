/*
 * oracle/nn_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's serial brute-force 1-NN (`v0::cudaCallback`,
 * /root/reference/sources/src/core.cu:27-62) and of the TA data generator
 * (/root/reference/sources/src/generator.h:14-50, driven in the order of
 * /root/reference/sources/src/main.cu:28-39 and 59-65).
 *
 * Nothing in the product path (multicore-hw2_b200/, include/) may link, import or call
 * this file.  It is used by tests/, by __graft_entry__.smoke() and by bench.py's
 * cpu_baseline / --impl reference legs, and only as the checker or the CPU arm.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this file against the eight
 * index lines of the reference's results.csv (tests/golden/ta_results.json) and against
 * outputs of the reference's own v0 compiled from /root/reference (oracle/_ref,
 * tests/golden/ref_v0_cases.npz).
 *
 * Build: /usr/bin/gcc -O2 -ffp-contract=off -fopenmp (see oracle/Makefile).  The flags
 * are part of the contract: IEEE round-to-nearest float sub/mul/add, no contraction to
 * FMA, no reassociation, strictly sequential accumulation over the dimensions.
 */
#define _GNU_SOURCE
#include <math.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Host threads this process may run on (its CPU affinity), NOT omp_get_max_threads(): torchrun
 * exports OMP_NUM_THREADS=1, which would make "all host threads" mean one on multi-GPU runs. */
static int host_threads(void)
{
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0)
        return CPU_COUNT(&set);
    return 1;
}

/* One query against the whole reference set.  Follows core.cu:39-56:
 *   start state (INFINITY, index 0)                                   core.cu:39-40
 *   squareSum = 0; for kInd: diff = s - r; squareSum += diff*diff     core.cu:44-49
 *   update iff minSquareSum > squareSum  (strict: lowest index wins)  core.cu:50-54
 * `volatile`-free on purpose: -ffp-contract=off is what forbids the FMA. */
static inline int nn_one_query(int k, int n, const float *q, const float *R, float *min_out)
{
    float minSquareSum = INFINITY;
    int minIndex = 0;
    for (int nInd = 0; nInd < n; ++nInd) {
        const float *r = R + (size_t)k * (size_t)nInd;
        float squareSum = 0;
        for (int kInd = 0; kInd < k; ++kInd) {
            const float diff = q[kInd] - r[kInd];
            squareSum += diff * diff;
        }
        if (minSquareSum > squareSum) {
            minSquareSum = squareSum;
            minIndex = nInd;
        }
    }
    if (min_out)
        *min_out = minSquareSum;
    return minIndex;
}

/* Serial v0 (core.cu:35-59).  `out` has m ints; `min_sq` (optional) gets the winning
 * squared distance of each query so tests can also assert the distance bits. */
void nn_oracle_v0(int k, int m, int n, const float *S, const float *R, int *out, float *min_sq)
{
    for (int mInd = 0; mInd < m; ++mInd)
        out[mInd] = nn_one_query(k, n, S + (size_t)k * (size_t)mInd, R,
                                 min_sq ? &min_sq[mInd] : NULL);
}

/* Same arithmetic, queries spread over `threads` OpenMP threads.  Queries are independent
 * in v0 (outer loop core.cu:37), so this is bit-identical to nn_oracle_v0 by construction.
 * Returns the number of threads actually used. */
int nn_oracle_v0_mt(int k, int m, int n, const float *S, const float *R, int *out,
                    float *min_sq, int threads)
{
    int used = 1;
#ifdef _OPENMP
    if (threads < 1)
        threads = host_threads();
    if (threads > m)
        threads = m > 0 ? m : 1;
    used = threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int mInd = 0; mInd < m; ++mInd)
        out[mInd] = nn_one_query(k, n, S + (size_t)k * (size_t)mInd, R,
                                 min_sq ? &min_sq[mInd] : NULL);
    return used;
}

/* Squared distance of one (query, reference) pair with v0's arithmetic (core.cu:44-49). */
float nn_oracle_sqdist(int k, const float *q, const float *r)
{
    float squareSum = 0;
    for (int kInd = 0; kInd < k; ++kInd) {
        const float diff = q[kInd] - r[kInd];
        squareSum += diff * diff;
    }
    return squareSum;
}

/* The packed key the device path reduces with: (float bits of d^2) << 32 | index, with
 * v0's start state (INFINITY, 0) when nothing beat it.  Restated here so tests can compare
 * keys, not only indices. */
void nn_oracle_keys(int k, int m, int n, const float *S, const float *R, uint64_t *keys)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(host_threads())
#endif
    for (int mInd = 0; mInd < m; ++mInd) {
        float d;
        int idx = nn_one_query(k, n, S + (size_t)k * (size_t)mInd, R, &d);
        uint32_t bits;
        memcpy(&bits, &d, 4);
        keys[mInd] = ((uint64_t)bits << 32) | (uint32_t)idx;
    }
}

/* AoS [n][k] -> SoA [k][n]: what mat_inv_kernel computes (core.cu:792-807):
 *   output[nInd + kInd * n] = input[nInd * k + kInd]. */
void nn_oracle_repack_soa(int k, int n, const float *in, float *out)
{
    for (int nInd = 0; nInd < n; ++nInd)
        for (int kInd = 0; kInd < k; ++kInd)
            out[(size_t)nInd + (size_t)kInd * (size_t)n] = in[(size_t)nInd * (size_t)k + kInd];
}

/* ---- TA generator (generator.h:14-50) ------------------------------------------------
 * getRandNum(): rand() / double(RAND_MAX) narrowed to float (generator.h:17-19).
 * getSample():  k*m query floats first, then k*n reference floats (generator.h:37-48).
 * test():       ONE srand(seed) then the samples in table order (main.cu:59-65), so sample i
 *               depends on every earlier sample having been drawn.
 * glibc rand() defines the stream; the gpurun image has the same libc. */
static const int ta_samples[8][3] = {
    /* main.cu:28-39 */
    {3, 1, 2}, {3, 2, 8}, {3, 1, 1024}, {3, 1, 65536},
    {16, 1, 65536}, {3, 1024, 1024}, {3, 1024, 65536}, {16, 1024, 65536},
};

int nn_oracle_ta_num_samples(void) { return 8; }

void nn_oracle_ta_shape(int sample, int *k, int *m, int *n)
{
    *k = ta_samples[sample][0];
    *m = ta_samples[sample][1];
    *n = ta_samples[sample][2];
}

/* Fills S (k*m floats) and R (k*n floats) of TA sample `sample` under seed `seed`
 * (the reference uses 1000, main.cu:43).  Earlier samples are drawn and discarded. */
void nn_oracle_ta_sample(int seed, int sample, float *S, float *R)
{
    const double DOUBLE_RAND_MAX = (double)RAND_MAX; /* generator.h:14 */
    srand((unsigned)seed);
    for (int i = 0; i <= sample; ++i) {
        const int k = ta_samples[i][0], m = ta_samples[i][1], n = ta_samples[i][2];
        const int last = (i == sample);
        for (long j = 0; j < (long)k * m; ++j) {
            float v = (float)(rand() / DOUBLE_RAND_MAX);
            if (last)
                S[j] = v;
        }
        for (long j = 0; j < (long)k * n; ++j) {
            float v = (float)(rand() / DOUBLE_RAND_MAX);
            if (last)
                R[j] = v;
        }
    }
}
