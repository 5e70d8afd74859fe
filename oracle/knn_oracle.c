/*
 * oracle/knn_oracle.c -- TEST INFRASTRUCTURE ONLY. NOT A PRODUCT PATH.
 *
 * CPU restatement (plain C, brute force) of the reference's batched exact KNN
 * index build.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file; the product
 * (ffb6d_b200/) never links, imports or falls back to it.
 *
 * What it restates (paths relative to /root/reference/ffb6d/models/RandLA/utils/
 * nearest_neighbors/):
 *   - cpp_knn_batch / cpp_knn_batch_omp           knn_.cxx:72-135
 *       per batch item, per query: K nearest support points, ascending
 *       squared distance, indices written as int64 (long) [B,Q,K].
 *   - L2_Adaptor::evalMetric, dim==3 tail loop    nanoflann.hpp:343-346
 *       result = 0; result += d0*d0; result += d1*d1; result += d2*d2;
 *       fp32 throughout, query minus support, NO fused multiply-add (the
 *       reference is built for baseline x86-64, NN/setup.py:12-13).
 *   - KNNResultSet::addPoint                      nanoflann.hpp:115-139
 *       insertion sort, strict `dists[i-1] > dist` => an equal distance is
 *       placed AFTER existing entries; a candidate equal to the worst of a
 *       full set is dropped.
 *   - K > S: slots >= S keep the value-initialised 0 (knn_.cxx:120-121).
 *
 * Difference by construction: the reference visits candidates in KD-tree
 * order, this file in ascending support index.  On inputs without exact
 * distance ties among the K+1 nearest the outputs are identical (verified
 * against oracle/_ref/libknn_ref.so by tests/test_oracle_vs_ref.py and the
 * committed tests/golden fixtures); on tied rows the sorted distances are
 * identical and the index order follows "lowest support index first".
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -march).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist3(const float *q, const float *s)
{
    /* nanoflann.hpp:343-346 -- sequential, unfused */
    float r = 0.0f;
    const float d0 = q[0] - s[0];
    r += d0 * d0;
    const float d1 = q[1] - s[1];
    r += d1 * d1;
    const float d2 = q[2] - s[2];
    r += d2 * d2;
    return r;
}

/* one query against one support cloud; dists/ids have room for K entries */
static void knn_one(const float *support, size_t S, const float *q, size_t K,
                    float *dists, int64_t *ids)
{
    size_t count = 0;
    for (size_t j = 0; j < K; ++j) { ids[j] = 0; dists[j] = 0.0f; }
    for (size_t s = 0; s < S; ++s) {
        const float d = sqdist3(q, support + 3 * s);
        /* leaf test of searchLevel (nanoflann.hpp:1361): strict < worst when full */
        if (count == K && !(d < dists[K - 1])) continue;
        size_t i;
        for (i = count; i > 0; --i) {           /* nanoflann.hpp:118-129 */
            if (dists[i - 1] > d) {
                if (i < K) { dists[i] = dists[i - 1]; ids[i] = ids[i - 1]; }
            } else break;
        }
        if (i < K) { dists[i] = d; ids[i] = (int64_t)s; }
        if (count < K) count++;
    }
}

/*
 * support [B,S,3] f32, query [B,Q,3] f32 -> idx [B,Q,K] int64 (and, when
 * dist_out != NULL, the matching squared distances [B,Q,K] f32; slots >= S are 0).
 * Same argument order and meaning as cpp_knn_batch_omp (knn_.h:14-16), dim fixed to 3.
 */
void oracle_knn_batch(const float *support, size_t B, size_t S,
                      const float *query, size_t Q, size_t K,
                      int64_t *idx_out, float *dist_out)
{
    if (K == 0) return;
#pragma omp parallel
    {
        float *dists = (float *)malloc(sizeof(float) * K);
        int64_t *ids = (int64_t *)malloc(sizeof(int64_t) * K);
#pragma omp for collapse(2) schedule(static)
        for (size_t b = 0; b < B; ++b) {
            for (size_t i = 0; i < Q; ++i) {
                knn_one(support + b * S * 3, S, query + (b * Q + i) * 3, K, dists, ids);
                memcpy(idx_out + (b * Q + i) * K, ids, sizeof(int64_t) * K);
                if (dist_out) memcpy(dist_out + (b * Q + i) * K, dists, sizeof(float) * K);
            }
        }
        free(dists);
        free(ids);
    }
}

/* squared distances of given (query, support-index) pairs, reference arithmetic.
 * Used by the tie-aware comparison in tests (sorted-distance equality). */
void oracle_sqdist_of_indices(const float *support, size_t B, size_t S,
                              const float *query, size_t Q, size_t K,
                              const int64_t *idx, float *dist_out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (size_t b = 0; b < B; ++b)
        for (size_t i = 0; i < Q; ++i)
            for (size_t k = 0; k < K; ++k) {
                const int64_t s = idx[(b * Q + i) * K + k];
                dist_out[(b * Q + i) * K + k] =
                    (s >= 0 && (size_t)s < S)
                        ? sqdist3(query + (b * Q + i) * 3, support + (b * S + (size_t)s) * 3)
                        : -1.0f;
            }
}
