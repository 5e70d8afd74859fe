/*
 * oracle/grid_oracle.c -- TEST INFRASTRUCTURE ONLY. NOT A PRODUCT PATH.
 *
 * CPU restatement (plain C) of the reference's voxel-grid barycentre
 * subsampling.  Paths relative to /root/reference/ffb6d/models/RandLA/utils/
 * cpp_wrappers/:
 *   - grid_subsampling()        cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106
 *   - SampledData::update_*     cpp_subsampling/grid_subsampling/grid_subsampling.h:42-79
 *   - min_point / max_point     cpp_utils/cloud/cloud.cpp:27-67
 *
 * Arithmetic restated exactly (all fp32 unless noted):
 *   origin = floor(minCorner * (1/dl)) * dl                          (:27)
 *   NX = (size_t)floor((max.x - origin.x) / dl) + 1, same for NY      (:30-31)
 *   iX = (size_t)floor((p.x - origin.x) / dl) ..., key = iX + NX*iY + NX*NY*iZ  (:53-56)
 *   per voxel, IN INPUT ORDER: count += 1; sum += p; feat_sum += f; label histogram
 *   barycentre = sum * (float)(1.0 / (double)count)                   (:87, double reciprocal)
 *   mean feature = feat_sum / (float)count                            (:90-94)
 *   label = a label with the maximal count                            (:99-101)
 *
 * Differences by construction: the reference emits voxels in libstdc++
 * unordered_map iteration order and, among labels tied for the maximal
 * count, picks the first in that order; this file emits voxels by ascending
 * key and picks the smallest tied label.  Tests therefore compare rows after
 * sorting and accept any label whose count is maximal.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t key; uint32_t i; } keyed_t;

static int cmp_keyed(const void *a, const void *b)
{
    const keyed_t *x = (const keyed_t *)a, *y = (const keyed_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->i < y->i ? -1 : (x->i > y->i);   /* keep input order inside a voxel */
}

static int cmp_int(const void *a, const void *b)
{
    const int x = *(const int *)a, y = *(const int *)b;
    return x < y ? -1 : (x > y);
}

/*
 * points [N,3] f32, features [N,fdim] f32 or NULL (fdim 0), classes [N,ldim] i32 or NULL (ldim 0).
 * Outputs must have room for N rows.  Returns the number M of occupied voxels, or -1.
 * keys_out (optional, [N] u64) receives each output row's voxel key.
 */
long oracle_grid_subsampling(const float *points, size_t N,
                             const float *features, size_t fdim,
                             const int *classes, size_t ldim,
                             float dl,
                             float *sub_points, float *sub_features, int *sub_classes,
                             uint64_t *keys_out)
{
    if (N == 0) return -1;
    float mn[3] = { points[0], points[1], points[2] };
    float mx[3] = { points[0], points[1], points[2] };
    for (size_t i = 0; i < N; ++i)
        for (int a = 0; a < 3; ++a) {
            const float v = points[3 * i + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    const float inv = 1 / dl;
    float org[3];
    for (int a = 0; a < 3; ++a) org[a] = floorf(mn[a] * inv) * dl;
    const size_t NX = (size_t)floorf((mx[0] - org[0]) / dl) + 1;
    const size_t NY = (size_t)floorf((mx[1] - org[1]) / dl) + 1;

    keyed_t *kv = (keyed_t *)malloc(sizeof(keyed_t) * N);
    if (!kv) return -1;
    for (size_t i = 0; i < N; ++i) {
        const size_t iX = (size_t)floorf((points[3 * i + 0] - org[0]) / dl);
        const size_t iY = (size_t)floorf((points[3 * i + 1] - org[1]) / dl);
        const size_t iZ = (size_t)floorf((points[3 * i + 2] - org[2]) / dl);
        kv[i].key = (uint64_t)(iX + NX * iY + NX * NY * iZ);
        kv[i].i = (uint32_t)i;
    }
    qsort(kv, N, sizeof(keyed_t), cmp_keyed);

    int *lab = ldim ? (int *)malloc(sizeof(int) * N) : NULL;
    long M = 0;
    size_t a = 0;
    while (a < N) {
        size_t b = a;
        while (b < N && kv[b].key == kv[a].key) ++b;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (size_t f = 0; f < fdim; ++f) sub_features[(size_t)M * fdim + f] = 0.f;
        for (size_t t = a; t < b; ++t) {
            const size_t i = kv[t].i;
            sx += points[3 * i + 0];
            sy += points[3 * i + 1];
            sz += points[3 * i + 2];
            for (size_t f = 0; f < fdim; ++f)
                sub_features[(size_t)M * fdim + f] += features[i * fdim + f];
        }
        const int count = (int)(b - a);
        const float r = (float)(1.0 / count);
        sub_points[3 * M + 0] = sx * r;
        sub_points[3 * M + 1] = sy * r;
        sub_points[3 * M + 2] = sz * r;
        const float fc = (float)count;
        for (size_t f = 0; f < fdim; ++f) sub_features[(size_t)M * fdim + f] /= fc;
        for (size_t l = 0; l < ldim; ++l) {
            for (size_t t = a; t < b; ++t) lab[t - a] = classes[(size_t)kv[t].i * ldim + l];
            qsort(lab, (size_t)count, sizeof(int), cmp_int);
            int best = lab[0], best_n = 0, run = 0;
            for (int t = 0; t < count; ++t) {
                run = (t > 0 && lab[t] == lab[t - 1]) ? run + 1 : 1;
                if (run > best_n) { best_n = run; best = lab[t]; }
            }
            sub_classes[(size_t)M * ldim + l] = best;
        }
        if (keys_out) keys_out[M] = kv[a].key;
        ++M;
        a = b;
    }
    free(kv);
    free(lab);
    return M;
}
