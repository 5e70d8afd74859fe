// oracle/grid_ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" doorway into the UNMODIFIED reference grid_subsampling()
// (declared at cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.h:84-91),
// which takes std::vector arguments and so cannot be reached through ctypes
// directly.  Compiled by oracle/Makefile together with the reference's own
// grid_subsampling.cpp and cloud.cpp where they lie under /root/reference;
// nothing from the reference is copied into this repository.  The marshalling
// mirrors what the reference CPython wrapper does (wrapper.cpp:202-221, 246-265).
#include "grid_subsampling/grid_subsampling.h"   // found through -I<reference>/cpp_subsampling
#include <cstring>

extern "C" long ref_grid_subsampling(const float* points, size_t N,
                                     const float* features, size_t fdim,
                                     const int* classes, size_t ldim,
                                     float dl,
                                     float* sub_points, float* sub_features, int* sub_classes)
{
    std::vector<PointXYZ> op((const PointXYZ*)points, (const PointXYZ*)points + N);
    std::vector<float> of;
    std::vector<int> oc;
    if (fdim) of.assign(features, features + N * fdim);
    if (ldim) oc.assign(classes, classes + N * ldim);
    std::vector<PointXYZ> sp;
    std::vector<float> sf;
    std::vector<int> sc;
    grid_subsampling(op, sp, of, sf, oc, sc, dl, 0);
    std::memcpy(sub_points, sp.data(), sp.size() * sizeof(PointXYZ));
    if (fdim) std::memcpy(sub_features, sf.data(), sf.size() * sizeof(float));
    if (ldim) std::memcpy(sub_classes, sc.data(), sc.size() * sizeof(int));
    return (long)sp.size();
}
