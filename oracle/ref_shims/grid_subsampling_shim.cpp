// extern "C" door into the REFERENCE's grid_subsampling() (utils/cpp_wrappers/cpp_subsampling/grid_subsampling/
// grid_subsampling.cpp, compiled untouched beside this file by oracle/build_ref.sh).  TEST INFRASTRUCTURE: the
// reference exposes the function only through a CPython module (wrapper.cpp) that needs numpy's C headers; this shim
// passes plain arrays instead.  Outputs come back in the reference's own (unordered_map iteration) order.
#include <cstring>
#include <vector>

#include "grid_subsampling/grid_subsampling.h"

extern "C" int ref_grid_subsampling(int n, const float *points, int fdim, const float *features, int ldim,
                                    const int *classes, float sample_dl, float *sub_points, float *sub_features,
                                    int *sub_classes) {
    std::vector<PointXYZ> op((size_t)n), sp;
    for (int i = 0; i < n; ++i) op[i] = PointXYZ(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    std::vector<float> of, sf;
    std::vector<int> oc, sc;
    if (fdim > 0) of.assign(features, features + (size_t)n * fdim);
    if (ldim > 0) oc.assign(classes, classes + (size_t)n * ldim);
    grid_subsampling(op, sp, of, sf, oc, sc, sample_dl, 0);
    for (size_t i = 0; i < sp.size(); ++i) {
        sub_points[3 * i] = sp[i].x;
        sub_points[3 * i + 1] = sp[i].y;
        sub_points[3 * i + 2] = sp[i].z;
    }
    if (fdim > 0) std::memcpy(sub_features, sf.data(), sf.size() * sizeof(float));
    if (ldim > 0) std::memcpy(sub_classes, sc.data(), sc.size() * sizeof(int));
    return (int)sp.size();
}
