/*
 * fpfh_engine.cpp -- CPU evaluation of the per-point functions the FGR-front-end kernels are built from (TEST INFRASTRUCTURE).
 *
 * csrc/fpfh_math.cuh holds the arithmetic of hybrid-radius normals, SPFH and FPFH for one point given its neighbour list,
 * written once for device and host.  This file runs exactly those functions over a whole cloud on the CPU, with the
 * neighbour lists supplied by the caller (the oracle's KD-tree), so that tests/test_fgr_oracle.py can hold them against the
 * independent restatement in oracle/fgr_oracle.c before any of it runs on a GPU.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../point-cloud-registration-with-global-refinement_b200/csrc/fpfh_math.cuh"

using namespace mg;

extern "C" int orc_fpfh_engine(const double *xyz, int64_t n, const int32_t *idx_n, const int32_t *cnt_n, int cap_n,
                               const int32_t *idx_f, const double *d2_f, const int32_t *cnt_f, int cap_f, double *normals,
                               double *fpfh) {
    if (n < 0 || cap_n < 1 || cap_f < 1) return 1;
    auto point_at = [&](int32_t j) { return v3(xyz[3 * (int64_t)j], xyz[3 * (int64_t)j + 1], xyz[3 * (int64_t)j + 2]); };
    auto normal_at = [&](int32_t j) { return v3(normals[3 * (int64_t)j], normals[3 * (int64_t)j + 1], normals[3 * (int64_t)j + 2]); };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double cov[6];
        hybrid_covariance(idx_n + (int64_t)cap_n * i, cnt_n[i], point_at, cov);
        const V3 nv = normal_from_cov(cov);
        normals[3 * i] = nv.x; normals[3 * i + 1] = nv.y; normals[3 * i + 2] = nv.z;
    }
    std::vector<double> spfh((size_t)n * 33, 0.0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i)
        spfh_point(idx_f + (int64_t)cap_f * i, cnt_f[i], point_at((int32_t)i), normal_at((int32_t)i), point_at, normal_at, &spfh[33 * i]);
    std::memset(fpfh, 0, sizeof(double) * (size_t)n * 33);
    auto spfh_at = [&](int32_t j) { return (const double *)&spfh[33 * (int64_t)j]; };
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i)
        fpfh_point(idx_f + (int64_t)cap_f * i, d2_f + (int64_t)cap_f * i, cnt_f[i], &spfh[33 * i], spfh_at, fpfh + 33 * i);
    return 0;
}
