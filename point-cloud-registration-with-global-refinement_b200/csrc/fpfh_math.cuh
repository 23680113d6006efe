// fpfh_math.cuh -- per-point arithmetic of the FGR front end's feature stage, usable from device code
// and, for the CPU check in oracle/fpfh_engine.cpp, from host code.  Restates Open3D's Feature.cpp (ComputePairFeatures,
// ComputeSPFHFeature, ComputeFPFHFeature) and the covariance step of EstimateNormals for one point given its hybrid-search
// neighbour list (ascending distance, entry 0 = the point itself), i.e. what
//     estimate_normals(KDTreeSearchParamHybrid(2 v, 20)) / compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(10 v, 200))
// (ALL_FUNCTIONS.py:181-187) compute per point.  Plain fp64, no FMA contraction on either side; atan2 comes from the
// platform's libm (host) or libdevice (device) and may differ in the last bit, which can only move a neighbour across a
// histogram-bin boundary it sits on; the acos that decides the frame is the deterministic one shared with the oracle.
#pragma once
#include "mgicp_math.cuh"

namespace mg {

// Open3D ComputePairFeatures: (atan2 angle, v . n2, n1 . d / |d| or its swapped counterpart, |d|)
MG_HD void fpfh_pair_features(const V3 &p1, const V3 &n1, const V3 &p2, const V3 &n2, double f[4]) {
    V3 dp = v3(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z);
    f[0] = f[1] = f[2] = f[3] = 0.0;
    const double len = sqrt(dp.x * dp.x + dp.y * dp.y + dp.z * dp.z);
    if (len == 0.0) return;
    V3 a = n1, b = n2;
    const double angle1 = (a.x * dp.x + a.y * dp.y + a.z * dp.z) / len;
    const double angle2 = (b.x * dp.x + b.y * dp.y + b.z * dp.z) / len;
    double f2;
    // Open3D compares acos(|angle1|) > acos(|angle2|) with its libm's acos.  For (nearly) parallel normals the two angles agree
    // to the last bits and the outcome -- which mirrors the feature -- hangs on how that acos rounds, so GPU and oracle both
    // use the deterministic fdlibm acos of mgicp_math.cuh (first GPU run with libdevice's acos: 13 % of the descriptors of an
    // NCLT cloud differed from the glibc oracle).  |angle| > 1 (rounding) gives NaN in Open3D: no swap.
    const double c1 = fabs(angle1), c2 = fabs(angle2);
    if (c1 <= 1.0 && c2 <= 1.0 && det_acos(c1) > det_acos(c2)) {
        // the normal with the smaller angle to the connecting line becomes the frame's first axis
        const V3 t = a; a = b; b = t;
        dp = v3(-dp.x, -dp.y, -dp.z);
        f2 = -angle2;
    } else f2 = angle1;
    V3 v = cross(dp, a);
    const double vn = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    if (vn == 0.0) return;
    v = v3(v.x / vn, v.y / vn, v.z / vn);
    const V3 w = cross(a, v);
    f[3] = len;
    f[2] = f2;
    f[1] = dot(v, b);
    f[0] = atan2(dot(w, b), dot(a, b));
}

MG_HD int fpfh_bin(double x) {
    int h = (int)floor(x);
    if (h < 0) h = 0;
    if (h >= 11) h = 10;
    return h;
}

// the three histogram bins one neighbour falls into
MG_HD void fpfh_bins(const double f[4], int bins[3]) {
    const double pi = 3.14159265358979323846;
    bins[0] = fpfh_bin(11.0 * (f[0] + pi) / (2.0 * pi));
    bins[1] = 11 + fpfh_bin(11.0 * (f[1] + 1.0) * 0.5);
    bins[2] = 22 + fpfh_bin(11.0 * (f[2] + 1.0) * 0.5);
}

// ComputeSPFHFeature for point i: `cnt` list entries (the point itself first), hist[33] zeroed by the caller.
// PointAt(j) -> V3 position, NormalAt(j) -> V3 normal of cloud point j.
template <class PointAt, class NormalAt>
MG_HD void spfh_point(const int32_t *idx, int cnt, const V3 &p, const V3 &n, PointAt point_at, NormalAt normal_at, double *hist) {
    if (cnt <= 1) return;
    const double incr = 100.0 / (double)(cnt - 1);
    for (int k = 1; k < cnt; ++k) {
        double f[4];
        int b[3];
        fpfh_pair_features(p, n, point_at(idx[k]), normal_at(idx[k]), f);
        fpfh_bins(f, b);
        hist[b[0]] += incr; hist[b[1]] += incr; hist[b[2]] += incr;
    }
}

// a / b correctly rounded, given y = RN(1 / b) (one real division per neighbour instead of 33): q0 = RN(a y) is within 2 ulp
// of the quotient, one residual correction makes it faithful, and the second one rounds it correctly (Markstein's theorem;
// the residuals a - b q are exact in an FMA).  Normal, finite operands far from the overflow / underflow thresholds: SPFH
// values are 0 or >= 0.5, squared distances lie in [1e-30, r^2].  Checked against the division on 2e8 operand pairs
// (random, neighbourhoods of exact quotients, perturbed quotients) by tests/test_fgr_oracle.py::test_division_by_reciprocal_is_the_division.
MG_HD double div_by_recip(double a, double b, double y) {
    double q = a * y;
    double r = fma(-b, q, a);
    q = fma(r, y, q);
    r = fma(-b, q, a);
    return fma(r, y, q);
}

// ComputeFPFHFeature for point i: neighbours' SPFH weighted by 1 / d^2 (squared distances, as Open3D passes them), each third
// renormalised to 100, own SPFH added.  SpfhAt(j) -> const double* (33 values).  out[33] zeroed by the caller.
template <class SpfhAt>
MG_HD void fpfh_point(const int32_t *idx, const double *d2, int cnt, const double *own_spfh, SpfhAt spfh_at, double *out) {
    if (cnt <= 1) return;
    double sum[3] = {0.0, 0.0, 0.0};
    for (int k = 1; k < cnt; ++k) {
        const double dist = d2[k];
        if (dist == 0.0) continue;
        const double *s = spfh_at(idx[k]);
        const double y = 1.0 / dist;
        for (int j = 0; j < 33; ++j) {
            const double val = div_by_recip(s[j], dist, y);          // == s[j] / dist
            sum[j / 11] += val;
            out[j] += val;
        }
    }
    for (int j = 0; j < 3; ++j)
        if (sum[j] != 0.0) sum[j] = 100.0 / sum[j];
    for (int j = 0; j < 33; ++j) { out[j] *= sum[j / 11]; out[j] += own_spfh[j]; }
}

// EstimateNormals' covariance for one point from its neighbour list (>= 3 entries, else identity), cumulant form in list order
template <class PointAt>
MG_HD void hybrid_covariance(const int32_t *idx, int cnt, PointAt point_at, double cov[6]) {
    cov[0] = 1; cov[1] = 0; cov[2] = 0; cov[3] = 1; cov[4] = 0; cov[5] = 1;
    if (cnt < 3) return;
    double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < cnt; ++j) {
        const V3 p = point_at(idx[j]);
        cu[0] += p.x; cu[1] += p.y; cu[2] += p.z;
        cu[3] += p.x * p.x; cu[4] += p.x * p.y; cu[5] += p.x * p.z;
        cu[6] += p.y * p.y; cu[7] += p.y * p.z; cu[8] += p.z * p.z;
    }
    for (int j = 0; j < 9; ++j) cu[j] /= (double)cnt;
    cov[0] = cu[3] - cu[0] * cu[0]; cov[1] = cu[4] - cu[0] * cu[1]; cov[2] = cu[5] - cu[0] * cu[2];
    cov[3] = cu[6] - cu[1] * cu[1]; cov[4] = cu[7] - cu[1] * cu[2]; cov[5] = cu[8] - cu[2] * cu[2];
}

}  // namespace mg
