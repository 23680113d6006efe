// fgr_math.cuh -- per-item arithmetic of Fast Global Registration (Open3D FastGlobalRegistration.cpp as the
// reference calls it, ALL_FUNCTIONS.py:189-202), usable from device code and, for the CPU check in oracle/fpfh_engine.cpp,
// from host code: feature-space distance, the tuple test with its counter-based generator, and one correspondence's
// contribution to the normal equations of the graduated-non-convexity loop.  Plain fp64, no FMA contraction on either side.
#pragma once
#include "mgicp_math.cuh"

namespace mg {

// squared L2 distance of two 33-bin descriptors, summed in bin order (what the oracle's brute-force search evaluates)
MG_HD double fgr_feat_dist2(const double *a, const double *b) {
    double s = 0.0;
    for (int k = 0; k < 33; ++k) { const double d = a[k] - b[k]; s += d * d; }
    return s;
}

// k-th output (k = 0, 1, ...) of the splitmix64 sequence started at `seed`, upper half: the generator is a counter, so trial t
// of the tuple test reads outputs 3t, 3t+1, 3t+2 without any sequential state
MG_HD uint32_t fgr_rng(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

MG_HD double fgr_dist3(const V3 &a, const V3 &b) {
    const double x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
    return sqrt(x * x + y * y + z * z);
}

// AdvancedMatching's tuple constraint: the three edge lengths of the triangle agree within the factor `scale` in both clouds
MG_HD bool fgr_tuple_ok(const V3 &pi0, const V3 &pi1, const V3 &pi2, const V3 &pj0, const V3 &pj1, const V3 &pj2, double scale) {
    // edge by edge: most trials fail on the first edge, the other four square roots and two divisions are then not needed
    const double li0 = fgr_dist3(pi0, pi1), lj0 = fgr_dist3(pj0, pj1);
    if (!(li0 * scale < lj0 && lj0 < li0 / scale)) return false;
    const double li1 = fgr_dist3(pi1, pi2), lj1 = fgr_dist3(pj1, pj2);
    if (!(li1 * scale < lj1 && lj1 < li1 / scale)) return false;
    const double li2 = fgr_dist3(pi2, pi0), lj2 = fgr_dist3(pj2, pj0);
    return li2 * scale < lj2 && lj2 < li2 / scale;
}

// OptimizePairwiseRegistration, one correspondence (p fixed, q the moving copy): line-process weight s = (par / (|p-q|^2 + par))^2,
// rows J_x = (0, -q.z, q.y, -1, 0, 0), J_y = (q.z, 0, -q.x, 0, -1, 0), J_z = (-q.y, q.x, 0, 0, 0, -1), residuals p - q;
// acc[0..20] += upper(J^T J) s (row-major), acc[21..26] += J^T r s, each sum receiving its rows in x, y, z order.
// Open3D forms all 27 products per row; two thirds of them have a literal zero factor and add +-0 to a sum that is never -0
// (it starts at +0), and a factor -1 only flips a sign: leaving those out gives the same bits with a third of the arithmetic
// (for finite coordinates; 0 * inf would have poisoned the sums).  Each product keeps Open3D's association (J_i * J_j) * s.
template <class Acc>
MG_HD void fgr_accumulate(const V3 &p, const V3 &q, double par, Acc &&acc) {
    const double rx = p.x - q.x, ry = p.y - q.y, rz = p.z - q.z;
    const double temp = par / (rx * rx + ry * ry + rz * rz + par);
    const double s = temp * temp;
    const double nqx = -q.x, nqy = -q.y, nqz = -q.z;
    // row x: columns 1, 2, 3 = (-q.z, q.y, -1)
    acc[6] += nqz * nqz * s;   acc[7] += nqz * q.y * s;   acc[8] += q.z * s;
    acc[11] += q.y * q.y * s;  acc[12] += nqy * s;        acc[15] += s;
    acc[22] += nqz * rx * s;   acc[23] += q.y * rx * s;   acc[24] += -rx * s;
    // row y: columns 0, 2, 4 = (q.z, -q.x, -1)
    acc[0] += q.z * q.z * s;   acc[2] += q.z * nqx * s;   acc[4] += nqz * s;
    acc[11] += nqx * nqx * s;  acc[13] += q.x * s;        acc[18] += s;
    acc[21] += q.z * ry * s;   acc[23] += nqx * ry * s;   acc[25] += -ry * s;
    // row z: columns 0, 1, 5 = (-q.y, q.x, -1)
    acc[0] += nqy * nqy * s;   acc[1] += nqy * q.x * s;   acc[5] += q.y * s;
    acc[6] += q.x * q.x * s;   acc[10] += nqx * s;        acc[20] += s;
    acc[21] += nqy * rz * s;   acc[22] += q.x * rz * s;   acc[26] += -rz * s;
}

// GetTransformationOriginalScale followed by the inversion FastGlobalRegistration applies: `trans` maps the (centred, scaled)
// target onto the source; the result maps the original source into the original target frame.
MG_HD void fgr_finalize(const double trans[16], const double mean_src[3], const double mean_tgt[3], double scale_global, double T_out[16]) {
    double R[9], t[3], to[3];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) R[3 * a + b] = trans[4 * a + b];
        t[a] = trans[4 * a + 3];
    }
    for (int a = 0; a < 3; ++a)
        to[a] = -(R[3 * a] * mean_tgt[0] + R[3 * a + 1] * mean_tgt[1] + R[3 * a + 2] * mean_tgt[2]) + t[a] * scale_global + mean_src[a];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) T_out[4 * a + b] = R[3 * b + a];
        T_out[4 * a + 3] = -(R[a] * to[0] + R[3 + a] * to[1] + R[6 + a] * to[2]);
    }
    T_out[12] = 0; T_out[13] = 0; T_out[14] = 0; T_out[15] = 1;
}

}  // namespace mg
