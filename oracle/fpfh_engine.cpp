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

/* ------------------------------------------------------------------------------------------------------------------
 * Fast Global Registration through the per-item functions of csrc/fgr_math.cuh (what the CUDA kernels call), in plain
 * sequential order: the cross-check of those functions and of the kernels' control flow (source/target swap, mutual
 * matches in ascending order, counter-based tuple test with the cap, graduated non-convexity, undoing the normalisation)
 * against the independent restatement in fgr_oracle.c.
 * ------------------------------------------------------------------------------------------------------------------ */
#include "../point-cloud-registration-with-global-refinement_b200/csrc/fgr_math.cuh"

struct orc_fgr_opts_mirror {          /* layout of orc_fgr_opts (fgr_oracle.c) */
    double division_factor; int32_t use_absolute_scale; int32_t decrease_mu; double maximum_correspondence_distance;
    int32_t iteration_number; double tuple_scale; int32_t maximum_tuple_count; uint64_t seed;
};

static int fgr_engine_impl(const double *src_xyz, int64_t ns, const double *tgt_xyz, int64_t nt, const double *src_feat,
                           const double *tgt_feat, const orc_fgr_opts_mirror *o, double T_out[16], int64_t *n_corres_out,
                           bool kernel_order) {
    const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::memcpy(T_out, I4, sizeof(I4));
    if (n_corres_out) *n_corres_out = 0;
    if (ns <= 0 || nt <= 0) return 0;
    const int64_t n[2] = {ns, nt};
    std::vector<V3> P[2];
    const double *X[2] = {src_xyz, tgt_xyz};
    double mean[2][3], scale = 0.0;
    for (int c = 0; c < 2; ++c) {
        double m[3] = {0, 0, 0};
        for (int64_t i = 0; i < n[c]; ++i) { m[0] += X[c][3 * i]; m[1] += X[c][3 * i + 1]; m[2] += X[c][3 * i + 2]; }
        for (int k = 0; k < 3; ++k) { m[k] /= (double)n[c]; mean[c][k] = m[k]; }
        P[c].resize((size_t)n[c]);
        double mx = 0.0;
        for (int64_t i = 0; i < n[c]; ++i) {
            const V3 p = v3(X[c][3 * i] - m[0], X[c][3 * i + 1] - m[1], X[c][3 * i + 2] - m[2]);
            P[c][i] = p;
            const double t = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
            if (t > mx) mx = t;
        }
        if (mx > scale) scale = mx;
    }
    const double scale_global = o->use_absolute_scale ? 1.0 : scale, scale_start = o->use_absolute_scale ? scale : 1.0;
    for (int c = 0; c < 2; ++c)
        for (auto &p : P[c]) p = v3(p.x / scale_global, p.y / scale_global, p.z / scale_global);
    int fi = 0, fj = 1;
    const bool swapped = n[1] > n[0];
    if (swapped) { fi = 1; fj = 0; }
    const double *F[2] = {src_feat, tgt_feat};
    const int64_t ni = n[fi], nj = n[fj];
    std::vector<int32_t> j2i((size_t)nj), i2j((size_t)ni);
    auto nn = [&](const double *A, int64_t na, const double *B, int64_t nb, std::vector<int32_t> &out) {
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t a = 0; a < na; ++a) {
            double best = INFINITY; int32_t bj = -1;
            for (int64_t b = 0; b < nb; ++b) {
                const double s = fgr_feat_dist2(A + 33 * a, B + 33 * b);
                if (s < best) { best = s; bj = (int32_t)b; }
            }
            out[a] = bj;
        }
    };
    nn(F[fj], nj, F[fi], ni, j2i);
    nn(F[fi], ni, F[fj], nj, i2j);
    std::vector<int32_t> cross;          /* (i, j) mutual nearest neighbours, ascending i */
    for (int64_t i = 0; i < ni; ++i) {
        const int32_t j = i2j[i];
        if (j >= 0 && j2i[j] == (int32_t)i) { cross.push_back((int32_t)i); cross.push_back(j); }
    }
    const int64_t ncross = (int64_t)cross.size() / 2, cap = o->maximum_tuple_count > 0 ? o->maximum_tuple_count : 0;
    std::vector<int32_t> cor;            /* (index in cloud 0, index in cloud 1) per correspondence */
    if (ncross > 0 && cap > 0) {
        int64_t accepted = 0;
        for (int64_t t = 0; t < ncross * 100 && accepted < cap; ++t) {
            const int64_t r0 = fgr_rng(o->seed, 3 * (uint64_t)t) % (uint64_t)ncross, r1 = fgr_rng(o->seed, 3 * (uint64_t)t + 1) % (uint64_t)ncross,
                          r2 = fgr_rng(o->seed, 3 * (uint64_t)t + 2) % (uint64_t)ncross;
            const int32_t i0 = cross[2 * r0], j0 = cross[2 * r0 + 1], i1 = cross[2 * r1], j1 = cross[2 * r1 + 1], i2 = cross[2 * r2], j2 = cross[2 * r2 + 1];
            if (!fgr_tuple_ok(P[fi][i0], P[fi][i1], P[fi][i2], P[fj][j0], P[fj][j1], P[fj][j2], o->tuple_scale)) continue;
            const int32_t tri[6] = {i0, j0, i1, j1, i2, j2};
            for (int k = 0; k < 3; ++k) { cor.push_back(swapped ? tri[2 * k + 1] : tri[2 * k]); cor.push_back(swapped ? tri[2 * k] : tri[2 * k + 1]); }
            ++accepted;
        }
    }
    const int64_t nc = (int64_t)cor.size() / 2;
    if (n_corres_out) *n_corres_out = nc;
    double trans[16];
    std::memcpy(trans, I4, sizeof(I4));
    if (nc >= 10) {
        double par = scale_start;
        std::vector<V3> Q = P[1];
        for (int itr = 0; itr < o->iteration_number; ++itr) {
            double acc[27] = {0};
            if (!kernel_order) {
                for (int64_t c = 0; c < nc; ++c) fgr_accumulate(P[0][cor[2 * c]], Q[cor[2 * c + 1]], par, acc);
            } else {
                /* the reduction tree of k_fgr_pair: 512 thread-strided partials, shuffle-down tree per warp, warps in order */
                const int NT = 512;
                std::vector<double> part((size_t)NT * 27, 0.0);
                for (int t = 0; t < NT; ++t)
                    for (int64_t c = t; c < nc; c += NT) fgr_accumulate(P[0][cor[2 * c]], Q[cor[2 * c + 1]], par, &part[(size_t)t * 27]);
                for (int a = 0; a < 27; ++a) {
                    double s = 0.0;
                    for (int w = 0; w < NT / 32; ++w) {
                        double v[32];
                        for (int l = 0; l < 32; ++l) v[l] = part[(size_t)(w * 32 + l) * 27 + a];
                        for (int off = 16; off > 0; off >>= 1)
                            for (int l = 0; l < off; ++l) v[l] = v[l] + v[l + off];
                        s += v[0];
                    }
                    acc[a] = s;
                }
            }
            double x[6], delta[16];
            ldlt_solve6(acc, x);          /* JTJ x = -JTr  ==  SolveLinearSystemPSD(-JTJ, JTr) */
            vec6_to_mat4(x, delta);
            mat4_mul(delta, trans, trans);
            for (auto &q : Q) q = transform_point(delta, q);
            if (o->decrease_mu && itr % 4 == 0 && par > o->maximum_correspondence_distance) par /= o->division_factor;
        }
    }
    fgr_finalize(trans, mean[0], mean[1], scale_global, T_out);
    return 0;
}

extern "C" int orc_fgr_engine(const double *src_xyz, int64_t ns, const double *tgt_xyz, int64_t nt, const double *src_feat,
                              const double *tgt_feat, const orc_fgr_opts_mirror *o, double T_out[16], int64_t *n_corres_out) {
    return fgr_engine_impl(src_xyz, ns, tgt_xyz, nt, src_feat, tgt_feat, o, T_out, n_corres_out, false);
}

/* the same with the 27 sums of every iteration reduced in the order of the CUDA kernel k_fgr_pair: what the GPU result should
 * equal bit for bit */
extern "C" int orc_fgr_engine_kernel_order(const double *src_xyz, int64_t ns, const double *tgt_xyz, int64_t nt, const double *src_feat,
                                           const double *tgt_feat, const orc_fgr_opts_mirror *o, double T_out[16], int64_t *n_corres_out) {
    return fgr_engine_impl(src_xyz, ns, tgt_xyz, nt, src_feat, tgt_feat, o, T_out, n_corres_out, true);
}



/* div_by_recip(a, b, RN(1 / b)) against a / b: returns the number of operand pairs on which they differ (expected: 0).
 * A third of the pairs are random (a in the range of SPFH sums, b in the range of squared distances), a third are built so that
 * the quotient is exact or a few ulps of the dividend beside an exact one, a third with that quotient perturbed by half an ulp. */
extern "C" int64_t orc_check_recip_div(int64_t n, uint64_t seed) {
    int64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint64_t x = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
        auto next = [&x]() { x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31; x += 0x9E3779B97F4A7C15ull; return x; };
        auto unit = [&]() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); };
        double a, b;
        const int kind = (int)(i % 3);
        b = ldexp(0.5 + 0.5 * unit(), (int)(next() % 60) - 50);                  /* 2^-51 .. 2^9 */
        if (kind == 0) a = (next() % 7 == 0) ? 0.0 : ldexp(0.5 + 0.5 * unit(), (int)(next() % 20) - 2);
        else {
            /* q with few significant bits times b is (nearly) exact: a = RN(q b) lands on or next to an exact quotient; adding
             * about half an ulp of q (kind 2) perturbs it */
            const int bits = 1 + (int)(next() % 52);
            double q = ldexp(floor(ldexp(0.5 + 0.5 * unit(), bits)), -bits + (int)(next() % 12) - 2);
            if (kind == 2) q += ldexp(q, -53) * (1.0 + 2.0 * (double)(next() % 2));
            a = q * b;
            const int nudge = (int)(next() % 5) - 2;
            for (int t = 0; t < (nudge < 0 ? -nudge : nudge); ++t) a = nextafter(a, nudge < 0 ? 0.0 : INFINITY);
        }
        const double y = 1.0 / b;
        if (div_by_recip(a, b, y) != a / b) ++bad;
    }
    return bad;
}
