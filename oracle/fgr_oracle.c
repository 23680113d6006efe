/*
 * fgr_oracle.c -- CPU restatement of the stage BEFORE the refinement: registro_FGR (TEST INFRASTRUCTURE).
 *
 * SURVEY.md 8(f) N3.  The reference's front end (ALL_FUNCTIONS.py:178-203 == 1_FGR_pairwise_registration_in_NCLT_dataset.py:41-66)
 * is four Open3D calls:
 *     estimate_normals(KDTreeSearchParamHybrid(radius = 2 v, max_nn = 20))                 -> orc_estimate_normals_hybrid
 *     compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius = 10 v, max_nn = 200))      -> orc_compute_fpfh
 *     registration_fgr_based_on_feature_matching(source, target, fpfh_s, fpfh_t,
 *         FastGlobalRegistrationOption(division_factor 1.4, use_absolute_scale True, decrease_mu True,
 *         maximum_correspondence_distance 2 v, iteration_number 300, tuple_scale 0.95,
 *         maximum_tuple_count int(0.2 * (n_s + n_t) / 2)))                                 -> orc_fgr
 * Open3D (unpinned, most plausibly 0.17; not installable offline) is restated from its published algorithm:
 *   - KDTreeFlann::SearchHybrid: the max_nn nearest points (query included), cut at d^2 < radius^2;
 *   - EstimateNormals: covariance of the neighbour set (>= 3 points, else identity), FastEigen3x3 -- shared with the
 *     refinement oracle (mgicp_oracle.c);
 *   - Feature.cpp: ComputePairFeatures / ComputeSPFHFeature / ComputeFPFHFeature (Rusu's FPFH with 3 x 11 bins, SPFH of
 *     the neighbours weighted by 1 / d^2, each third renormalised to 100, own SPFH added);
 *   - FastGlobalRegistration.cpp (Zhou, Park, Koltun 2016): AdvancedMatching (nearest neighbours in feature space both
 *     ways, cross check, random tuple test), NormalizePointCloud, OptimizePairwiseRegistration (graduated non-convexity
 *     with the Geman-McClure line process, 6 x 6 solves), transformation mapped back to the original scale and inverted.
 * The tuple test draws random triples: Open3D's generator and seed are its own, so FGR results are comparable only
 * within FGR's own run-to-run scatter.  PARITY UNPINNED at any fixed tolerance; the soft pin is the reference's shipped
 * FGR poses (relative_poses_FGR/NCLT), see oracle/pin_fgr_against_goldens.py and tests/test_fgr_oracle.py.
 *
 * Nothing under the product package may call into this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_EINVAL 1
#define ORC_ENOMEM 2

/* shared with mgicp_oracle.c */
int orc_knn(const double *xyz, int64_t n, const double *queries, int64_t nq, int k, int32_t *idx_out, double *d2_out, int32_t *cnt_out);
void orc_fast_eigen3x3(const double cov[6], double out[3]);
void orc_ldlt_solve6(const double A_in[36], const double b_in[6], double x[6]);
void orc_vec6_to_mat4(const double x[6], double T[16]);
double orc_det_acos(double x);

/* ------------------------------------------------------------------------------------------ */
/* KDTreeFlann::SearchHybrid for every point of the cloud: counts + ascending (idx, d2) lists     */
/* ------------------------------------------------------------------------------------------ */
static int hybrid_lists(const double *xyz, int64_t n, double radius, int max_nn, int32_t **idx_out, double **d2_out, int32_t **cnt_out) {
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * (size_t)max_nn);
    double *d2 = (double *)malloc(sizeof(double) * (size_t)n * (size_t)max_nn);
    int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    if (!idx || !d2 || !cnt) { free(idx); free(d2); free(cnt); return ORC_ENOMEM; }
    int rc = orc_knn(xyz, n, xyz, n, max_nn, idx, d2, cnt);
    if (rc) { free(idx); free(d2); free(cnt); return rc; }
    const double r2 = radius * radius;
    for (int64_t i = 0; i < n; ++i) {
        int c = 0;                                  /* std::lower_bound(d2, d2 + k, r2): first entry >= r2 */
        while (c < cnt[i] && d2[(int64_t)max_nn * i + c] < r2) ++c;
        cnt[i] = c;
    }
    *idx_out = idx; *d2_out = d2; *cnt_out = cnt;
    return ORC_OK;
}

/* estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) on a cloud without normals */
int orc_estimate_normals_hybrid(const double *xyz, int64_t n, double radius, int max_nn, double *normals) {
    if (!(radius > 0.0) || max_nn < 1) return ORC_EINVAL;
    if (n == 0) return ORC_OK;
    int32_t *idx, *cnt; double *d2;
    int rc = hybrid_lists(xyz, n, radius, max_nn, &idx, &d2, &cnt);
    if (rc) return rc;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double cov[6] = {1, 0, 0, 1, 0, 1};         /* identity with fewer than 3 neighbours */
        const int c = cnt[i];
        if (c >= 3) {
            double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < c; ++j) {
                const double *p = xyz + 3 * (int64_t)idx[(int64_t)max_nn * i + j];
                cu[0] += p[0]; cu[1] += p[1]; cu[2] += p[2];
                cu[3] += p[0] * p[0]; cu[4] += p[0] * p[1]; cu[5] += p[0] * p[2];
                cu[6] += p[1] * p[1]; cu[7] += p[1] * p[2]; cu[8] += p[2] * p[2];
            }
            for (int j = 0; j < 9; ++j) cu[j] /= (double)c;
            cov[0] = cu[3] - cu[0] * cu[0]; cov[1] = cu[4] - cu[0] * cu[1]; cov[2] = cu[5] - cu[0] * cu[2];
            cov[3] = cu[6] - cu[1] * cu[1]; cov[4] = cu[7] - cu[1] * cu[2]; cov[5] = cu[8] - cu[2] * cu[2];
        }
        double nv[3];
        orc_fast_eigen3x3(cov, nv);
        if (sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]) == 0.0) { nv[0] = 0; nv[1] = 0; nv[2] = 1; }
        normals[3 * i] = nv[0]; normals[3 * i + 1] = nv[1]; normals[3 * i + 2] = nv[2];
    }
    free(idx); free(d2); free(cnt);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* FPFH                                                                                          */
/* ------------------------------------------------------------------------------------------ */
static void pair_features(const double *p1, const double *n1, const double *p2, const double *n2, double f[4]) {
    double dp[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    f[0] = f[1] = f[2] = f[3] = 0.0;
    const double len = sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
    if (len == 0.0) return;
    double a[3] = {n1[0], n1[1], n1[2]}, b[3] = {n2[0], n2[1], n2[2]};
    const double angle1 = (a[0] * dp[0] + a[1] * dp[1] + a[2] * dp[2]) / len;
    const double angle2 = (b[0] * dp[0] + b[1] * dp[1] + b[2] * dp[2]) / len;
    double f2;
    /* Open3D: acos(fabs(angle1)) > acos(fabs(angle2)) with its libm; here the deterministic fdlibm acos the GPU path uses
     * too (for nearly parallel normals the outcome hangs on the last bit of acos); |angle| > 1 is NaN there: no swap */
    const double c1 = fabs(angle1), c2 = fabs(angle2);
    if (c1 <= 1.0 && c2 <= 1.0 && orc_det_acos(c1) > orc_det_acos(c2)) {
        /* the normal with the smaller angle to the connecting line becomes the frame's first axis */
        for (int i = 0; i < 3; ++i) { const double t = a[i]; a[i] = b[i]; b[i] = t; dp[i] = -dp[i]; }
        f2 = -angle2;
    } else f2 = angle1;
    double v[3] = {dp[1] * a[2] - dp[2] * a[1], dp[2] * a[0] - dp[0] * a[2], dp[0] * a[1] - dp[1] * a[0]};
    const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (vn == 0.0) return;
    v[0] /= vn; v[1] /= vn; v[2] /= vn;
    const double w[3] = {a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0]};
    f[3] = len;
    f[2] = f2;
    f[1] = v[0] * b[0] + v[1] * b[1] + v[2] * b[2];
    f[0] = atan2(w[0] * b[0] + w[1] * b[1] + w[2] * b[2], a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
}

static inline int bin11(double x) {
    int h = (int)floor(x);
    if (h < 0) h = 0;
    if (h >= 11) h = 10;
    return h;
}

/* compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius, max_nn)): out is [n][33] (Open3D stores 33 x n column-major:
 * the same memory layout) */
int orc_compute_fpfh(const double *xyz, const double *nrm, int64_t n, double radius, int max_nn, double *out) {
    if (!(radius > 0.0) || max_nn < 1) return ORC_EINVAL;
    if (n == 0) return ORC_OK;
    int32_t *idx, *cnt; double *d2;
    int rc = hybrid_lists(xyz, n, radius, max_nn, &idx, &d2, &cnt);
    if (rc) return rc;
    double *spfh = (double *)calloc((size_t)n * 33, sizeof(double));
    if (!spfh) { free(idx); free(d2); free(cnt); return ORC_ENOMEM; }
    const double pi = 3.14159265358979323846;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        const int c = cnt[i];
        if (c <= 1) continue;                         /* only the point itself: no histogram */
        const double incr = 100.0 / (double)(c - 1);
        double *h = spfh + 33 * i;
        for (int k = 1; k < c; ++k) {                 /* entry 0 is the query point itself */
            const int64_t j = idx[(int64_t)max_nn * i + k];
            double f[4];
            pair_features(xyz + 3 * i, nrm + 3 * i, xyz + 3 * j, nrm + 3 * j, f);
            h[bin11(11.0 * (f[0] + pi) / (2.0 * pi))] += incr;
            h[11 + bin11(11.0 * (f[1] + 1.0) * 0.5)] += incr;
            h[22 + bin11(11.0 * (f[2] + 1.0) * 0.5)] += incr;
        }
    }
    memset(out, 0, sizeof(double) * (size_t)n * 33);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        const int c = cnt[i];
        if (c <= 1) continue;
        double sum[3] = {0, 0, 0};
        double *o = out + 33 * i;
        for (int k = 1; k < c; ++k) {
            const double dist = d2[(int64_t)max_nn * i + k];       /* squared distance, as in Open3D */
            if (dist == 0.0) continue;
            const double *s = spfh + 33 * (int64_t)idx[(int64_t)max_nn * i + k];
            for (int j = 0; j < 33; ++j) {
                const double val = s[j] / dist;
                sum[j / 11] += val;
                o[j] += val;
            }
        }
        for (int j = 0; j < 3; ++j) if (sum[j] != 0.0) sum[j] = 100.0 / sum[j];
        for (int j = 0; j < 33; ++j) { o[j] *= sum[j / 11]; o[j] += spfh[33 * i + j]; }
    }
    free(spfh); free(idx); free(d2); free(cnt);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* Fast Global Registration                                                                      */
/* ------------------------------------------------------------------------------------------ */
/* nearest neighbour of every row of A (na x 33) among the rows of B (nb x 33), squared L2, ties to the lower index
 * (FLANN's exact KD-tree search is replaced by brute force: same answer up to exact ties) */
static void feature_nn(const double *A, int64_t na, const double *B, int64_t nb, int32_t *out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < na; ++i) {
        const double *a = A + 33 * i;
        double best = INFINITY; int32_t bj = -1;
        for (int64_t j = 0; j < nb; ++j) {
            const double *b = B + 33 * j;
            double s = 0.0;
            for (int k = 0; k < 33; ++k) { const double d = a[k] - b[k]; s += d * d; }
            if (s < best) { best = s; bj = (int32_t)j; }
        }
        out[i] = bj;
    }
}

typedef struct { uint64_t s; } rng_t;
static inline uint32_t rng_next(rng_t *r) {          /* splitmix64, upper half */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

static inline double dist3(const double *a, const double *b) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrt(x * x + y * y + z * z);
}

typedef struct {
    double division_factor;                  /* 1.4 */
    int32_t use_absolute_scale;              /* reference: True */
    int32_t decrease_mu;                     /* reference: True */
    double maximum_correspondence_distance;  /* 2 v */
    int32_t iteration_number;                /* 300 */
    double tuple_scale;                      /* 0.95 */
    int32_t maximum_tuple_count;             /* int(0.2 * (n_s + n_t) / 2) */
    uint64_t seed;                           /* tuple-test generator (Open3D: its global engine) */
} orc_fgr_opts;

/* registration_fgr_based_on_feature_matching: T_out maps SOURCE into the TARGET frame (row-major 4 x 4);
 * n_corres_out = number of correspondences that entered the optimisation (3 per accepted tuple). */
int orc_fgr(const double *src_xyz, int64_t ns, const double *tgt_xyz, int64_t nt, const double *src_feat, const double *tgt_feat,
            const orc_fgr_opts *o, double T_out[16], int64_t *n_corres_out) {
    static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(T_out, I4, sizeof(I4));
    if (n_corres_out) *n_corres_out = 0;
    if (ns <= 0 || nt <= 0) return ORC_OK;
    /* point_cloud_vec = {source, target}: i = 0, j = 1 */
    const int64_t n[2] = {ns, nt};
    double *P[2];
    P[0] = (double *)malloc(sizeof(double) * 3 * (size_t)ns);
    P[1] = (double *)malloc(sizeof(double) * 3 * (size_t)nt);
    if (!P[0] || !P[1]) { free(P[0]); free(P[1]); return ORC_ENOMEM; }
    memcpy(P[0], src_xyz, sizeof(double) * 3 * (size_t)ns);
    memcpy(P[1], tgt_xyz, sizeof(double) * 3 * (size_t)nt);

    /* ---- NormalizePointCloud ---- */
    double mean[2][3], scale = 0.0;
    for (int c = 0; c < 2; ++c) {
        double m[3] = {0, 0, 0};
        for (int64_t i = 0; i < n[c]; ++i) { m[0] += P[c][3 * i]; m[1] += P[c][3 * i + 1]; m[2] += P[c][3 * i + 2]; }
        for (int k = 0; k < 3; ++k) { m[k] /= (double)n[c]; mean[c][k] = m[k]; }
        double mx = 0.0;
        for (int64_t i = 0; i < n[c]; ++i) {
            double *p = P[c] + 3 * i;
            p[0] -= m[0]; p[1] -= m[1]; p[2] -= m[2];
            const double t = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            if (t > mx) mx = t;
        }
        if (mx > scale) scale = mx;
    }
    double scale_global, scale_start;
    if (o->use_absolute_scale) { scale_global = 1.0; scale_start = scale; } else { scale_global = scale; scale_start = 1.0; }
    for (int c = 0; c < 2; ++c)
        for (int64_t i = 0; i < 3 * n[c]; ++i) P[c][i] /= scale_global;

    /* ---- AdvancedMatching ---- */
    int fi = 0, fj = 1, swapped = 0;
    if (n[fj] > n[fi]) { fi = 1; fj = 0; swapped = 1; }
    const double *F[2] = {src_feat, tgt_feat};
    const int64_t ni = n[fi], nj = n[fj];
    int32_t *j2i = (int32_t *)malloc(sizeof(int32_t) * (size_t)nj);     /* nearest i-feature of every j */
    int32_t *i2j = (int32_t *)malloc(sizeof(int32_t) * (size_t)ni);     /* nearest j-feature of every i */
    uint8_t *hit = (uint8_t *)calloc((size_t)ni, 1);
    if (!j2i || !i2j || !hit) { free(P[0]); free(P[1]); free(j2i); free(i2j); free(hit); return ORC_ENOMEM; }
    feature_nn(F[fj], nj, F[fi], ni, j2i);
    feature_nn(F[fi], ni, F[fj], nj, i2j);
    /* corres_ji = (nn_i(j), j) for every j; corres_ij = (i, nn_j(i)) for every i that is some j's nearest neighbour.  The
     * cross check (Mi from corres_ij, Mj from corres_ji) keeps (i, j) iff it occurs in both: mutual nearest neighbours in
     * feature space, emitted in ascending i. */
    for (int64_t j = 0; j < nj; ++j) hit[j2i[j]] = 1;
    int64_t ncross = 0;
    for (int64_t j = 0; j < nj; ++j) { const int32_t i = j2i[j]; if (i2j[i] == (int32_t)j) ++ncross; }
    int32_t *cross = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(ncross > 0 ? ncross : 1));
    if (!cross) { free(P[0]); free(P[1]); free(j2i); free(i2j); free(hit); return ORC_ENOMEM; }
    ncross = 0;
    for (int64_t i = 0; i < ni; ++i) {               /* ascending i, like the reference's outer loop */
        if (!hit[i]) continue;
        const int32_t j = i2j[i];
        if (j2i[j] == (int32_t)i) { cross[2 * ncross] = (int32_t)i; cross[2 * ncross + 1] = j; ++ncross; }
    }
    free(j2i); free(i2j); free(hit);
    /* tuple test */
    int64_t ntuple = 0;
    const int64_t cap = o->maximum_tuple_count > 0 ? o->maximum_tuple_count : 0;
    int32_t *cor = (int32_t *)malloc(sizeof(int32_t) * 2 * 3 * (size_t)(cap > 0 ? cap : 1));
    if (!cor) { free(P[0]); free(P[1]); free(cross); return ORC_ENOMEM; }
    if (ncross > 0 && cap > 0) {
        rng_t rng = {o->seed};
        const double sc = o->tuple_scale;
        const int64_t trials = ncross * 100;
        for (int64_t t = 0; t < trials; ++t) {
            const int64_t r0 = rng_next(&rng) % (uint64_t)ncross, r1 = rng_next(&rng) % (uint64_t)ncross, r2 = rng_next(&rng) % (uint64_t)ncross;
            const int32_t i0 = cross[2 * r0], j0 = cross[2 * r0 + 1], i1 = cross[2 * r1], j1 = cross[2 * r1 + 1], i2 = cross[2 * r2], j2 = cross[2 * r2 + 1];
            const double li0 = dist3(P[fi] + 3 * (int64_t)i0, P[fi] + 3 * (int64_t)i1), li1 = dist3(P[fi] + 3 * (int64_t)i1, P[fi] + 3 * (int64_t)i2),
                         li2 = dist3(P[fi] + 3 * (int64_t)i2, P[fi] + 3 * (int64_t)i0);
            const double lj0 = dist3(P[fj] + 3 * (int64_t)j0, P[fj] + 3 * (int64_t)j1), lj1 = dist3(P[fj] + 3 * (int64_t)j1, P[fj] + 3 * (int64_t)j2),
                         lj2 = dist3(P[fj] + 3 * (int64_t)j2, P[fj] + 3 * (int64_t)j0);
            if (li0 * sc < lj0 && lj0 < li0 / sc && li1 * sc < lj1 && lj1 < li1 / sc && li2 * sc < lj2 && lj2 < li2 / sc) {
                int32_t *c = cor + 6 * ntuple;
                c[0] = i0; c[1] = j0; c[2] = i1; c[3] = j1; c[4] = i2; c[5] = j2;
                ++ntuple;
            }
            if (ntuple >= cap) break;
        }
    }
    free(cross);
    const int64_t nc = 3 * ntuple;
    if (swapped)
        for (int64_t c = 0; c < nc; ++c) { const int32_t t = cor[2 * c]; cor[2 * c] = cor[2 * c + 1]; cor[2 * c + 1] = t; }
    if (n_corres_out) *n_corres_out = nc;

    /* ---- OptimizePairwiseRegistration: moves cloud 1 (target) onto cloud 0 (source) ---- */
    double trans[16];
    memcpy(trans, I4, sizeof(I4));
    if (nc >= 10) {
        double par = scale_start;
        double *Q = (double *)malloc(sizeof(double) * 3 * (size_t)nt);    /* point_cloud_copy_j */
        if (!Q) { free(P[0]); free(P[1]); free(cor); return ORC_ENOMEM; }
        memcpy(Q, P[1], sizeof(double) * 3 * (size_t)nt);
        for (int itr = 0; itr < o->iteration_number; ++itr) {
            double JTJ[36], JTr[6];
            memset(JTJ, 0, sizeof(JTJ)); memset(JTr, 0, sizeof(JTr));
            for (int64_t c = 0; c < nc; ++c) {
                const double *p = P[0] + 3 * (int64_t)cor[2 * c], *q = Q + 3 * (int64_t)cor[2 * c + 1];
                const double rpq[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
                const double temp = par / (rpq[0] * rpq[0] + rpq[1] * rpq[1] + rpq[2] * rpq[2] + par);
                const double s = temp * temp;
                const double Jr[3][6] = {{0, -q[2], q[1], -1, 0, 0}, {q[2], 0, -q[0], 0, -1, 0}, {-q[1], q[0], 0, 0, 0, -1}};
                for (int r = 0; r < 3; ++r) {
                    for (int a = 0; a < 6; ++a) {
                        for (int b = 0; b < 6; ++b) JTJ[6 * a + b] += Jr[r][a] * Jr[r][b] * s;
                        JTr[a] += Jr[r][a] * rpq[r] * s;
                    }
                }
            }
            /* SolveLinearSystemPSD(-JTJ, JTr) */
            double A[36], x[6], delta[16], tmp[16];
            for (int a = 0; a < 36; ++a) A[a] = -JTJ[a];
            orc_ldlt_solve6(A, JTr, x);
            orc_vec6_to_mat4(x, delta);
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) {
                    double sacc = 0;
                    for (int k = 0; k < 4; ++k) sacc += delta[4 * a + k] * trans[4 * k + b];
                    tmp[4 * a + b] = sacc;
                }
            memcpy(trans, tmp, sizeof(tmp));
            for (int64_t i = 0; i < nt; ++i) {
                double *q = Q + 3 * i;
                const double x0 = q[0], y0 = q[1], z0 = q[2];
                q[0] = delta[0] * x0 + delta[1] * y0 + delta[2] * z0 + delta[3];
                q[1] = delta[4] * x0 + delta[5] * y0 + delta[6] * z0 + delta[7];
                q[2] = delta[8] * x0 + delta[9] * y0 + delta[10] * z0 + delta[11];
            }
            if (o->decrease_mu && itr % 4 == 0 && par > o->maximum_correspondence_distance) par /= o->division_factor;
        }
        free(Q);
    }
    free(cor); free(P[0]); free(P[1]);

    /* ---- GetTransformationOriginalScale, then inverse (trans maps target -> source) ---- */
    double R[9], t[3];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) R[3 * a + b] = trans[4 * a + b];
        t[a] = trans[4 * a + 3];
    }
    double to[3];
    for (int a = 0; a < 3; ++a)
        to[a] = -(R[3 * a] * mean[1][0] + R[3 * a + 1] * mean[1][1] + R[3 * a + 2] * mean[1][2]) + t[a] * scale_global + mean[0][a];
    /* inverse of [R | to]: [R^T | -R^T to] */
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) T_out[4 * a + b] = R[3 * b + a];
        T_out[4 * a + 3] = -(R[a] * to[0] + R[3 + a] * to[1] + R[6 + a] * to[2]);
    }
    T_out[12] = 0; T_out[13] = 0; T_out[14] = 0; T_out[15] = 1;
    return ORC_OK;
}

/* registro_FGR(source, target, voxel_size) (ALL_FUNCTIONS.py:178-203): the three stages chained with the reference's
 * parameters; the clouds are taken as given (the scripts pass the pre-processed NCLT clouds). */
int orc_registro_fgr(const double *src_xyz, int64_t ns, const double *tgt_xyz, int64_t nt, double voxel_size, uint64_t seed,
                     double T_out[16], int64_t *n_corres_out) {
    if (!(voxel_size > 0.0)) return ORC_EINVAL;
    double *sn = (double *)malloc(sizeof(double) * 3 * (size_t)(ns > 0 ? ns : 1)), *tn = (double *)malloc(sizeof(double) * 3 * (size_t)(nt > 0 ? nt : 1));
    double *sf = (double *)malloc(sizeof(double) * 33 * (size_t)(ns > 0 ? ns : 1)), *tf = (double *)malloc(sizeof(double) * 33 * (size_t)(nt > 0 ? nt : 1));
    int rc = (!sn || !tn || !sf || !tf) ? ORC_ENOMEM : ORC_OK;
    if (!rc) rc = orc_estimate_normals_hybrid(src_xyz, ns, 2.0 * voxel_size, 20, sn);
    if (!rc) rc = orc_estimate_normals_hybrid(tgt_xyz, nt, 2.0 * voxel_size, 20, tn);
    if (!rc) rc = orc_compute_fpfh(src_xyz, sn, ns, 10.0 * voxel_size, 200, sf);
    if (!rc) rc = orc_compute_fpfh(tgt_xyz, tn, nt, 10.0 * voxel_size, 200, tf);
    if (!rc) {
        orc_fgr_opts o;
        o.division_factor = 1.4; o.use_absolute_scale = 1; o.decrease_mu = 1;
        o.maximum_correspondence_distance = 2.0 * voxel_size; o.iteration_number = 300; o.tuple_scale = 0.95;
        const int64_t n_pontos = (ns + nt) / 2;                       /* int((len(source.points) + len(target.points)) / 2) */
        o.maximum_tuple_count = (int32_t)((double)n_pontos * 0.2);    /* int(n_pontos * 0.2) */
        o.seed = seed;
        rc = orc_fgr(src_xyz, ns, tgt_xyz, nt, sf, tf, &o, T_out, n_corres_out);
    }
    free(sn); free(tn); free(sf); free(tf);
    return rc;
}
