/*
 * mgicp_oracle.c -- CPU ORACLE for the multiscale Generalized-ICP refinement path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path (the CUDA
 * library behind include/mgicp.h) never calls into this file.
 *
 * What it restates.  The reference's hot path is `Multiscale_GICP`
 *   /root/reference/ALL_FUNCTIONS.py:272-313  and
 *   /root/reference/2_MGICP_refinement_in_NCLT_dataset.py:128-164
 * whose arithmetic lives entirely inside Open3D (PyPI `open3d`, un-vendored, un-pinned in the
 * reference; API evidence says 0.13-0.17, most plausibly 0.17.0).  Open3D is not installable here,
 * so this file restates the published algorithms of the Open3D calls the reference makes, in
 * plain fp64 C, in the same operation order:
 *   voxel_down_sample            <- AF:293-294  (Open3D geometry/PointCloud.cpp VoxelDownSample)
 *   remove_statistical_outlier   <- AF:297-298  (RemoveStatisticalOutliers)
 *   estimate_normals(KNN)        <- AF:301-302  (EstimateNormals.cpp, FastEigen3x3)
 *   registration_generalized_icp <- AF:304-311  (GeneralizedICP.cpp, Registration.cpp, RobustKernel.cpp)
 * Parity pin: the reference ships no tests; the only golden vectors are the pose files under
 * relative_poses_FGR_GICP/NCLT, which this oracle reproduces at the millimetre level (see
 * oracle/pin_against_goldens.py and tests/golden/).  At the 1e-4 tolerance: PARITY UNPINNED.
 *
 * Determinism: every reduction runs in a fixed order that does not depend on the thread count.
 * The down-sampled cloud is emitted in canonical order (sorted by voxel index x, then y, then z)
 * because Open3D's own order is std::unordered_map iteration order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_EINVAL 1
#define ORC_ENOMEM 2

/* ------------------------------------------------------------------------------------------ */
/* KD-tree (exact kNN; plays the role of nanoflann in Open3D's KDTreeFlann, leaf size 15).     */
/* Result order: ascending (d2, index) -- nanoflann returns ascending d2; ties are arbitrary    */
/* there, here they are broken by the smaller index so the oracle is deterministic.             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t left, right; /* children, -1 for leaf */
    int32_t lo, hi;      /* [lo,hi) range in perm for leaves */
    int32_t dim;
    double split;
} kd_node;

typedef struct {
    const double *xyz; /* borrowed, n x 3 */
    int64_t n;
    int32_t *perm;
    kd_node *nodes;
    int32_t n_nodes, cap_nodes;
} kd_tree;

#define KD_LEAF 15

static void kd_swap(int32_t *a, int32_t *b) { int32_t t = *a; *a = *b; *b = t; }

/* quickselect on perm[lo,hi) so that element at position mid is in sorted place along dim */
static void kd_select(const double *xyz, int32_t *perm, int32_t lo, int32_t hi, int32_t mid, int dim) {
    while (hi - lo > 1) {
        /* median of three pivot */
        int32_t a = lo, b = lo + (hi - lo) / 2, c = hi - 1;
        double va = xyz[3 * (int64_t)perm[a] + dim], vb = xyz[3 * (int64_t)perm[b] + dim], vc = xyz[3 * (int64_t)perm[c] + dim];
        int32_t p = (va < vb) ? ((vb < vc) ? b : (va < vc ? c : a)) : ((va < vc) ? a : (vb < vc ? c : b));
        double pv = xyz[3 * (int64_t)perm[p] + dim];
        int32_t ppi = perm[p];
        kd_swap(&perm[p], &perm[hi - 1]);
        int32_t s = lo;
        for (int32_t i = lo; i < hi - 1; ++i) {
            double v = xyz[3 * (int64_t)perm[i] + dim];
            if (v < pv || (v == pv && perm[i] < ppi)) { kd_swap(&perm[i], &perm[s]); ++s; }
        }
        kd_swap(&perm[s], &perm[hi - 1]);
        if (s == mid) return;
        if (mid < s) hi = s; else lo = s + 1;
    }
}

static int32_t kd_build_rec(kd_tree *t, int32_t lo, int32_t hi) {
    int32_t id = t->n_nodes++;
    kd_node *nd = &t->nodes[id];
    nd->lo = lo; nd->hi = hi; nd->left = nd->right = -1; nd->dim = 0; nd->split = 0;
    if (hi - lo <= KD_LEAF) return id;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int32_t i = lo; i < hi; ++i)
        for (int d = 0; d < 3; ++d) {
            double v = t->xyz[3 * (int64_t)t->perm[i] + d];
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    int dim = 0;
    if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
    if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
    if (!(mx[dim] - mn[dim] > 0)) return id; /* all points identical: keep as (big) leaf */
    int32_t mid = lo + (hi - lo) / 2;
    kd_select(t->xyz, t->perm, lo, hi, mid, dim);
    double split = t->xyz[3 * (int64_t)t->perm[mid] + dim];
    int32_t l = kd_build_rec(t, lo, mid);
    int32_t r = kd_build_rec(t, mid, hi);
    nd = &t->nodes[id]; /* nodes array is preallocated, pointer stays valid; re-read for clarity */
    nd->dim = dim; nd->split = split; nd->left = l; nd->right = r;
    return id;
}

static int kd_build(kd_tree *t, const double *xyz, int64_t n) {
    memset(t, 0, sizeof(*t));
    t->xyz = xyz; t->n = n;
    if (n <= 0) return ORC_OK;
    t->perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    t->cap_nodes = (int32_t)(2 * n + 2);
    t->nodes = (kd_node *)malloc(sizeof(kd_node) * (size_t)t->cap_nodes);
    if (!t->perm || !t->nodes) return ORC_ENOMEM;
    for (int64_t i = 0; i < n; ++i) t->perm[i] = (int32_t)i;
    kd_build_rec(t, 0, (int32_t)n);
    return ORC_OK;
}

static void kd_free(kd_tree *t) { free(t->perm); free(t->nodes); memset(t, 0, sizeof(*t)); }

typedef struct {
    int k, cnt;
    double *d2;   /* ascending */
    int32_t *idx;
} kd_result;

static inline double kd_worst(const kd_result *r) { return r->cnt < r->k ? INFINITY : r->d2[r->k - 1]; }

static inline void kd_push(kd_result *r, double d2, int32_t idx) {
    /* insert keeping ascending (d2, idx) */
    int pos;
    if (r->cnt < r->k) pos = r->cnt++;
    else {
        int last = r->k - 1;
        if (d2 > r->d2[last] || (d2 == r->d2[last] && idx > r->idx[last])) return;
        pos = last;
    }
    while (pos > 0 && (r->d2[pos - 1] > d2 || (r->d2[pos - 1] == d2 && r->idx[pos - 1] > idx))) {
        r->d2[pos] = r->d2[pos - 1];
        r->idx[pos] = r->idx[pos - 1];
        --pos;
    }
    r->d2[pos] = d2; r->idx[pos] = idx;
}

static void kd_search_rec(const kd_tree *t, int32_t id, const double q[3], kd_result *r) {
    const kd_node *nd = &t->nodes[id];
    if (nd->left < 0) {
        for (int32_t i = nd->lo; i < nd->hi; ++i) {
            int32_t j = t->perm[i];
            const double *p = t->xyz + 3 * (int64_t)j;
            /* nanoflann L2 metric, dim 3: ((dx*dx) + dy*dy) + dz*dz with d = query - point */
            double dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
            double d2 = dx * dx + dy * dy + dz * dz;
            kd_push(r, d2, j);
        }
        return;
    }
    double diff = q[nd->dim] - nd->split;
    int32_t near = diff < 0 ? nd->left : nd->right;
    int32_t far = diff < 0 ? nd->right : nd->left;
    kd_search_rec(t, near, q, r);
    if (diff * diff <= kd_worst(r)) kd_search_rec(t, far, q, r);
}

/* returns number found (<= k) */
static int kd_knn(const kd_tree *t, const double q[3], int k, double *d2, int32_t *idx) {
    kd_result r = {k, 0, d2, idx};
    if (t->n > 0 && k > 0) kd_search_rec(t, 0, q, &r);
    return r.cnt;
}


/* ------------------------------------------------------------------------------------------ */
/* Deterministic sin / cos / acos (fdlibm kernels; <= 1 ulp from glibc, tests/test_oracle.py).   */
/* Open3D calls std::acos / std::cos / std::sin, whose last bit differs between libm builds; the  */
/* CUDA engine and this oracle evaluate the same fdlibm kernels so that this 1-ulp freedom does    */
/* not get amplified by the L1-IRLS iteration ("L1 chaos").                         */
/* ------------------------------------------------------------------------------------------ */
static double det_ksin(double x, double y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x, v = z * x;
    double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
static double det_kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double ax = fabs(x);
    if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
    double qx = ax > 0.78125 ? 0.28125 : (double)(float)(0.25 * ax);
    double hz = 0.5 * z - qx, a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}
static int det_reduce(double x, double *y0, double *y1) {
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    double n = rint(x * invpio2);
    double r0 = x - n * pio2_1, w = n * pio2_1t;
    *y0 = r0 - w;
    *y1 = (r0 - *y0) - w;
    return (int)n & 3;
}
double orc_det_sin(double x) {
    double y0, y1;
    switch (det_reduce(x, &y0, &y1)) {
        case 0: return det_ksin(y0, y1);
        case 1: return det_kcos(y0, y1);
        case 2: return -det_ksin(y0, y1);
        default: return -det_kcos(y0, y1);
    }
}
double orc_det_cos(double x) {
    double y0, y1;
    switch (det_reduce(x, &y0, &y1)) {
        case 0: return det_kcos(y0, y1);
        case 1: return -det_ksin(y0, y1);
        case 2: return -det_kcos(y0, y1);
        default: return det_ksin(y0, y1);
    }
}
static double det_acos_pq(double z) {
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
                 pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01,
                 qS4 = 7.70381505559019352791e-02;
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return p / q;
}
double orc_det_acos(double x) {
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17, pi = 3.14159265358979311600e+00;
    if (x >= 1.0) return 0.0;
    if (x <= -1.0) return pi + 2.0 * pio2_lo;
    if (fabs(x) < 0.5) {
        double r = det_acos_pq(x * x);
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (x < 0) {
        double z = (1.0 + x) * 0.5, s = sqrt(z), r = det_acos_pq(z);
        double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    double z = (1.0 - x) * 0.5, s = sqrt(z);
    double df = (double)(float)s;
    double c = (z - df * df) / (s + df);
    double r = det_acos_pq(z);
    double w = r * s + c;
    return 2.0 * (df + w);
}

/* opaque KD-tree handle for oracle/engine_order.cpp */
void *orc_kd_create(const double *xyz, int64_t n) {
    kd_tree *t = (kd_tree *)malloc(sizeof(kd_tree));
    if (!t) return NULL;
    if (kd_build(t, xyz, n)) { kd_free(t); free(t); return NULL; }
    return t;
}
void orc_kd_free(void *t) { if (t) { kd_free((kd_tree *)t); free(t); } }
/* nearest neighbour; returns index or -1 when the tree is empty */
int32_t orc_kd_nn(const void *t, const double q[3], double *d2_out) {
    double d2; int32_t j;
    int c = kd_knn((const kd_tree *)t, q, 1, &d2, &j);
    if (c <= 0) return -1;
    *d2_out = d2;
    return j;
}

/* ------------------------------------------------------------------------------------------ */
/* exported: exact kNN for tests (neighbour sets of the CUDA grid search are compared to this)   */
/* ------------------------------------------------------------------------------------------ */
int orc_knn(const double *xyz, int64_t n, const double *queries, int64_t nq, int k, int32_t *idx_out, double *d2_out,
            int32_t *cnt_out) {
    kd_tree t;
    int rc = kd_build(&t, xyz, n);
    if (rc) { kd_free(&t); return rc; }
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < nq; ++i) {
        int c = kd_knn(&t, queries + 3 * i, k, d2_out + (int64_t)k * i, idx_out + (int64_t)k * i);
        for (int j = c; j < k; ++j) { idx_out[(int64_t)k * i + j] = -1; d2_out[(int64_t)k * i + j] = INFINITY; }
        if (cnt_out) cnt_out[i] = c;
    }
    kd_free(&t);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* A.1 VoxelDownSample  (reference call sites AF:293-294, S2:146-147)                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t ix, iy, iz; int32_t src; } vox_rec;

static int vox_cmp(const void *a, const void *b) {
    const vox_rec *p = (const vox_rec *)a, *q = (const vox_rec *)b;
    if (p->ix != q->ix) return p->ix < q->ix ? -1 : 1;
    if (p->iy != q->iy) return p->iy < q->iy ? -1 : 1;
    if (p->iz != q->iz) return p->iz < q->iz ? -1 : 1;
    return p->src < q->src ? -1 : (p->src > q->src ? 1 : 0); /* keep input order inside a voxel */
}

/* out_xyz capacity n x 3; out_vox (optional) capacity n x 3 int32 voxel indices; returns m via *m_out */
int orc_voxel_down_sample(const double *xyz, int64_t n, double voxel, double *out_xyz, int32_t *out_vox, int64_t *m_out) {
    *m_out = 0;
    if (!(voxel > 0.0)) return ORC_EINVAL; /* Open3D: "voxel_size <= 0" */
    if (n == 0) return ORC_OK;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            double v = xyz[3 * i + d];
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    double org[3], ext = 0;
    for (int d = 0; d < 3; ++d) {
        org[d] = mn[d] - voxel * 0.5;                 /* voxel_min_bound = min_bound - voxel_size*0.5 */
        double e = (mx[d] + voxel * 0.5) - org[d];    /* voxel_max_bound - voxel_min_bound */
        if (e > ext) ext = e;
    }
    if (voxel * (double)INT32_MAX < ext) return ORC_EINVAL; /* "voxel_size is too small" */
    vox_rec *rec = (vox_rec *)malloc(sizeof(vox_rec) * (size_t)n);
    if (!rec) return ORC_ENOMEM;
    for (int64_t i = 0; i < n; ++i) {
        rec[i].ix = (int32_t)floor((xyz[3 * i + 0] - org[0]) / voxel);
        rec[i].iy = (int32_t)floor((xyz[3 * i + 1] - org[1]) / voxel);
        rec[i].iz = (int32_t)floor((xyz[3 * i + 2] - org[2]) / voxel);
        rec[i].src = (int32_t)i;
    }
    qsort(rec, (size_t)n, sizeof(vox_rec), vox_cmp);
    int64_t m = 0, i = 0;
    while (i < n) {
        int64_t j = i;
        double sx = 0, sy = 0, sz = 0;
        while (j < n && rec[j].ix == rec[i].ix && rec[j].iy == rec[i].iy && rec[j].iz == rec[i].iz) {
            const double *p = xyz + 3 * (int64_t)rec[j].src;
            sx += p[0]; sy += p[1]; sz += p[2];      /* AccumulatedPoint::AddPoint, input order */
            ++j;
        }
        double cnt = (double)(j - i);
        out_xyz[3 * m + 0] = sx / cnt; out_xyz[3 * m + 1] = sy / cnt; out_xyz[3 * m + 2] = sz / cnt;
        if (out_vox) { out_vox[3 * m + 0] = rec[i].ix; out_vox[3 * m + 1] = rec[i].iy; out_vox[3 * m + 2] = rec[i].iz; }
        ++m; i = j;
    }
    free(rec);
    *m_out = m;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* A.2 RemoveStatisticalOutliers(k, ratio)  (AF:297-298, constants AF:280-281)                   */
/* keep[i] in {0,1}; avg_out optional (mean neighbour distance, -1 if no neighbours)             */
/* ------------------------------------------------------------------------------------------ */
int orc_remove_statistical_outlier(const double *xyz, int64_t n, int k, double ratio, uint8_t *keep, double *avg_out,
                                   int64_t *kept_out, double *thresh_out) {
    *kept_out = 0;
    if (k < 1 || !(ratio > 0.0)) return ORC_EINVAL;
    if (n == 0) return ORC_OK;
    kd_tree t;
    int rc = kd_build(&t, xyz, n);
    if (rc) { kd_free(&t); return rc; }
    double *avg = (double *)malloc(sizeof(double) * (size_t)n);
    if (!avg) { kd_free(&t); return ORC_ENOMEM; }
#pragma omp parallel
    {
        double *d2 = (double *)malloc(sizeof(double) * (size_t)k);
        int32_t *ix = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            int c = kd_knn(&t, xyz + 3 * i, k, d2, ix);
            double mean = -1.0;
            if (c > 0) {
                double s = 0.0;
                for (int j = 0; j < c; ++j) s += sqrt(d2[j]); /* std::accumulate over sqrt'd dists, ascending */
                mean = s / (double)c;
            }
            avg[i] = mean;
        }
        free(d2); free(ix);
    }
    int64_t valid = n; /* every point finds at least itself */
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) if (avg[i] > 0) sum += avg[i];
    double cloud_mean = sum / (double)valid;
    double sq = 0.0;
    for (int64_t i = 0; i < n; ++i) if (avg[i] > 0) sq += (avg[i] - cloud_mean) * (avg[i] - cloud_mean);
    double std_dev = sqrt(sq / (double)(valid - 1));
    double thr = cloud_mean + ratio * std_dev;
    int64_t kept = 0;
    for (int64_t i = 0; i < n; ++i) {
        uint8_t kp = (avg[i] > 0 && avg[i] < thr) ? 1 : 0;
        keep[i] = kp; kept += kp;
        if (avg_out) avg_out[i] = avg[i];
    }
    if (thresh_out) *thresh_out = thr;
    *kept_out = kept;
    free(avg); kd_free(&t);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* A.3 FastEigen3x3 + EstimateNormals(KNN k)  (AF:301-302)                                        */
/* ------------------------------------------------------------------------------------------ */
static void cross3(const double a[3], const double b[3], double o[3]) {
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}

/* A symmetric stored as a00,a01,a02,a11,a12,a22 */
static void eigvec0(const double A[6], double ev, double out[3]) {
    double r0[3] = {A[0] - ev, A[1], A[2]};
    double r1[3] = {A[1], A[3] - ev, A[4]};
    double r2[3] = {A[2], A[4], A[5] - ev};
    double c01[3], c02[3], c12[3];
    cross3(r0, r1, c01); cross3(r0, r2, c02); cross3(r1, r2, c12);
    double d0 = c01[0] * c01[0] + c01[1] * c01[1] + c01[2] * c01[2];
    double d1 = c02[0] * c02[0] + c02[1] * c02[1] + c02[2] * c02[2];
    double d2 = c12[0] * c12[0] + c12[1] * c12[1] + c12[2] * c12[2];
    double dmax = d0; int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    const double *c = imax == 0 ? c01 : (imax == 1 ? c02 : c12);
    double dd = imax == 0 ? d0 : (imax == 1 ? d1 : d2);
    double s = sqrt(dd);
    out[0] = c[0] / s; out[1] = c[1] / s; out[2] = c[2] / s;
}

static void eigvec1(const double A[6], const double e0[3], double ev1, double out[3]) {
    double U[3], V[3];
    if (fabs(e0[0]) > fabs(e0[1])) {
        double inv = 1.0 / sqrt(e0[0] * e0[0] + e0[2] * e0[2]);
        U[0] = -e0[2] * inv; U[1] = 0; U[2] = e0[0] * inv;
    } else {
        double inv = 1.0 / sqrt(e0[1] * e0[1] + e0[2] * e0[2]);
        U[0] = 0; U[1] = e0[2] * inv; U[2] = -e0[1] * inv;
    }
    cross3(e0, U, V);
    double AU[3] = {A[0] * U[0] + A[1] * U[1] + A[2] * U[2], A[1] * U[0] + A[3] * U[1] + A[4] * U[2],
                    A[2] * U[0] + A[4] * U[1] + A[5] * U[2]};
    double AV[3] = {A[0] * V[0] + A[1] * V[1] + A[2] * V[2], A[1] * V[0] + A[3] * V[1] + A[4] * V[2],
                    A[2] * V[0] + A[4] * V[1] + A[5] * V[2]};
    double m00 = U[0] * AU[0] + U[1] * AU[1] + U[2] * AU[2] - ev1;
    double m01 = U[0] * AV[0] + U[1] * AV[1] + U[2] * AV[2];
    double m11 = V[0] * AV[0] + V[1] * AV[1] + V[2] * AV[2] - ev1;
    double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        double mx = a00 > a01 ? a00 : a01;
        if (mx > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m00; }
            else            { m00 /= m01; m01 = 1.0 / sqrt(1.0 + m00 * m00); m00 *= m01; }
            for (int d = 0; d < 3; ++d) out[d] = m01 * U[d] - m00 * V[d];
        } else { out[0] = U[0]; out[1] = U[1]; out[2] = U[2]; }
    } else {
        double mx = a11 > a01 ? a11 : a01;
        if (mx > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m11; }
            else            { m11 /= m01; m01 = 1.0 / sqrt(1.0 + m11 * m11); m11 *= m01; }
            for (int d = 0; d < 3; ++d) out[d] = m11 * U[d] - m01 * V[d];
        } else { out[0] = U[0]; out[1] = U[1]; out[2] = U[2]; }
    }
}

/* cov: c00,c01,c02,c11,c12,c22.  Returns the (unnormalised-by-us) normal as Open3D's FastEigen3x3 */
void orc_fast_eigen3x3(const double cov[6], double out[3]) {
    double A[6];
    double mc = cov[0];
    for (int i = 1; i < 6; ++i) if (cov[i] > mc) mc = cov[i]; /* Eigen maxCoeff (signed) */
    if (mc == 0) { out[0] = out[1] = out[2] = 0; return; }
    for (int i = 0; i < 6; ++i) A[i] = cov[i] / mc;
    double norm = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
    if (norm > 0) {
        double q = (A[0] + A[3] + A[5]) / 3.0;
        double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
        double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0);
        double c00 = b11 * b22 - A[4] * A[4];
        double c01 = A[1] * b22 - A[4] * A[2];
        double c02 = A[1] * A[4] - b11 * A[2];
        double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        double angle = orc_det_acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = orc_det_cos(angle) * 2.0;
        double beta0 = orc_det_cos(angle + two_thirds_pi) * 2.0;
        double beta1 = -(beta0 + beta2);
        double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        double v0[3], v1[3], v2[3];
        if (half_det >= 0) {
            eigvec0(A, e2, v2);
            if (e2 < e0 && e2 < e1) { out[0] = v2[0]; out[1] = v2[1]; out[2] = v2[2]; return; }
            eigvec1(A, v2, e1, v1);
            if (e1 < e0 && e1 < e2) { out[0] = v1[0]; out[1] = v1[1]; out[2] = v1[2]; return; }
            cross3(v1, v2, v0);
            out[0] = v0[0]; out[1] = v0[1]; out[2] = v0[2];
        } else {
            eigvec0(A, e0, v0);
            if (e0 < e1 && e0 < e2) { out[0] = v0[0]; out[1] = v0[1]; out[2] = v0[2]; return; }
            eigvec1(A, v0, e1, v1);
            if (e1 < e0 && e1 < e2) { out[0] = v1[0]; out[1] = v1[1]; out[2] = v1[2]; return; }
            cross3(v0, v1, v2);
            out[0] = v2[0]; out[1] = v2[1]; out[2] = v2[2];
        }
    } else {
        /* diagonal matrix (A*max_coeff restores cov; comparisons are scale-invariant for mc>0) */
        if (cov[0] < cov[3] && cov[0] < cov[5]) { out[0] = 1; out[1] = 0; out[2] = 0; }
        else if (cov[3] < cov[0] && cov[3] < cov[5]) { out[0] = 0; out[1] = 1; out[2] = 0; }
        else { out[0] = 0; out[1] = 0; out[2] = 1; }
    }
}

/* covariance of the neighbour set in returned (ascending-distance) order, cumulant form */
static void knn_covariance(const double *xyz, const int32_t *idx, int c, double cov[6]) {
    double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < c; ++j) {
        const double *p = xyz + 3 * (int64_t)idx[j];
        cu[0] += p[0]; cu[1] += p[1]; cu[2] += p[2];
        cu[3] += p[0] * p[0]; cu[4] += p[0] * p[1]; cu[5] += p[0] * p[2];
        cu[6] += p[1] * p[1]; cu[7] += p[1] * p[2]; cu[8] += p[2] * p[2];
    }
    for (int j = 0; j < 9; ++j) cu[j] /= (double)c;
    cov[0] = cu[3] - cu[0] * cu[0];
    cov[1] = cu[4] - cu[0] * cu[1];
    cov[2] = cu[5] - cu[0] * cu[2];
    cov[3] = cu[6] - cu[1] * cu[1];
    cov[4] = cu[7] - cu[1] * cu[2];
    cov[5] = cu[8] - cu[2] * cu[2];
}

int orc_estimate_normals(const double *xyz, int64_t n, int k, double *normals) {
    if (k < 1) return ORC_EINVAL;
    if (n == 0) return ORC_OK;
    kd_tree t;
    int rc = kd_build(&t, xyz, n);
    if (rc) { kd_free(&t); return rc; }
#pragma omp parallel
    {
        double *d2 = (double *)malloc(sizeof(double) * (size_t)k);
        int32_t *ix = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            int c = kd_knn(&t, xyz + 3 * i, k, d2, ix);
            double cov[6] = {1, 0, 0, 1, 0, 1}; /* Identity when < 3 neighbours */
            if (c >= 3) knn_covariance(xyz, ix, c, cov);
            double nv[3];
            orc_fast_eigen3x3(cov, nv);
            double nn = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
            if (nn == 0.0) { nv[0] = 0; nv[1] = 0; nv[2] = 1; }
            normals[3 * i + 0] = nv[0]; normals[3 * i + 1] = nv[1]; normals[3 * i + 2] = nv[2];
        }
        free(d2); free(ix);
    }
    kd_free(&t);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* A.4 covariances from normals (InitializePointCloudForGeneralizedICP, epsilon)                 */
/* ------------------------------------------------------------------------------------------ */
static void mat3_mul(const double A[9], const double B[9], double C[9]) {
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i + 0] * B[0 + j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, T, sizeof(T));
}
static void mat3_mul_bt(const double A[9], const double B[9], double C[9]) { /* A * B^T */
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i + 0] * B[3 * j + 0] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
    memcpy(C, T, sizeof(T));
}

void orc_covariance_from_normal(const double nrm[3], double eps, double C[9]) {
    /* GetRotationFromE1ToX: v = e1 x n, c = e1 . n ; if c < -0.99 return I */
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double c = nrm[0];
    if (!(c < -0.99)) {
        double v[3] = {0.0, -nrm[2], nrm[1]}; /* e1 x n = (0*nz - 0*ny, 0*nx - 1*nz, 1*ny - 0*nx) */
        double sv[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        double sv2[9];
        mat3_mul(sv, sv, sv2);
        double f = 1.0 / (1.0 + c);
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + sv[i] + sv2[i] * f;
    }
    double D[9] = {eps, 0, 0, 0, 1, 0, 0, 0, 1};
    double RD[9];
    mat3_mul(R, D, RD);
    mat3_mul_bt(RD, R, C);
}

/* ------------------------------------------------------------------------------------------ */
/* A.7 helpers: 3x3 inverse (adjugate/det like Eigen's fixed-size inverse), principal sqrt of a  */
/* symmetric PD 3x3 via cyclic Jacobi, 6x6 LDL^T with diagonal pivoting (Eigen ldlt()).          */
/* ------------------------------------------------------------------------------------------ */
static void mat3_inverse(const double M[9], double I[9]) {
    double c00 = M[4] * M[8] - M[5] * M[7];
    double c01 = M[5] * M[6] - M[3] * M[8];
    double c02 = M[3] * M[7] - M[4] * M[6];
    double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
    double id = 1.0 / det;
    I[0] = c00 * id; I[1] = (M[2] * M[7] - M[1] * M[8]) * id; I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    I[3] = c01 * id; I[4] = (M[0] * M[8] - M[2] * M[6]) * id; I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    I[6] = c02 * id; I[7] = (M[1] * M[6] - M[0] * M[7]) * id; I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

static void sym3_sqrt(const double S_in[9], double out[9]) {
    double S[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S[3 * i + j] = 0.5 * (S_in[3 * i + j] + S_in[3 * j + i]);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = S[1] * S[1] + S[2] * S[2] + S[5] * S[5];
        double dia = S[0] * S[0] + S[4] * S[4] + S[8] * S[8];
        if (off <= 1e-34 * dia || off == 0.0) break;
        static const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
        for (int r = 0; r < 3; ++r) {
            int p = P[r], q = Q[r];
            double apq = S[3 * p + q];
            if (apq == 0.0) continue;
            double theta = (S[3 * q + q] - S[3 * p + p]) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) { /* S <- S * G */
                double skp = S[3 * k + p], skq = S[3 * k + q];
                S[3 * k + p] = c * skp - s * skq;
                S[3 * k + q] = s * skp + c * skq;
            }
            for (int k = 0; k < 3; ++k) { /* S <- G^T * S */
                double spk = S[3 * p + k], sqk = S[3 * q + k];
                S[3 * p + k] = c * spk - s * sqk;
                S[3 * q + k] = s * spk + c * sqk;
            }
            for (int k = 0; k < 3; ++k) {
                double vkp = V[3 * k + p], vkq = V[3 * k + q];
                V[3 * k + p] = c * vkp - s * vkq;
                V[3 * k + q] = s * vkp + c * vkq;
            }
        }
    }
    double l[3] = {sqrt(S[0]), sqrt(S[4]), sqrt(S[8])};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            out[3 * i + j] = V[3 * i + 0] * l[0] * V[3 * j + 0] + V[3 * i + 1] * l[1] * V[3 * j + 1] + V[3 * i + 2] * l[2] * V[3 * j + 2];
}

/* solve A x = b, A symmetric 6x6 (full storage), LDL^T with symmetric diagonal pivoting.
 * No definiteness checks (Open3D SolveLinearSystemPSD defaults).  Returns x (may be non-finite). */
void orc_ldlt_solve6(const double A_in[36], const double b_in[6], double x[6]) {
    double A[36]; int perm[6]; double b[6];
    memcpy(A, A_in, sizeof(A));
    for (int i = 0; i < 6; ++i) perm[i] = i;
    for (int k = 0; k < 6; ++k) {
        int piv = k; double best = fabs(A[7 * k]);
        for (int i = k + 1; i < 6; ++i) if (fabs(A[7 * i]) > best) { best = fabs(A[7 * i]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < 6; ++j) { double t = A[6 * k + j]; A[6 * k + j] = A[6 * piv + j]; A[6 * piv + j] = t; }
            for (int j = 0; j < 6; ++j) { double t = A[6 * j + k]; A[6 * j + k] = A[6 * j + piv]; A[6 * j + piv] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        double d = A[7 * k];
        double col[6];
        for (int i = k + 1; i < 6; ++i) col[i] = A[6 * i + k];
        for (int i = k + 1; i < 6; ++i)
            for (int j = k + 1; j <= i; ++j) A[6 * i + j] -= col[i] * (col[j] / d);
        for (int i = k + 1; i < 6; ++i) A[6 * i + k] = col[i] / d;
        /* keep the trailing block symmetric in full storage */
        for (int i = k + 1; i < 6; ++i)
            for (int j = i + 1; j < 6; ++j) A[6 * i + j] = A[6 * j + i];
    }
    for (int i = 0; i < 6; ++i) b[i] = b_in[perm[i]];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) b[i] -= A[6 * i + j] * b[j];
    for (int i = 0; i < 6; ++i) b[i] /= A[7 * i];
    for (int i = 5; i >= 0; --i) for (int j = i + 1; j < 6; ++j) b[i] -= A[6 * j + i] * b[j];
    for (int i = 0; i < 6; ++i) x[perm[i]] = b[i];
}

static void mat4_mul(const double A[16], const double B[16], double C[16]) {
    double T[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            T[4 * i + j] = s;
        }
    memcpy(C, T, sizeof(T));
}

/* TransformVector6dToMatrix4d: R = Rz(x2) * Ry(x1) * Rx(x0), t = x3..5 */
void orc_vec6_to_mat4(const double x[6], double T[16]) {
    double ca = orc_det_cos(x[0]), sa = orc_det_sin(x[0]), cb = orc_det_cos(x[1]), sb = orc_det_sin(x[1]), cg = orc_det_cos(x[2]), sg = orc_det_sin(x[2]);
    double Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca};
    double Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb};
    double Rz[9] = {cg, -sg, 0, sg, cg, 0, 0, 0, 1};
    double RzRy[9], R[9];
    mat3_mul(Rz, Ry, RzRy);
    mat3_mul(RzRy, Rx, R);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = x[3 + i];
    }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

/* robust kernel weights (RobustKernel.cpp).  loss: 0 L2, 1 L1, 2 Huber, 3 Cauchy, 4 GM, 5 Tukey */
static inline double kernel_weight(int loss, double k, double r) {
    switch (loss) {
        case 0: return 1.0;
        case 1: return 1.0 / fabs(r);
        case 2: { double e = fabs(r); return k / (e > k ? e : k); }
        case 3: return 1.0 / (1.0 + (r / k) * (r / k));
        case 4: return k / ((k + r * r) * (k + r * r));
        case 5: { double e = fabs(r) / k; if (e > 1.0) e = 1.0; double q = 1.0 - e * e; return q * q; }
        default: return 1.0;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* A.5-A.7 RegistrationGeneralizedICP  (AF:304-311)                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double epsilon;       /* GICP plane regularisation, Open3D default 1e-3 */
    int loss;             /* see kernel_weight; the reference uses L1 (AF:284) */
    double loss_k;        /* scale parameter for Huber/Cauchy/GM/Tukey */
    double rel_fitness;   /* ICPConvergenceCriteria.relative_fitness (AF:309) -- an ABSOLUTE difference in Open3D */
    double rel_rmse;      /* ICPConvergenceCriteria.relative_rmse (AF:310) */
    int max_iteration;
} orc_gicp_opts;

static int64_t g_chunk = 1024; /* summation chunk of the normal equations (fixed => thread-count independent) */
void orc_set_sum_chunk(int64_t c) { if (c > 0) g_chunk = c; }
#define ORC_CHUNK g_chunk

static void transform_points(double *p, double *C, int64_t n, const double T[16]) {
    double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
        double nx = T[0] * x + T[1] * y + T[2] * z + T[3] * 1.0;
        double ny = T[4] * x + T[5] * y + T[6] * z + T[7] * 1.0;
        double nz = T[8] * x + T[9] * y + T[10] * z + T[11] * 1.0;
        double nw = T[12] * x + T[13] * y + T[14] * z + T[15] * 1.0;
        p[3 * i] = nx / nw; p[3 * i + 1] = ny / nw; p[3 * i + 2] = nz / nw;
        double RC[9];
        mat3_mul(R, C + 9 * i, RC);
        mat3_mul_bt(RC, R, C + 9 * i);
    }
}

static void correspond(const kd_tree *tt, const double *p, int64_t ns, double max_d, int32_t *corr, double *cd2,
                       int64_t *K_out, double *fitness, double *rmse) {
    double r2 = max_d * max_d;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < ns; ++i) {
        double d2; int32_t j;
        int c = kd_knn(tt, p + 3 * i, 1, &d2, &j);
        if (c > 0 && d2 < r2) { corr[i] = j; cd2[i] = d2; } /* lower_bound on r^2 -> strict */
        else { corr[i] = -1; cd2[i] = 0.0; }
    }
    int64_t K = 0; double e2 = 0.0;
    for (int64_t i = 0; i < ns; ++i) if (corr[i] >= 0) { ++K; e2 += cd2[i]; }
    *K_out = K;
    if (K == 0) { *fitness = 0.0; *rmse = 0.0; }
    else { *fitness = (double)K / (double)ns; *rmse = sqrt(e2 / (double)K); }
}

/* trace (optional): per pass i (0..iters): fitness, rmse, K  -> 3 doubles per pass, capacity max_iteration+1 */
int orc_gicp(const double *src_xyz, const double *src_nrm, int64_t ns, const double *tgt_xyz, const double *tgt_nrm,
             int64_t nt, double max_d, const double T_init[16], const orc_gicp_opts *o, double T_out[16],
             double *fitness_out, double *rmse_out, int32_t *iters_out, int64_t *ncorr_out, double *trace,
             double *sys_trace /* optional 27 doubles per iteration: 21 upper JTJ + 6 JTr */) {
    if (!(max_d > 0.0)) return ORC_EINVAL; /* Open3D: "Invalid max_correspondence_distance" */
    memcpy(T_out, T_init, sizeof(double) * 16);
    *fitness_out = 0; *rmse_out = 0; *iters_out = 0; *ncorr_out = 0;
    if (ns == 0 || nt == 0) return ORC_OK;
    double *p = (double *)malloc(sizeof(double) * 3 * (size_t)ns);
    double *Cs = (double *)malloc(sizeof(double) * 9 * (size_t)ns);
    double *Ct = (double *)malloc(sizeof(double) * 9 * (size_t)nt);
    int32_t *corr = (int32_t *)malloc(sizeof(int32_t) * (size_t)ns);
    double *cd2 = (double *)malloc(sizeof(double) * (size_t)ns);
    int64_t nchunk = (ns + ORC_CHUNK - 1) / ORC_CHUNK;
    double *part = (double *)malloc(sizeof(double) * 27 * (size_t)nchunk);
    if (!p || !Cs || !Ct || !corr || !cd2 || !part) return ORC_ENOMEM;
    memcpy(p, src_xyz, sizeof(double) * 3 * (size_t)ns);
    for (int64_t i = 0; i < ns; ++i) orc_covariance_from_normal(src_nrm + 3 * i, o->epsilon, Cs + 9 * i);
    for (int64_t i = 0; i < nt; ++i) orc_covariance_from_normal(tgt_nrm + 3 * i, o->epsilon, Ct + 9 * i);
    kd_tree tt;
    int rc = kd_build(&tt, tgt_xyz, nt);
    if (rc) return rc;
    int is_identity = 1;
    for (int i = 0; i < 16; ++i) if (T_init[i] != ((i % 5 == 0) ? 1.0 : 0.0)) is_identity = 0;
    if (!is_identity) transform_points(p, Cs, ns, T_init);
    int64_t K; double fit, rmse;
    correspond(&tt, p, ns, max_d, corr, cd2, &K, &fit, &rmse);
    if (trace) { trace[0] = fit; trace[1] = rmse; trace[2] = (double)K; }
    int it = 0;
    for (it = 0; it < o->max_iteration; ++it) {
        double U[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        if (K > 0) {
            /* ComputeTransformation: JTJ = sum w J^T J, JTr = sum w J^T r over 3 rows per correspondence.
             * Fixed chunking over source index => independent of thread count. */
#pragma omp parallel for schedule(dynamic, 1)
            for (int64_t c = 0; c < nchunk; ++c) {
                double acc[27];
                for (int a = 0; a < 27; ++a) acc[a] = 0.0;
                int64_t lo = c * ORC_CHUNK, hi = lo + ORC_CHUNK < ns ? lo + ORC_CHUNK : ns;
                for (int64_t i = lo; i < hi; ++i) {
                    int32_t j = corr[i];
                    if (j < 0) continue;
                    const double *vs = p + 3 * i, *vt = tgt_xyz + 3 * (int64_t)j;
                    double d[3] = {vs[0] - vt[0], vs[1] - vt[1], vs[2] - vt[2]};
                    double M[9], Mi[9], W[9];
                    for (int a = 0; a < 9; ++a) M[a] = Ct[9 * (int64_t)j + a] + Cs[9 * i + a];
                    mat3_inverse(M, Mi);
                    sym3_sqrt(Mi, W);
                    /* J = W * [ -skew(vs) | I ] ; skew(v) = [[0,-z,y],[z,0,-x],[-y,x,0]] */
                    double S[9] = {0, vs[2], -vs[1], -vs[2], 0, vs[0], vs[1], -vs[0], 0}; /* = -skew(vs) */
                    double WS[9];
                    mat3_mul(W, S, WS);
                    for (int row = 0; row < 3; ++row) {
                        double Jr[6] = {WS[3 * row], WS[3 * row + 1], WS[3 * row + 2], W[3 * row], W[3 * row + 1], W[3 * row + 2]};
                        double r = W[3 * row] * d[0] + W[3 * row + 1] * d[1] + W[3 * row + 2] * d[2];
                        double w = kernel_weight(o->loss, o->loss_k, r);
                        int a = 0;
                        for (int u = 0; u < 6; ++u)
                            for (int v = u; v < 6; ++v) acc[a++] += Jr[u] * w * Jr[v];
                        for (int u = 0; u < 6; ++u) acc[21 + u] += Jr[u] * w * r;
                    }
                }
                memcpy(part + 27 * c, acc, sizeof(acc));
            }
            double tot[27];
            for (int a = 0; a < 27; ++a) tot[a] = 0.0;
            for (int64_t c = 0; c < nchunk; ++c)
                for (int a = 0; a < 27; ++a) tot[a] += part[27 * c + a];
            if (sys_trace) memcpy(sys_trace + 27 * it, tot, sizeof(tot));
            double A[36], b[6], x[6];
            int a = 0;
            for (int u = 0; u < 6; ++u)
                for (int v = u; v < 6; ++v) { A[6 * u + v] = tot[a]; A[6 * v + u] = tot[a]; ++a; }
            for (int u = 0; u < 6; ++u) b[u] = -tot[21 + u];
            orc_ldlt_solve6(A, b, x);
            orc_vec6_to_mat4(x, U);
        } else if (sys_trace) {
            for (int a = 0; a < 27; ++a) sys_trace[27 * it + a] = 0.0;
        }
        mat4_mul(U, T_out, T_out);
        transform_points(p, Cs, ns, U);
        double bf = fit, br = rmse;
        correspond(&tt, p, ns, max_d, corr, cd2, &K, &fit, &rmse);
        if (trace) { trace[3 * (it + 1)] = fit; trace[3 * (it + 1) + 1] = rmse; trace[3 * (it + 1) + 2] = (double)K; }
        if (fabs(bf - fit) < o->rel_fitness && fabs(br - rmse) < o->rel_rmse) { ++it; break; }
    }
    *fitness_out = fit; *rmse_out = rmse; *iters_out = it; *ncorr_out = K;
    free(p); free(Cs); free(Ct); free(corr); free(cd2); free(part); kd_free(&tt);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* Multiscale_GICP body (AF:286-312 / S2:140-163) with explicit schedules.                       */
/* stats_out (optional): per scale 8 doubles: M_src, M_tgt, M'_src, M'_tgt, iters, K, fitness, rmse */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int sor_k;        /* knn_filtro = 30 (AF:280) */
    double sor_std;   /* std_filtro = 1.0 (AF:281) */
    int normal_k;     /* KDTreeSearchParamKNN(knn=20) (AF:301) */
    orc_gicp_opts gicp;
} orc_opts;

/* wall-clock split of orc_multiscale_gicp, accumulated since the last orc_stage_reset():
 * [0] voxel_down_sample, [1] remove_statistical_outlier, [2] estimate_normals, [3] registration_generalized_icp, then the
 * ICP seconds and the ICP passes (iterations + 1) per scale index 0..7.  Calls are made from one host thread. */
static double g_stage_s[4 + 16];
void orc_stage_reset(void) { memset(g_stage_s, 0, sizeof(g_stage_s)); }
void orc_stage_seconds(double *out /* [20] */) { memcpy(out, g_stage_s, sizeof(g_stage_s)); }

static int preprocess(const double *xyz, int64_t n, double voxel, const orc_opts *o, double **pts_out, double **nrm_out,
                      int64_t *m_ds, int64_t *m_out) {
    double *ds = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    if (!ds) return ORC_ENOMEM;
    int64_t m = 0;
    double t0 = omp_get_wtime();
    int rc = orc_voxel_down_sample(xyz, n, voxel, ds, NULL, &m);
    g_stage_s[0] += omp_get_wtime() - t0;
    if (rc) { free(ds); return rc; }
    uint8_t *keep = (uint8_t *)malloc((size_t)(m > 0 ? m : 1));
    int64_t kept = 0;
    t0 = omp_get_wtime();
    rc = orc_remove_statistical_outlier(ds, m, o->sor_k, o->sor_std, keep, NULL, &kept, NULL);
    g_stage_s[1] += omp_get_wtime() - t0;
    if (rc) { free(ds); free(keep); return rc; }
    double *pts = (double *)malloc(sizeof(double) * 3 * (size_t)(kept > 0 ? kept : 1));
    double *nrm = (double *)malloc(sizeof(double) * 3 * (size_t)(kept > 0 ? kept : 1));
    int64_t w = 0;
    for (int64_t i = 0; i < m; ++i)
        if (keep[i]) { pts[3 * w] = ds[3 * i]; pts[3 * w + 1] = ds[3 * i + 1]; pts[3 * w + 2] = ds[3 * i + 2]; ++w; }
    t0 = omp_get_wtime();
    rc = orc_estimate_normals(pts, kept, o->normal_k, nrm);
    g_stage_s[2] += omp_get_wtime() - t0;
    free(ds); free(keep);
    if (rc) { free(pts); free(nrm); return rc; }
    *pts_out = pts; *nrm_out = nrm; *m_ds = m; *m_out = kept;
    return ORC_OK;
}

int orc_multiscale_gicp(const double *src_xyz, int64_t ns, const double *tgt_xyz, int64_t nt, int n_scales,
                        const double *voxel_sizes, const double *max_dists, const int32_t *max_iters,
                        const double T_init[16], const orc_opts *o, double T_out[16], double *fitness_out,
                        double *rmse_out, int32_t *iters_out /* n_scales */, int64_t *ncorr_out, double *stats_out) {
    double T[16];
    memcpy(T, T_init, sizeof(T));
    *fitness_out = 0; *rmse_out = 0; *ncorr_out = 0;
    for (int s = 0; s < n_scales; ++s) {
        double *sp = NULL, *sn = NULL, *tp = NULL, *tn = NULL;
        int64_t ms_ds = 0, ms = 0, mt_ds = 0, mt = 0;
        int rc = preprocess(src_xyz, ns, voxel_sizes[s], o, &sp, &sn, &ms_ds, &ms);
        if (rc) return rc;
        rc = preprocess(tgt_xyz, nt, voxel_sizes[s], o, &tp, &tn, &mt_ds, &mt);
        if (rc) { free(sp); free(sn); return rc; }
        orc_gicp_opts g = o->gicp;
        g.max_iteration = max_iters[s];
        double Tn[16]; int32_t it = 0; int64_t K = 0; double fit = 0, rmse = 0;
        const double t0 = omp_get_wtime();
        rc = orc_gicp(sp, sn, ms, tp, tn, mt, max_dists[s], T, &g, Tn, &fit, &rmse, &it, &K, NULL, NULL);
        const double dt = omp_get_wtime() - t0;
        g_stage_s[3] += dt;
        if (s < 8) { g_stage_s[4 + s] += dt; g_stage_s[12 + s] += (double)(it + 1); }
        free(sp); free(sn); free(tp); free(tn);
        if (rc) return rc;
        memcpy(T, Tn, sizeof(T));
        iters_out[s] = it;
        *fitness_out = fit; *rmse_out = rmse; *ncorr_out = K;
        if (stats_out) {
            double *st = stats_out + 8 * s;
            st[0] = (double)ms_ds; st[1] = (double)mt_ds; st[2] = (double)ms; st[3] = (double)mt;
            st[4] = (double)it; st[5] = (double)K; st[6] = fit; st[7] = rmse;
        }
    }
    memcpy(T_out, T, sizeof(T));
    return ORC_OK;
}

/* ---- evaluate_registration / get_information_matrix_from_point_clouds (App. A.9) ----
 * o3d.pipelines.registration.evaluate_registration(source, target, max_d, T)        ALL_FUNCTIONS.py:809-822
 * o3d.pipelines.registration.get_information_matrix_from_point_clouds(source, target, max_d, T)
 *                                                   ALL_FUNCTIONS.py:327-331, 3_Global_Refinement...py:317-320
 * Both work on the clouds AS GIVEN (no down-sampling): pcd = source; if (!T.isIdentity()) pcd.Transform(T);
 * GetRegistrationResultAndCorrespondences(pcd, target, kdtree(target), max_d).  The information matrix is
 * GTG = sum over correspondences of the three rows G_r G_r^T built from the TARGET point (x, y, z):
 *   (0, z, -y, 1, 0, 0), (-z, 0, x, 0, 1, 0), (y, -x, 0, 0, 0, 1).
 * corr_out (optional): target index per source point, -1 = none.  gtg_out (optional): 6x6 row-major. */
int orc_evaluate_registration(const double *src_xyz, int64_t ns, const double *tgt_xyz, int64_t nt, double max_d,
                              const double T[16], double *fitness_out, double *rmse_out, int64_t *ncorr_out,
                              int32_t *corr_out, double *gtg_out) {
    if (!(max_d > 0.0)) return ORC_EINVAL;
    *fitness_out = 0; *rmse_out = 0; *ncorr_out = 0;
    if (gtg_out) memset(gtg_out, 0, sizeof(double) * 36);
    if (corr_out) for (int64_t i = 0; i < ns; ++i) corr_out[i] = -1;
    if (ns == 0 || nt == 0) return ORC_OK;
    double *p = (double *)malloc(sizeof(double) * 3 * (size_t)ns);
    int32_t *corr = (int32_t *)malloc(sizeof(int32_t) * (size_t)ns);
    double *cd2 = (double *)malloc(sizeof(double) * (size_t)ns);
    if (!p || !corr || !cd2) return ORC_ENOMEM;
    memcpy(p, src_xyz, sizeof(double) * 3 * (size_t)ns);
    int is_identity = 1;
    for (int i = 0; i < 16; ++i) if (T[i] != ((i % 5 == 0) ? 1.0 : 0.0)) is_identity = 0;
    if (!is_identity) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < ns; ++i) {
            double x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
            double nx = T[0] * x + T[1] * y + T[2] * z + T[3] * 1.0;
            double ny = T[4] * x + T[5] * y + T[6] * z + T[7] * 1.0;
            double nz = T[8] * x + T[9] * y + T[10] * z + T[11] * 1.0;
            double nw = T[12] * x + T[13] * y + T[14] * z + T[15] * 1.0;
            p[3 * i] = nx / nw; p[3 * i + 1] = ny / nw; p[3 * i + 2] = nz / nw;
        }
    }
    kd_tree tt;
    int rc = kd_build(&tt, tgt_xyz, nt);
    if (rc) { free(p); free(corr); free(cd2); return rc; }
    int64_t K; double fit, rmse;
    correspond(&tt, p, ns, max_d, corr, cd2, &K, &fit, &rmse);
    *fitness_out = fit; *rmse_out = rmse; *ncorr_out = K;
    if (corr_out) memcpy(corr_out, corr, sizeof(int32_t) * (size_t)ns);
    if (gtg_out) {
        for (int64_t i = 0; i < ns; ++i) {
            if (corr[i] < 0) continue;
            const double *q = tgt_xyz + 3 * (int64_t)corr[i];
            const double x = q[0], y = q[1], z = q[2];
            const double G[3][6] = {{0.0, z, -y, 1.0, 0.0, 0.0}, {-z, 0.0, x, 0.0, 1.0, 0.0}, {y, -x, 0.0, 0.0, 0.0, 1.0}};
            for (int r = 0; r < 3; ++r)
                for (int a = 0; a < 6; ++a)
                    for (int b = 0; b < 6; ++b) gtg_out[6 * a + b] += G[r][a] * G[r][b];
        }
    }
    kd_free(&tt);
    free(p); free(corr); free(cd2);
    return ORC_OK;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
