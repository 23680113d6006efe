/*
 * engine_order.cpp -- CPU emulation of the CUDA ICP kernel's ARITHMETIC ORDER (TEST INFRASTRUCTURE).
 *
 * Why it exists.  With the reference's L1 robust kernel (ALL_FUNCTIONS.py:284) the IRLS iteration of
 * registration_generalized_icp is chaotic: re-associating the 27 normal-equation sums (which Open3D
 * itself does from run to run, its OpenMP reduction order is unspecified) moves the final pose by
 * 1e-5..1e-3 m (measured on the oracle alone, tests/test_oracle.py::test_l1_self_sensitivity).  A
 * fixed 1e-4 tolerance against an independently ordered implementation is therefore not a
 * meaningful pass/fail signal for the L1 loop.  This file evaluates the SAME per-correspondence
 * closed forms (csrc/mgicp_math.cuh, compiled for the host without FMA contraction) in the SAME
 * reduction tree as the kernel k_icp (512 threads per block, a gang of CL blocks per pair: thread-strided
 * partial sums, per accumulator lane l adds threads l, l + 32, ... in order, shuffle-down tree over the lanes, gang ranks
 * in order), so the CUDA loop can be
 * checked bit for bit.  Nearest neighbours come from the oracle's KD-tree (mgicp_oracle.c), i.e. the
 * search structure stays independent of the GPU's spatial hash.
 *
 * The reference-faithful restatement (full 3x3 covariances, M.inverse().sqrt(), sequential sums) stays in
 * mgicp_oracle.c; the two agree to 1e-15 under the contractive L2 kernel and to 1e-10 on single passes.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../point-cloud-registration-with-global-refinement_b200/csrc/mgicp_math.cuh"

extern "C" {
void *orc_kd_create(const double *xyz, int64_t n);
void orc_kd_free(void *t);
int32_t orc_kd_nn(const void *t, const double q[3], double *d2_out);
}

using namespace mg;

namespace {
constexpr int NT = 512, NACC = 29;

// the reduction tree of pair_reduce<CL> in mgicp.cu
void reduce_like_kernel(const std::vector<double> &acc /* [nthr][NACC] */, int cl, double tot[NACC]) {
    const int nthr = cl * NT;
    std::vector<double> part((size_t)cl * NACC);
    for (int r = 0; r < cl; ++r) {
        // transposed reduction of block_reduce_acc: lane l adds the partials of threads l, l + 32, ... in order, then the
        // shuffle-down tree over the lanes
        for (int a = 0; a < NACC; ++a) {
            double v[32];
            for (int l = 0; l < 32; ++l) {
                v[l] = acc[((size_t)r * NT + l) * NACC + a];
                for (int i = 1; i < NT / 32; ++i) v[l] = v[l] + acc[((size_t)r * NT + l + 32 * i) * NACC + a];
            }
            for (int o = 16; o > 0; o >>= 1)
                for (int l = 0; l < o; ++l) v[l] = v[l] + v[l + o];
            part[(size_t)r * NACC + a] = v[0];
        }
    }
    (void)nthr;
    for (int a = 0; a < NACC; ++a) {
        if (cl == 1) { tot[a] = part[a]; continue; }
        double s = 0.0;
        for (int r = 0; r < cl; ++r) s += part[(size_t)r * NACC + a];
        tot[a] = s;
    }
}
}  // namespace

extern "C" int orc_gicp_engine_order(const double *src_xyz, const double *src_nrm, int64_t ns, const double *tgt_xyz,
                                     const double *tgt_nrm, int64_t nt, double max_d, const double T_init[16], double epsilon,
                                     int loss, double loss_k, double rel_fitness, double rel_rmse, int max_iteration, int cl,
                                     double T_out[16], double *fitness_out, double *rmse_out, int32_t *iters_out,
                                     int64_t *ncorr_out, double *trace /* optional (max_iteration+1) x 3 */) {
    if (!(max_d > 0.0) || cl < 1 || cl > 1024) return 1;
    std::memcpy(T_out, T_init, sizeof(double) * 16);
    *fitness_out = 0; *rmse_out = 0; *iters_out = 0; *ncorr_out = 0;
    if (ns == 0 || nt == 0) return 0;
    const int nthr = cl * NT;
    const double r2 = max_d * max_d, k = 1.0 - epsilon;
    double T[16];
    std::memcpy(T, T_init, sizeof(T));
    bool ident = true;
    for (int i = 0; i < 16; ++i) ident &= (T[i] == ((i % 5 == 0) ? 1.0 : 0.0));
    std::vector<V3> p(ns), m(ns);
    for (int64_t i = 0; i < ns; ++i) {
        V3 pp = v3(src_xyz[3 * i], src_xyz[3 * i + 1], src_xyz[3 * i + 2]);
        V3 mm = effective_normal(v3(src_nrm[3 * i], src_nrm[3 * i + 1], src_nrm[3 * i + 2]));
        if (!ident) { pp = transform_point(T, pp); mm = rotate_vec(T, mm); }
        p[i] = pp; m[i] = mm;
    }
    void *tree = orc_kd_create(tgt_xyz, nt);
    if (!tree) return 2;
    std::vector<int32_t> corr(ns);
    std::vector<double> cd2(ns);
    std::vector<double> acc((size_t)nthr * NACC);
    double U[16], fit = 0, rmse = 0, pfit = 0, prmse = 0, Klast = 0;
    int iters = 0;
    for (int pass = 0;; ++pass) {
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < ns; ++i) {
            if (pass > 0) { p[i] = transform_point(U, p[i]); m[i] = rotate_vec(U, m[i]); }
            double q[3] = {p[i].x, p[i].y, p[i].z}, d2 = 0;
            int32_t j = orc_kd_nn(tree, q, &d2);
            if (j >= 0) d2 = dist2(p[i].x, p[i].y, p[i].z, tgt_xyz[3 * j], tgt_xyz[3 * j + 1], tgt_xyz[3 * j + 2]);
            if (j >= 0 && d2 < r2) { corr[i] = j; cd2[i] = d2; } else { corr[i] = -1; cd2[i] = 0; }
        }
        std::fill(acc.begin(), acc.end(), 0.0);
        // owner mapping of k_icp: warp gw owns Q consecutive queries per round (Q = 8/16/32 by the kernel's rule)
        const int nwarps = nthr / 32;
        int Q = 32;
        if ((ns + 7) / 8 <= nwarps) Q = 8;
        else if ((ns + 15) / 16 <= nwarps) Q = 16;
#pragma omp parallel for schedule(static)
        for (int t = 0; t < nthr; ++t) {
            double *a = &acc[(size_t)t * NACC];
            const int gw = t / 32, ln = t % 32;
            if (ln >= Q) continue;
            for (int64_t i = (int64_t)gw * Q + ln; i < ns; i += (int64_t)nwarps * Q) {
                int32_t j = corr[i];
                if (j < 0) continue;
                V3 q = v3(tgt_xyz[3 * (int64_t)j], tgt_xyz[3 * (int64_t)j + 1], tgt_xyz[3 * (int64_t)j + 2]);
                V3 mt = effective_normal(v3(tgt_nrm[3 * (int64_t)j], tgt_nrm[3 * (int64_t)j + 1], tgt_nrm[3 * (int64_t)j + 2]));
                gicp_accumulate(p[i], q, m[i], mt, k, loss, loss_k, a);
                a[27] += 1.0;
                a[28] += cd2[i];
            }
        }
        double tot[NACC];
        reduce_like_kernel(acc, cl, tot);
        const double K = tot[27], e2 = tot[28];
        Klast = K;
        if (K > 0.0) { fit = K / (double)ns; rmse = std::sqrt(e2 / K); } else { fit = 0; rmse = 0; }
        if (trace) { trace[3 * pass] = fit; trace[3 * pass + 1] = rmse; trace[3 * pass + 2] = K; }
        iters = pass;
        if (pass > 0 && std::fabs(pfit - fit) < rel_fitness && std::fabs(prmse - rmse) < rel_rmse) break;
        if (pass >= max_iteration) break;
        double Um[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        if (K > 0.0) {
            double x[6];
            ldlt_solve6(tot, x);
            vec6_to_mat4(x, Um);
        }
        mat4_mul(Um, T, T);
        std::memcpy(U, Um, sizeof(U));
        pfit = fit; prmse = rmse;
    }
    orc_kd_free(tree);
    std::memcpy(T_out, T, sizeof(T));
    *fitness_out = fit; *rmse_out = rmse; *iters_out = iters; *ncorr_out = (int64_t)Klast;
    return 0;
}

/* host probes of the shared math header, for tests/test_host_math.py */
extern "C" void probe_fast_eigen3x3(const double cov[6], double out[3]) {
    V3 n = fast_eigen3x3(cov);
    out[0] = n.x; out[1] = n.y; out[2] = n.z;
}
extern "C" void probe_weight_matrix(const double a[3], const double b[3], double k, double W[6]) {
    gicp_weight_matrix(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2]), k, W);
}
extern "C" void probe_ldlt_solve6(const double sums[27], double x[6]) { ldlt_solve6(sums, x); }
extern "C" void probe_vec6_to_mat4(const double x[6], double T[16]) { vec6_to_mat4(x, T); }
extern "C" void probe_trig(double x, double out[3]) { out[0] = det_sin(x); out[1] = det_cos(x); out[2] = det_acos(fmin(fmax(x, -1.0), 1.0)); }
