/*
 * mgicp.h -- C ABI of the B200-native multiscale Generalized-ICP refinement engine.
 *
 * Drop-in boundary.  The reference has no FFI: its boundary is the Python function
 *   Multiscale_GICP(source, target, n_scales, itera_escala, T_ini)
 *     /root/reference/ALL_FUNCTIONS.py:272-313
 *     /root/reference/2_MGICP_refinement_in_NCLT_dataset.py:128-164
 * whose body is a fixed sequence of Open3D calls.  Each entry point below names the Open3D call
 * (and the reference line that makes it) it replaces.  The Python mirror of the reference
 * interface lives in point-cloud-registration-with-global-refinement_b200/registration.py and binds
 * these symbols with ctypes.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary; every call returns an mgicp_status.
 *   - pointers marked DEVICE are CUDA device pointers valid on the handle's device, HOST are host
 *     pointers; work is enqueued on `stream` (a cudaStream_t passed as void*) and is stream-ordered
 *     unless stated otherwise; the caller owns every buffer it passes in.
 *   - 4x4 matrices are row-major double[16]; T maps SOURCE points into the TARGET frame.
 *   - one handle per (device, host thread); calls on one handle must be serialised by the caller.
 *   - there is NO CPU fallback: without a CUDA device mgicp_create fails.
 */
#ifndef MGICP_H_
#define MGICP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgicp_handle_s *mgicp_handle;

typedef enum {
    MGICP_OK = 0,
    MGICP_E_INVALID = 1,   /* bad argument: voxel_size <= 0, max_correspondence_distance <= 0 (Open3D raises RuntimeError) */
    MGICP_E_CUDA = 2,      /* CUDA runtime error, see mgicp_last_error */
    MGICP_E_NOMEM = 3,     /* workspace allocation failed */
    MGICP_E_RANGE = 4,     /* voxel / cell index range exceeded (extent / voxel_size >= 2^21; Open3D's limit is INT_MAX) */
    MGICP_E_OVERFLOW = 5,  /* internal hash table overflow (should not happen; reported, never silent) */
    MGICP_E_STATE = 6      /* stage accessor called before the stage ran */
} mgicp_status;

enum { MGICP_F32 = 0, MGICP_F64 = 1 };

/* robust kernels of Open3D's RobustKernel.cpp; the reference uses L1Loss (ALL_FUNCTIONS.py:284) */
enum { MGICP_LOSS_L2 = 0, MGICP_LOSS_L1 = 1, MGICP_LOSS_HUBER = 2, MGICP_LOSS_CAUCHY = 3, MGICP_LOSS_GM = 4, MGICP_LOSS_TUKEY = 5 };

typedef struct {
    int32_t sor_k;        /* remove_statistical_outlier nb_neighbors   (knn_filtro = 30, ALL_FUNCTIONS.py:280) */
    double  sor_std;      /* remove_statistical_outlier std_ratio      (std_filtro = 1.0, ALL_FUNCTIONS.py:281) */
    int32_t normal_k;     /* KDTreeSearchParamKNN(knn=20)              (ALL_FUNCTIONS.py:301) */
    double  epsilon;      /* TransformationEstimationForGeneralizedICP epsilon (Open3D default 1e-3) */
    int32_t loss;         /* MGICP_LOSS_*                              (ALL_FUNCTIONS.py:284) */
    double  loss_k;       /* scale of Huber/Cauchy/GM/Tukey */
    double  rel_fitness;  /* ICPConvergenceCriteria.relative_fitness   (ALL_FUNCTIONS.py:309) */
    double  rel_rmse;     /* ICPConvergenceCriteria.relative_rmse      (ALL_FUNCTIONS.py:310) */
    double  cell_factor;  /* tuning: kNN spatial-hash cell edge = cell_factor * voxel_size (<= 0: default 12) */
    double  icp_cell_factor; /* tuning: ICP spatial-hash cell edge = icp_cell_factor * voxel_size (<= 0: default 3.5; mgicp_run_batch: mgicp_auto_icp_cell_factor) */
    int32_t ctas_per_pair;/* tuning: thread-block cluster size cooperating on one pair's ICP loop (0: auto) */
    int32_t debug;        /* != 0: keep kNN neighbour lists and per-iteration traces for the stage accessors */
} mgicp_opts;

/* fills *o with the reference's constants listed above */
void mgicp_default_opts(mgicp_opts *o);

/* lifetime -------------------------------------------------------------------------------------- */
int mgicp_create(int device, mgicp_handle *out);
int mgicp_destroy(mgicp_handle h);
const char *mgicp_last_error(mgicp_handle h);
/* library version string, and the number of kernels launched by this handle so far (for gpu_launches) */
const char *mgicp_version(void);
int64_t mgicp_kernel_launches(mgicp_handle h);

/* K0: per-cloud axis-aligned bounds.  Replaces get_min_bound()/get_max_bound()
 * (ALL_FUNCTIONS.py:1093-1098, radius_from_cloud_pair) and the bounds inside VoxelDownSample.
 *   xyz        DEVICE  concatenated clouds, n_total x 3, float or double (xyz_dtype)
 *   cloud_off  HOST    int64[n_clouds + 1] point offsets
 *   bounds_out DEVICE  double[n_clouds * 6] = min xyz, max xyz
 */
int mgicp_cloud_bounds(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                       int32_t xyz_dtype, double *bounds_out);

/* Per-scale preprocessing of every cloud, all (cloud, scale) jobs batched per launch.  Replaces, per
 * cloud and scale, voxel_down_sample (ALL_FUNCTIONS.py:293-294), remove_statistical_outlier
 * (:297-298), estimate_normals (:301-302) and builds the spatial hash that replaces KDTreeFlann(target)
 * inside registration_generalized_icp.  Results stay in the handle's workspace for
 * mgicp_register_batch / mgicp_get_stage.
 *   voxel_sizes HOST double[n_scales]
 */
int mgicp_preprocess(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                     int32_t xyz_dtype, int32_t n_scales, const double *voxel_sizes, const mgicp_opts *opts);

/* The ICP loops of registration_generalized_icp (ALL_FUNCTIONS.py:304-311) for every pair and every
 * scale, coarse to fine with the transform chained (ALL_FUNCTIONS.py:312), in ONE launch: the
 * iteration loop, the 6x6 solve and the convergence test run on the device.
 * Needs a prior mgicp_preprocess on the same handle.
 *   pair_src, pair_tgt HOST   int32[n_pairs] cloud indices
 *   max_dists          HOST   double[n_pairs * n_scales] max_correspondence_distance per pair and scale
 *   max_iters          HOST   int32[n_scales]            ICPConvergenceCriteria.max_iteration per scale
 *   T_init             DEVICE double[n_pairs * 16]
 *   T_out              DEVICE double[n_pairs * 16]   result.transformation of the last scale
 *   fitness, rmse      DEVICE double[n_pairs]        result.fitness / result.inlier_rmse of the last scale
 *   iters              DEVICE int32[n_pairs * n_scales]  iterations executed per scale            (may be NULL)
 *   ncorr              DEVICE int32[n_pairs]         len(result.correspondence_set) of the last scale (may be NULL)
 *   stats              DEVICE double[n_pairs * n_scales * 8]: M'_src, M'_tgt, iterations, K_last, fitness, rmse,
 *                             sum over passes of K, passes                                          (may be NULL)
 */
int mgicp_register_batch(mgicp_handle h, void *stream, int32_t n_pairs, const int32_t *pair_src, const int32_t *pair_tgt,
                         const double *max_dists, const int32_t *max_iters, const mgicp_opts *opts, const double *T_init,
                         double *T_out, double *fitness, double *rmse, int32_t *iters, int32_t *ncorr, double *stats);

/* The cell edge of the ICP grid, in voxels, that suits a schedule: the largest max_dists[p][s] / voxel_sizes[s] clamped to
 * [3.5, 16] (3.5 for the script-2 schedule, 2_MGICP...py:110-120; 16 for the ALL_FUNCTIONS one whose search radius is the cloud's
 * size, ALL_FUNCTIONS.py:260-278, 1092-1101).  mgicp_run_batch applies it when opts->icp_cell_factor == 0; callers of
 * mgicp_preprocess + mgicp_register_batch set opts->icp_cell_factor themselves (0 there means 3.5).  Any value gives the same
 * nearest neighbours; only the point order inside the grid, hence the summation order, depends on it.  HOST arrays. */
double mgicp_auto_icp_cell_factor(int32_t n_scales, const double *voxel_sizes, int32_t n_pairs, const double *max_dists);

/* Convenience: mgicp_preprocess + mgicp_register_batch.  This is the call that replaces the body of
 * Multiscale_GICP (ALL_FUNCTIONS.py:286-312) for a batch of pairs that share a cloud list. */
int mgicp_run_batch(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                    int32_t xyz_dtype, int32_t n_scales, const double *voxel_sizes, int32_t n_pairs,
                    const int32_t *pair_src, const int32_t *pair_tgt, const double *max_dists, const int32_t *max_iters,
                    const mgicp_opts *opts, const double *T_init, double *T_out, double *fitness, double *rmse,
                    int32_t *iters, int32_t *ncorr, double *stats);

/* Synchronous: waits for the device and returns the first device-side error flag of the last mgicp_preprocess
 * (MGICP_E_RANGE, MGICP_E_OVERFLOW) or MGICP_OK.  Stream-ordered calls cannot report those themselves. */
int mgicp_check(mgicp_handle h);

/* Stream-ordered form of mgicp_check: *err_out (DEVICE int32) receives MGICP_OK or the first device-side error flag
 * (MGICP_E_RANGE, MGICP_E_OVERFLOW) of the jobs of the last mgicp_preprocess / mgicp_evaluate_clouds, so that a pipelined
 * caller can bring it home together with the results instead of synchronising the device. */
int mgicp_job_errors(mgicp_handle h, void *stream, int32_t *err_out);

/* Optional stage timing for the measurement rows of SURVEY 8(d) / BASELINE.md 2.1 item 3 (per-stage split, time per
 * scale).  With timing on, mgicp_preprocess / mgicp_register_batch record CUDA events at their stage boundaries on the
 * stream they are given and the ICP kernel stamps %globaltimer at the start and end of every pair's scale.
 * mgicp_get_timing (synchronous) returns, for the last preprocess (+ register) on this handle, HOST double[16] in ms:
 *   [0] voxel_down_sample (bounds, voxel hash, centroids)   [1] kNN spatial hash build
 *   [2] remove_statistical_outlier (kNN 30 + selection)     [3] estimate_normals (+ certificate lists)
 *   [4] ICP spatial hash build                              [5] registration_generalized_icp, all pairs and scales
 *   [8 + s], s < 8: mean over the pairs of the wall time between a pair's first and last pass at scale s (in a batch this
 *   includes the time the pair's tasks wait in the queue; for a single pair it is the latency of the scale). */
int mgicp_set_timing(mgicp_handle h, int32_t on);
int mgicp_get_timing(mgicp_handle h, double *ms_out);

/* One correspondence pass at a given pose: replaces evaluate_registration (ALL_FUNCTIONS.py:809-822) on the
 * preprocessed clouds of `scale`; also returns the 27 normal-equation sums (21 upper-triangular JTJ terms then
 * 6 JTr terms) of TransformationEstimationForGeneralizedICP::ComputeTransformation at that pose.
 *   T DEVICE double[n_pairs*16]; out DEVICE double[n_pairs * 32]: fitness, rmse, K, sum d^2, 27 sums, pad
 */
int mgicp_evaluate_batch(mgicp_handle h, void *stream, int32_t scale, int32_t n_pairs, const int32_t *pair_src,
                         const int32_t *pair_tgt, const double *max_dists, const mgicp_opts *opts, const double *T,
                         double *out);

/* evaluate_registration (ALL_FUNCTIONS.py:809-822, calculate_RMSE_and_fitness) and
 * get_information_matrix_from_point_clouds (ALL_FUNCTIONS.py:327-331; 3_Global_Refinement...py:317-320, 331-334) for a
 * batch of pairs, on the clouds AS GIVEN (no down-sampling): source transformed by T, nearest target point accepted iff
 * d < max_dist, fitness = K / N_source, inlier_rmse = sqrt(sum d^2 / K), GTG = sum over correspondences of G^T G built
 * from the target point.  Builds its own spatial hash per target cloud in the handle's workspace: a previous
 * mgicp_preprocess on this handle is invalidated.  Stream-ordered; mgicp_check reports range errors afterwards.
 *   xyz, cloud_off  as in mgicp_preprocess (DEVICE / HOST)
 *   max_dists HOST double[n_pairs];  T HOST double[n_pairs * 16]
 *   out   DEVICE double[n_pairs * 32]: fitness, rmse, K, sum d^2, the 21 upper-triangular terms of GTG (row-major), pad
 *   corr  DEVICE int32[sum over pairs of N_source(pair)] target index per source point (-1: none), pairs back to back;
 *         may be NULL (the reference reads correspondence_set only for visualisation, ALL_FUNCTIONS.py:1064)
 */
int mgicp_evaluate_clouds(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                          int32_t xyz_dtype, int32_t n_pairs, const int32_t *pair_src, const int32_t *pair_tgt,
                          const double *max_dists, const double *T, double *out, int32_t *corr);

/* FGR front end, feature stage (first, unoptimised CUDA path, see csrc/mgicp_fgr.cuh): for every cloud as given,
 *   estimate_normals(KDTreeSearchParamHybrid(radius_normals, max_nn_normals))       ALL_FUNCTIONS.py:181-183, 1_FGR...py:44-46
 *   compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius_fpfh, max_nn_fpfh))    ALL_FUNCTIONS.py:185-187, 1_FGR...py:48-50
 * Hybrid search = the max_nn nearest points (the query included) with d^2 < radius^2.  Builds its own spatial hash per
 * cloud in the handle's workspace: a previous mgicp_preprocess on this handle is invalidated.  Stream-ordered.
 *   xyz, cloud_off  as in mgicp_preprocess (DEVICE / HOST); cloud_off[0] == 0
 *   normals_out DEVICE double[total_points * 3]   unit normals (Open3D's orientation: none), original point order
 *   fpfh_out    DEVICE double[total_points * 33]  one 33-bin descriptor per point (Open3D's Feature.data column)
 */
int mgicp_fpfh_clouds(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                      int32_t xyz_dtype, double radius_normals, int32_t max_nn_normals, double radius_fpfh,
                      int32_t max_nn_fpfh, double *normals_out, double *fpfh_out);

/* FGR front end, registration stage (see csrc/mgicp_fgr.cuh):
 *   registration_fgr_based_on_feature_matching(source, target, source_fpfh, target_fpfh, FastGlobalRegistrationOption(...))
 *   for a batch of pairs                                                       ALL_FUNCTIONS.py:189-202, 1_FGR...py:52-65
 * Nearest neighbours in descriptor space both ways, cross check, tuple test (counter-based generator, one seed per pair:
 * Open3D draws from its own global engine, results are comparable within FGR's run-to-run scatter only), graduated
 * non-convexity, transformation mapped back to the original scale and inverted: T maps SOURCE into the TARGET frame.
 *   xyz, cloud_off  as in mgicp_preprocess (DEVICE / HOST); cloud_off[0] == 0;  feat DEVICE double[total_points * 33]
 *   pair_src, pair_tgt, seeds HOST [n_pairs];  T_out DEVICE double[n_pairs * 16] row-major;  ncorr_out DEVICE int32[n_pairs]
 *   tuple_counts HOST int32[n_pairs] maximum_tuple_count per pair (the reference derives it from the pair's sizes), or NULL:
 *                opts->maximum_tuple_count for every pair
 */
typedef struct {
    double division_factor;                 /* Open3D default 1.4 */
    int32_t use_absolute_scale;             /* default 0; the reference passes True */
    int32_t decrease_mu;                    /* default 0 (Open3D >= 0.13: 1); the reference passes True */
    double maximum_correspondence_distance; /* default 0.025; the reference passes 2 * voxel_size */
    int32_t iteration_number;               /* default 64; the reference passes 300 */
    double tuple_scale;                     /* default 0.95 */
    int32_t maximum_tuple_count;            /* default 1000; the reference passes int(0.2 * (n_source + n_target) / 2) */
} mgicp_fgr_opts;
int mgicp_fgr_pairs(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                    int32_t xyz_dtype, const double *feat, int32_t n_pairs, const int32_t *pair_src, const int32_t *pair_tgt,
                    const mgicp_fgr_opts *opts, const int32_t *tuple_counts, const uint64_t *seeds, double *T_out,
                    int32_t *ncorr_out);

/* Stage accessors for the parity tests (synchronous; copy from the workspace into HOST memory).
 * `what` selects the array; `dst` has room for `cap` elements of the array's element type; *count receives the
 * number of ROWS (points) written. */
enum {
    MGICP_STAGE_DOWNSAMPLED = 0, /* double[M*3]   voxel centroids, canonical (hash-slot) order          */
    MGICP_STAGE_GRID_POINTS = 1, /* double[M*3]   same points in spatial-hash order (the order of 2,3,6)   */
    MGICP_STAGE_SOR_AVG     = 2, /* double[M]     mean kNN distance per point                              */
    MGICP_STAGE_SOR_KEEP    = 3, /* uint8[M]      1 = survives remove_statistical_outlier                  */
    MGICP_STAGE_POINTS      = 4, /* double[M'*3]  final points (after SOR), order of 5 and 7               */
    MGICP_STAGE_NORMALS     = 5, /* double[M'*3]  estimate_normals result                                  */
    MGICP_STAGE_KNN_SOR     = 6, /* int32[M*sor_k]     neighbour indices into 1 (debug != 0), -1 padded    */
    MGICP_STAGE_KNN_NORMAL  = 7, /* int32[M'*normal_k] neighbour indices into 4 (debug != 0), -1 padded    */
    MGICP_STAGE_BOUNDS      = 8, /* double[6]     min xyz, max xyz of the raw cloud                        */
    MGICP_STAGE_ICP_POINTS  = 9, /* double[M'*3]  final points in the order the ICP kernel walks them      */
    MGICP_STAGE_ICP_NORMALS = 10 /* double[M'*3]  their normals                                            */
};
int mgicp_get_stage(mgicp_handle h, int32_t cloud, int32_t scale, int32_t what, void *dst, int64_t cap, int64_t *count);

/* RegistrationResult.correspondence_set of the last scale (read by the reference only for drawing, ALL_FUNCTIONS.py:1064) for
 * pair `pair` of the last mgicp_register_batch on this handle.  Synchronous.  dst HOST int32[cap * 2] receives rows
 * (source index, target index) into the clouds registration_generalized_icp saw at the last scale, i.e. MGICP_STAGE_POINTS of
 * (source cloud, last scale) and (target cloud, last scale), ordered by source index; *count = number of rows
 * (= ncorr of that pair).  Valid until the next call that uses the handle's workspace. */
int mgicp_get_correspondences(mgicp_handle h, int32_t pair, int32_t *dst, int64_t cap, int64_t *count);

#ifdef __cplusplus
}
#endif
#endif /* MGICP_H_ */
