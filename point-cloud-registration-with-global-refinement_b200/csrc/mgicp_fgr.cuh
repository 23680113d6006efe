// mgicp_fgr.cuh -- first CUDA path of the FGR front end's feature stage: hybrid-radius normals and FPFH
// on the clouds as given.  Included at the end of mgicp.cu (it re-uses the raw-cloud spatial-hash build of
// mgicp_evaluate_clouds and the handle's workspace).
//
// STATUS: the per-point arithmetic (csrc/fpfh_math.cuh) is shared with the host and checked bit for bit against the oracle on
// the CPU (oracle/fpfh_engine.cpp); on a B200 normals and descriptors of the NCLT fixtures equal the oracle's bit for bit
// (tests/test_gpu_fgr.py).  Deliberately simple -- one thread per point everywhere, unmeasured: the warp-cooperative versions
// and the rest of FGR (matching, tuple test, optimisation) are the next steps.
#pragma once
#include "fpfh_math.cuh"

struct FgrArgs {
    const Job *jobs;                 // one job per cloud, ICP grid (which = 2) built over the raw cloud, cell >= every radius
    const int64_t *cloud_off;        // [clouds + 1] point offsets
    double r2;                       // squared search radius
    int cap;                         // max_nn
    int32_t *idx;                    // [points][cap] neighbour indices (cloud-local, original order), ascending (d2, idx)
    double *d2;                      // [points][cap]
    int32_t *cnt;                    // [points]
    const int32_t *idx_n; const int32_t *cnt_n; int cap_n;     // the normals' lists (k_hybrid_normals)
    double *normals;                 // [points][3]
    double *spfh;                    // [points][33]
    double *fpfh;                    // [points][33]
};

struct FgrPointAt {
    const double4 *pts;
    MG_HD V3 operator()(int32_t j) const { const double4 q = pts[j]; return v3(q.x, q.y, q.z); }
};
struct FgrNormalAt {
    const double *nrm;
    MG_HD V3 operator()(int32_t j) const { return v3(nrm[3 * (size_t)j], nrm[3 * (size_t)j + 1], nrm[3 * (size_t)j + 2]); }
};
struct FgrSpfhAt {
    const double *spfh;
    MG_HD const double *operator()(int32_t j) const { return spfh + 33 * (size_t)j; }
};

// grid (chunks, clouds), one thread per query: KDTreeFlann::SearchHybrid(p, r, max_nn) = the max_nn nearest points with
// d^2 < r^2, ascending (the query itself first).  27 cells of a grid whose cell edge is >= r; a bounded sorted list per query
// in global memory (insertion from the back).
__global__ void __launch_bounds__(128) k_hybrid_lists(FgrArgs A, int only_flagged) {
    const Job &J = A.jobs[blockIdx.y];
    if (J.err) return;
    const GridView g = make_view(J, 2);
    const int64_t base = A.cloud_off[blockIdx.y];
    const int cap = A.cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
        if (only_flagged && A.cnt[base + i] != -1) continue;       // k_hybrid_lists_w settled this query
        const double4 p = J.pts[i];
        int32_t *li = A.idx + (size_t)(base + i) * cap;
        double *ld = A.d2 + (size_t)(base + i) * cap;
        int cnt = 0;
        const int cx = cell_coord(p.x, g.org[0], g.cell), cy = cell_coord(p.y, g.org[1], g.cell), cz = cell_coord(p.z, g.org[2], g.cell);
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int x = cx + dx, y = cy + dy, z = cz + dz;
                    if (x < 0 || x >= g.dim[0] || y < 0 || y >= g.dim[1] || z < 0 || z >= g.dim[2]) continue;
                    int s, c;
                    if (!cell_find(g.tab, g.bits, pack_key(x, y, z), s, c)) continue;
                    for (int t = s; t < s + c; ++t) {
                        const double4 q = ldg4(g.pts + t);
                        const double d = dist2(p.x, p.y, p.z, q.x, q.y, q.z);
                        if (!(d < A.r2)) continue;
                        const int32_t o = J.i2a[t];
                        if (cnt == cap && !(d < ld[cap - 1] || (d == ld[cap - 1] && o < li[cap - 1]))) continue;
                        int pos = cnt < cap ? cnt : cap - 1;                     // slot that opens up
                        while (pos > 0 && (d < ld[pos - 1] || (d == ld[pos - 1] && o < li[pos - 1]))) {
                            ld[pos] = ld[pos - 1]; li[pos] = li[pos - 1];
                            --pos;
                        }
                        ld[pos] = d; li[pos] = o;
                        if (cnt < cap) ++cnt;
                    }
                }
        A.cnt[base + i] = cnt;
    }
}

// grid (chunks, clouds), one WARP per query: the same lists, built the way the candidates come -- the 27 cells are looked up one
// per lane, their points packed over the lanes, the ones inside the radius appended (ballot + prefix) to a per-warp list in
// shared memory; one bitonic sort by (d2, index) per query, the first max_nn entries written out.  The thread-per-query
// kernel above keeps a sorted list in GLOBAL memory and shifts it at every insertion (8.7 ms for two NCLT clouds at the FPFH
// radius; this one: see profiles/).  A query with more than HW_CAP points inside the radius is flagged (cnt = -1) and left to
// the kernel above.
constexpr int HW_WARPS = 8, HW_CAP = 512;
struct HwWarp { double d[HW_CAP]; int32_t o[HW_CAP]; };
__global__ void __launch_bounds__(HW_WARPS * 32) k_hybrid_lists_w(FgrArgs A) {
    extern __shared__ __align__(16) unsigned char hw_raw[];
    const Job &J = A.jobs[blockIdx.y];
    if (J.err) return;
    const GridView g = make_view(J, 2);
    const int64_t base = A.cloud_off[blockIdx.y];
    const int cap = A.cap, lane = threadIdx.x & 31;
    HwWarp &W = reinterpret_cast<HwWarp *>(hw_raw)[threadIdx.x >> 5];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < g.n; i += nwarp) {
        const double4 p = J.pts[i];
        const int cx = cell_coord(p.x, g.org[0], g.cell), cy = cell_coord(p.y, g.org[1], g.cell), cz = cell_coord(p.z, g.org[2], g.cell);
        int s = 0, c = 0;
        if (lane < 27) {
            const int dz = lane / 9, r = lane - 9 * dz, dy = r / 3, dx = r - 3 * dy;
            const int x = cx + dx - 1, y = cy + dy - 1, z = cz + dz - 1;
            if (x >= 0 && x < g.dim[0] && y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2])
                if (!cell_find(g.tab, g.bits, pack_key(x, y, z), s, c)) c = 0;
        }
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        int n = 0;                                                    // entries in the list (warp-uniform)
        bool overflow = false;
        for (int r0 = 0; r0 < total && !overflow; r0 += 32) {
            const int gi = r0 + lane;
            const int owner = lane_of_slot(incl, gi);
            const int t = __shfl_sync(FULL, s, owner) + (gi - (__shfl_sync(FULL, incl, owner) - __shfl_sync(FULL, c, owner)));
            double d = INFINITY;
            if (gi < total) {
                const double4 q = ldg4(g.pts + t);
                d = dist2(p.x, p.y, p.z, q.x, q.y, q.z);
            }
            const bool in = gi < total && d < A.r2;
            const unsigned m = __ballot_sync(FULL, in);
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            if (n + __popc(m) > HW_CAP) { overflow = true; break; }
            if (in) { W.d[pos] = d; W.o[pos] = J.i2a[t]; }
            n += __popc(m);
        }
        if (overflow) { if (lane == 0) A.cnt[base + i] = -1; __syncwarp(); continue; }
        // bitonic sort of the list padded to a power of two with +inf
        int N = 32;
        while (N < n) N <<= 1;
        for (int e = n + lane; e < N; e += 32) { W.d[e] = INFINITY; W.o[e] = 0x7fffffff; }
        __syncwarp();
        for (int k2 = 2; k2 <= N; k2 <<= 1)
            for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                for (int t = lane; t < (N >> 1); t += 32) {
                    const int a = ((t & ~(j2 - 1)) << 1) | (t & (j2 - 1)), b = a | j2;      // a < b, partner pair of this step
                    const bool up = (a & k2) == 0;
                    const double da = W.d[a], db = W.d[b];
                    const int32_t oa = W.o[a], ob = W.o[b];
                    const bool b_first = db < da || (db == da && ob < oa);
                    if (b_first == up) { W.d[a] = db; W.o[a] = ob; W.d[b] = da; W.o[b] = oa; }
                }
                __syncwarp();
            }
        const int cnt = min(n, cap);
        int32_t *li = A.idx + (size_t)(base + i) * cap;
        double *ld = A.d2 + (size_t)(base + i) * cap;
        for (int e = lane; e < cnt; e += 32) { li[e] = W.o[e]; ld[e] = W.d[e]; }
        if (lane == 0) A.cnt[base + i] = cnt;
        __syncwarp();
    }
}

// grid (chunks, clouds), one thread per point: EstimateNormals on a cloud without normals
__global__ void __launch_bounds__(128) k_hybrid_normals(FgrArgs A) {
    const Job &J = A.jobs[blockIdx.y];
    if (J.err) return;
    const int64_t base = A.cloud_off[blockIdx.y];
    const FgrPointAt point_at{J.pts};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < J.Mf; i += gridDim.x * blockDim.x) {
        double cov[6];
        hybrid_covariance(A.idx_n + (size_t)(base + i) * A.cap_n, A.cnt_n[base + i], point_at, cov);
        const V3 nv = normal_from_cov(cov);
        double *o = A.normals + 3 * (size_t)(base + i);
        o[0] = nv.x; o[1] = nv.y; o[2] = nv.z;
    }
}

// grid (chunks, clouds), one thread per point: ComputeSPFHFeature
__global__ void __launch_bounds__(128) k_spfh(FgrArgs A) {
    const Job &J = A.jobs[blockIdx.y];
    if (J.err) return;
    const int64_t base = A.cloud_off[blockIdx.y];
    const FgrPointAt point_at{J.pts};
    const FgrNormalAt normal_at{A.normals + 3 * (size_t)base};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < J.Mf; i += gridDim.x * blockDim.x) {
        double hist[33];
#pragma unroll
        for (int b = 0; b < 33; ++b) hist[b] = 0.0;
        spfh_point(A.idx + (size_t)(base + i) * A.cap, A.cnt[base + i], point_at(i), normal_at(i), point_at, normal_at, hist);
        double *o = A.spfh + 33 * (size_t)(base + i);
        for (int b = 0; b < 33; ++b) o[b] = hist[b];
    }
}

// grid (chunks, clouds), one thread per point: ComputeFPFHFeature
__global__ void __launch_bounds__(128) k_fpfh(FgrArgs A) {
    const Job &J = A.jobs[blockIdx.y];
    if (J.err) return;
    const int64_t base = A.cloud_off[blockIdx.y];
    const FgrSpfhAt spfh_at{A.spfh + 33 * (size_t)base};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < J.Mf; i += gridDim.x * blockDim.x) {
        double out[33];
#pragma unroll
        for (int b = 0; b < 33; ++b) out[b] = 0.0;
        fpfh_point(A.idx + (size_t)(base + i) * A.cap, A.d2 + (size_t)(base + i) * A.cap, A.cnt[base + i], spfh_at(i), spfh_at, out);
        double *o = A.fpfh + 33 * (size_t)(base + i);
        for (int b = 0; b < 33; ++b) o[b] = out[b];
    }
}

extern "C" int mgicp_fpfh_clouds(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                                 int32_t xyz_dtype, double radius_normals, int32_t max_nn_normals, double radius_fpfh,
                                 int32_t max_nn_fpfh, double *normals_out, double *fpfh_out) {
    if (!h) return MGICP_E_INVALID;
    h->preprocessed = false;        // the workspace is reused: a previous mgicp_preprocess is gone after this call
    if (n_clouds <= 0 || !xyz || !cloud_off || !normals_out || !fpfh_out || (xyz_dtype != MGICP_F32 && xyz_dtype != MGICP_F64) ||
        !(radius_normals > 0.0) || !(radius_fpfh > 0.0) || max_nn_normals < 1 || max_nn_fpfh < 1 || max_nn_normals > 4096 ||
        max_nn_fpfh > 4096) { h->err = "mgicp_fpfh_clouds: bad arguments"; return MGICP_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    const size_t esz = xyz_dtype == MGICP_F32 ? 4 : 8;
    // one grid per cloud whose cell edge covers both radii: a neighbour within r lies in the 27 cells around the query's
    const double cell = std::max(radius_normals, radius_fpfh) * (1.0 + 1e-6);
    std::vector<Job> jobs(n_clouds);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o_ = off; off += align_up(bytes); return o_; };
    const size_t o_jobs = take(sizeof(Job) * n_clouds), o_benc = take(sizeof(u64) * BENC_W * n_clouds), o_coff = take(sizeof(int64_t) * (n_clouds + 1));
    int64_t maxn = 0;
    const int64_t total = cloud_off[n_clouds] - cloud_off[0];
    if (cloud_off[0] != 0 || total < 0) { h->err = "mgicp_fpfh_clouds: cloud_off must start at 0 and ascend"; return MGICP_E_INVALID; }
    std::vector<size_t> offs((size_t)n_clouds * 8, 0);
    for (int c = 0; c < n_clouds; ++c) {
        Job &j = jobs[c];
        memset(&j, 0, sizeof(Job));
        const int64_t n = cloud_off[c + 1] - cloud_off[c];
        if (n < 0 || n > (int64_t)1 << 30) { h->err = "cloud too large"; return MGICP_E_INVALID; }
        j.xyz = (const char *)xyz + (size_t)cloud_off[c] * 3 * esz;
        j.dtype = xyz_dtype; j.cloud = c; j.n = n;
        j.voxel = cell; j.cell = cell; j.cell_i = cell;
        int cb = 10; while (((int64_t)1 << cb) < 4 * n) ++cb;
        j.vbits = 10; j.cbits_max = cb;
        maxn = std::max(maxn, n);
        const size_t m = (size_t)std::max<int64_t>(n, 1), ccap = ((size_t)1 << cb) + TAB_PAD;
        size_t *o_ = &offs[(size_t)c * 8];
        o_[0] = take(sizeof(CellSlot) * ccap);   // itab
        o_[1] = take(sizeof(int32_t) * ccap);    // ccursor
        o_[2] = take(sizeof(int32_t) * m);       // pslot
        o_[3] = take(sizeof(int32_t) * m);       // order
        o_[4] = take(sizeof(double4) * m);       // pts
        o_[5] = take(sizeof(double4) * m);       // nrm
        o_[6] = take(sizeof(double4) * m * 2);   // ipts, inrm
        o_[7] = take(sizeof(int32_t) * m * 2);   // a2i, i2a
    }
    const size_t tp = (size_t)std::max<int64_t>(total, 1);
    const size_t o_in = take(sizeof(int32_t) * tp * max_nn_normals), o_dn = take(sizeof(double) * tp * max_nn_normals), o_cn = take(sizeof(int32_t) * tp);
    const size_t o_if = take(sizeof(int32_t) * tp * max_nn_fpfh), o_df = take(sizeof(double) * tp * max_nn_fpfh), o_cf = take(sizeof(int32_t) * tp);
    const size_t o_sp = take(sizeof(double) * tp * 33);
    int rc = grow(h, &h->arena, &h->arena_bytes, off);
    if (rc) return rc;
    char *base = h->arena;
    for (int c = 0; c < n_clouds; ++c) {
        Job &j = jobs[c];
        const size_t m = (size_t)std::max<int64_t>(j.n, 1);
        size_t *o_ = &offs[(size_t)c * 8];
        j.itab = (CellSlot *)(base + o_[0]); j.ccursor = (int32_t *)(base + o_[1]); j.pslot = (int32_t *)(base + o_[2]);
        j.order = (int32_t *)(base + o_[3]); j.pts = (double4 *)(base + o_[4]); j.nrm = (double4 *)(base + o_[5]);
        j.ipts = (double4 *)(base + o_[6]); j.inrm = j.ipts + m; j.a2i = (int32_t *)(base + o_[7]); j.i2a = j.a2i + m;
    }
    Job *jobs_dev = (Job *)(base + o_jobs);
    u64 *benc = (u64 *)(base + o_benc);
    int64_t *coff_dev = (int64_t *)(base + o_coff);
    CK(cudaMemcpyAsync(jobs_dev, jobs.data(), sizeof(Job) * n_clouds, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(coff_dev, cloud_off, sizeof(int64_t) * (n_clouds + 1), cudaMemcpyHostToDevice, st));
    const int cx_raw = chunks_for(maxn, 256 * 8, 256), cx_pts = chunks_for(maxn, 256 * 2, 1024);
    k_bounds_init<<<(n_clouds * BENC_W + 127) / 128, 128, 0, st>>>(benc, n_clouds);
    k_bounds<<<dim3(chunks_for(maxn, 256 * 8, 64), n_clouds), 256, 0, st>>>(xyz, xyz_dtype, coff_dev, benc);
    k_job_setup<<<(n_clouds + 127) / 128, 128, 0, st>>>(jobs_dev, n_clouds, benc);
    k_raw_load<<<dim3(cx_raw, n_clouds), 256, 0, st>>>(jobs_dev);
    k_table_clear<<<dim3(cx_pts, n_clouds), 256, 0, st>>>(jobs_dev, 2);
    k_cell_insert<<<dim3(cx_pts, n_clouds), 256, 0, st>>>(jobs_dev, 2);
    k_cell_count<<<dim3(cx_pts, n_clouds), 256, 0, st>>>(jobs_dev, 2);
    k_cell_scan<<<n_clouds, 1024, 0, st>>>(jobs_dev, 2);
    k_cell_scatter<<<dim3(cx_pts, n_clouds), 256, 0, st>>>(jobs_dev, 2);
    k_cell_gather<<<dim3(cx_pts, n_clouds), 256, 0, st>>>(jobs_dev, 2);
    FgrArgs A;
    A.jobs = jobs_dev; A.cloud_off = coff_dev;
    A.idx_n = (const int32_t *)(base + o_in); A.cnt_n = (const int32_t *)(base + o_cn); A.cap_n = max_nn_normals;
    A.normals = normals_out; A.spfh = (double *)(base + o_sp); A.fpfh = fpfh_out;
    const dim3 grid(chunks_for(maxn, 128, 4096), n_clouds);
    // normals' lists, then the features' lists
    A.r2 = radius_normals * radius_normals; A.cap = max_nn_normals;
    A.idx = (int32_t *)(base + o_in); A.d2 = (double *)(base + o_dn); A.cnt = (int32_t *)(base + o_cn);
    int lists_mode = 1;                   // 1: warp per query + shared-memory sort (flagged leftovers: thread per query), 0: thread per query
    if (const char *e = getenv("MGICP_FGR_LISTS")) lists_mode = atoi(e);                                    // A/B experiments
    const dim3 grid_w(chunks_for(maxn, HW_WARPS * 4, 2048), n_clouds);
    CK(cudaFuncSetAttribute(k_hybrid_lists_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(HW_WARPS * sizeof(HwWarp))));
    if (lists_mode) k_hybrid_lists_w<<<grid_w, HW_WARPS * 32, HW_WARPS * sizeof(HwWarp), st>>>(A);
    k_hybrid_lists<<<grid, 128, 0, st>>>(A, lists_mode);
    k_hybrid_normals<<<grid, 128, 0, st>>>(A);
    A.r2 = radius_fpfh * radius_fpfh; A.cap = max_nn_fpfh;
    A.idx = (int32_t *)(base + o_if); A.d2 = (double *)(base + o_df); A.cnt = (int32_t *)(base + o_cf);
    if (lists_mode) k_hybrid_lists_w<<<grid_w, HW_WARPS * 32, HW_WARPS * sizeof(HwWarp), st>>>(A);
    k_hybrid_lists<<<grid, 128, 0, st>>>(A, lists_mode);
    k_spfh<<<grid, 128, 0, st>>>(A);
    k_fpfh<<<grid, 128, 0, st>>>(A);
    h->launches += lists_mode ? 17 : 15;
    CK(cudaGetLastError());
    h->eval_jobs = jobs_dev; h->eval_n = n_clouds;
    return MGICP_OK;
}

// =============================================================================================
// Fast Global Registration on given descriptors: registration_fgr_based_on_feature_matching for a batch of pairs.
// STATUS: the per-item arithmetic (csrc/fgr_math.cuh) reproduces the oracle bit for bit on the CPU (oracle/fpfh_engine.cpp,
// its FGR check); on the B200 the kernels give the oracle's correspondences and, with the oracle's sums taken in the kernel's
// order, its pose bit for bit (tests/test_gpu_fgr.py, profiles/r2_first/r2_fgr_first.log).
// Matching: tensor cores + exact re-check (mgicp_fgr_tc.cuh); k_fgr_nn below is the brute-force fp64 version it must equal
// (one thread per query, targets staged through shared memory; MGICP_FGR_MATCH=0).  Then ONE block per pair for everything
// sequential in nature (normalisation, mutual matches, tuple test, 300 solves).
// =============================================================================================
#include "fgr_math.cuh"

struct FgrPair {                       // per pair, device pointers into the workspace
    int32_t src, tgt;                  // cloud indices
    int32_t fi, fj;                    // 0 = source, 1 = target: "i" is the larger cloud (AdvancedMatching's swap)
    int64_t ni, nj;
    V3 *P[2];                          // centred (and scaled) points of source / target
    V3 *Pc, *Qc;                       // [3 * cap] per correspondence: its source point, and the moving copy of its target point
    int32_t *j2i, *i2j;                // nearest i-descriptor of every j / nearest j-descriptor of every i
    int32_t *cross;                    // [min(ni, nj)][2] mutual matches, ascending i
    int32_t *cor;                      // [3 * cap][2] (source index, target index)
};

struct FgrRunArgs {
    const void *xyz; int dtype; const int64_t *cloud_off;
    const double *feat;                // [points][33]
    FgrPair *pairs;
    double division_factor, maximum_correspondence_distance, tuple_scale;
    int use_absolute_scale, decrease_mu, iteration_number, maximum_tuple_count;
    const uint64_t *seeds;             // [pairs]
    const int32_t *caps;               // [pairs] maximum_tuple_count per pair, or null: maximum_tuple_count for all
    double *T_out;                     // [pairs][16]
    int32_t *ncorr_out;                // [pairs]
};

constexpr int FGR_NN_NT = 128, FGR_NN_TILE = 32;

// grid (chunks of queries, 2 * pairs): direction 0 fills j2i (queries = descriptors of cloud j, searched among cloud i's),
// direction 1 fills i2j.  Exact fp64 distances summed in bin order, targets scanned in ascending index with a strict '<':
// the nearest neighbour with ties to the lower index, exactly what the oracle's brute force returns.
__global__ void __launch_bounds__(FGR_NN_NT) k_fgr_nn(FgrRunArgs A, const int32_t *only_if_count = nullptr, int only_if_above = 0) {
    __shared__ double tile[FGR_NN_TILE][33];
    if (only_if_count && only_if_count[blockIdx.y] <= only_if_above) return;      // fallback of the tensor-core matcher: rarely needed
    const FgrPair &pr = A.pairs[blockIdx.y >> 1];
    const int dir = blockIdx.y & 1;
    const int cq = dir == 0 ? pr.fj : pr.fi, ct = dir == 0 ? pr.fi : pr.fj;           // which of (source, target) queries / is searched
    const int64_t nq = dir == 0 ? pr.nj : pr.ni, nt = dir == 0 ? pr.ni : pr.nj;
    const double *Fq = A.feat + 33 * A.cloud_off[cq == 0 ? pr.src : pr.tgt];
    const double *Ft = A.feat + 33 * A.cloud_off[ct == 0 ? pr.src : pr.tgt];
    int32_t *out = dir == 0 ? pr.j2i : pr.i2j;
    for (int64_t q0 = (int64_t)blockIdx.x * FGR_NN_NT; q0 < nq; q0 += (int64_t)gridDim.x * FGR_NN_NT) {   // block-uniform trip count
        const int64_t q = q0 + threadIdx.x;
        double qf[33];
#pragma unroll
        for (int k = 0; k < 33; ++k) qf[k] = q < nq ? Fq[33 * q + k] : 0.0;
        double best = INFINITY;
        int32_t bj = -1;
        for (int64_t t0 = 0; t0 < nt; t0 += FGR_NN_TILE) {
            __syncthreads();
            for (int e = threadIdx.x; e < FGR_NN_TILE * 33; e += FGR_NN_NT) {
                const int64_t t = t0 + e / 33;
                tile[e / 33][e % 33] = t < nt ? Ft[33 * t + e % 33] : 0.0;
            }
            __syncthreads();
            const int m = (int)min((int64_t)FGR_NN_TILE, nt - t0);
            for (int b = 0; b < m; ++b) {
                const double s = fgr_feat_dist2(qf, tile[b]);
                if (s < best) { best = s; bj = (int32_t)(t0 + b); }
            }
        }
        if (q < nq) out[q] = bj;
    }
}

#include "mgicp_fgr_tc.cuh"

#ifndef MGICP_FGR_TU
#define MGICP_FGR_TU 8
#endif
constexpr int FGR_NT = 512, FGR_TU = MGICP_FGR_TU;

// x % n without the 64-bit division: q = floor(x * floor((2^64 - 1) / n) / 2^64) is floor(x / n) or up to 2 below it
__device__ __forceinline__ uint64_t fgr_mod(uint64_t x, uint64_t n, uint64_t barrett) {
    if (n <= 1) return 0;
    uint64_t r = x - __umul64hi(x, barrett) * n;
    while (r >= n) r -= n;
    return r;
}

// one block per pair: NormalizePointCloud, cross check, tuple test, OptimizePairwiseRegistration, original scale + inverse
__global__ void __launch_bounds__(FGR_NT) k_fgr_pair(FgrRunArgs A) {
    __shared__ double s_mean[2][3], s_max[2], s_red[FGR_NT / 32][27], s_tot[27], s_delta[16], s_trans[16], s_par;
    __shared__ double s_tile[2][3][FGR_NT];
    __shared__ int s_scan[33], s_count, s_stop;
    const FgrPair &pr = A.pairs[blockIdx.x];
    const int64_t n[2] = {A.cloud_off[pr.src + 1] - A.cloud_off[pr.src], A.cloud_off[pr.tgt + 1] - A.cloud_off[pr.tgt]};
    const int64_t off[2] = {A.cloud_off[pr.src], A.cloud_off[pr.tgt]};
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#ifdef FGR_PHASES
    long long ph[8], gph[3] = {0, 0, 0}; int phn = 0;
#define FGR_STAMP() do { if (tid == 0 && blockIdx.x == 0) ph[phn++] = clock64(); } while (0)
#else
#define FGR_STAMP() do {} while (0)
#endif
    FGR_STAMP();
    // ---- NormalizePointCloud: the means are summed sequentially (one thread per coordinate), like Open3D's loop, so that
    // the centred points -- and with them every discrete decision of the tuple test -- equal the oracle's bit for bit
    // (tiles of FGR_NT points of both clouds staged in shared memory by the whole block; lanes 0-2 / 3-5 of warp 0 add them up in order)
    {
        const int c = tid / 3, k = tid % 3;
        double m = 0.0;
        const int64_t nmax = n[0] > n[1] ? n[0] : n[1];
        for (int64_t i0 = 0; i0 < nmax; i0 += FGR_NT) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
                if (i0 + tid < n[cc]) {
                    double x, y, z;
                    load_point(A.xyz, A.dtype, off[cc] + i0 + tid, x, y, z);
                    s_tile[cc][0][tid] = x; s_tile[cc][1][tid] = y; s_tile[cc][2][tid] = z;
                }
            __syncthreads();
            if (tid < 6) {
                const int64_t left = n[c] - i0;
                const int cnt = left < FGR_NT ? (left > 0 ? (int)left : 0) : FGR_NT;
                const double *col = s_tile[c][k];
                for (int i = 0; i < cnt; ++i) m += col[i];
            }
            __syncthreads();
        }
        if (tid < 6) s_mean[c][k] = m / (double)n[c];
    }
    if (tid < 2) s_max[tid] = 0.0;
    __syncthreads();
    FGR_STAMP();
    for (int c = 0; c < 2; ++c) {
        double mx = 0.0;
        for (int64_t i = tid; i < n[c]; i += FGR_NT) {
            double x, y, z;
            load_point(A.xyz, A.dtype, off[c] + i, x, y, z);
            const V3 p = v3(x - s_mean[c][0], y - s_mean[c][1], z - s_mean[c][2]);
            pr.P[c][i] = p;
            mx = fmax(mx, sqrt(p.x * p.x + p.y * p.y + p.z * p.z));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned long long *>(&s_max[c]), (unsigned long long)__double_as_longlong(mx));   // mx >= 0
    }
    __syncthreads();
    const double scale = fmax(s_max[0], s_max[1]);
    const double scale_global = A.use_absolute_scale ? 1.0 : scale, scale_start = A.use_absolute_scale ? scale : 1.0;
    for (int c = 0; c < 2; ++c)
        for (int64_t i = tid; i < n[c]; i += FGR_NT) {
            const V3 p = pr.P[c][i];
            pr.P[c][i] = v3(p.x / scale_global, p.y / scale_global, p.z / scale_global);
        }
    if (tid == 0) s_count = 0;
    __syncthreads();
    FGR_STAMP();
    // ---- cross check: (i, j) with nn_j(i) = j and nn_i(j) = i, in ascending i (ordered compaction, 512 at a time)
    for (int64_t i0 = 0; i0 < pr.ni; i0 += FGR_NT) {
        const int64_t i = i0 + tid;
        int32_t j = -1;
        if (i < pr.ni) { j = pr.i2j[i]; if (j >= 0 && pr.j2i[j] != (int32_t)i) j = -1; }
        int total;
        const int pos = block_excl_scan(j >= 0 ? 1 : 0, s_scan, &total);
        const int base = s_count;
        if (j >= 0) { pr.cross[2 * (size_t)(base + pos)] = (int32_t)i; pr.cross[2 * (size_t)(base + pos) + 1] = j; }
        __syncthreads();
        if (tid == 0) s_count = base + total;
        __syncthreads();
    }
    const int64_t ncross = s_count;
    __syncthreads();
    FGR_STAMP();
    // ---- tuple test: trial t draws outputs 3t, 3t+1, 3t+2 of the counter-based generator; accepted trials are kept in trial
    // order until maximum_tuple_count is reached (the sequential loop's break)
    const int cap_in = A.caps ? A.caps[blockIdx.x] : A.maximum_tuple_count;
    const int cap = cap_in > 0 ? cap_in : 0;
    const V3 *Pi = pr.P[pr.fi], *Pj = pr.P[pr.fj];
    const bool swapped = pr.fi == 1;
    const uint64_t seed = A.seeds[blockIdx.x];
    const uint64_t barrett = ncross > 1 ? ~0ull / (uint64_t)ncross : 0;      // floor((2^64 - 1) / n)
    if (tid == 0) { s_count = 0; s_stop = (ncross == 0 || cap == 0) ? 1 : 0; }
    __syncthreads();
    // (FGR_TU consecutive trials per thread and round: the gathers of a round overlap; the result does not depend on the round size)
    for (int64_t t0 = 0; t0 < ncross * 100 && !s_stop; t0 += FGR_NT * FGR_TU) {
        int32_t tri[FGR_TU][6];
        bool ok[FGR_TU];
        int mine = 0;
#pragma unroll
        for (int u = 0; u < FGR_TU; ++u) {
            const int64_t t = t0 + (int64_t)tid * FGR_TU + u;
            ok[u] = false;
#pragma unroll
            for (int k = 0; k < 6; ++k) tri[u][k] = 0;
            if (t < ncross * 100) {
                const int64_t r0 = fgr_mod(fgr_rng(seed, 3 * (uint64_t)t), (uint64_t)ncross, barrett),
                              r1 = fgr_mod(fgr_rng(seed, 3 * (uint64_t)t + 1), (uint64_t)ncross, barrett),
                              r2 = fgr_mod(fgr_rng(seed, 3 * (uint64_t)t + 2), (uint64_t)ncross, barrett);
                tri[u][0] = pr.cross[2 * r0]; tri[u][1] = pr.cross[2 * r0 + 1]; tri[u][2] = pr.cross[2 * r1]; tri[u][3] = pr.cross[2 * r1 + 1];
                tri[u][4] = pr.cross[2 * r2]; tri[u][5] = pr.cross[2 * r2 + 1];
            }
        }
#pragma unroll
        for (int u = 0; u < FGR_TU; ++u) {
            const int64_t t = t0 + (int64_t)tid * FGR_TU + u;
            if (t < ncross * 100)
                ok[u] = fgr_tuple_ok(Pi[tri[u][0]], Pi[tri[u][2]], Pi[tri[u][4]], Pj[tri[u][1]], Pj[tri[u][3]], Pj[tri[u][5]], A.tuple_scale);
            mine += ok[u] ? 1 : 0;
        }
        int total;
        int pos = block_excl_scan(mine, s_scan, &total);
        const int base = s_count;
#pragma unroll
        for (int u = 0; u < FGR_TU; ++u)
            if (ok[u]) {
                if (base + pos < cap) {
                    int32_t *c = pr.cor + 6 * (size_t)(base + pos);
#pragma unroll
                    for (int k = 0; k < 3; ++k) { c[2 * k] = swapped ? tri[u][2 * k + 1] : tri[u][2 * k]; c[2 * k + 1] = swapped ? tri[u][2 * k] : tri[u][2 * k + 1]; }
                }
                ++pos;
            }
        __syncthreads();
        if (tid == 0) { s_count = min(base + total, cap); if (s_count >= cap) s_stop = 1; }
        __syncthreads();
    }
    const int64_t nc = 3 * (int64_t)s_count;
    FGR_STAMP();
    // ---- OptimizePairwiseRegistration: moves the copy Q of the target onto the source
    if (tid < 16) s_trans[tid] = s_delta[tid] = (tid % 5 == 0) ? 1.0 : 0.0;
    if (tid == 0) s_par = scale_start;
    // Open3D transforms the whole moving copy of the target every iteration; only the points of the correspondences are read, so
    // every correspondence carries its own copy (same operations on the same values: same bits), transformed by the previous
    // iteration's update right before it is used -- coalesced, no gathers, no separate sweep
    for (int64_t c = tid; c < nc; c += FGR_NT) { pr.Pc[c] = pr.P[0][pr.cor[2 * c]]; pr.Qc[c] = pr.P[1][pr.cor[2 * c + 1]]; }
    __syncthreads();
    if (nc >= 10) {
        for (int itr = 0; itr < A.iteration_number; ++itr) {
            const double par = s_par;
            double acc[27];
#ifdef FGR_PHASES
            const long long g0 = clock64();
#endif
#pragma unroll
            for (int a = 0; a < 27; ++a) acc[a] = 0.0;
            {
                // the loads of the next correspondence are issued before the arithmetic of this one (the store to Qc would
                // otherwise keep the compiler from hoisting them)
                double d[16];
#pragma unroll
                for (int a = 0; a < 16; ++a) d[a] = s_delta[a];
                const V3 *__restrict__ Pc = pr.Pc;
                V3 *__restrict__ Qc = pr.Qc;
                int64_t c = tid;
                V3 pn = v3(0.0, 0.0, 0.0), qn = pn;
                if (c < nc) { pn = Pc[c]; qn = Qc[c]; }
                for (; c < nc; c += FGR_NT) {
                    const V3 p = pn;
                    V3 q = qn;
                    const int64_t c2 = c + FGR_NT;
                    if (c2 < nc) { pn = Pc[c2]; qn = Qc[c2]; }
                    if (itr > 0) { q = transform_point(d, q); Qc[c] = q; }
                    fgr_accumulate(p, q, par, acc);
                }
            }
#ifdef FGR_PHASES
            const long long g1 = clock64();
#endif
            // deterministic block reduction: shuffle tree, then the warps in order
#pragma unroll
            for (int a = 0; a < 27; ++a) {
                double v = acc[a];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0) s_red[w][a] = v;
            }
            __syncthreads();
#ifdef FGR_PHASES
            const long long g2 = clock64();
#endif
            if (w == 0) {
                if (lane < 27) {
                    double s_ = 0.0;
                    for (int i = 0; i < FGR_NT / 32; ++i) s_ += s_red[i][lane];
                    s_tot[lane] = s_;
                }
                __syncwarp();
                if (lane == 0) {
                    double sums[27], x[6], delta[16], tn[16];
                    for (int a = 0; a < 27; ++a) sums[a] = s_tot[a];
                    ldlt_solve6(sums, x);                       // JTJ x = -JTr  ==  SolveLinearSystemPSD(-JTJ, JTr)
                    vec6_to_mat4(x, delta);
                    double told[16];
                    for (int a = 0; a < 16; ++a) told[a] = s_trans[a];
                    mat4_mul(delta, told, tn);
                    for (int a = 0; a < 16; ++a) { s_trans[a] = tn[a]; s_delta[a] = delta[a]; }
                    if (A.decrease_mu && itr % 4 == 0 && par > A.maximum_correspondence_distance) s_par = par / A.division_factor;
                }
            }
#ifdef FGR_PHASES
            if (tid == 0 && blockIdx.x == 0) { const long long g3 = clock64(); gph[0] += g1 - g0; gph[1] += g2 - g1; gph[2] += g3 - g2; }
#endif
            __syncthreads();
        }
#ifdef FGR_PHASES
        if (tid == 0 && blockIdx.x == 0) printf("gnc (kcycles): accumulate %lld reduce+sync %lld solve %lld\n", gph[0] / 1000, gph[1] / 1000, gph[2] / 1000);
#endif
    }
    FGR_STAMP();
#ifdef FGR_PHASES
    if (tid == 0 && blockIdx.x == 0)
        printf("k_fgr_pair phases (kcycles): means %lld normalise %lld cross %lld tuples %lld (ncross %lld, accepted %d) gnc %lld (nc %lld, n1 %lld)\n",
               (ph[1] - ph[0]) / 1000, (ph[2] - ph[1]) / 1000, (ph[3] - ph[2]) / 1000, (ph[4] - ph[3]) / 1000, (long long)ncross, s_count,
               (ph[5] - ph[4]) / 1000, (long long)nc, (long long)n[1]);
#endif
    if (tid == 0) {
        double tr[16], T[16];
        for (int a = 0; a < 16; ++a) tr[a] = s_trans[a];
        fgr_finalize(tr, s_mean[0], s_mean[1], scale_global, T);
        for (int a = 0; a < 16; ++a) A.T_out[16 * (size_t)blockIdx.x + a] = T[a];
        A.ncorr_out[blockIdx.x] = (int32_t)nc;
    }
}

extern "C" int mgicp_fgr_pairs(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                               int32_t xyz_dtype, const double *feat, int32_t n_pairs, const int32_t *pair_src,
                               const int32_t *pair_tgt, const mgicp_fgr_opts *o, const int32_t *tuple_counts, const uint64_t *seeds,
                               double *T_out, int32_t *ncorr_out) {
    if (!h) return MGICP_E_INVALID;
    if (n_clouds <= 0 || n_pairs <= 0 || !xyz || !cloud_off || !feat || !pair_src || !pair_tgt || !o || !seeds || !T_out || !ncorr_out ||
        (xyz_dtype != MGICP_F32 && xyz_dtype != MGICP_F64) || cloud_off[0] != 0 || !(o->division_factor > 1.0) ||
        !(o->tuple_scale > 0.0 && o->tuple_scale <= 1.0) || o->iteration_number < 0 || o->maximum_tuple_count < 0) {
        h->err = "mgicp_fgr_pairs: bad arguments"; return MGICP_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    // the scratch arena (the ICP scratch of mgicp_register_batch) holds everything: descriptors and clouds belong to the caller
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o_ = off; off += align_up(bytes); return o_; };
    const size_t o_pairs = take(sizeof(FgrPair) * n_pairs), o_coff = take(sizeof(int64_t) * (n_clouds + 1)), o_seed = take(sizeof(uint64_t) * n_pairs);
    const size_t o_caps = take(sizeof(int32_t) * n_pairs);
    std::vector<FgrPair> pairs(n_pairs);
    std::vector<size_t> offs((size_t)n_pairs * 8);
    int64_t max_q = 1;
    for (int p = 0; p < n_pairs; ++p) {
        if (tuple_counts && tuple_counts[p] < 0) { h->err = "mgicp_fgr_pairs: negative tuple count"; return MGICP_E_INVALID; }
        const size_t cap = (size_t)std::max(tuple_counts ? tuple_counts[p] : o->maximum_tuple_count, 1);
        const int s = pair_src[p], t = pair_tgt[p];
        if (s < 0 || s >= n_clouds || t < 0 || t >= n_clouds) { h->err = "pair index out of range"; return MGICP_E_INVALID; }
        const int64_t ns = cloud_off[s + 1] - cloud_off[s], nt = cloud_off[t + 1] - cloud_off[t];
        if (ns <= 0 || nt <= 0 || ns > (int64_t)1 << 30 || nt > (int64_t)1 << 30) { h->err = "mgicp_fgr_pairs: empty or oversized cloud"; return MGICP_E_INVALID; }
        FgrPair &pr = pairs[p];
        pr.src = s; pr.tgt = t;
        pr.fi = nt > ns ? 1 : 0; pr.fj = 1 - pr.fi;
        pr.ni = pr.fi == 0 ? ns : nt; pr.nj = pr.fi == 0 ? nt : ns;
        max_q = std::max(max_q, std::max(ns, nt));
        size_t *o_ = &offs[(size_t)p * 8];
        o_[0] = take(sizeof(V3) * ns); o_[1] = take(sizeof(V3) * nt); o_[2] = take(sizeof(V3) * 3 * cap); o_[7] = take(sizeof(V3) * 3 * cap);
        o_[3] = take(sizeof(int32_t) * pr.nj); o_[4] = take(sizeof(int32_t) * pr.ni);
        o_[5] = take(sizeof(int32_t) * 2 * std::min(pr.ni, pr.nj)); o_[6] = take(sizeof(int32_t) * 6 * cap);
    }
    int rc = grow(h, &h->scratch, &h->scratch_bytes, off);
    if (rc) return rc;
    char *base = h->scratch;
    for (int p = 0; p < n_pairs; ++p) {
        FgrPair &pr = pairs[p];
        size_t *o_ = &offs[(size_t)p * 8];
        pr.P[0] = (V3 *)(base + o_[0]); pr.P[1] = (V3 *)(base + o_[1]); pr.Pc = (V3 *)(base + o_[2]); pr.Qc = (V3 *)(base + o_[7]);
        pr.j2i = (int32_t *)(base + o_[3]); pr.i2j = (int32_t *)(base + o_[4]); pr.cross = (int32_t *)(base + o_[5]); pr.cor = (int32_t *)(base + o_[6]);
    }
    CK(cudaMemcpyAsync(base + o_pairs, pairs.data(), sizeof(FgrPair) * n_pairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_coff, cloud_off, sizeof(int64_t) * (n_clouds + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_seed, seeds, sizeof(uint64_t) * n_pairs, cudaMemcpyHostToDevice, st));
    if (tuple_counts) CK(cudaMemcpyAsync(base + o_caps, tuple_counts, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));            // `pairs` is a local vector: the copy must be done before it goes away
    FgrRunArgs A;
    A.xyz = xyz; A.dtype = xyz_dtype; A.cloud_off = (const int64_t *)(base + o_coff); A.feat = feat; A.pairs = (FgrPair *)(base + o_pairs);
    A.division_factor = o->division_factor; A.maximum_correspondence_distance = o->maximum_correspondence_distance; A.tuple_scale = o->tuple_scale;
    A.use_absolute_scale = o->use_absolute_scale; A.decrease_mu = o->decrease_mu; A.iteration_number = o->iteration_number;
    A.maximum_tuple_count = o->maximum_tuple_count; A.seeds = (const uint64_t *)(base + o_seed); A.caps = tuple_counts ? (const int32_t *)(base + o_caps) : nullptr; A.T_out = T_out; A.ncorr_out = ncorr_out;
    int match_mode = 1;                   // 1: tensor cores (tcgen05) + exact re-check, 0: brute-force fp64 on the CUDA cores
    if (const char *e = getenv("MGICP_FGR_MATCH")) match_mode = atoi(e);                                    // A/B experiments
    if (match_mode == 0) {
        k_fgr_nn<<<dim3(chunks_for(max_q, FGR_NN_NT, 4096), 2 * n_pairs), FGR_NN_NT, 0, st>>>(A, nullptr, 0);
        h->launches += 1;
    } else {
        // operands packed once per cloud (split fp16, canonical UMMA tiles), in the arena (free during FGR)
        size_t poff = 0;
        auto ptake = [&](size_t bytes) { size_t o_ = poff; poff += align_up(bytes, 1024); return o_; };
        std::vector<size_t> po((size_t)n_clouds * 4);
        for (int c = 0; c < n_clouds; ++c) {
            const int64_t n = cloud_off[c + 1] - cloud_off[c];
            const size_t npa = (size_t)(n + fgrtc::TM - 1) / fgrtc::TM * fgrtc::TM, npb = (size_t)(n + fgrtc::TN - 1) / fgrtc::TN * fgrtc::TN;
            po[4 * c] = ptake(npa * fgrtc::KP * 2); po[4 * c + 1] = ptake(npb * fgrtc::KP * 2); po[4 * c + 2] = ptake(npb * 4); po[4 * c + 3] = ptake(16);
        }
        const size_t o_cp = ptake(sizeof(fgrtc::CloudPack) * n_clouds), o_ptr = ptake(sizeof(void *) * 4 * n_clouds);
        const size_t o_fbc = ptake(sizeof(int32_t) * 2 * n_pairs), o_fbl = ptake(sizeof(int32_t) * 2 * (size_t)n_pairs * max_q);
        const size_t o_part = ptake(sizeof(fgrtc::FbPart) * 2 * (size_t)n_pairs * fgrtc::FB_CAP * fgrtc::FB_SLICES);
        const size_t o_seedfb = ptake(sizeof(fgrtc::FbPart) * 2 * (size_t)n_pairs * fgrtc::FB_CAP);
        rc = grow(h, &h->arena, &h->arena_bytes, poff);
        if (rc) return rc;
        h->preprocessed = false;          // the arena is reused: a previous mgicp_preprocess is gone after this call
        h->eval_jobs = nullptr;
        char *pb = h->arena;
        std::vector<fgrtc::CloudPack> cps(n_clouds);
        std::vector<void *> ptrs((size_t)4 * n_clouds);
        for (int c = 0; c < n_clouds; ++c) {
            cps[c].PA = (const __half *)(pb + po[4 * c]); cps[c].PB = (const __half *)(pb + po[4 * c + 1]);
            cps[c].nbf = (const float *)(pb + po[4 * c + 2]); cps[c].nmax = (const float *)(pb + po[4 * c + 3]);
            cps[c].n = (int32_t)(cloud_off[c + 1] - cloud_off[c]);
            ptrs[c] = pb + po[4 * c]; ptrs[n_clouds + c] = pb + po[4 * c + 1]; ptrs[2 * n_clouds + c] = pb + po[4 * c + 2]; ptrs[3 * n_clouds + c] = pb + po[4 * c + 3];
            CK(cudaMemsetAsync(pb + po[4 * c + 3], 0, 16, st));
        }
        CK(cudaMemcpyAsync(pb + o_cp, cps.data(), sizeof(fgrtc::CloudPack) * n_clouds, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(pb + o_ptr, ptrs.data(), sizeof(void *) * 4 * n_clouds, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(pb + o_fbc, 0, sizeof(int32_t) * 2 * n_pairs, st));
        CK(cudaStreamSynchronize(st));        // cps / ptrs are local vectors
        fgrtc::PackArgs PA_;
        PA_.feat = feat; PA_.cloud_off = A.cloud_off;
        PA_.PA = (__half *const *)(pb + o_ptr); PA_.PB = PA_.PA + n_clouds;
        PA_.nbf = (float *const *)(pb + o_ptr + sizeof(void *) * 2 * n_clouds); PA_.nmax = PA_.nbf + n_clouds;
        fgrtc::MatchArgs MA;
        MA.packs = (const fgrtc::CloudPack *)(pb + o_cp); MA.pairs = A.pairs; MA.feat = feat; MA.cloud_off = A.cloud_off;
        MA.fb_list = (int32_t *)(pb + o_fbl); MA.fb_count = (int32_t *)(pb + o_fbc); MA.fb_stride = max_q; MA.fb_seed = (fgrtc::FbPart *)(pb + o_seedfb);
        CK(cudaFuncSetAttribute(fgrtc::k_fgr_match_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fgrtc::SMEM));
        fgrtc::k_fgr_pack<<<dim3(chunks_for(max_q * (fgrtc::KP / 8), 256, 1024), n_clouds), 256, 0, st>>>(PA_);
        fgrtc::k_fgr_match_tc<<<dim3((unsigned)((max_q + fgrtc::TM - 1) / fgrtc::TM), 2 * n_pairs), fgrtc::NT, fgrtc::SMEM, st>>>(MA);
        fgrtc::k_fgr_match_fb<<<dim3(fgrtc::FB_SLICES, 2 * n_pairs), fgrtc::FB_NT, 0, st>>>(MA, (fgrtc::FbPart *)(pb + o_part));
        fgrtc::k_fgr_match_fb2<<<dim3(8, 2 * n_pairs), 128, 0, st>>>(MA, (const fgrtc::FbPart *)(pb + o_part));
        // more unsettled rows than the queue holds (thousands of duplicate descriptors): the plain brute force for that direction
        k_fgr_nn<<<dim3(chunks_for(max_q, FGR_NN_NT, 4096), 2 * n_pairs), FGR_NN_NT, 0, st>>>(A, MA.fb_count, fgrtc::FB_CAP);
        h->launches += 5;
        if (getenv("MGICP_FGR_STATS")) {                          // diagnostics: rows the tensor-core pass could not settle
            std::vector<int32_t> fc((size_t)2 * n_pairs);
            CK(cudaMemcpyAsync(fc.data(), pb + o_fbc, sizeof(int32_t) * 2 * n_pairs, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            long long tot = 0; int mx = 0;
            for (int32_t v : fc) { tot += v; mx = std::max(mx, (int)v); }
            fprintf(stderr, "mgicp_fgr_pairs: %d directions, rows queued for the fp64 search: %lld (mean %.1f, max %d, cap %d)\n", 2 * n_pairs, tot,
                    (double)tot / (2 * n_pairs), mx, fgrtc::FB_CAP);
        }
    }
    k_fgr_pair<<<n_pairs, FGR_NT, 0, st>>>(A);
    h->launches += 1;
    CK(cudaGetLastError());
    return MGICP_OK;
}
