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
__global__ void __launch_bounds__(128) k_hybrid_lists(FgrArgs A) {
    const Job &J = A.jobs[blockIdx.y];
    if (J.err) return;
    const GridView g = make_view(J, 2);
    const int64_t base = A.cloud_off[blockIdx.y];
    const int cap = A.cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
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
    const size_t o_jobs = take(sizeof(Job) * n_clouds), o_benc = take(sizeof(u64) * 6 * n_clouds), o_coff = take(sizeof(int64_t) * (n_clouds + 1));
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
    k_bounds_init<<<(n_clouds * 6 + 127) / 128, 128, 0, st>>>(benc, n_clouds);
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
    k_hybrid_lists<<<grid, 128, 0, st>>>(A);
    k_hybrid_normals<<<grid, 128, 0, st>>>(A);
    A.r2 = radius_fpfh * radius_fpfh; A.cap = max_nn_fpfh;
    A.idx = (int32_t *)(base + o_if); A.d2 = (double *)(base + o_df); A.cnt = (int32_t *)(base + o_cf);
    k_hybrid_lists<<<grid, 128, 0, st>>>(A);
    k_spfh<<<grid, 128, 0, st>>>(A);
    k_fpfh<<<grid, 128, 0, st>>>(A);
    h->launches += 15;
    CK(cudaGetLastError());
    h->eval_jobs = jobs_dev; h->eval_n = n_clouds;
    return MGICP_OK;
}
