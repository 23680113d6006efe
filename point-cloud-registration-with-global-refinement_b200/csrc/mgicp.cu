// mgicp.cu -- B200 (sm_100a) kernels and the C ABI (include/mgicp.h) of the multiscale
// Generalized-ICP refinement engine.  The path replaced is the body of the reference's
// Multiscale_GICP (/root/reference/ALL_FUNCTIONS.py:286-312, 2_MGICP_refinement_in_NCLT_dataset.py:140-163):
//   K0 bounds -> K1 hashed voxel down-sample -> K2a statistical outlier removal (kNN 30)
//   -> K2b kNN(20) covariance + closed-form eigen normals -> K3+K4+K5 fused on-device ICP loop.
// All (cloud, scale) jobs of a batch go through each preprocessing kernel together
// (grid = chunks x jobs); the ICP loop of a pair, all scales and all iterations, is one launch.
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/mgicp.h"
#include "mgicp_device.cuh"

namespace cg = cooperative_groups;
using namespace mg;

// =============================================================================================
// K0: bounds
// =============================================================================================
__device__ __forceinline__ u64 enc_double(double d) {
    long long b = __double_as_longlong(d);
    return b < 0 ? ~(u64)b : ((u64)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_double(u64 e) {
    long long b = (e & 0x8000000000000000ull) ? (long long)(e & 0x7FFFFFFFFFFFFFFFull) : (long long)~e;
    return __longlong_as_double(b);
}

__device__ __forceinline__ void load_point(const void *xyz, int dtype, int64_t i, double &x, double &y, double &z) {
    if (dtype == MGICP_F32) {
        const float *p = reinterpret_cast<const float *>(xyz) + 3 * i;
        x = (double)__ldg(p); y = (double)__ldg(p + 1); z = (double)__ldg(p + 2);
    } else {
        const double *p = reinterpret_cast<const double *>(xyz) + 3 * i;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    }
}

// benc: per cloud 6 encoded bounds + 1 word that stays 1 while every coordinate seen is float32-representable
constexpr int BENC_W = 7;
__global__ void k_bounds_init(u64 *benc, int n_clouds) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_clouds * BENC_W) benc[i] = (i % BENC_W) == 6 ? 1ull : ((i % BENC_W) < 3 ? enc_double(INFINITY) : enc_double(-INFINITY));
}

// grid (chunks, clouds)
__global__ void __launch_bounds__(256) k_bounds(const void *xyz, int dtype, const int64_t *cloud_off, u64 *benc) {
    const int c = blockIdx.y;
    const int64_t lo = cloud_off[c], n = cloud_off[c + 1] - lo;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool f32ok = true;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        load_point(xyz, dtype, lo + i, x, y, z);
        mn[0] = fmin(mn[0], x); mn[1] = fmin(mn[1], y); mn[2] = fmin(mn[2], z);
        mx[0] = fmax(mx[0], x); mx[1] = fmax(mx[1], y); mx[2] = fmax(mx[2], z);
        f32ok &= x == (double)(float)x && y == (double)(float)y && z == (double)(float)z;
    }
    if (!f32ok) benc[c * BENC_W + 6] = 0ull;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fmin(mn[d], __shfl_down_sync(0xffffffffu, mn[d], o));
            mx[d] = fmax(mx[d], __shfl_down_sync(0xffffffffu, mx[d], o));
        }
    }
    if ((threadIdx.x & 31) == 0 && n > 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&benc[c * BENC_W + d], enc_double(mn[d]));
            atomicMax(&benc[c * BENC_W + 3 + d], enc_double(mx[d]));
        }
    }
}

__global__ void k_bounds_decode(const u64 *benc, double *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = dec_double(benc[(i / 6) * BENC_W + i % 6]);
}

// one thread per job: grid origin, grid dimensions, range checks
__global__ void k_job_setup(Job *jobs, int n_jobs, const u64 *benc) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    Job &J = jobs[j];
    J.M = 0; J.Mf = 0; J.fb_count = 0; J.err = ERR_NONE; J.ordered = (J.dtype == MGICP_F64 && benc[J.cloud * BENC_W + 6] == 0ull) ? 1 : 0; J.cbits = 10; J.ibits = 10; J.sor_thresh = 0.0;
    J.gdim[0] = J.gdim[1] = J.gdim[2] = 0;
    J.idim[0] = J.idim[1] = J.idim[2] = 0;
    if (J.n <= 0) { J.org[0] = J.org[1] = J.org[2] = 0.0; return; }
    for (int d = 0; d < 3; ++d) {
        double mn = dec_double(benc[J.cloud * BENC_W + d]), mx = dec_double(benc[J.cloud * BENC_W + 3 + d]);
        double org = mn - J.voxel * 0.5;            // Open3D: voxel_min_bound = min_bound - voxel_size * 0.5
        J.org[d] = org;
        double nv = floor((mx - org) / J.voxel);
        double nc = floor((mx - org) / J.cell), ni = floor((mx - org) / J.cell_i);
        if (!(nv < (double)(COORD_LIMIT - 1)) || !(nc < (double)(COORD_LIMIT - 1)) || !(ni < (double)(COORD_LIMIT - 1))) {
            J.err = ERR_RANGE; nc = 0; ni = 0;
        }
        J.gdim[d] = (int)nc + 1;
        J.idim[d] = (int)ni + 1;
    }
}

// =============================================================================================
// K1: sort-free hashed-grid voxel down-sample
// =============================================================================================
// grid (chunks, jobs): insert every point's voxel key into the job's ordered hash table
__global__ void __launch_bounds__(256) k_vox_insert(Job *jobs) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const int64_t n = J.n;
    const double ox = J.org[0], oy = J.org[1], oz = J.org[2], v = J.voxel;
    bool ok = true;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        load_point(J.xyz, J.dtype, i, x, y, z);
        int ix = (int)floor((x - ox) / v), iy = (int)floor((y - oy) / v), iz = (int)floor((z - oz) / v);
        ok &= ordered_insert<1>(J.vkeys, J.vbits, pack_key(ix, iy, iz));
    }
    if (!ok) J.err = ERR_OVERFLOW;
}

// one CTA per job: rank the occupied slots (canonical voxel ids), size and clear the cell table,
// zero the accumulators
__global__ void __launch_bounds__(1024) k_vox_scan(Job *jobs) {
    __shared__ int sm[33];
    Job &J = jobs[blockIdx.x];
    if (J.err || J.n <= 0) return;
    const int cap = (1 << J.vbits) + TAB_PAD;
    const u64 *vk = J.vkeys;
    int32_t *vr = J.vrank;
    const int M = block_region_scan(cap, sm, [&](int i) { return vk[i] != EMPTY_KEY ? 1 : 0; },
                                    [&](int i, int pre, int v) { if (v) vr[i] = pre; });
    int cbits = 10;
    while (cbits < J.cbits_max && (1 << cbits) < 4 * M) ++cbits;
    if (threadIdx.x == 0) { J.M = M; J.cbits = cbits; }
}

// grid (chunks, jobs): clear the cell table about to be filled (which = 0: kNN grid, 2: ICP grid), its scatter cursors, and,
// for the kNN grid, the per-voxel accumulators
__global__ void __launch_bounds__(256) k_table_clear(Job *jobs, int which) {
    Job &J = jobs[blockIdx.y];
    if (J.err || J.n <= 0) return;
    CellSlot *tab = which == 0 ? J.ctab : J.itab;
    const int cap = (1 << (which == 0 ? J.cbits : J.ibits)) + TAB_PAD;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    CellSlot e; e.key = EMPTY_KEY; e.start = 0; e.count = 0;
    for (int i = t0; i < cap; i += nt) { tab[i] = e; J.ccursor[i] = 0; }
    if (which == 0) {
        const int M = J.M;
        for (int i = t0; i < 3 * M; i += nt) J.vsum[i] = 0.0;
        for (int i = t0; i < M; i += nt) J.vcnt[i] = 0;
    }
}

// grid (chunks, jobs): the kNN grid of the surviving points = the same cells with their ranges mapped through newidx
__global__ void __launch_bounds__(256) k_ftab_build(Job *jobs) {
    Job &J = jobs[blockIdx.y];
    if (J.err || J.n <= 0) return;
    const int cap = (1 << J.cbits) + TAB_PAD;
    for (int sI = blockIdx.x * blockDim.x + threadIdx.x; sI < cap; sI += gridDim.x * blockDim.x) {
        CellSlot e = J.ctab[sI];
        if (e.key != EMPTY_KEY) {
            int a = J.newidx[e.start], b = J.newidx[e.start + e.count];
            e.start = a; e.count = b - a;
        }
        J.ftab[sI] = e;
    }
}

// grid (chunks, jobs): accumulate coordinate sums per voxel.  fp64 sums of float32-sourced coordinates are exact
// (SURVEY App. B), so neither the atomics nor the association below make the result order dependent for PCD inputs.
// Consecutive points of a scan mostly fall into the same voxel: every run of equal voxels inside a warp is summed with a
// segmented shuffle scan and only the last lane of the run touches memory (4 atomics per run instead of per point).
__global__ void __launch_bounds__(256) k_vox_accum(Job *jobs) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const int64_t n = J.n;
    const double ox = J.org[0], oy = J.org[1], oz = J.org[2], v = J.voxel;
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {   // warp-uniform
        const int64_t i = i0 + lane;
        double x = 0.0, y = 0.0, z = 0.0;
        int r = -1, c = 0;
        if (i < n) {
            load_point(J.xyz, J.dtype, i, x, y, z);
            int ix = (int)floor((x - ox) / v), iy = (int)floor((y - oy) / v), iz = (int)floor((z - oz) / v);
            int slot = ordered_find<1>(J.vkeys, J.vbits, pack_key(ix, iy, iz));
            if (slot < 0) J.err = ERR_OVERFLOW;
            else { r = J.vrank[slot]; c = 1; }
        }
        const int rprev = __shfl_up_sync(0xffffffffu, r, 1);
        const bool head = lane == 0 || rprev != r;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const int dist = lane - (31 - __clz(heads & (0xffffffffu >> (31 - lane))));     // lanes since the head of this run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double tx = __shfl_up_sync(0xffffffffu, x, o), ty = __shfl_up_sync(0xffffffffu, y, o), tz = __shfl_up_sync(0xffffffffu, z, o);
            const int tc = __shfl_up_sync(0xffffffffu, c, o);
            if (dist >= o) { x += tx; y += ty; z += tz; c += tc; }
        }
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (tail && r >= 0) {
            atomicAdd(&J.vsum[3 * r + 0], x);
            atomicAdd(&J.vsum[3 * r + 1], y);
            atomicAdd(&J.vsum[3 * r + 2], z);
            atomicAdd(&J.vcnt[r], c);
        }
    }
}

// ---- ordered voxel sums for genuine float64 clouds -----------------------------------------------------------------------
// fp64 atomics add the coordinates of a voxel in whatever order the threads arrive: exact (hence order-free) for
// float32-sourced clouds such as the reference's PCD files, last-bit order-dependent for coordinates that need all 53 bits --
// and under the L1 kernel a last bit forks the pose.  For MGICP_F64 clouds with at least one coordinate that is not float32-
// representable (k_bounds looks at every coordinate anyway) the sums are therefore redone in INPUT ORDER, the
// order in which Open3D's VoxelDownSample adds the points of a voxel (AccumulatedPoint::AddPoint in a loop over the cloud):
// counting sort of the point indices by voxel (counts from k_vox_accum), each voxel's list put in ascending order, one
// sequential sum per voxel.  Scratch: buffers that are not in use yet at this stage (pslot, order, newidx, avg).
// one CTA per job: list offsets of the voxels, cursors cleared
__global__ void __launch_bounds__(1024) k_vox_ord_scan(Job *jobs) {
    __shared__ int sm[33];
    Job &J = jobs[blockIdx.x];
    if (J.err || J.n <= 0 || !J.ordered) return;
    const int M = J.M;
    const int32_t *cnt = J.vcnt;
    int32_t *vstart = J.newidx, *vcur = reinterpret_cast<int32_t *>(J.avg);
    block_region_scan(M, sm, [&](int i) { return cnt[i]; }, [&](int i, int pre, int) { vstart[i] = pre; vcur[i] = 0; });
}
// grid (chunks, jobs): point indices into their voxel's list (order inside a list fixed by k_vox_ord_sum)
__global__ void __launch_bounds__(256) k_vox_ord_scatter(Job *jobs) {
    Job &J = jobs[blockIdx.y];
    if (J.err || !J.ordered) return;
    const int64_t n = J.n;
    const double ox = J.org[0], oy = J.org[1], oz = J.org[2], v = J.voxel;
    int32_t *vcur = reinterpret_cast<int32_t *>(J.avg);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        load_point(J.xyz, J.dtype, i, x, y, z);
        const int ix = (int)floor((x - ox) / v), iy = (int)floor((y - oy) / v), iz = (int)floor((z - oz) / v);
        const int slot = ordered_find<1>(J.vkeys, J.vbits, pack_key(ix, iy, iz));
        if (slot < 0) { J.err = ERR_OVERFLOW; continue; }
        const int r = J.vrank[slot];
        J.pslot[J.newidx[r] + atomicAdd(&vcur[r], 1)] = (int32_t)i;
    }
}
// grid (chunks, jobs), one warp per voxel: the list in ascending index order (rank = number of smaller indices), then the
// coordinates summed in that order by one lane
__global__ void __launch_bounds__(256) k_vox_ord_sum(Job *jobs) {
    Job &J = jobs[blockIdx.y];
    if (J.err || !J.ordered) return;
    const int M = J.M;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < M; r += nwarp) {
        const int s = J.newidx[r], c = J.vcnt[r];
        const int32_t *list = J.pslot + s;
        int32_t *sorted = J.order + s;
        for (int e = lane; e < c; e += 32) {
            const int idx = list[e];
            int rank = 0;
            for (int f = 0; f < c; ++f) rank += (list[f] < idx) ? 1 : 0;
            sorted[rank] = idx;
        }
        __syncwarp();
        if (lane == 0) {
            double sx = 0.0, sy = 0.0, sz = 0.0;
            for (int t = 0; t < c; ++t) {
                double x, y, z;
                load_point(J.xyz, J.dtype, sorted[t], x, y, z);
                sx += x; sy += y; sz += z;
            }
            J.vsum[3 * r] = sx; J.vsum[3 * r + 1] = sy; J.vsum[3 * r + 2] = sz;
        }
        __syncwarp();
    }
}

// grid (chunks, jobs): centroid = sum / count; insert the centroid's cell into the spatial hash
__global__ void __launch_bounds__(256) k_vox_final(Job *jobs) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const int M = J.M;
    bool ok = true;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M; r += gridDim.x * blockDim.x) {
        double c = (double)J.vcnt[r];
        double x = J.vsum[3 * r] / c, y = J.vsum[3 * r + 1] / c, z = J.vsum[3 * r + 2] / c;
        J.ds[3 * r] = x; J.ds[3 * r + 1] = y; J.ds[3 * r + 2] = z;
        u64 key = pack_key(cell_coord(x, J.org[0], J.cell), cell_coord(y, J.org[1], J.cell), cell_coord(z, J.org[2], J.cell));
        ok &= ordered_insert<2>(reinterpret_cast<u64 *>(J.ctab), J.cbits, key);
    }
    if (!ok) J.err = ERR_OVERFLOW;
}

// ---- spatial-hash build, shared by the kNN grid (which = 0, over ds[]) and the ICP grid (which = 2, over pts[]) ------
struct BuildView {
    CellSlot *tab; int bits; double cell; int n;
};
__device__ __forceinline__ BuildView build_view(Job &J, int which) {
    BuildView b;
    b.tab = which == 0 ? J.ctab : J.itab;
    b.bits = which == 0 ? J.cbits : J.ibits;
    b.cell = which == 0 ? J.cell : J.cell_i;
    b.n = which == 0 ? J.M : J.Mf;
    return b;
}
__device__ __forceinline__ void build_point(const Job &J, int which, int r, double &x, double &y, double &z) {
    if (which == 0) { x = J.ds[3 * r]; y = J.ds[3 * r + 1]; z = J.ds[3 * r + 2]; }
    else { const double4 p = J.pts[r]; x = p.x; y = p.y; z = p.z; }
}

// grid (chunks, jobs): insert the cells of the final cloud into the ICP grid
__global__ void __launch_bounds__(256) k_cell_insert(Job *jobs, int which) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const BuildView b = build_view(J, which);
    bool ok = true;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n; r += gridDim.x * blockDim.x) {
        double x, y, z;
        build_point(J, which, r, x, y, z);
        u64 key = pack_key(cell_coord(x, J.org[0], b.cell), cell_coord(y, J.org[1], b.cell), cell_coord(z, J.org[2], b.cell));
        ok &= ordered_insert<2>(reinterpret_cast<u64 *>(b.tab), b.bits, key);
    }
    if (!ok) J.err = ERR_OVERFLOW;
}

// grid (chunks, jobs): count points per cell, remember each point's slot
__global__ void __launch_bounds__(256) k_cell_count(Job *jobs, int which) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const BuildView b = build_view(J, which);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n; r += gridDim.x * blockDim.x) {
        double x, y, z;
        build_point(J, which, r, x, y, z);
        u64 key = pack_key(cell_coord(x, J.org[0], b.cell), cell_coord(y, J.org[1], b.cell), cell_coord(z, J.org[2], b.cell));
        int slot = ordered_find<2>(reinterpret_cast<const u64 *>(b.tab), b.bits, key);
        if (slot < 0) { J.err = ERR_OVERFLOW; J.pslot[r] = 0; continue; }
        J.pslot[r] = slot;
        atomicAdd(&b.tab[slot].count, 1);
    }
}

// one CTA per job: exclusive scan of the per-slot counts -> cell start offsets
__global__ void __launch_bounds__(1024) k_cell_scan(Job *jobs, int which) {
    __shared__ int sm[33];
    Job &J = jobs[blockIdx.x];
    if (J.err || J.n <= 0) return;
    const BuildView b = build_view(J, which);
    const int cap = (1 << b.bits) + TAB_PAD;
    CellSlot *tab = b.tab;
    block_region_scan(cap, sm, [&](int i) { return tab[i].count; }, [&](int i, int pre, int v) { tab[i].start = pre; });
}

// grid (chunks, jobs): scatter point ids into their cell's range (order inside a cell fixed later)
__global__ void __launch_bounds__(256) k_cell_scatter(Job *jobs, int which) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const BuildView b = build_view(J, which);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n; r += gridDim.x * blockDim.x) {
        int slot = J.pslot[r];
        int pos = b.tab[slot].start + atomicAdd(&J.ccursor[slot], 1);
        J.order[pos] = r;
    }
}

// grid (chunks, jobs): per occupied cell, order its points by ascending id (deterministic) and gather them.
// A warp scans 32 slots at a time and handles each occupied one cooperatively: rank of a point = number of smaller ids.
__global__ void __launch_bounds__(256) k_cell_gather(Job *jobs, int which) {
    Job &J = jobs[blockIdx.y];
    if (J.err || J.n <= 0) return;
    const BuildView b = build_view(J, which);
    const int cap = (1 << b.bits) + TAB_PAD;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int s0 = warp * 32; s0 < cap; s0 += nwarp * 32) {
        const int sl = s0 + lane;
        int cnt = 0, st = 0;
        if (sl < cap) { cnt = b.tab[sl].count; st = b.tab[sl].start; }
        unsigned occ = __ballot_sync(0xffffffffu, cnt > 0);
        while (occ) {
            const int src = __ffs(occ) - 1;
            occ &= occ - 1;
            const int c = __shfl_sync(0xffffffffu, cnt, src), s = __shfl_sync(0xffffffffu, st, src);
            const int *o = J.order + s;
            for (int e = lane; e < c; e += 32) {
                const int r = o[e];
                int rank = 0;
                for (int f = 0; f < c; ++f) rank += (o[f] < r) ? 1 : 0;
                if (which == 0) J.gpts[s + rank] = make_double4(J.ds[3 * r], J.ds[3 * r + 1], J.ds[3 * r + 2], (double)r);
                else {
                    double4 p = J.pts[r];
                    const double4 nn = J.nrm[r];
                    p.w = nn.w;                               // safe radius^2 rides with the point (one 32-byte load in the seed check)
                    J.ipts[s + rank] = p;
                    J.inrm[s + rank] = nn;
                    J.a2i[r] = s + rank;
                    J.i2a[s + rank] = r;
                }
            }
        }
    }
}

// =============================================================================================
// K2a / K2b: exact grid kNN -> mean neighbour distance (outlier filter) / covariance + normal
// =============================================================================================
// ---- outlier-filter kNN (k = sor_k <= 32) over the down-sampled cloud --------------------------------------------------
// One warp per query, one candidate per lane (every distance is computed once), but no running top-k list: maintaining the
// sorted list (ballot + shuffle insertions, ~100 per query) was most of the instructions of k_knn.  Selection by histogram:
//   pass 1  ring by ring (own cell, then the 26 cells around it, then the shells of rings 2 and 3 in the sparse far field of
//           a LiDAR scan): the cells of a shell are looked up one per lane, their points packed densely over the lanes, and
//           every candidate's squared distance is dropped into a per-warp histogram in shared memory with log-spaced bins
//           (float exponent + 3 mantissa bits: 8 bins per octave, scale-free); the candidate's index and bin are kept in
//           shared memory.  The bin b* at which the cumulative count reaches k bounds the k-th distance from above: once the
//           own cell holds k points, shell cells farther than that bound are not even looked at, and the search ends when
//           the upper edge of b* lies inside the examined cells;
//   pass 2  the cached (index, bin) records with bin <= b* (typically k + 1 of them) are placed in a list at the positions the
//           histogram's prefix sum reserves for their bin (counting sort), their distances recomputed;
//   sort    rank of an entry = start of its bin + the entries of the same bin that precede it in (d^2, index) order: the k
//           nearest in ascending order, exactly the list knn_warp returns (same candidates, same order, same tie-break).
// Exactness does not depend on the bins: every candidate with d^2 below the edge of b* is collected, that edge is at least
// the k-th distance, and the cells examined (or skipped with proof) cover the ball of that radius.  Queries that need more than
// three rings, more than KH_NC candidates or a longer list than KH_CAP take the knn_warp path in place.
// (Tried and dropped, all bit-identical but slower on the B200: one THREAD per query with per-thread histograms and lists
// in shared memory -- 15 of 32 lanes active, 12 warps per SM; one warp per CELL sharing the flattened candidate list among the
// cell's queries, with and without the candidates staged in shared memory -- 1450-1500 instead of 2350 instructions per query
// but 9-13 warps per SM and a dependent chain per query: 25 % of the issue slots against 63 % here.)
#ifndef MGICP_KH_BLOCKS
#define MGICP_KH_BLOCKS 4      // 64 registers, four blocks per SM: 3 % more pairs/s over the whole step than 80 registers / three blocks
#endif
constexpr int KH_WARPS = 8;      // warps (queries in flight) per block
constexpr int KH_NC = 1024;      // cached candidates per query
constexpr int KH_CAP = 40;       // list entries per query (k <= 32 plus the rest of the last bin)
constexpr int KH_NB = 96;        // histogram bins: 12 octaves of d^2 below the largest certifiable radius, 3 bins per lane
constexpr int KH_RMAX = 3;
struct __align__(16) KhWarp {
    unsigned hist[KH_NB];        // counts; after the selection: list cursor of every bin
    int bstart[KH_NB];           // first list position of every bin
    int cidx[KH_NC];
    unsigned char cbin[KH_NC];
    double key[KH_CAP];
    int idx[KH_CAP];
    unsigned char ebin[KH_CAP];
    double skey[KH_CAP];         // sorted
    int sidx[KH_CAP];
};

__device__ __forceinline__ int kh_bin(const double d2, const int base) {
    const int b = (__float_as_int(__double2float_rn(d2)) >> 20) - base;
    return max(0, min(KH_NB - 1, b));
}

// cumulative histogram across the warp: lane l owns bins 3l, 3l+1, 3l+2.  Returns the smallest bin at which the cumulative
// count reaches k (KH_NB - 1 if it never does), cum_at = cumulative count at that bin (the total if it never does);
// ex = exclusive prefix of the lane's first bin, c0..c2 its counts.
__device__ __forceinline__ int kh_select(const KhWarp &W, const int k, const int lane, int &cum_at, int &ex, int &c0, int &c1, int &c2) {
    c0 = (int)W.hist[3 * lane]; c1 = (int)W.hist[3 * lane + 1]; c2 = (int)W.hist[3 * lane + 2];
    const int sum = c0 + c1 + c2;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    ex = incl - sum;
    const unsigned m = __ballot_sync(FULL, incl >= k);
    if (m == 0u) { cum_at = __shfl_sync(FULL, incl, 31); return KH_NB - 1; }
    const int src = __ffs(m) - 1;
    int bs = 3 * lane, cu = ex + c0;                 // the crossing lane: which of its three bins
    if (cu < k) { bs += 1; cu += c1; if (cu < k) { bs += 1; cu += c2; } }
    cum_at = __shfl_sync(FULL, cu, src);
    return __shfl_sync(FULL, bs, src);
}

// grid (chunks, jobs), one warp per query
__global__ void __launch_bounds__(KH_WARPS * 32, MGICP_KH_BLOCKS) k_knn_hist(Job *jobs, int k) {
    extern __shared__ __align__(16) unsigned char kh_raw[];
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const GridView g = make_view(J, 0);
    const int lane = threadIdx.x & 31;
    KhWarp &W = reinterpret_cast<KhWarp *>(kh_raw)[threadIdx.x >> 5];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    // bin KH_NB - 2 holds the largest squared radius that can be certified (ring KH_RMAX), KH_NB - 1 everything beyond
    const double top = (KH_RMAX + 1.0) * g.cell;
    const int base = (__float_as_int(__double2float_rn(top * top)) >> 20) - (KH_NB - 2);
    const double cell = g.cell, slack = CELL_SLACK * g.cell, inv_cell = 1.0 / g.cell;
    for (int i = warp; i < g.n; i += nwarp) {
        const double4 p = ldg4(g.pts + i);
        const double px = p.x, py = p.y, pz = p.z;
        // The cell the walk starts from.  It only has to be the query's cell up to rounding (the multiplication by the reciprocal
        // may put a point that sits on a cell face one cell off): every face distance below is measured from THIS cell's box, a
        // query slightly outside it gets a (slightly) negative distance, i.e. a smaller certified radius, never a larger one.
        const int cx = min(max((int)floor((px - g.org[0]) * inv_cell), 0), g.dim[0] - 1);
        const int cy = min(max((int)floor((py - g.org[1]) * inv_cell), 0), g.dim[1] - 1);
        const int cz = min(max((int)floor((pz - g.org[2]) * inv_cell), 0), g.dim[2] - 1);
        const double bx = g.org[0] + (double)cx * cell, by = g.org[1] + (double)cy * cell, bz = g.org[2] + (double)cz * cell;
        const double flx = px - bx - slack, fhx = (bx + cell) - px - slack;
        const double fly = py - by - slack, fhy = (by + cell) - py - slack;
        const double flz = pz - bz - slack, fhz = (bz + cell) - pz - slack;
        W.hist[lane] = 0u; W.hist[lane + 32] = 0u; W.hist[lane + 64] = 0u;
        __syncwarp();
        int ncand = 0, bstar = KH_NB - 1, nsel = 0;
        double bound2 = INFINITY;          // upper bound of the k-th squared distance (edge of the current b*)
        bool certified = false, overflow = false;
        int ex = 0, c0 = 0, c1 = 0, c2 = 0;
        for (int R = 0; R <= KH_RMAX && !certified && !overflow; ++R) {
            const int side = 2 * R + 1, vol = side * side * side;
            for (int cb = 0; cb < vol && !overflow; cb += 32) {
                const int e = cb + lane;
                int s = 0, c = 0;
                if (R == 0) {
                    if (lane == 0 && !cell_find(g.tab, g.bits, pack_key(cx, cy, cz), s, c)) c = 0;
                } else if (e < vol) {
                    int dx, dy, dz;             // e = (dz * side + dy) * side + dx, divisions by compile-time constants
                    switch (R) {
                        case 1: { dz = e / 9; const int r = e - 9 * dz; dy = r / 3; dx = r - 3 * dy; break; }
                        case 2: { dz = e / 25; const int r = e - 25 * dz; dy = r / 5; dx = r - 5 * dy; break; }
                        default: { dz = e / 49; const int r = e - 49 * dz; dy = r / 7; dx = r - 7 * dy; break; }
                    }
                    dx -= R; dy -= R; dz -= R;
                    const int x = cx + dx, y = cy + dy, z = cz + dz;
                    const bool shell = max(max(abs(dx), abs(dy)), abs(dz)) == R;
                    if (shell && x >= 0 && x < g.dim[0] && y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
                        // a cell farther than the current bound of the k-th distance cannot contribute
                        bool far = false;
                        if (bound2 < INFINITY) {
                            const double gx = dx == 0 ? 0.0 : fmax((dx < 0 ? flx : fhx) + (double)(abs(dx) - 1) * cell, 0.0);
                            const double gy = dy == 0 ? 0.0 : fmax((dy < 0 ? fly : fhy) + (double)(abs(dy) - 1) * cell, 0.0);
                            const double gz = dz == 0 ? 0.0 : fmax((dz < 0 ? flz : fhz) + (double)(abs(dz) - 1) * cell, 0.0);
                            far = gx * gx + gy * gy + gz * gz > bound2;
                        }
                        if (!far && !cell_find(g.tab, g.bits, pack_key(x, y, z), s, c)) c = 0;
                    }
                }
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                if (ncand + total > KH_NC) { overflow = true; break; }
                // two batches of 32 candidates per turn: both gathers are in flight before either is used (the gather's latency
                // was the kernel's largest stall)
                for (int r0 = 0; r0 < total; r0 += 64) {
                    const int giA = r0 + lane, giB = r0 + 32 + lane;
                    const int ownerA = lane_of_slot(incl, giA);
                    const int tA = __shfl_sync(FULL, s, ownerA) + (giA - (__shfl_sync(FULL, incl, ownerA) - __shfl_sync(FULL, c, ownerA)));
                    int tB = 0;
                    if (r0 + 32 < total) {              // warp-uniform
                        const int ownerB = lane_of_slot(incl, giB);
                        tB = __shfl_sync(FULL, s, ownerB) + (giB - (__shfl_sync(FULL, incl, ownerB) - __shfl_sync(FULL, c, ownerB)));
                    }
                    double4 qA = make_double4(0, 0, 0, 0), qB = qA;
                    if (giA < total) qA = ldg4(g.pts + tA);
                    if (giB < total) qB = ldg4(g.pts + tB);
                    if (giA < total) {
                        const int b = kh_bin(dist2(px, py, pz, qA.x, qA.y, qA.z), base);
                        atomicAdd(&W.hist[b], 1u);
                        W.cidx[ncand + giA] = tA;
                        W.cbin[ncand + giA] = (unsigned char)b;
                    }
                    if (giB < total) {
                        const int b = kh_bin(dist2(px, py, pz, qB.x, qB.y, qB.z), base);
                        atomicAdd(&W.hist[b], 1u);
                        W.cidx[ncand + giB] = tB;
                        W.cbin[ncand + giB] = (unsigned char)b;
                    }
                }
                ncand += total;
            }
            if (overflow) break;
            // after the own cell alone nothing can be certified (unless it is the whole grid); a bound for the pruning of ring 1
            // is only worth a selection when the own cell holds k points
            const bool whole = g.dim[0] <= side && g.dim[1] <= side && g.dim[2] <= side;
            if (R == 0 && !whole && ncand < k) continue;
            __syncwarp();
            int cum_at;
            const int bs = kh_select(W, k, lane, cum_at, ex, c0, c1, c2);
            // every candidate of a bin <= bs has float(d^2) < edge, i.e. d^2 < edge * (1 + 2^-24)
            const double edge = bs >= KH_NB - 1 ? INFINITY : (double)__int_as_float((bs + base + 1) << 20) * (1.0 + 1e-7);
            if (cum_at >= k) bound2 = edge;
            if (R == 0 && !whole) continue;
            // distance to the nearest face beyond which cells are still unexamined
            double gmin = INFINITY;
            const double Rc = (double)R * cell;
            if (cx - R > 0) gmin = fmin(gmin, flx + Rc);
            if (cx + R < g.dim[0] - 1) gmin = fmin(gmin, fhx + Rc);
            if (cy - R > 0) gmin = fmin(gmin, fly + Rc);
            if (cy + R < g.dim[1] - 1) gmin = fmin(gmin, fhy + Rc);
            if (cz - R > 0) gmin = fmin(gmin, flz + Rc);
            if (cz + R < g.dim[2] - 1) gmin = fmin(gmin, fhz + Rc);
            if (gmin == INFINITY) {                       // the whole grid has been examined
                certified = true;
                bstar = (cum_at >= k && edge < INFINITY) ? bs : KH_NB - 1;
            } else if (cum_at >= k && gmin > 0.0 && edge < gmin * gmin) {
                certified = true;
                bstar = bs;
            }
            if (certified) nsel = bstar == bs ? cum_at : ncand;
        }
        if (certified && nsel <= KH_CAP) {
            // ---- pass 2: counting sort of the selected records by bin ----
            W.bstart[3 * lane] = ex; W.bstart[3 * lane + 1] = ex + c0; W.bstart[3 * lane + 2] = ex + c0 + c1;
            W.hist[3 * lane] = (unsigned)ex; W.hist[3 * lane + 1] = (unsigned)(ex + c0); W.hist[3 * lane + 2] = (unsigned)(ex + c0 + c1);
            __syncwarp();
            const unsigned *cb4 = reinterpret_cast<const unsigned *>(W.cbin);
            for (int r0 = 0; r0 < ncand; r0 += 128) {            // four cached bins per lane and load (bins past ncand are stale)
                const int g0 = r0 + 4 * lane;
                unsigned w4 = g0 < ncand ? cb4[g0 >> 2] : 0xffffffffu;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int b = (int)(w4 & 0xffu), gi = g0 + u;
                    w4 >>= 8;
                    if (b <= bstar && gi < ncand) {
                        const int pos = (int)atomicAdd(&W.hist[b], 1u);
                        W.idx[pos] = W.cidx[gi];
                        W.ebin[pos] = (unsigned char)b;
                    }
                }
            }
            __syncwarp();
            // the selected candidates' distances, all lanes gathering at once (inside the scan above it was one lane at a time)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = lane + 32 * h;
                if (e < nsel) {
                    const double4 q = ldg4(g.pts + W.idx[e]);
                    W.key[e] = dist2(px, py, pz, q.x, q.y, q.z);
                }
            }
            __syncwarp();
            // ---- rank inside the bin by (d^2, index): lane l places entries l and l + 32 ----
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = lane + 32 * h;
                if (e < nsel) {
                    const double kd = W.key[e];
                    const int ki = W.idx[e], b = (int)W.ebin[e];
                    const int lo = W.bstart[b], hi = (int)W.hist[b];
                    int rank = lo;
                    for (int j = lo; j < hi; ++j) rank += before(W.key[j], W.idx[j], kd, ki) ? 1 : 0;
                    W.skey[rank] = kd;
                    W.sidx[rank] = ki;
                }
            }
            __syncwarp();
            const int cnt = min(nsel, k);
            // RemoveStatisticalOutliers: mean of sqrt(d2) over the neighbours in ascending order (std::accumulate)
            if (lane < cnt) W.key[lane] = sqrt(W.skey[lane]);
            // the neighbour list is kept: the normals pass derives its k nearest SURVIVORS from it
            if (lane < k) J.knn_sor[(size_t)i * k + lane] = lane < cnt ? W.sidx[lane] : -1;
            __syncwarp();
            if (lane == 0) {
                double sum = 0.0;
                if (cnt == 30) {                      // the reference's nb_neighbors: unrolled (30 loads in flight, then the ordered adds)
#pragma unroll
                    for (int t = 0; t < 30; ++t) sum += W.key[t];
                } else {
                    for (int t = 0; t < cnt; ++t) sum += W.key[t];
                }
                J.avg[i] = cnt > 0 ? sum / (double)cnt : -1.0;
            }
        } else {
            double ld2; int lidx, cnt;
            knn_warp(g, px, py, pz, k, ld2, lidx, cnt);
            const double sq = sqrt(ld2);
            double sum = 0.0;
            for (int t = 0; t < cnt; ++t) sum += __shfl_sync(FULL, sq, t);
            if (lane == 0) J.avg[i] = cnt > 0 ? sum / (double)cnt : -1.0;
            if (lane < k) J.knn_sor[(size_t)i * k + lane] = lane < cnt ? lidx : -1;
        }
        __syncwarp();
    }
}

// grid (chunks, jobs), one warp per query: the warp-cooperative search with a running top-k list (the first version of the
// outlier-filter kNN, kept for A/B runs: MGICP_KNN_MODE=0; `queued`: only the queries listed in fb_list)
__global__ void __launch_bounds__(256) k_knn(Job *jobs, int k, int queued) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const GridView g = make_view(J, 0);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    const int nq = queued ? J.fb_count : g.n;
    for (int e = warp; e < nq; e += nwarp) {
        const int i = queued ? J.fb_list[e] : e;
        const double4 p = ldg4(g.pts + i);
        double ld2; int lidx, cnt;
        knn_warp(g, p.x, p.y, p.z, k, ld2, lidx, cnt);
        // RemoveStatisticalOutliers: mean of sqrt(d2) over the neighbours in ascending order (std::accumulate)
        const double sq = sqrt(ld2);
        double sum = 0.0;
        for (int t = 0; t < cnt; ++t) sum += __shfl_sync(FULL, sq, t);
        if (lane == 0) J.avg[i] = cnt > 0 ? sum / (double)cnt : -1.0;
        // the neighbour list is kept: the normals pass derives its k nearest SURVIVORS from it
        if (lane < k) J.knn_sor[(size_t)i * k + lane] = lane < cnt ? lidx : -1;
    }
}

// grid (chunks, jobs), one THREAD per query: estimate_normals(KNN k) over the outlier-filtered cloud.
// The k nearest neighbours among the survivors come from the outlier filter's list of the k1 nearest points of the
// unfiltered cloud: if at least k of them survived (or the list already covers the whole cloud), the first k survivors
// of that list ARE the answer -- a survivor outside the list is farther than every list entry -- in the same ascending
// order.  Otherwise (rare) the query is queued for k_normals_search.  The covariance (cumulants in ascending-distance
// order) and the closed-form eigenvector are per-thread work; lists live in shared memory.
constexpr int NRM_NT = 256, NRM_LD = 33;
__global__ void __launch_bounds__(NRM_NT) k_normals(Job *jobs, int k, int k1, int debug) {
    __shared__ int32_t s_list[NRM_NT * NRM_LD];
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const GridView g = make_view(J, 1);
    const int lane = threadIdx.x & 31;
    int32_t *list = s_list + threadIdx.x * NRM_LD;
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < g.n; i0 += gridDim.x * blockDim.x) {   // warp-uniform
        const int i = i0 + lane;
        const bool have = i < g.n;
        bool continue_ = false;
        double4 p = make_double4(0, 0, 0, 0);
        int cnt = 0, listed = 0;
        if (have) {
            p = ldg4(g.pts + i);
            const int32_t *src = J.knn_sor + (size_t)(int)p.w * k1;       // p.w = index of this point in the unfiltered (grid) order
            for (int u = 0; u < k1; ++u) {
                const int t = __ldg(src + u);
                if (t >= 0) {
                    ++listed;
                    if (J.keep[t] && cnt < k) list[cnt++] = J.newidx[t];
                }
            }
        }
        if (have && cnt < k && listed == k1) {
            // too few survivors in the list: queue the query for the warp-per-query search of k_normals_search
            J.fb_list[atomicAdd(&J.fb_count, 1)] = i;
            continue_ = true;
        }
        if (have && !continue_) {
            double cov[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
            if (cnt >= 3) {
                Cumulants cu;
                cu.clear();
                for (int u = 0; u < cnt; ++u) {
                    const double4 q = ldg4(g.pts + list[u]);
                    cu.add(q.x, q.y, q.z);
                }
                cu.covariance(cnt, cov);
            }
            const V3 nv = normal_from_cov(cov);
            // Neighbour-walk certificate for the ICP loop: a query closer to this point than half the distance to this
            // point's 9th nearest other point has its exact nearest neighbour among this point and its 8 nearest ones
            // (anything else is at least that 9th distance away from this point).  If the list covers the whole cloud the
            // radius is unbounded.
            double safe2 = cnt < k ? INFINITY : 0.0;
            if (cnt >= 10) {
                const double4 q = ldg4(g.pts + list[9]);
                safe2 = 0.25 * dist2(p.x, p.y, p.z, q.x, q.y, q.z) * (1.0 - 1e-9);
            }
            J.nrm[i] = make_double4(nv.x, nv.y, nv.z, safe2);
            int nb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) nb[u] = u + 1 < cnt ? list[u + 1] : -1;
            int4 *dst = reinterpret_cast<int4 *>(J.nbrA + (size_t)i * 8);
            dst[0] = make_int4(nb[0], nb[1], nb[2], nb[3]);
            dst[1] = make_int4(nb[4], nb[5], nb[6], nb[7]);
            if (debug) for (int u = 0; u < k; ++u) J.knn_nrm[(size_t)i * k + u] = u < cnt ? list[u] : -1;
        }
        __syncwarp();
    }
}

// grid (chunks, jobs), one warp per queued query: the queries of k_normals whose outlier-filter list held fewer than k
// survivors get an exact kNN search over the filtered cloud's grid (order of the queue is irrelevant: every entry is
// independent and writes only its own row)
__global__ void __launch_bounds__(256) k_normals_search(Job *jobs, int k, int debug) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const GridView g = make_view(J, 1);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    const int nfb = J.fb_count;
    for (int e = warp; e < nfb; e += nwarp) {
        const int i = J.fb_list[e];
        const double4 p = ldg4(g.pts + i);
        double ld2; int lidx, cnt;
        knn_warp(g, p.x, p.y, p.z, k, ld2, lidx, cnt);
        double4 q = make_double4(0, 0, 0, 0);
        if (lane < cnt) q = ldg4(g.pts + lidx);
        double cov[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
        if (cnt >= 3) {
            Cumulants cu;
            cu.clear();
            for (int u = 0; u < cnt; ++u)
                cu.add(__shfl_sync(FULL, q.x, u), __shfl_sync(FULL, q.y, u), __shfl_sync(FULL, q.z, u));
            cu.covariance(cnt, cov);
        }
        const V3 nv = normal_from_cov(cov);      // every lane computes the same value; lane 0 stores it
        const double far2 = __shfl_sync(FULL, ld2, 9);
        const double safe2 = cnt >= 10 ? 0.25 * far2 * (1.0 - 1e-9) : (cnt < k ? INFINITY : 0.0);
        if (lane == 0) J.nrm[i] = make_double4(nv.x, nv.y, nv.z, safe2);
        if (lane >= 1 && lane <= 8) J.nbrA[(size_t)i * 8 + (lane - 1)] = lane < cnt ? lidx : -1;
        if (debug && lane < k) J.knn_nrm[(size_t)i * k + lane] = lane < cnt ? lidx : -1;
    }
}

// one CTA per job: RemoveStatisticalOutliers statistics, keep mask, order-preserving compaction, and the
// spatial hash of the surviving points (same cells, re-ranged)
__global__ void __launch_bounds__(1024) k_sor_select(Job *jobs, double ratio) {
    __shared__ int sm[33];
    __shared__ double sd[32];
    Job &J = jobs[blockIdx.x];
    if (J.err || J.n <= 0) return;
    const int M = J.M;
    double s = 0.0;
    for (int i = threadIdx.x; i < M; i += 1024) { double a = J.avg[i]; if (a > 0) s += a; }
    const double mean = block_sum(s, sd) / (double)M;          // valid_distances == M: every point finds itself
    double q = 0.0;
    for (int i = threadIdx.x; i < M; i += 1024) { double a = J.avg[i]; if (a > 0) q += (a - mean) * (a - mean); }
    const double stdev = sqrt(block_sum(q, sd) / (double)(M - 1));
    const double thr = mean + ratio * stdev;
    const double *avg = J.avg;
    uint8_t *keep = J.keep;
    int32_t *newidx = J.newidx;
    const double4 *gpts = J.gpts;
    double4 *pts = J.pts;
    const int Mf = block_region_scan(M, sm, [&](int i) { const double a = avg[i]; return (a > 0 && a < thr) ? 1 : 0; },
                                     [&](int i, int pre, int v) {
                                         keep[i] = (uint8_t)v;
                                         newidx[i] = pre;
                                         if (v) { double4 p = gpts[i]; p.w = (double)i; pts[pre] = p; }
                                     });
    if (threadIdx.x == 0) { J.newidx[M] = Mf; J.Mf = Mf; J.sor_thresh = thr; J.fb_count = 0; }     // the kNN queue is done: k_normals reuses it
    // size the ICP grid (cleared and filled by later kernels)
    int ibits = 10;
    while (ibits < J.cbits_max && (1 << ibits) < 4 * Mf) ++ibits;
    if (threadIdx.x == 0) J.ibits = ibits;
}

// grid (chunks, jobs): neighbour lists re-indexed to ICP-grid order
__global__ void __launch_bounds__(256) k_nbr_remap(Job *jobs) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const int n8 = J.Mf * 8;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n8; e += gridDim.x * blockDim.x) {
        const int pos = e >> 3, u = e & 7;
        const int t = J.nbrA[(size_t)J.i2a[pos] * 8 + u];
        J.inbr[e] = t >= 0 ? J.a2i[t] : -1;
    }
}

// =============================================================================================
// K3 + K4 + K5: the ICP loop of registration_generalized_icp, all scales, one launch.
// One thread block, or a gang of G co-resident blocks synchronised through global memory, per pair.
// =============================================================================================
struct PairState;
struct IcpArgs {
    const Job *jobs;
    int n_scales;
    const int32_t *pair_src, *pair_tgt;    // device
    const double *max_d;                   // device [pairs * scales]
    const int32_t *max_it;                 // device [scales]
    const double *T_init;                  // device [pairs * 16]
    double *T_out, *fitness, *rmse;
    int32_t *iters, *ncorr;
    double *stats;
    double *pcur, *mcur;                   // scratch, per pair at 3 * scratch_off[pair]: packed xyz per source point
    double4 *anchor;                       // scratch, per pair at scratch_off[pair]
    int2 *prev;                            // per source point: last correspondence (-1: none) and its margin (float bits)
    const int64_t *scratch_off;
    double k;                              // 1 - epsilon
    int loss; double loss_k;
    double rel_fitness, rel_rmse;
    int gang;                              // thread blocks per pair (static gangs) / pass chunks per pair (task mode)
    double *gpart;                         // gang: [pairs][2][gang][NACC], tasks: [pairs][chunks][NACC] cross-block partial sums
    unsigned int *gsync;                   // [pairs][2] barrier state (zeroed before the launch)
    int eval_scale;                        // >= 0: single evaluation pass at that scale (mgicp_evaluate_batch)
    double *eval_out;
    // task mode
    int n_pairs, n_ctas;
    PairState *ps;                         // [pairs]
    int *queue;                            // task slots, zero = not yet published
    unsigned int *qctl;                    // [0] head (next ticket), [1] tail (next free slot), [2] pairs finished, [3] next pair to start
    int active_pairs;                      // pairs in flight at any time (bounds the working set that has to stay in L2)
    int vmax;                              // task ids are pair * vmax + chunk + 1
    long long v_num, v_den;                // chunks per pass of a scale with ns source points: v_den == 0 ? gang :
                                           // clamp((ns * v_num + v_den / 2) / v_den, 1, vmax)
    unsigned long long *scale_t;           // optional [pairs][scales][2]: %globaltimer (ns) at the start / end of a pair's scale
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

#ifndef MGICP_NT
#define MGICP_NT 512
#endif
constexpr int ICP_NT = MGICP_NT;
constexpr int ICP_MIN_CTAS = ICP_NT <= 512 ? 512 / ICP_NT : 1;
constexpr int NACC = 29;   // 21 JTJ + 6 JTr + K + sum d2

// Sense-reversing barrier among the G thread blocks ("gang") that share one pair.  The launch is cooperative when G > 1,
// so all blocks are co-resident.  sync[0] = arrival counter, sync[1] = generation.
__device__ __forceinline__ void gang_barrier(unsigned int *sync, const int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int *gen = sync + 1;
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(sync, 1u) == (unsigned int)(G - 1)) {
            sync[0] = 0u;
            __threadfence();
            atomicAdd(sync + 1, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// The 27 normal-equation sums of a thread live in shared memory (column `threadIdx.x` of a [27][ICP_NT] array: consecutive
// threads, consecutive addresses), which frees 54 registers for the search; K and sum d^2 stay in registers.
constexpr int NSUM = 27;
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr size_t ICP_DYN_SMEM = sizeof(double) * NACC * ICP_NT;     // 27 sums + the K and sum d^2 rows of the block reduction
struct SAcc {
    double *col;
    __device__ __forceinline__ SAcc(double *base) : col(base + threadIdx.x) {}
    __device__ __forceinline__ SAcc(double *base, const int column) : col(base + column) {}
    __device__ __forceinline__ double &operator[](const int a) { return col[a * ICP_NT]; }
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int a = 0; a < NSUM; ++a) col[a * ICP_NT] = 0.0;
    }
};

// Reduce the NACC per-thread sums across the block.  They already sit in shared memory as [NACC][ICP_NT] rows (K and
// sum d^2 are appended), so warp w sums rows w, w + 16: lane l adds columns l, l + 32, ... in order, then a shuffle-down
// tree over the lanes -- ~95 instructions per warp where 29 per-warp shuffle trees took ~520 (single pair: ICP 3.69 -> 3.26 ms).
// On return threads < NACC hold the block total of accumulator threadIdx.x (other threads: 0).
__device__ __forceinline__ double block_reduce_acc(SAcc &acc, const double accK, const double accD, double *red /* [NACC] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    acc.col[NSUM * ICP_NT] = accK;
    acc.col[(NSUM + 1) * ICP_NT] = accD;
    __syncthreads();
    const double *base = acc.col - threadIdx.x;
    for (int a = w; a < NACC; a += ICP_NT / 32) {
        const double *row = base + a * ICP_NT;
        double v = row[lane];
#pragma unroll
        for (int i = 1; i < ICP_NT / 32; ++i) v += row[lane + 32 * i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[a] = v;
    }
    __syncthreads();
    return threadIdx.x < NACC ? red[threadIdx.x] : 0.0;
}

// ... and across the G blocks of a static gang (partials through global memory, summed in rank order by every block:
// identical totals everywhere)
__device__ __forceinline__ void pair_reduce(SAcc &acc, const double accK, const double accD, double *red,
                                            double *gpart /* [2][G][NACC] */, unsigned int *sync, const int G, const int rank,
                                            const int phase, double *tot) {
    const double s = block_reduce_acc(acc, accK, accD, red);
    if (G == 1) {
        if (threadIdx.x < NACC) tot[threadIdx.x] = s;
        __syncthreads();
    } else {
        if (threadIdx.x < NACC) __stcg(gpart + ((size_t)phase * G + rank) * NACC + threadIdx.x, s);
        gang_barrier(sync, G);
        if (threadIdx.x < NACC) {
            const double *all = gpart + (size_t)phase * G * NACC;
            double t = 0.0;
            for (int r = 0; r < G; ++r) t += __ldcg(all + (size_t)r * NACC + threadIdx.x);   // fixed rank order
            tot[threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// Per-pair scratch accesses.  In task mode the chunks of consecutive passes run on different SMs, so the evolving
// per-point state must bypass the (non-coherent) L1: ld.cg / st.cg.  A static gang always maps a point to the same thread.
template <bool COH> __device__ __forceinline__ double4 ld_d4(const double4 *p) {
    if (COH) return ldcg4(p);
    return *p;
}
template <bool COH> __device__ __forceinline__ void st_d4(double4 *p, const double4 v) {
    if (COH) stcg4(p, v);
    else *p = v;
}
// per-point state (transformed source point / effective normal): packed 3 doubles, 24 bytes per point
template <bool COH> __device__ __forceinline__ V3 ld_v3(const double *base, const int i) {
    const double *p = base + 3 * (size_t)i;
    if (COH) return v3(__ldcg(p), __ldcg(p + 1), __ldcg(p + 2));
    return v3(p[0], p[1], p[2]);
}
template <bool COH> __device__ __forceinline__ void st_v3(double *base, const int i, const V3 &v) {
    double *p = base + 3 * (size_t)i;
    if (COH) { __stcg(p, v.x); __stcg(p + 1, v.y); __stcg(p + 2, v.z); }
    else { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
}
// last correspondence of a source point + its margin (see icp_pass_steady), one 8-byte record
template <bool COH> __device__ __forceinline__ int2 ld_i2(const int2 *p) { return COH ? __ldcg(p) : *p; }
template <bool COH> __device__ __forceinline__ void st_i2(int2 *p, const int2 v) { if (COH) __stcg(p, v); else *p = v; }

// One pass of GetRegistrationResultAndCorrespondences + the linearisation of ComputeTransformation over the share of the
// source points owned by (rank `tid / ICP_NT` of `nthr / ICP_NT`).
// Queries per warp: 32 normally; when there are more warps than 32-query chunks, shorter chunks (16 or 8 owners per warp,
// all 32 lanes still cooperate in the search) cut the latency of the slowest warp, which is what a pass of a
// latency-bound single pair waits for.  Deterministic function of (ns, nthr): the oracle's emulation mirrors it.
__device__ __forceinline__ int queries_per_warp(const int ns, const int nwarps) {
    if ((ns + 7) / 8 <= nwarps) return 8;
    if ((ns + 15) / 16 <= nwarps) return 16;
    return 32;
}

// second half of a point's turn: cooperative search if still needed, bookkeeping for the next pass
template <bool COH, bool FIRST>
__device__ __forceinline__ void icp_resolve(const GridView &g, WarpSearch &ws, const bool have, const bool need, const V3 &p, const double r,
                                            const double r2, const double rs, const double rs2, double &d2, int &j, double4 *anchor_i,
                                            int2 *prev_i, const float margin) {
    const bool searched = need;
    nn_search_coop(g, ws, need, p.x, p.y, p.z, rs2, d2, j);
    const bool matched = j >= 0 && d2 < r2;                // accepted iff d2 < r2 (strict), like SearchHybrid
    if (searched && !matched) {
        // no neighbour within r: remember the proof radius for the following passes
        st_d4<COH>(anchor_i, make_double4(p.x, p.y, p.z, j >= 0 ? sqrt(d2) * (1.0 - 1e-9) : rs));
    } else if (FIRST && have) {
        st_d4<COH>(anchor_i, make_double4(0.0, 0.0, 0.0, 0.0));
    }
    if (!matched) j = -1;
    if (have) st_i2<COH>(prev_i, make_int2(j, __float_as_int(searched || !matched ? 0.0f : margin)));
}

__device__ __forceinline__ void icp_linearise(const IcpArgs &A, const Job &JT, const V3 &p, const V3 &m, const int j, const double d2,
                                              SAcc &acc, double &accK, double &accD) {
    const double4 q = ldg4(JT.ipts + j);
    const double4 nq = ldg4(JT.inrm + j);
    const V3 mt = effective_normal(v3(nq.x, nq.y, nq.z));
    gicp_accumulate(p, v3(q.x, q.y, q.z), m, mt, A.k, A.loss, A.loss_k, acc);
    accK += 1.0;
    accD += d2;
}

// what a pass does with a point once its correspondence is settled: linearise on the spot (the fused kernels) ...
struct FusedSink {
    const IcpArgs &A; const Job &JT; SAcc &acc; double &accK; double &accD;
    __device__ __forceinline__ void operator()(const bool ok, const V3 &p, const V3 &m, const int j, const double d2) {
        if (ok) icp_linearise(A, JT, p, m, j, d2, acc, accK, accD);
    }
};

// First pass of a scale: pcd = source; if (!init.isIdentity()) pcd.Transform(init) -- points and covariances -- with
// M = the running transformation, then a full search for every point.
template <bool COH, class Sink>
__device__ __forceinline__ void icp_pass_first(const IcpArgs &A, const Job &JS, const Job &JT, const GridView &g, const int ns,
                                               const double r, const double *M /* shared memory */, const int tid, const int nthr,
                                               double *pcur, double *mcur, double4 *anchor, int2 *prev, WarpSearch &ws, Sink &sink) {
    const int lane = threadIdx.x & 31;
    const double r2 = r * r;
    const double rs = 1.5 * r, rs2 = rs * rs;      // search radius for points without a correspondence
    const int nwarps = nthr >> 5, gw = tid >> 5;
    const int Q = queries_per_warp(ns, nwarps);
    bool ident = true;
#pragma unroll
    for (int i = 0; i < 16; ++i) ident &= (M[i] == ((i % 5 == 0) ? 1.0 : 0.0));
    for (int ib = gw * Q; ib < ns; ib += nwarps * Q) {       // warp-uniform trip count
        const int i = ib + lane;
        const bool have = lane < Q && i < ns;
        V3 p = v3(0, 0, 0), m = v3(1, 0, 0);
        if (have) {
            const double4 p0 = ldg4(JS.ipts + i);
            const double4 n0 = ldg4(JS.inrm + i);
            p = v3(p0.x, p0.y, p0.z);
            m = effective_normal(v3(n0.x, n0.y, n0.z));
            if (!ident) { p = transform_point(M, p); m = rotate_vec(M, m); }
            st_v3<COH>(pcur, i, p);
            st_v3<COH>(mcur, i, m);
        }
        double d2 = rs2;
        int j = -1;
        icp_resolve<COH, true>(g, ws, have, have, p, r, r2, rs, rs2, d2, j, anchor + i, prev + i, 0.0f);
        sink(have && j >= 0, p, m, j, d2);
    }
}

// Later passes: pcd.Transform(update) with M = the last update on the stored state, then the correspondence of every
// point is re-established from what the previous pass left behind:
//  * matched points carry their last correspondence as seed, together with a margin: a lower bound (float, rounded down)
//    on the distance from the point's previous position to every OTHER target point.  If |p - seed| + (distance moved by
//    the update) < margin, the seed is still the unique nearest neighbour and nothing else is loaded; the margin shrinks
//    by the distance moved.  Typical for the many small-update passes of the L1 loop;
//  * otherwise, if the query sits inside the seed's safe ball, the exact nearest neighbour is the seed or one of its 8
//    nearest points (neighbour walk), which also yields a fresh margin: the second best of the nine, or what the safe
//    ball guarantees for all the others;
//  * unmatched points were searched with a radius 1.5 r.  Whatever that search found (nothing, or the nearest point at
//    distance >= r) is a lower bound `lb` on the nearest-neighbour distance at that position (the anchor); while the point
//    stays within lb - r of its anchor it provably has no neighbour within r and the search is skipped.  An anchor is a
//    fact about the target cloud, so it stays valid for the whole scale;
//  * everything else goes through the warp-cooperative grid search.
// (Issuing the first two links of the next turn's load chain during the current one -- software pipelining through
// registers -- was measured slower: 41.9 vs 39.4 ms at 148 pairs; the extra live registers spill.)
template <bool COH, class Sink>
__device__ __forceinline__ void icp_pass_steady(const IcpArgs &A, const Job &JS, const Job &JT, const GridView &g, const int ns,
                                                const double r, const double *M /* shared memory */, const int tid, const int nthr,
                                                double *pcur, double *mcur, double4 *anchor, int2 *prev, WarpSearch &ws, Sink &sink) {
    const int lane = threadIdx.x & 31;
    const double r2 = r * r;
    const double rs = 1.5 * r, rs2 = rs * rs;
    const int nwarps = nthr >> 5, gw = tid >> 5;
    const int Q = queries_per_warp(ns, nwarps);
    // The turn of a point is a chain of dependent loads (state -> seed point + neighbour list -> neighbour points ->
    // target normal).  The next turn's seed record is fetched one turn ahead so that the lines the next turn will need can be
    // requested with register-free prefetches while this turn computes, and the next turn's state is loaded into registers
    // right before this turn's linearisation (the long fp64 stretch of a turn), so that its latency is covered too.
    // (Looking two turns ahead and also requesting the eight neighbour points was measured 6 % slower.)
    const int stride = nwarps * Q;
    int2 rec_next = make_int2(-1, 0);
    V3 p_next = v3(0, 0, 0), m_next = v3(1, 0, 0);
    {
        const int i0 = gw * Q + lane;
        if (lane < Q && i0 < ns) {
            rec_next = ld_i2<COH>(prev + i0);
            p_next = ld_v3<COH>(pcur, i0); m_next = ld_v3<COH>(mcur, i0);
        }
    }
    for (int ib = gw * Q; ib < ns; ib += nwarps * Q) {       // warp-uniform trip count
        const int i = ib + lane;
        const bool have = lane < Q && i < ns;
        V3 p = v3(0, 0, 0), m = v3(1, 0, 0);
        const int seed = rec_next.x;
        const float lb = __int_as_float(rec_next.y);
        rec_next = make_int2(-1, 0);
        const int i_n = i + stride;
        const bool have_n = lane < Q && i_n < ns;
        if (have_n) {
            rec_next = ld_i2<COH>(prev + i_n);
            prefetch_l2(pcur + 3 * (size_t)i_n);
            prefetch_l2(mcur + 3 * (size_t)i_n);
        }
        float moved = 0.0f;
        if (have) {
            p = transform_point(M, p_next);
            m = rotate_vec(M, m_next);
            moved = __fsqrt_ru(__double2float_ru(dist2(p.x, p.y, p.z, p_next.x, p_next.y, p_next.z)));   // >= |p - p_old|
            st_v3<COH>(pcur, i, p);
            st_v3<COH>(mcur, i, m);
        }
        double d2 = rs2;
        int j = -1;
        bool need = have;
        float margin = 0.0f;
        if (have && seed < 0) {
            const double4 an = ld_d4<COH>(anchor + i);
            const double slackd = (an.w - r) * (1.0 - 1e-9);
            if (slackd > 0.0 && dist2(p.x, p.y, p.z, an.x, an.y, an.z) < slackd * slackd) need = false;
        }
        if (seed >= 0) {
            const double4 q = ldg4(g.pts + seed);
            const double d = dist2(p.x, p.y, p.z, q.x, q.y, q.z);
            const float df = __fsqrt_ru(__double2float_ru(d));                       // >= |p - seed|
            if (__fmul_ru(__fadd_ru(df, moved), 1.000001f) < lb) {
                // margin certificate: every other target point was farther than lb from the previous position, so it is
                // farther than lb - moved > |p - seed| from this one: same nearest neighbour, no walk.  Every float is
                // rounded to the safe side; the 1e-6 guard covers the fp64 rounding (1e-15) of the squared distances
                // the floats were derived from
                need = false;
                margin = __fsub_rd(lb, moved);
                if (d < r2) { d2 = d; j = seed; }
            } else if (d < q.w) {
                double bd = d, sd = INFINITY;                                          // best and second best of the nine
                int bj = seed;
                const int4 *nb = reinterpret_cast<const int4 *>(JT.inbr + (size_t)seed * 8);
                const int4 n0 = __ldg(nb), n1 = __ldg(nb + 1);
                const int cand[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
                double4 qq[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) qq[u] = ldg4(g.pts + max(cand[u], 0));     // 8 independent loads in flight
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double dd = dist2(p.x, p.y, p.z, qq[u].x, qq[u].y, qq[u].z);
                    if (cand[u] >= 0) {
                        if (dd < bd || (dd == bd && cand[u] < bj)) { sd = bd; bd = dd; bj = cand[u]; }
                        else sd = fmin(sd, dd);
                    }
                }
                need = false;
                if (bd < r2) { d2 = bd; j = bj; }
                // lower bound on the distance from p to every target point but bj: the second best of the nine, and
                // (9th-neighbour distance of the seed = 2 x its safe radius) - |p - seed| for all the others
                margin = fminf(__fsqrt_rd(__double2float_rd(sd)), __fsub_rd(2.0f * __fsqrt_rd(__double2float_rd(q.w)), df));
            } else if (d < rs2) { d2 = d; j = seed; }
        }
        icp_resolve<COH, false>(g, ws, have, need, p, r, r2, rs, rs2, d2, j, anchor + i, prev + i, margin);
        if (rec_next.x >= 0) {
            // the seed point and its normal; the neighbour list is not requested: the walk is rare once the margins hold
            // (with the list: +1.4 % time and 12 % more DRAM traffic; all three into L2 only: no difference to L1)
            prefetch_l1(g.pts + rec_next.x);
            prefetch_l1(JT.inrm + rec_next.x);
        } else if (have_n) prefetch_l2(anchor + i_n);
        if (have_n) { p_next = ld_v3<COH>(pcur, i_n); m_next = ld_v3<COH>(mcur, i_n); }
        sink(have && j >= 0, p, m, j, d2);
    }
}

// ComputeTransformation's solve + "transformation = update * transformation" (thread 0 only)
__device__ __forceinline__ void solve_and_update(const double *tot, const double K, const double *Told, double *Uout, double *Tout) {
    double Um[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (K > 0.0) {
        double sums[27], x[6];
#pragma unroll
        for (int a = 0; a < 27; ++a) sums[a] = tot[a];
        ldlt_solve6(sums, x);
        vec6_to_mat4(x, Um);
    }
    double Tn[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) Tn[i] = Told[i];
    double To[16];
    mat4_mul(Um, Tn, To);
#pragma unroll
    for (int i = 0; i < 16; ++i) { Uout[i] = Um[i]; Tout[i] = To[i]; }
}

// ---- static mode: one thread block, or a gang of G co-resident blocks synchronised through global memory, per pair ----
__global__ void __launch_bounds__(ICP_NT, ICP_MIN_CTAS) k_icp(IcpArgs A) {
    __shared__ double sT[16], sU[16], tot[32];
    __shared__ double red[32];
    __shared__ WarpSearch wsm[ICP_NT / 32];
    extern __shared__ double s_sums[];
    SAcc acc(s_sums);
    const int G = A.gang;
    const int pair = blockIdx.x / G;
    const int rank = blockIdx.x % G;
    const int tid = rank * ICP_NT + threadIdx.x, nthr = G * ICP_NT;
    double *gpart = A.gpart + (size_t)pair * 2 * G * NACC;
    unsigned int *gsync = A.gsync + 2 * pair;
    const int S = A.n_scales;
    const int sc = A.pair_src[pair], tc = A.pair_tgt[pair];
    double *pcur = A.pcur + 3 * A.scratch_off[pair];
    double *mcur = A.mcur + 3 * A.scratch_off[pair];
    double4 *anchor = A.anchor + A.scratch_off[pair];
    int2 *prev = A.prev + A.scratch_off[pair];
    if (threadIdx.x < 16) sT[threadIdx.x] = A.T_init[pair * 16 + threadIdx.x];
    __syncthreads();
    int phase = 0;
    double fit = 0.0, rmse = 0.0, Klast = 0.0;
    const int s_begin = A.eval_scale >= 0 ? A.eval_scale : 0, s_end = A.eval_scale >= 0 ? A.eval_scale + 1 : S;
    for (int s = s_begin; s < s_end; ++s) {
        const Job &JS = A.jobs[sc * S + s];
        const Job &JT = A.jobs[tc * S + s];
        const int ns = JS.Mf, nt = JT.Mf;
        const double r = A.max_d[pair * S + s];
        const int max_it = A.eval_scale >= 0 ? 0 : A.max_it[s];
        const GridView g = make_view(JT, 2);
        int iters = 0;
        double sumK = 0.0;
        int passes = 0;
        double pfit = 0.0, prmse = 0.0;
        fit = 0.0; rmse = 0.0; Klast = 0.0;
        if (A.scale_t && rank == 0 && threadIdx.x == 0) A.scale_t[((size_t)pair * S + s) * 2] = global_ns();
        if (ns > 0 && nt > 0) {
            for (int pass = 0;; ++pass) {
                double accK = 0.0, accD = 0.0;
                acc.clear();
                FusedSink sink{A, JT, acc, accK, accD};
                if (pass == 0) icp_pass_first<false>(A, JS, JT, g, ns, r, sT, tid, nthr, pcur, mcur, anchor, prev, wsm[threadIdx.x >> 5], sink);
                else icp_pass_steady<false>(A, JS, JT, g, ns, r, sU, tid, nthr, pcur, mcur, anchor, prev, wsm[threadIdx.x >> 5], sink);
                pair_reduce(acc, accK, accD, red, gpart, gsync, G, rank, phase, tot);
                phase ^= 1;
                ++passes;
                const double K = tot[27], e2 = tot[28];
                sumK += K;
                Klast = K;
                if (K > 0.0) { fit = K / (double)ns; rmse = sqrt(e2 / K); } else { fit = 0.0; rmse = 0.0; }
                if (A.eval_scale >= 0) {
                    if (rank == 0 && threadIdx.x < 27) A.eval_out[pair * 32 + 4 + threadIdx.x] = tot[threadIdx.x];
                    if (rank == 0 && threadIdx.x == 0) {
                        A.eval_out[pair * 32 + 0] = fit; A.eval_out[pair * 32 + 1] = rmse;
                        A.eval_out[pair * 32 + 2] = K; A.eval_out[pair * 32 + 3] = e2;
                    }
                    break;
                }
                iters = pass;
                if (pass > 0 && fabs(pfit - fit) < A.rel_fitness && fabs(prmse - rmse) < A.rel_rmse) break;
                if (pass >= max_it) break;
                __syncthreads();
                if (threadIdx.x == 0) solve_and_update(tot, K, sT, sU, sT);
                __syncthreads();
                pfit = fit; prmse = rmse;
            }
        }
        __syncthreads();
        if (rank == 0 && threadIdx.x == 0 && A.eval_scale < 0) {
            if (A.scale_t) A.scale_t[((size_t)pair * S + s) * 2 + 1] = global_ns();
            if (A.iters) A.iters[pair * S + s] = iters;
            if (A.stats) {
                double *st = A.stats + ((size_t)pair * S + s) * 8;
                st[0] = (double)ns; st[1] = (double)nt; st[2] = (double)iters; st[3] = Klast;
                st[4] = fit; st[5] = rmse; st[6] = sumK; st[7] = (double)passes;
            }
        }
    }
    if (rank == 0 && A.eval_scale < 0) {
        if (threadIdx.x < 16) A.T_out[pair * 16 + threadIdx.x] = sT[threadIdx.x];
        if (threadIdx.x == 0) {
            A.fitness[pair] = fit;
            A.rmse[pair] = rmse;
            if (A.ncorr) A.ncorr[pair] = (int32_t)Klast;
        }
    }
}

// ---- task mode: dynamic load balancing for batches -------------------------------------------------------------------
// Pairs need anything from ~60 to ~300 passes, so one block per pair leaves a third of the SM-time idle once the short
// pairs are done.  Here a pass of a pair is cut into V fixed chunks (chunk c owns exactly the points rank c of a static
// gang of V would own, so the arithmetic -- including the summation order -- is that of a gang of V), and persistent
// thread blocks pull (pair, chunk) tasks from a ticket queue.  The block that completes the last chunk of a pass sums the
// partials in rank order, runs the 6x6 solve and the convergence test, publishes the pair's new state, enqueues chunks
// 1..V-1 of the next pass and continues with chunk 0 itself.
struct PairState {
    double T[16];          // running transformation
    double U[16];          // last update
    double pfit, prmse;    // fitness / rmse of the previous pass
    double sumK;           // correspondences summed over the passes of this scale (roofline accounting)
    int32_t scale, pass;
    unsigned int done;     // chunks of the current pass finished so far
    int32_t V;             // chunks per pass at the current scale
};

// Chunks per pass.  Fixed (explicit opts.ctas_per_pair < 0), or sized so that a chunk is worth its scheduling overhead:
// coarse scales (few points) run as one or two chunks, the finest (most points, most passes, last to finish) as many,
// which is also what shortens the tail of a batch.  A function of the point count and launch constants only.
__device__ __forceinline__ int chunks_for_scale(const IcpArgs &A, const int ns) {
    if (A.v_den == 0) return A.gang;
    const long long v = ((long long)ns * A.v_num + A.v_den / 2) / A.v_den;
    return (int)max(1LL, min(v, (long long)A.vmax));
}

__device__ __forceinline__ void queue_push_range(const IcpArgs &A, const int first_value, const int count, const int step) {
    const unsigned int pos = atomicAdd(&A.qctl[1], (unsigned int)count);
    volatile int *q = A.queue;
    for (int c = 0; c < count; ++c) q[pos + c] = first_value + c * step;
}

// record the (empty) result of scales that cannot run, return the first scale >= s with points on both sides (or S)
__device__ __forceinline__ int next_runnable_scale(const IcpArgs &A, const int pair, int s) {
    const int S = A.n_scales, sc = A.pair_src[pair], tc = A.pair_tgt[pair];
    for (; s < S; ++s) {
        const int ns = A.jobs[sc * S + s].Mf, nt = A.jobs[tc * S + s].Mf;
        if (ns > 0 && nt > 0) break;
        if (A.iters) A.iters[pair * S + s] = 0;
        if (A.stats) {
            double *st = A.stats + ((size_t)pair * S + s) * 8;
            st[0] = (double)ns; st[1] = (double)nt; st[2] = 0.0; st[3] = 0.0; st[4] = 0.0; st[5] = 0.0; st[6] = 0.0; st[7] = 0.0;
        }
    }
    return s;
}

// one thread: the pair is finished.  The last pair to finish releases every block with an exit token; otherwise the
// index of the next waiting pair (to be passed to start_pair) or -1 is returned.
__device__ __forceinline__ int finish_pair(const IcpArgs &A, const int pair, const double *T, const double fit, const double rmse,
                                           const double Klast) {
    for (int i = 0; i < 16; ++i) A.T_out[pair * 16 + i] = T[i];
    A.fitness[pair] = fit;
    A.rmse[pair] = rmse;
    if (A.ncorr) A.ncorr[pair] = (int32_t)Klast;
    if (atomicAdd(&A.qctl[2], 1u) == (unsigned int)(A.n_pairs - 1)) { queue_push_range(A, -1, A.n_ctas, 0); return -1; }
    const unsigned int next = atomicAdd(&A.qctl[3], 1u);
    return next < (unsigned int)A.n_pairs ? (int)next : -1;
}

// one thread: initial state of pairs `pair`, ... and the V tasks of the first pass; a pair none of whose scales can run
// finishes on the spot, which starts the next waiting one
__device__ __forceinline__ void start_pairs(const IcpArgs &A, int pair) {
    while (pair >= 0) {
        PairState &P = A.ps[pair];
        for (int i = 0; i < 16; ++i) { P.T[i] = A.T_init[pair * 16 + i]; P.U[i] = (i % 5 == 0) ? 1.0 : 0.0; }
        P.pfit = 0.0; P.prmse = 0.0; P.sumK = 0.0; P.pass = 0; P.done = 0u; P.V = 1;
        const int s = next_runnable_scale(A, pair, 0);
        P.scale = s;
        if (s >= A.n_scales) { pair = finish_pair(A, pair, A.T_init + pair * 16, 0.0, 0.0, 0.0); continue; }
        const int V = chunks_for_scale(A, A.jobs[A.pair_src[pair] * A.n_scales + s].Mf);
        P.V = V;
        __threadfence();
        queue_push_range(A, pair * A.vmax + 1, V, 1);
        pair = -1;
    }
}

// the first `active_pairs` pairs start right away, the others as pairs finish
__global__ void k_icp_task_init(IcpArgs A) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair < A.active_pairs) start_pairs(A, pair);
}

#ifndef MGICP_TASK_MAXREG
#define MGICP_TASK_MAXREG 0
#endif
#if MGICP_TASK_MAXREG
__global__ void __maxnreg__(MGICP_TASK_MAXREG) k_icp_tasks(IcpArgs A) {
#else
__global__ void __launch_bounds__(ICP_NT, ICP_MIN_CTAS) k_icp_tasks(IcpArgs A) {
#endif
    __shared__ double sM[16], tot[32];
    __shared__ double red[32];
    __shared__ WarpSearch wsm[ICP_NT / 32];
    __shared__ int s_task, s_flag;
    extern __shared__ double s_sums[];
    SAcc acc(s_sums);
    const int S = A.n_scales;
    for (;;) {
        if (threadIdx.x == 0) {
            const unsigned int ticket = atomicAdd(&A.qctl[0], 1u);
            volatile int *slot = A.queue + ticket;
            int v;
            unsigned int backoff = 20;
            while ((v = *slot) == 0) { __nanosleep(backoff); if (backoff < 400) backoff += backoff; }
            __threadfence();
            s_task = v;
        }
        __syncthreads();
        const int task = s_task;
        if (task < 0) break;
        const int pair = (task - 1) / A.vmax;
        int chunk = (task - 1) % A.vmax;
        PairState *P = A.ps + pair;
        const int sc = A.pair_src[pair], tc = A.pair_tgt[pair];
        double *pcur = A.pcur + 3 * A.scratch_off[pair];
        double *mcur = A.mcur + 3 * A.scratch_off[pair];
        double4 *anchor = A.anchor + A.scratch_off[pair];
        int2 *prev = A.prev + A.scratch_off[pair];
        double *gpart = A.gpart + (size_t)pair * A.vmax * NACC;
        for (;;) {      // the block that completes a pass continues with chunk 0 of the next one
            const int s = __ldcg(&P->scale), pass = __ldcg(&P->pass), V = __ldcg(&P->V);
            if (threadIdx.x < 16) sM[threadIdx.x] = __ldcg(pass == 0 ? &P->T[threadIdx.x] : &P->U[threadIdx.x]);
            __syncthreads();
            const Job &JS = A.jobs[sc * S + s];
            const Job &JT = A.jobs[tc * S + s];
            const int ns = JS.Mf;
            const double r = A.max_d[pair * S + s];
            if (A.scale_t && pass == 0 && chunk == 0 && threadIdx.x == 0) A.scale_t[((size_t)pair * S + s) * 2] = global_ns();
            {
                const GridView g = make_view(JT, 2);
                double accK = 0.0, accD = 0.0;
                acc.clear();
                const int tid = chunk * ICP_NT + threadIdx.x;
                FusedSink sink{A, JT, acc, accK, accD};
                if (pass == 0) icp_pass_first<true>(A, JS, JT, g, ns, r, sM, tid, V * ICP_NT, pcur, mcur, anchor, prev, wsm[threadIdx.x >> 5], sink);
                else icp_pass_steady<true>(A, JS, JT, g, ns, r, sM, tid, V * ICP_NT, pcur, mcur, anchor, prev, wsm[threadIdx.x >> 5], sink);
                const double part = block_reduce_acc(acc, accK, accD, red);
                if (V == 1) {
                    if (threadIdx.x < NACC) tot[threadIdx.x] = part;
                } else {
                    if (threadIdx.x < NACC) __stcg(gpart + (size_t)chunk * NACC + threadIdx.x, part);
                }
            }
            __syncthreads();
            if (V > 1) {
                if (threadIdx.x == 0) {
                    __threadfence();
                    const bool last = atomicAdd(&P->done, 1u) == (unsigned int)(V - 1);
                    if (last) { *((volatile unsigned int *)&P->done) = 0u; }
                    __threadfence();
                    s_flag = last ? 1 : 0;
                }
                __syncthreads();
                if (!s_flag) break;                         // somebody else completes this pass: next task
                if (threadIdx.x < NACC) {
                    double t = 0.0;
                    for (int c = 0; c < V; ++c) t += __ldcg(gpart + (size_t)c * NACC + threadIdx.x);   // fixed rank order
                    tot[threadIdx.x] = t;
                }
                __syncthreads();
            }
            // ---- this block completed the pass: registration result, convergence test, update ----
            const double K = tot[27], e2 = tot[28];
            double fit = 0.0, rmse = 0.0;
            if (K > 0.0) { fit = K / (double)ns; rmse = sqrt(e2 / K); }
            const double pfit = __ldcg(&P->pfit), prmse = __ldcg(&P->prmse);
            const int max_it = A.max_it[s];
            const bool stop = (pass > 0 && fabs(pfit - fit) < A.rel_fitness && fabs(prmse - rmse) < A.rel_rmse) || pass >= max_it;
            if (threadIdx.x == 0) {
                const double sumK = __ldcg(&P->sumK) + K;
                int finished = 0, Vnext = V;
                if (!stop) {
                    double Told[16], Un[16], Tn[16];
                    for (int i = 0; i < 16; ++i) Told[i] = __ldcg(&P->T[i]);
                    solve_and_update(tot, K, Told, Un, Tn);
                    for (int i = 0; i < 16; ++i) { __stcg(&P->U[i], Un[i]); __stcg(&P->T[i], Tn[i]); }
                    __stcg(&P->pfit, fit); __stcg(&P->prmse, rmse); __stcg(&P->sumK, sumK);
                    __stcg(&P->pass, pass + 1);
                } else {
                    if (A.scale_t) A.scale_t[((size_t)pair * S + s) * 2 + 1] = global_ns();
                    if (A.iters) A.iters[pair * S + s] = pass;
                    if (A.stats) {
                        double *st = A.stats + ((size_t)pair * S + s) * 8;
                        st[0] = (double)ns; st[1] = (double)JT.Mf; st[2] = (double)pass; st[3] = K;
                        st[4] = fit; st[5] = rmse; st[6] = sumK; st[7] = (double)(pass + 1);
                    }
                    const int s2 = next_runnable_scale(A, pair, s + 1);
                    if (s2 >= S) {
                        double T[16];
                        for (int i = 0; i < 16; ++i) T[i] = __ldcg(&P->T[i]);
                        // a trailing scale that could not run reports an empty result, like the static kernel
                        const bool tail_empty = s + 1 < S;
                        start_pairs(A, finish_pair(A, pair, T, tail_empty ? 0.0 : fit, tail_empty ? 0.0 : rmse, tail_empty ? 0.0 : K));
                        finished = 1;
                    } else {
                        __stcg(&P->pfit, 0.0); __stcg(&P->prmse, 0.0); __stcg(&P->sumK, 0.0);
                        __stcg(&P->scale, s2); __stcg(&P->pass, 0);
                        Vnext = chunks_for_scale(A, A.jobs[sc * S + s2].Mf);
                        __stcg(&P->V, Vnext);
                    }
                }
                if (!finished && Vnext > 1) {
                    __threadfence();
                    queue_push_range(A, pair * A.vmax + 2, Vnext - 1, 1);      // chunks 1..V-1 of the next pass
                }
                s_flag = finished;
            }
            __syncthreads();
            if (s_flag) break;
            chunk = 0;
        }
        __syncthreads();      // s_task / s_flag are rewritten by thread 0 at the top of the loop
    }
}

// ---- task mode, warp-specialised (experiment, MGICP_WS=1; measured SLOWER than k_icp_tasks: 65.3 ms against 40.0 ms at 296 pairs) ----
// The fused turn of a point needs ~170 registers (128 with spills), which caps a block at 16 warps; the correspondence half of a
// turn (state, certificates, search) is a chain of dependent loads that wants MANY warps, the linearisation half is dense fp64 that
// wants many REGISTERS.  Here a block is 24 warps: 16 "search" warps (setmaxnreg down) are the 16 virtual warps of the 512-thread
// layout and do everything up to the settled correspondence, which they hand -- point, effective normal, j, d^2, 32 lanes at a
// time -- through a shared-memory ring to 8 "linearise" warps (setmaxnreg up, two rings each), which accumulate into the SAME
// per-virtual-thread columns in the same order: the arithmetic, the sums and their order are those of k_icp_tasks bit for bit
// (tests/test_gpu_parity.py::test_warp_specialised_task_kernel).  Result on the B200: 24 warps resident (38 % instead of 25 %),
// but the 8 linearise warps cannot keep up -- the search warps spend a third of their instructions spinning on full rings -- and
// the register pool of a block is fixed at launch (768 x 80): 128 registers for the linearise warps leave 56 for the search warps,
// which spill.  Kept as a correct, selectable variant and as the record of the experiment; not the default.
constexpr int WS_L = 8, WS_S = 16, WS_NT = (WS_L + WS_S) * 32, WS_DEPTH = 2;
struct WsSlot { double px[32], py[32], pz[32], mx[32], my[32], mz[32], d2[32]; int j[32]; };
struct WsRing { WsSlot slot[WS_DEPTH]; volatile unsigned head, tail; unsigned pad[2]; };
constexpr size_t WS_DYN_SMEM = sizeof(double) * NACC * ICP_NT + sizeof(WsRing) * WS_S;

struct RingSink {
    WsRing &R; unsigned &n; const Job &JT; const int lane;
    __device__ __forceinline__ void operator()(const bool ok, const V3 &p, const V3 &m, const int j, const double d2) {
        while ((int)(n - R.tail) >= WS_DEPTH) { }                  // the consumer is WS_DEPTH slots behind: wait
        WsSlot &S = R.slot[n % WS_DEPTH];
        S.px[lane] = p.x; S.py[lane] = p.y; S.pz[lane] = p.z; S.mx[lane] = m.x; S.my[lane] = m.y; S.mz[lane] = m.z;
        S.d2[lane] = d2; S.j[lane] = ok ? j : -1;
        if (ok) { prefetch_l1(JT.ipts + j); prefetch_l1(JT.inrm + j); }
        __syncwarp();
        ++n;
        if (lane == 0) { __threadfence_block(); R.head = n; }
    }
};

// one linearise warp: the slots of its two rings as they come, na / nb of them this pass
__device__ __forceinline__ void ws_consume(const IcpArgs &A, const Job &JT, WsRing &RA, WsRing &RB, unsigned &ca, unsigned &cb, int na, int nb,
                                           SAcc &accA, SAcc &accB, double &kA, double &dA, double &kB, double &dB, const int lane) {
    while (na | nb) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            WsRing &R = h ? RB : RA;
            unsigned &c = h ? cb : ca;
            int &left = h ? nb : na;
            if (left && (int)(R.head - c) > 0) {
                __threadfence_block();
                const WsSlot &S = R.slot[c % WS_DEPTH];
                const int j = S.j[lane];
                if (j >= 0) {
                    const V3 p = v3(S.px[lane], S.py[lane], S.pz[lane]), m = v3(S.mx[lane], S.my[lane], S.mz[lane]);
                    const double d2 = S.d2[lane];
                    if (h) icp_linearise(A, JT, p, m, j, d2, accB, kB, dB); else icp_linearise(A, JT, p, m, j, d2, accA, kA, dA);
                }
                __syncwarp();
                ++c; --left;
                if (lane == 0) R.tail = c;
            }
        }
    }
}

template <bool PRODUCER>
__device__ __forceinline__ void ws_task_loop(const IcpArgs &A, double *sM, double *tot, double *red, WarpSearch *wsm, int &s_task, int &s_flag,
                                             double *s_sums, WsRing *rings) {
    const int S = A.n_scales;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int vw = PRODUCER ? warp - WS_L : 0;                       // the virtual warp a search warp stands for
    unsigned npush = 0, ca = 0, cb = 0;                              // ring positions, running over passes and tasks
    for (;;) {
        if (!PRODUCER && threadIdx.x == 0) {
            const unsigned int ticket = atomicAdd(&A.qctl[0], 1u);
            volatile int *slot = A.queue + ticket;
            int v;
            unsigned int backoff = 20;
            while ((v = *slot) == 0) { __nanosleep(backoff); if (backoff < 400) backoff += backoff; }
            __threadfence();
            s_task = v;
        }
        __syncthreads();
        const int task = s_task;
        if (task < 0) break;
        const int pair = (task - 1) / A.vmax;
        int chunk = (task - 1) % A.vmax;
        PairState *P = A.ps + pair;
        const int sc = A.pair_src[pair], tc = A.pair_tgt[pair];
        double *pcur = A.pcur + 3 * A.scratch_off[pair];
        double *mcur = A.mcur + 3 * A.scratch_off[pair];
        double4 *anchor = A.anchor + A.scratch_off[pair];
        int2 *prev = A.prev + A.scratch_off[pair];
        double *gpart = A.gpart + (size_t)pair * A.vmax * NACC;
        for (;;) {      // the block that completes a pass continues with chunk 0 of the next one
            const int s = __ldcg(&P->scale), pass = __ldcg(&P->pass), V = __ldcg(&P->V);
            if (threadIdx.x < 16) sM[threadIdx.x] = __ldcg(pass == 0 ? &P->T[threadIdx.x] : &P->U[threadIdx.x]);
            for (int e = threadIdx.x; e < NSUM * ICP_NT; e += WS_NT) s_sums[e] = 0.0;
            __syncthreads();
            const Job &JS = A.jobs[sc * S + s];
            const Job &JT = A.jobs[tc * S + s];
            const int ns = JS.Mf;
            const double r = A.max_d[pair * S + s];
            if (!PRODUCER && A.scale_t && pass == 0 && chunk == 0 && threadIdx.x == 0) A.scale_t[((size_t)pair * S + s) * 2] = global_ns();
            const int nwarps = V * (ICP_NT / 32);
            const int Q = queries_per_warp(ns, nwarps);
            if (PRODUCER) {
                const GridView g = make_view(JT, 2);
                RingSink sink{rings[vw], npush, JT, lane};
                const int tid = chunk * ICP_NT + vw * 32 + lane;
                if (pass == 0) icp_pass_first<true>(A, JS, JT, g, ns, r, sM, tid, V * ICP_NT, pcur, mcur, anchor, prev, wsm[vw], sink);
                else icp_pass_steady<true>(A, JS, JT, g, ns, r, sM, tid, V * ICP_NT, pcur, mcur, anchor, prev, wsm[vw], sink);
            } else {
                // turns of virtual warp gw this pass: ib = gw * Q, gw * Q + nwarps * Q, ... < ns
                const int gwa = chunk * (ICP_NT / 32) + 2 * warp, gwb = gwa + 1;
                const int na = gwa * Q < ns ? (ns - gwa * Q + nwarps * Q - 1) / (nwarps * Q) : 0;
                const int nb = gwb * Q < ns ? (ns - gwb * Q + nwarps * Q - 1) / (nwarps * Q) : 0;
                SAcc accA(s_sums, (2 * warp) * 32 + lane), accB(s_sums, (2 * warp + 1) * 32 + lane);
                double kA = 0.0, dA = 0.0, kB = 0.0, dB = 0.0;
                ws_consume(A, JT, rings[2 * warp], rings[2 * warp + 1], ca, cb, na, nb, accA, accB, kA, dA, kB, dB, lane);
                accA.col[NSUM * ICP_NT] = kA; accA.col[(NSUM + 1) * ICP_NT] = dA;
                accB.col[NSUM * ICP_NT] = kB; accB.col[(NSUM + 1) * ICP_NT] = dB;
            }
            __syncthreads();
            // block reduction over the [NACC][512] rows: the arithmetic of block_reduce_acc, rows dealt to all 24 warps
            for (int a = warp; a < NACC; a += WS_NT / 32) {
                const double *row = s_sums + a * ICP_NT;
                double v = row[lane];
#pragma unroll
                for (int i = 1; i < ICP_NT / 32; ++i) v += row[lane + 32 * i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0) red[a] = v;
            }
            __syncthreads();
            const double part = threadIdx.x < NACC ? red[threadIdx.x] : 0.0;
            if (V == 1) {
                if (threadIdx.x < NACC) tot[threadIdx.x] = part;
            } else {
                if (threadIdx.x < NACC) __stcg(gpart + (size_t)chunk * NACC + threadIdx.x, part);
            }
            __syncthreads();
            if (V > 1) {
                if (!PRODUCER && threadIdx.x == 0) {
                    __threadfence();
                    const bool last = atomicAdd(&P->done, 1u) == (unsigned int)(V - 1);
                    if (last) { *((volatile unsigned int *)&P->done) = 0u; }
                    __threadfence();
                    s_flag = last ? 1 : 0;
                }
                __syncthreads();
                if (!s_flag) break;                         // somebody else completes this pass: next task
                if (threadIdx.x < NACC) {
                    double t = 0.0;
                    for (int c = 0; c < V; ++c) t += __ldcg(gpart + (size_t)c * NACC + threadIdx.x);   // fixed rank order
                    tot[threadIdx.x] = t;
                }
                __syncthreads();
            }
            // ---- this block completed the pass: registration result, convergence test, update ----
            if (!PRODUCER && threadIdx.x == 0) {
                const double K = tot[27], e2 = tot[28];
                double fit = 0.0, rmse = 0.0;
                if (K > 0.0) { fit = K / (double)ns; rmse = sqrt(e2 / K); }
                const double pfit = __ldcg(&P->pfit), prmse = __ldcg(&P->prmse);
                const int max_it = A.max_it[s];
                const bool stop = (pass > 0 && fabs(pfit - fit) < A.rel_fitness && fabs(prmse - rmse) < A.rel_rmse) || pass >= max_it;
                const double sumK = __ldcg(&P->sumK) + K;
                int finished = 0, Vnext = V;
                if (!stop) {
                    double Told[16], Un[16], Tn[16];
                    for (int i = 0; i < 16; ++i) Told[i] = __ldcg(&P->T[i]);
                    solve_and_update(tot, K, Told, Un, Tn);
                    for (int i = 0; i < 16; ++i) { __stcg(&P->U[i], Un[i]); __stcg(&P->T[i], Tn[i]); }
                    __stcg(&P->pfit, fit); __stcg(&P->prmse, rmse); __stcg(&P->sumK, sumK);
                    __stcg(&P->pass, pass + 1);
                } else {
                    if (A.scale_t) A.scale_t[((size_t)pair * S + s) * 2 + 1] = global_ns();
                    if (A.iters) A.iters[pair * S + s] = pass;
                    if (A.stats) {
                        double *st = A.stats + ((size_t)pair * S + s) * 8;
                        st[0] = (double)ns; st[1] = (double)JT.Mf; st[2] = (double)pass; st[3] = K;
                        st[4] = fit; st[5] = rmse; st[6] = sumK; st[7] = (double)(pass + 1);
                    }
                    const int s2 = next_runnable_scale(A, pair, s + 1);
                    if (s2 >= S) {
                        double T[16];
                        for (int i = 0; i < 16; ++i) T[i] = __ldcg(&P->T[i]);
                        const bool tail_empty = s + 1 < S;
                        start_pairs(A, finish_pair(A, pair, T, tail_empty ? 0.0 : fit, tail_empty ? 0.0 : rmse, tail_empty ? 0.0 : K));
                        finished = 1;
                    } else {
                        __stcg(&P->pfit, 0.0); __stcg(&P->prmse, 0.0); __stcg(&P->sumK, 0.0);
                        __stcg(&P->scale, s2); __stcg(&P->pass, 0);
                        Vnext = chunks_for_scale(A, A.jobs[sc * S + s2].Mf);
                        __stcg(&P->V, Vnext);
                    }
                }
                if (!finished && Vnext > 1) {
                    __threadfence();
                    queue_push_range(A, pair * A.vmax + 2, Vnext - 1, 1);      // chunks 1..V-1 of the next pass
                }
                s_flag = finished;
            }
            __syncthreads();
            if (s_flag) break;
            chunk = 0;
        }
        __syncthreads();      // s_task / s_flag are rewritten by thread 0 at the top of the loop
    }
}

__global__ void __launch_bounds__(WS_NT, 1) k_icp_tasks_ws(IcpArgs A) {
    __shared__ double sM[16], tot[32];
    __shared__ double red[32];
    __shared__ WarpSearch wsm[WS_S];
    __shared__ int s_task, s_flag;
    extern __shared__ double s_sums[];
    WsRing *rings = reinterpret_cast<WsRing *>(s_sums + NACC * ICP_NT);
    if (threadIdx.x < WS_S) { rings[threadIdx.x].head = 0u; rings[threadIdx.x].tail = 0u; }
    __syncthreads();
    if ((threadIdx.x >> 5) < WS_L) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        ws_task_loop<false>(A, sM, tot, red, wsm, s_task, s_flag, s_sums, rings);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        ws_task_loop<true>(A, sM, tot, red, wsm, s_task, s_flag, s_sums, rings);
    }
}

// =============================================================================================
// evaluate_registration / get_information_matrix_from_point_clouds on the clouds as given
// =============================================================================================
// grid (chunks, jobs): the raw cloud becomes the "final cloud" of its job (no down-sampling, no filter, no normals), so
// that the ICP-grid build kernels can hash it
__global__ void __launch_bounds__(256) k_raw_load(Job *jobs) {
    Job &J = jobs[blockIdx.y];
    if (J.err) return;
    const int64_t n = J.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        load_point(J.xyz, J.dtype, i, x, y, z);
        J.pts[i] = make_double4(x, y, z, (double)i);
        J.nrm[i] = make_double4(0.0, 0.0, 0.0, 0.0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int ibits = 10;
        while (ibits < J.cbits_max && ((int64_t)1 << ibits) < 4 * n) ++ibits;
        J.ibits = ibits;
        J.Mf = (int32_t)n;
    }
}

__global__ void k_eval_set_n(Job *jobs, const int64_t *cloud_off, int n_clouds) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_clouds) jobs[c].n = cloud_off[c + 1] - cloud_off[c];
}

constexpr int EVAL_NT = 256;
constexpr int NEV = 23;   // K, sum d^2, 21 upper-triangular terms of GTG

struct EvalArgs {
    const Job *jobs;                       // one job per cloud
    const int32_t *pair_src, *pair_tgt;
    const double *max_d;                   // [pairs]
    const double *T;                       // [pairs * 16]
    double *part;                          // [pairs][chunks][NEV]
    double *out;                           // [pairs * 32]
    int32_t *corr;                         // optional [sum of source sizes]: target index per source point, -1 = none
    const int64_t *corr_off;               // [pairs]
};

// grid (chunks, pairs): GetRegistrationResultAndCorrespondences at pose T on the raw clouds; thread-strided partial sums,
// fixed shuffle tree, warps in order -> one partial per block
__global__ void __launch_bounds__(EVAL_NT) k_eval_clouds(EvalArgs E) {
    __shared__ WarpSearch wsm[EVAL_NT / 32];
    __shared__ double red[EVAL_NT / 32][NEV];
    __shared__ double sT[16];
    const int pair = blockIdx.y;
    const Job &JS = E.jobs[E.pair_src[pair]];
    const Job &JT = E.jobs[E.pair_tgt[pair]];
    if (threadIdx.x < 16) sT[threadIdx.x] = E.T[pair * 16 + threadIdx.x];
    __syncthreads();
    bool ident = true;
#pragma unroll
    for (int i = 0; i < 16; ++i) ident &= (sT[i] == ((i % 5 == 0) ? 1.0 : 0.0));
    const GridView g = make_view(JT, 2);
    const double r = E.max_d[pair], r2 = r * r;
    const int64_t ns = JS.n;
    const int nt = JT.Mf;
    double acc[NEV];
#pragma unroll
    for (int a = 0; a < NEV; ++a) acc[a] = 0.0;
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = (int64_t)blockIdx.x * EVAL_NT + threadIdx.x - lane; i0 < ns; i0 += (int64_t)gridDim.x * EVAL_NT) {   // warp-uniform
        const int64_t i = i0 + lane;
        const bool have = i < ns && nt > 0;
        V3 p = v3(0, 0, 0);
        if (i < ns) {
            load_point(JS.xyz, JS.dtype, i, p.x, p.y, p.z);
            if (!ident) p = transform_point(sT, p);
        }
        double d2 = r2;
        int j = -1;
        nn_search_coop(g, wsm[threadIdx.x >> 5], have, p.x, p.y, p.z, r2, d2, j);
        const bool matched = have && j >= 0 && d2 < r2;
        if (i < ns && E.corr) E.corr[E.corr_off[pair] + i] = matched ? JT.i2a[j] : -1;
        if (matched) {
            const double4 q = ldg4(g.pts + j);
            const double x = q.x, y = q.y, z = q.z;
            acc[0] += 1.0;
            acc[1] += d2;
            // rows (0, z, -y, 1, 0, 0), (-z, 0, x, 0, 1, 0), (y, -x, 0, 0, 0, 1): upper triangle of sum G_r G_r^T, row-major
            acc[2] += z * z + y * y;  acc[3] += -(x * y);        acc[4] += -(x * z);        acc[5] += 0.0;  acc[6] += -z;   acc[7] += y;
            acc[8] += z * z + x * x;  acc[9] += -(y * z);        acc[10] += z;              acc[11] += 0.0; acc[12] += -x;
            acc[13] += y * y + x * x; acc[14] += -y;             acc[15] += x;              acc[16] += 0.0;
            acc[17] += 1.0;           acc[18] += 0.0;            acc[19] += 0.0;
            acc[20] += 1.0;           acc[21] += 0.0;
            acc[22] += 1.0;
        }
    }
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NEV; ++a) {
        double v = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[w][a] = v;
    }
    __syncthreads();
    if (threadIdx.x < NEV) {
        double t = 0.0;
        for (int i = 0; i < EVAL_NT / 32; ++i) t += red[i][threadIdx.x];
        E.part[((size_t)pair * gridDim.x + blockIdx.x) * NEV + threadIdx.x] = t;
    }
}

// one warp per pair: block partials in order -> fitness, rmse, K, sum d^2, GTG
__global__ void k_eval_finish(EvalArgs E, int chunks) {
    const int pair = blockIdx.x;
    if (threadIdx.x >= NEV) return;
    double t = 0.0;
    for (int c = 0; c < chunks; ++c) t += E.part[((size_t)pair * chunks + c) * NEV + threadIdx.x];
    double *o = E.out + (size_t)pair * 32;
    if (threadIdx.x >= 2) o[4 + (threadIdx.x - 2)] = t;
    const double K = __shfl_sync(0x7fffffu, t, 0), e2 = __shfl_sync(0x7fffffu, t, 1);
    if (threadIdx.x == 0) {
        const double ns = (double)E.jobs[E.pair_src[pair]].n;
        o[0] = K > 0.0 ? K / ns : 0.0;
        o[1] = K > 0.0 ? sqrt(e2 / K) : 0.0;
        o[2] = K; o[3] = e2;
        for (int a = 25; a < 32; ++a) o[a] = 0.0;
    }
}

// =============================================================================================
// host side
// =============================================================================================
struct mgicp_handle_s {
    int device = 0;
    std::string err;
    int64_t launches = 0;
    // workspace
    char *arena = nullptr; size_t arena_bytes = 0;
    char *scratch = nullptr; size_t scratch_bytes = 0;
    char *small = nullptr; size_t small_bytes = 0;    // per-call small device arrays
    // state of the last preprocess
    int n_clouds = 0, n_scales = 0;
    std::vector<Job> jobs_host;      // static part mirror
    std::vector<int64_t> cloud_n;
    Job *jobs_dev = nullptr;
    u64 *benc = nullptr;
    int64_t *cloud_off_dev = nullptr;
    bool preprocessed = false;
    Job *eval_jobs = nullptr; int eval_n = 0;   // jobs of the last mgicp_evaluate_clouds (error flags for mgicp_check)
    mgicp_opts opts;
    // optional stage timing (mgicp_set_timing): events at the stage boundaries of the last preprocess + register
    bool timing = false;
    cudaEvent_t tev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int tev_n = 0;                              // events recorded by the last preprocess (+ register)
    unsigned long long *scale_t = nullptr;      // device [pairs * scales * 2] globaltimer at the start / end of a pair's scale
    int scale_t_pairs = 0;
    // the last mgicp_register_batch: where each pair's final correspondences live (mgicp_get_correspondences)
    const int2 *last_prev = nullptr;
    std::vector<int64_t> last_soff;
    std::vector<int32_t> last_src, last_tgt;
};

static void tmark(mgicp_handle h, cudaStream_t st, int i) {
    if (!h->timing) return;
    if (!h->tev[i]) cudaEventCreate(&h->tev[i]);
    cudaEventRecord(h->tev[i], st);
    h->tev_n = i + 1;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return MGICP_E_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

extern "C" void mgicp_default_opts(mgicp_opts *o) {
    o->sor_k = 30; o->sor_std = 1.0; o->normal_k = 20; o->epsilon = 1e-3; o->loss = MGICP_LOSS_L1; o->loss_k = 1.0;
    o->rel_fitness = 1e-6; o->rel_rmse = 1e-6; o->cell_factor = 0.0; o->icp_cell_factor = 0.0; o->ctas_per_pair = 0; o->debug = 0;
}

extern "C" const char *mgicp_version(void) { return "mgicp-b200 0.1 (sm_100a)"; }

extern "C" int mgicp_create(int device, mgicp_handle *out) {
    if (!out) return MGICP_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return MGICP_E_CUDA;
    mgicp_handle h = new mgicp_handle_s();
    h->device = device;
    mgicp_default_opts(&h->opts);
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return MGICP_E_CUDA; }
    *out = h;
    return MGICP_OK;
}

extern "C" int mgicp_destroy(mgicp_handle h) {
    if (!h) return MGICP_OK;
    cudaSetDevice(h->device);
    cudaFree(h->arena); cudaFree(h->scratch); cudaFree(h->small);
    for (cudaEvent_t e : h->tev) if (e) cudaEventDestroy(e);
    delete h;
    return MGICP_OK;
}

extern "C" const char *mgicp_last_error(mgicp_handle h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t mgicp_kernel_launches(mgicp_handle h) { return h ? h->launches : 0; }

static int grow(mgicp_handle h, char **buf, size_t *have, size_t need) {
    if (need <= *have) return MGICP_OK;
    CK(cudaDeviceSynchronize());
    if (*buf) CK(cudaFree(*buf));
    *buf = nullptr; *have = 0;
    size_t want = need + need / 8;
    cudaError_t e = cudaMalloc((void **)buf, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = need;
        e = cudaMalloc((void **)buf, want);
    }
    if (e != cudaSuccess) { cudaGetLastError(); h->err = "workspace allocation failed"; return MGICP_E_NOMEM; }
    *have = want;
    return MGICP_OK;
}

static int chunks_for(int64_t n, int per_block, int cap) {
    int64_t c = (n + per_block - 1) / per_block;
    if (c < 1) c = 1;
    if (c > cap) c = cap;
    return (int)c;
}

extern "C" int mgicp_cloud_bounds(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                                  int32_t xyz_dtype, double *bounds_out) {
    if (!h) return MGICP_E_INVALID;
    if (n_clouds <= 0 || !cloud_off || !bounds_out) { h->err = "mgicp_cloud_bounds: bad arguments"; return MGICP_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    size_t need = align_up(sizeof(u64) * BENC_W * n_clouds) + align_up(sizeof(int64_t) * (n_clouds + 1));
    int rc = grow(h, &h->small, &h->small_bytes, need);
    if (rc) return rc;
    u64 *benc = (u64 *)h->small;
    int64_t *off = (int64_t *)(h->small + align_up(sizeof(u64) * BENC_W * n_clouds));
    CK(cudaMemcpyAsync(off, cloud_off, sizeof(int64_t) * (n_clouds + 1), cudaMemcpyHostToDevice, st));
    int64_t maxn = 0;
    for (int c = 0; c < n_clouds; ++c) maxn = std::max(maxn, cloud_off[c + 1] - cloud_off[c]);
    k_bounds_init<<<(n_clouds * BENC_W + 127) / 128, 128, 0, st>>>(benc, n_clouds);
    k_bounds<<<dim3(chunks_for(maxn, 256 * 8, 64), n_clouds), 256, 0, st>>>(xyz, xyz_dtype, off, benc);
    k_bounds_decode<<<(n_clouds * 6 + 127) / 128, 128, 0, st>>>(benc, bounds_out, n_clouds * 6);
    h->launches += 3;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));   // `small` may be reused by the next call
    return MGICP_OK;
}

extern "C" int mgicp_preprocess(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                                int32_t xyz_dtype, int32_t n_scales, const double *voxel_sizes, const mgicp_opts *opts_in) {
    if (!h) return MGICP_E_INVALID;
    h->preprocessed = false;
    h->eval_jobs = nullptr;
    mgicp_opts o;
    if (opts_in) o = *opts_in; else mgicp_default_opts(&o);
    if (n_clouds <= 0 || n_scales <= 0 || !cloud_off || !voxel_sizes || (xyz_dtype != MGICP_F32 && xyz_dtype != MGICP_F64)) {
        h->err = "mgicp_preprocess: bad arguments"; return MGICP_E_INVALID;
    }
    for (int s = 0; s < n_scales; ++s)
        if (!(voxel_sizes[s] > 0.0)) { h->err = "voxel_size <= 0"; return MGICP_E_INVALID; }
    if (o.sor_k < 1 || o.sor_k > 32 || !(o.sor_std > 0.0) || o.normal_k < 1 || o.normal_k > 32) {
        h->err = "sor_k and normal_k must be in 1..32 (the warp-wide neighbour list), sor_std > 0"; return MGICP_E_INVALID;
    }
    const double cf = o.cell_factor > 0.0 ? o.cell_factor : 12.0;         // kNN grid: measured with k_knn_hist, step ms for factors 8 / 10 / 12 / 14 / 16 / 20: 81.0 / 77.6 / 76.2 / 76.6 / 77.5 / 80.0
    const double cfi = o.icp_cell_factor > 0.0 ? o.icp_cell_factor : 3.5;  // ICP grid: search radius <= 3 voxels in script 2's schedule; ICP ms for factors 2 / 2.5 / 3 / 3.5 / 4.5: 44.2 / 42.2 / 40.7 / 39.7 / 39.9
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    const int J = n_clouds * n_scales;
    const size_t esz = xyz_dtype == MGICP_F32 ? 4 : 8;
    // ---- lay the workspace out --------------------------------------------------------------
    std::vector<Job> jobs(J);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o_ = off; off += align_up(bytes); return o_; };
    const size_t o_jobs = take(sizeof(Job) * J);
    const size_t o_benc = take(sizeof(u64) * BENC_W * n_clouds);
    const size_t o_coff = take(sizeof(int64_t) * (n_clouds + 1));
    const size_t o_vkeys_begin = off;
    std::vector<size_t> o_vkeys(J);
    int64_t maxn = 0;
    for (int c = 0; c < n_clouds; ++c) {
        const int64_t n = cloud_off[c + 1] - cloud_off[c];
        if (n < 0 || n > (int64_t)1 << 30) { h->err = "cloud too large"; return MGICP_E_INVALID; }
        maxn = std::max(maxn, n);
        for (int s = 0; s < n_scales; ++s) {
            Job &j = jobs[c * n_scales + s];
            memset(&j, 0, sizeof(Job));
            j.xyz = (const char *)xyz + (size_t)cloud_off[c] * 3 * esz;
            j.dtype = xyz_dtype; j.cloud = c; j.n = n;
            j.voxel = voxel_sizes[s]; j.cell = cf * voxel_sizes[s]; j.cell_i = cfi * voxel_sizes[s];
            int vb = 10; while (((int64_t)1 << vb) < 2 * n) ++vb;
            int cb = 10; while (((int64_t)1 << cb) < 4 * n) ++cb;
            j.vbits = vb; j.cbits_max = cb;
            o_vkeys[c * n_scales + s] = take(sizeof(u64) * (((size_t)1 << vb) + TAB_PAD));
        }
    }
    const size_t o_vkeys_end = off;
    std::vector<size_t> offs(J * 20);
    for (int jx = 0; jx < J; ++jx) {
        Job &j = jobs[jx];
        const size_t n = (size_t)std::max<int64_t>(j.n, 1);
        const size_t vcap = ((size_t)1 << j.vbits) + TAB_PAD, ccap = ((size_t)1 << j.cbits_max) + TAB_PAD;
        size_t *o_ = &offs[jx * 20];
        o_[0] = take(sizeof(int32_t) * vcap);      // vrank
        o_[1] = take(sizeof(double) * 3 * n);      // vsum
        o_[2] = take(sizeof(int32_t) * n);         // vcnt
        o_[3] = take(sizeof(double) * 3 * n);      // ds
        o_[4] = take(sizeof(CellSlot) * ccap);     // ctab
        o_[5] = take(sizeof(CellSlot) * ccap);     // ftab
        o_[6] = take(sizeof(int32_t) * ccap);      // ccursor
        o_[7] = take(sizeof(int32_t) * n);         // pslot
        o_[8] = take(sizeof(int32_t) * n);         // order
        o_[9] = take(sizeof(double4) * n);         // gpts
        o_[10] = take(sizeof(double) * n);         // avg
        o_[11] = take(n);                          // keep
        o_[12] = take(sizeof(int32_t) * (n + 1));  // newidx
        o_[13] = take(sizeof(double4) * n);        // pts
        o_[14] = take(sizeof(double4) * n);        // nrm
        o_[18] = take(sizeof(CellSlot) * ccap);    // itab
        o_[19] = take(sizeof(double4) * n * 2);    // ipts, inrm
        o_[16] = take(sizeof(int32_t) * n * o.sor_k);
        o_[15] = take(sizeof(int32_t) * n * 18);   // nbrA (8n), inbr (8n), a2i (n), i2a (n)
        o_[17] = o.debug ? take(sizeof(int32_t) * n * o.normal_k) : 0;
    }
    int rc = grow(h, &h->arena, &h->arena_bytes, off);
    if (rc) return rc;
    char *base = h->arena;
    auto n_of = [](const Job &j) { return (size_t)std::max<int64_t>(j.n, 1); };
    for (int jx = 0; jx < J; ++jx) {
        Job &j = jobs[jx];
        size_t *o_ = &offs[jx * 20];
        j.vkeys = (u64 *)(base + o_vkeys[jx]);
        j.vrank = (int32_t *)(base + o_[0]); j.vsum = (double *)(base + o_[1]); j.vcnt = (int32_t *)(base + o_[2]);
        j.ds = (double *)(base + o_[3]); j.ctab = (CellSlot *)(base + o_[4]); j.ftab = (CellSlot *)(base + o_[5]);
        j.ccursor = (int32_t *)(base + o_[6]); j.pslot = (int32_t *)(base + o_[7]); j.order = (int32_t *)(base + o_[8]);
        j.gpts = (double4 *)(base + o_[9]); j.avg = (double *)(base + o_[10]); j.keep = (uint8_t *)(base + o_[11]);
        j.newidx = (int32_t *)(base + o_[12]); j.pts = (double4 *)(base + o_[13]); j.nrm = (double4 *)(base + o_[14]);
        j.fb_list = (int32_t *)(base + o_[7]);      // shares pslot: the down-sampled cloud's slots are dead once its grid is built
        j.nbrA = (int32_t *)(base + o_[15]); j.inbr = j.nbrA + 8 * n_of(j); j.a2i = j.inbr + 8 * n_of(j); j.i2a = j.a2i + n_of(j);
        j.itab = (CellSlot *)(base + o_[18]);
        j.ipts = (double4 *)(base + o_[19]); j.inrm = j.ipts + n_of(j);
        j.knn_sor = (int32_t *)(base + o_[16]);
        j.knn_nrm = o.debug ? (int32_t *)(base + o_[17]) : nullptr;
    }
    h->jobs_dev = (Job *)(base + o_jobs);
    h->benc = (u64 *)(base + o_benc);
    h->cloud_off_dev = (int64_t *)(base + o_coff);
    h->jobs_host = jobs;
    h->n_clouds = n_clouds; h->n_scales = n_scales; h->opts = o;
    h->cloud_n.assign(n_clouds, 0);
    for (int c = 0; c < n_clouds; ++c) h->cloud_n[c] = cloud_off[c + 1] - cloud_off[c];
    // jobs_host / cloud_off are pageable host memory: cudaMemcpyAsync from pageable memory returns after staging
    CK(cudaMemcpyAsync(h->jobs_dev, jobs.data(), sizeof(Job) * J, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->cloud_off_dev, cloud_off, sizeof(int64_t) * (n_clouds + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(base + o_vkeys_begin, 0xFF, o_vkeys_end - o_vkeys_begin, st));
    // ---- launches ---------------------------------------------------------------------------
    const int cx_raw = chunks_for(maxn, 256 * 8, 128);
    const int cx_pts = chunks_for(maxn, 256 * 2, 256);
    const int cx_knn = std::max(1, std::min(chunks_for(maxn, 8 * 4, 2048), std::max(16, 16384 / J)));   // 8 warps per block, >= 4 queries per warp
    int knn_mode = 1;                                     // 1: k_knn_hist (histogram selection), 0: k_knn (running top-k list)
    if (const char *e = getenv("MGICP_KNN_MODE")) knn_mode = atoi(e);                                      // A/B experiments
    CK(cudaFuncSetAttribute(k_knn_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KH_WARPS * sizeof(KhWarp))));
    tmark(h, st, 0);
    k_bounds_init<<<(n_clouds * BENC_W + 127) / 128, 128, 0, st>>>(h->benc, n_clouds);
    k_bounds<<<dim3(chunks_for(maxn, 256 * 8, 64), n_clouds), 256, 0, st>>>(xyz, xyz_dtype, h->cloud_off_dev, h->benc);
    k_job_setup<<<(J + 127) / 128, 128, 0, st>>>(h->jobs_dev, J, h->benc);
    k_vox_insert<<<dim3(cx_raw, J), 256, 0, st>>>(h->jobs_dev);
    k_vox_scan<<<J, 1024, 0, st>>>(h->jobs_dev);
    k_table_clear<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 0);
    k_vox_accum<<<dim3(cx_raw, J), 256, 0, st>>>(h->jobs_dev);
    if (xyz_dtype == MGICP_F64) {          // genuine float64 coordinates: the sums redone in input order (deterministic, Open3D's order)
        k_vox_ord_scan<<<J, 1024, 0, st>>>(h->jobs_dev);
        k_vox_ord_scatter<<<dim3(cx_raw, J), 256, 0, st>>>(h->jobs_dev);
        k_vox_ord_sum<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev);
        h->launches += 3;
    }
    k_vox_final<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev);
    tmark(h, st, 1);
    k_cell_count<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 0);
    k_cell_scan<<<J, 1024, 0, st>>>(h->jobs_dev, 0);
    k_cell_scatter<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 0);
    k_cell_gather<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 0);
    h->launches += 12;
    tmark(h, st, 2);
    if (knn_mode == 0) {
        k_knn<<<dim3(cx_knn, J), 256, 0, st>>>(h->jobs_dev, o.sor_k, 0);
    } else {
        k_knn_hist<<<dim3(cx_knn, J), KH_WARPS * 32, KH_WARPS * sizeof(KhWarp), st>>>(h->jobs_dev, o.sor_k);
    }
    k_sor_select<<<J, 1024, 0, st>>>(h->jobs_dev, o.sor_std);
    k_ftab_build<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev);
    k_table_clear<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 2);
    tmark(h, st, 3);
    k_normals<<<dim3(cx_pts, J), NRM_NT, 0, st>>>(h->jobs_dev, o.normal_k, o.sor_k, o.debug);
    k_normals_search<<<dim3(std::min(cx_knn, 64), J), 256, 0, st>>>(h->jobs_dev, o.normal_k, o.debug);
    h->launches += 1;
    tmark(h, st, 4);
    // ICP grid over the final cloud (its own cell size); points and normals are re-gathered into its order
    k_cell_insert<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 2);
    k_cell_count<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 2);
    k_cell_scan<<<J, 1024, 0, st>>>(h->jobs_dev, 2);
    k_cell_scatter<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 2);
    k_cell_gather<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev, 2);
    k_nbr_remap<<<dim3(cx_pts, J), 256, 0, st>>>(h->jobs_dev);
    h->launches += 11;
    tmark(h, st, 5);
    CK(cudaGetLastError());
    h->preprocessed = true;
    return MGICP_OK;
}

static int icp_launch(mgicp_handle h, cudaStream_t st, int32_t n_pairs, const int32_t *pair_src, const int32_t *pair_tgt,
                      const double *max_dists, const int32_t *max_iters, const mgicp_opts *opts_in, const double *T_init,
                      double *T_out, double *fitness, double *rmse, int32_t *iters, int32_t *ncorr, double *stats,
                      int eval_scale, double *eval_out) {
    if (!h->preprocessed) { h->err = "mgicp_register_batch: call mgicp_preprocess first"; return MGICP_E_STATE; }
    mgicp_opts o;
    if (opts_in) o = *opts_in; else o = h->opts;
    const int S = h->n_scales;
    const int ns_md = eval_scale >= 0 ? 1 : S;
    if (n_pairs <= 0 || !pair_src || !pair_tgt || !max_dists || !T_init) { h->err = "register: bad arguments"; return MGICP_E_INVALID; }
    for (int i = 0; i < n_pairs * ns_md; ++i)
        if (!(max_dists[i] > 0.0)) { h->err = "max_correspondence_distance <= 0"; return MGICP_E_INVALID; }
    for (int i = 0; i < n_pairs; ++i)
        if (pair_src[i] < 0 || pair_src[i] >= h->n_clouds || pair_tgt[i] < 0 || pair_tgt[i] >= h->n_clouds) {
            h->err = "pair index out of range"; return MGICP_E_INVALID;
        }
    CK(cudaSetDevice(h->device));
    // Few pairs: every pair gets a static gang of blocks (cooperative launch => co-resident => the global-memory barrier is
    // safe), which minimises the latency of a pass.  Batches: task mode, passes cut into V chunks pulled from a queue by
    // persistent blocks (dynamic load balancing; arithmetic identical to a static gang of V).
    // opts.ctas_per_pair: > 0 static gang of that size, < 0 task mode with V = -ctas_per_pair, 0 automatic.
    int dev_sms = 0, occ = 0, occ_t = 0;
    CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device));
    CK(cudaFuncSetAttribute(k_icp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ICP_DYN_SMEM));
    CK(cudaFuncSetAttribute(k_icp_tasks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ICP_DYN_SMEM));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_icp, ICP_NT, ICP_DYN_SMEM));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, k_icp_tasks, ICP_NT, ICP_DYN_SMEM));
    const int resident = std::max(1, dev_sms * std::max(1, occ));
    if (const char *e = getenv("MGICP_TASK_OCC")) occ_t = std::max(1, std::min(occ_t, atoi(e)));                  // co-residency experiments
    const int resident_t = std::max(1, dev_sms * std::max(1, occ_t));
    int chunk_points = 4096;
    if (const char *e = getenv("MGICP_CHUNK_POINTS")) chunk_points = std::max(256, atoi(e));                        // tuning experiments
    int gang = o.ctas_per_pair;
    bool tasks = false, adaptive = false;
    if (eval_scale >= 0) {
        gang = std::max(1, std::min(resident / n_pairs, 96));
    } else if (gang < 0) {
        tasks = true; gang = std::min(-gang, 64);
    } else if (gang == 0) {
        if ((long long)n_pairs * 8 <= resident) gang = std::min(resident / n_pairs, 96);
        else {
            // gang = the cap of the per-scale chunk count (no scale has more points than the largest raw cloud)
            tasks = true; adaptive = true;
            int64_t max_n = 1;
            for (int i = 0; i < n_pairs; ++i) max_n = std::max(max_n, h->cloud_n[pair_src[i]]);
            const long long num = n_pairs >= resident_t ? 1 : resident_t;
            const long long den = n_pairs >= resident_t ? 2LL * chunk_points : (long long)chunk_points * n_pairs;
            gang = (int)std::max(1LL, std::min(16LL, ((long long)max_n * num + den / 2) / den));
        }
    }
    if (!tasks && gang > 1 && (long long)gang * n_pairs > resident) gang = std::max(1, resident / n_pairs);
    // scratch: per pair, capacity = source cloud size
    std::vector<int64_t> soff(n_pairs + 1, 0);
    for (int i = 0; i < n_pairs; ++i) soff[i + 1] = soff[i] + std::max<int64_t>(h->cloud_n[pair_src[i]], 1);
    const size_t tot_pts = (size_t)soff[n_pairs];
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o_ = off; off += align_up(bytes); return o_; };
    const size_t o_p = take(sizeof(double) * 3 * tot_pts), o_m = take(sizeof(double) * 3 * tot_pts), o_prev = take(sizeof(int2) * tot_pts);
    const size_t o_an = take(sizeof(double4) * tot_pts);
    const size_t o_soff = take(sizeof(int64_t) * (n_pairs + 1));
    const size_t o_ps = take(sizeof(int32_t) * n_pairs), o_pt = take(sizeof(int32_t) * n_pairs);
    const size_t o_md = take(sizeof(double) * n_pairs * S), o_mi = take(sizeof(int32_t) * S);
    const size_t o_sync = take(sizeof(unsigned int) * 2 * n_pairs);
    const size_t o_gpart = take(sizeof(double) * (size_t)n_pairs * 2 * gang * NACC);
    // task mode: pair states, control words, and a slot per task ever published (no wrap-around: V per pair at the start,
    // V-1 per further pass, one exit token per block)
    long long pass_cap = 0;
    if (max_iters) for (int s = 0; s < S; ++s) pass_cap += (long long)std::max(max_iters[s], 0) + 1;
    const int n_ctas = tasks ? (int)std::min<long long>(resident_t, (long long)n_pairs * gang) : 0;
    const size_t q_slots = tasks ? (size_t)n_pairs * gang + (size_t)n_pairs * (gang - 1) * (size_t)pass_cap + n_ctas + 64 : 0;
    const size_t o_scalet = take(h->timing && eval_scale < 0 ? sizeof(unsigned long long) * 2 * (size_t)n_pairs * S : 0);
    const size_t o_pstate = take(tasks ? sizeof(PairState) * n_pairs : 0);
    const size_t o_qctl = take(tasks ? 256 : 0);
    const size_t o_queue = take(sizeof(int) * q_slots);
    int rc = grow(h, &h->scratch, &h->scratch_bytes, off);
    if (rc) return rc;
    char *b = h->scratch;
    std::vector<double> md(n_pairs * S, 1.0);
    if (eval_scale >= 0) { for (int i = 0; i < n_pairs; ++i) md[i * S + eval_scale] = max_dists[i]; }
    else md.assign(max_dists, max_dists + (size_t)n_pairs * S);
    std::vector<int32_t> mi(S, 0);
    if (max_iters) mi.assign(max_iters, max_iters + S);
    CK(cudaMemcpyAsync(b + o_soff, soff.data(), sizeof(int64_t) * (n_pairs + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_ps, pair_src, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_pt, pair_tgt, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_md, md.data(), sizeof(double) * n_pairs * S, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_mi, mi.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice, st));
    IcpArgs A;
    A.jobs = h->jobs_dev; A.n_scales = S;
    A.pair_src = (const int32_t *)(b + o_ps); A.pair_tgt = (const int32_t *)(b + o_pt);
    A.max_d = (const double *)(b + o_md); A.max_it = (const int32_t *)(b + o_mi);
    A.T_init = T_init; A.T_out = T_out; A.fitness = fitness; A.rmse = rmse; A.iters = iters; A.ncorr = ncorr; A.stats = stats;
    A.pcur = (double *)(b + o_p); A.mcur = (double *)(b + o_m); A.prev = (int2 *)(b + o_prev); A.anchor = (double4 *)(b + o_an);
    A.scratch_off = (const int64_t *)(b + o_soff);
    A.k = 1.0 - o.epsilon; A.loss = o.loss; A.loss_k = o.loss_k; A.rel_fitness = o.rel_fitness; A.rel_rmse = o.rel_rmse;
    A.eval_scale = eval_scale; A.eval_out = eval_out;
    A.gang = gang;
    CK(cudaMemsetAsync(b + o_sync, 0, sizeof(unsigned int) * 2 * n_pairs, st));
    A.gpart = (double *)(b + o_gpart);
    A.gsync = (unsigned int *)(b + o_sync);
    A.n_pairs = n_pairs; A.n_ctas = n_ctas;
    A.vmax = gang;
    // adaptive: with at least one pair per block a chunk is ~8192 source points (measured best at 148, 296 and 592 pairs:
    // the coarse scales run as one chunk, the finest as two); with fewer pairs than blocks proportionally more, smaller
    // chunks keep the blocks busy
    const bool many = n_pairs >= resident_t;
    A.v_num = adaptive ? (many ? 1 : resident_t) : 0;
    A.v_den = adaptive ? (many ? 2LL * chunk_points : (long long)chunk_points * n_pairs) : 0;
    A.ps = (PairState *)(b + o_pstate); A.qctl = (unsigned int *)(b + o_qctl); A.queue = (int *)(b + o_queue);
    h->last_prev = nullptr;
    if (eval_scale < 0) {
        h->last_prev = A.prev; h->last_soff = soff;
        h->last_src.assign(pair_src, pair_src + n_pairs); h->last_tgt.assign(pair_tgt, pair_tgt + n_pairs);
    }
    A.scale_t = nullptr;
    if (h->timing && eval_scale < 0) {
        A.scale_t = (unsigned long long *)(b + o_scalet);
        CK(cudaMemsetAsync(A.scale_t, 0, sizeof(unsigned long long) * 2 * (size_t)n_pairs * S, st));
        h->scale_t = A.scale_t; h->scale_t_pairs = n_pairs;
        tmark(h, st, 6);
    }
    if (tasks) {
        // qctl and the queue are adjacent: one memset publishes "no tasks yet"
        CK(cudaMemsetAsync(b + o_qctl, 0, (o_queue - o_qctl) + sizeof(int) * q_slots, st));
        int active = n_pairs;
        if (const char *e = getenv("MGICP_ACTIVE_PAIRS")) active = std::max(1, std::min(n_pairs, atoi(e)));       // tuning experiments
        A.active_pairs = active;
        const unsigned int first_waiting = (unsigned int)active;
        CK(cudaMemcpyAsync(b + o_qctl + 3 * sizeof(unsigned int), &first_waiting, sizeof(unsigned int), cudaMemcpyHostToDevice, st));
        k_icp_task_init<<<(active + 127) / 128, 128, 0, st>>>(A);
        // an ordinary launch: a block that has not started yet holds nothing another block could wait for (tasks are only
        // ever taken by running blocks, and the exit tokens are still there when a late block arrives)
        const int ws_mode = getenv("MGICP_WS") ? atoi(getenv("MGICP_WS")) : 0;                                 // experiment, see k_icp_tasks_ws
        if (ws_mode && ICP_NT == 512) {
            CK(cudaFuncSetAttribute(k_icp_tasks_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_DYN_SMEM));
            k_icp_tasks_ws<<<n_ctas, WS_NT, WS_DYN_SMEM, st>>>(A);
        } else
        k_icp_tasks<<<n_ctas, ICP_NT, ICP_DYN_SMEM, st>>>(A);
        h->launches += 2;
    } else if (gang > 1) {
        void *args[] = {&A};
        CK(cudaLaunchCooperativeKernel((void *)k_icp, dim3(n_pairs * gang), dim3(ICP_NT), args, ICP_DYN_SMEM, st));
        h->launches += 1;
    } else {
        k_icp<<<n_pairs, ICP_NT, ICP_DYN_SMEM, st>>>(A);
        h->launches += 1;
    }
    if (h->timing && eval_scale < 0) tmark(h, st, 7);
    CK(cudaGetLastError());
    return MGICP_OK;
}

extern "C" int mgicp_register_batch(mgicp_handle h, void *stream, int32_t n_pairs, const int32_t *pair_src, const int32_t *pair_tgt,
                                    const double *max_dists, const int32_t *max_iters, const mgicp_opts *opts, const double *T_init,
                                    double *T_out, double *fitness, double *rmse, int32_t *iters, int32_t *ncorr, double *stats) {
    if (!h) return MGICP_E_INVALID;
    if (!T_out || !fitness || !rmse || !max_iters) { h->err = "register: null output"; return MGICP_E_INVALID; }
    return icp_launch(h, (cudaStream_t)stream, n_pairs, pair_src, pair_tgt, max_dists, max_iters, opts, T_init, T_out, fitness, rmse,
                      iters, ncorr, stats, -1, nullptr);
}

extern "C" int mgicp_evaluate_batch(mgicp_handle h, void *stream, int32_t scale, int32_t n_pairs, const int32_t *pair_src,
                                    const int32_t *pair_tgt, const double *max_dists, const mgicp_opts *opts, const double *T,
                                    double *out) {
    if (!h) return MGICP_E_INVALID;
    if (!out || scale < 0 || scale >= h->n_scales) { h->err = "evaluate: bad arguments"; return MGICP_E_INVALID; }
    return icp_launch(h, (cudaStream_t)stream, n_pairs, pair_src, pair_tgt, max_dists, nullptr, opts, T, nullptr, nullptr, nullptr,
                      nullptr, nullptr, nullptr, scale, out);
}

extern "C" int mgicp_evaluate_clouds(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                                     int32_t xyz_dtype, int32_t n_pairs, const int32_t *pair_src, const int32_t *pair_tgt,
                                     const double *max_dists, const double *T, double *out, int32_t *corr) {
    if (!h) return MGICP_E_INVALID;
    h->preprocessed = false;        // the workspace is reused: a previous mgicp_preprocess is gone after this call
    if (n_clouds <= 0 || n_pairs <= 0 || !cloud_off || !pair_src || !pair_tgt || !max_dists || !T || !out ||
        (xyz_dtype != MGICP_F32 && xyz_dtype != MGICP_F64)) { h->err = "mgicp_evaluate_clouds: bad arguments"; return MGICP_E_INVALID; }
    double cell = 0.0;
    for (int i = 0; i < n_pairs; ++i) {
        if (!(max_dists[i] > 0.0)) { h->err = "max_correspondence_distance <= 0"; return MGICP_E_INVALID; }
        if (pair_src[i] < 0 || pair_src[i] >= n_clouds || pair_tgt[i] < 0 || pair_tgt[i] >= n_clouds) {
            h->err = "pair index out of range"; return MGICP_E_INVALID;
        }
        cell = std::max(cell, max_dists[i]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    const size_t esz = xyz_dtype == MGICP_F32 ? 4 : 8;
    // only clouds that serve as a target get a grid
    std::vector<char> is_tgt(n_clouds, 0);
    for (int i = 0; i < n_pairs; ++i) is_tgt[pair_tgt[i]] = 1;
    std::vector<Job> jobs(n_clouds);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o_ = off; off += align_up(bytes); return o_; };
    const size_t o_jobs = take(sizeof(Job) * n_clouds), o_benc = take(sizeof(u64) * BENC_W * n_clouds), o_coff = take(sizeof(int64_t) * (n_clouds + 1));
    int64_t maxn = 0, max_src = 0;
    std::vector<size_t> offs((size_t)n_clouds * 10, 0);
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
        if (!is_tgt[c]) continue;
        maxn = std::max(maxn, n);
        const size_t m = (size_t)std::max<int64_t>(n, 1), ccap = ((size_t)1 << cb) + TAB_PAD;
        size_t *o_ = &offs[(size_t)c * 10];
        o_[0] = take(sizeof(CellSlot) * ccap);   // itab
        o_[1] = take(sizeof(int32_t) * ccap);    // ccursor
        o_[2] = take(sizeof(int32_t) * m);       // pslot
        o_[3] = take(sizeof(int32_t) * m);       // order
        o_[4] = take(sizeof(double4) * m);       // pts
        o_[5] = take(sizeof(double4) * m);       // nrm
        o_[6] = take(sizeof(double4) * m * 2);   // ipts, inrm
        o_[7] = take(sizeof(int32_t) * m * 2);   // a2i, i2a
    }
    std::vector<int64_t> coff(n_pairs + 1, 0);
    for (int i = 0; i < n_pairs; ++i) {
        const int64_t n = cloud_off[pair_src[i] + 1] - cloud_off[pair_src[i]];
        max_src = std::max(max_src, n);
        coff[i + 1] = coff[i] + n;
    }
    const int chunks = chunks_for(max_src, EVAL_NT * 8, 512);     // a function of the input sizes only: results do not depend on the GPU
    const size_t o_ps = take(sizeof(int32_t) * n_pairs), o_pt = take(sizeof(int32_t) * n_pairs), o_md = take(sizeof(double) * n_pairs);
    const size_t o_T = take(sizeof(double) * 16 * n_pairs), o_part = take(sizeof(double) * (size_t)n_pairs * chunks * NEV);
    const size_t o_co = take(sizeof(int64_t) * (n_pairs + 1));
    int rc = grow(h, &h->arena, &h->arena_bytes, off);
    if (rc) return rc;
    char *base = h->arena;
    for (int c = 0; c < n_clouds; ++c) {
        if (!is_tgt[c]) continue;
        Job &j = jobs[c];
        const size_t m = (size_t)std::max<int64_t>(j.n, 1);
        size_t *o_ = &offs[(size_t)c * 10];
        j.itab = (CellSlot *)(base + o_[0]); j.ccursor = (int32_t *)(base + o_[1]); j.pslot = (int32_t *)(base + o_[2]);
        j.order = (int32_t *)(base + o_[3]); j.pts = (double4 *)(base + o_[4]); j.nrm = (double4 *)(base + o_[5]);
        j.ipts = (double4 *)(base + o_[6]); j.inrm = j.ipts + m; j.a2i = (int32_t *)(base + o_[7]); j.i2a = j.a2i + m;
    }
    // jobs of clouds that are only sources keep n but build nothing: give the build kernels an empty job
    std::vector<Job> build = jobs;
    for (int c = 0; c < n_clouds; ++c) if (!is_tgt[c]) build[c].n = 0;
    Job *jobs_dev = (Job *)(base + o_jobs);
    u64 *benc = (u64 *)(base + o_benc);
    int64_t *coff_dev = (int64_t *)(base + o_coff);
    CK(cudaMemcpyAsync(jobs_dev, build.data(), sizeof(Job) * n_clouds, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(coff_dev, cloud_off, sizeof(int64_t) * (n_clouds + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_ps, pair_src, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_pt, pair_tgt, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_md, max_dists, sizeof(double) * n_pairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_T, T, sizeof(double) * 16 * n_pairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(base + o_co, coff.data(), sizeof(int64_t) * (n_pairs + 1), cudaMemcpyHostToDevice, st));
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
    // the evaluation reads every cloud's raw points as a source: restore the true sizes
    k_eval_set_n<<<(n_clouds + 127) / 128, 128, 0, st>>>(jobs_dev, coff_dev, n_clouds);
    EvalArgs E;
    E.jobs = jobs_dev; E.pair_src = (const int32_t *)(base + o_ps); E.pair_tgt = (const int32_t *)(base + o_pt);
    E.max_d = (const double *)(base + o_md); E.T = (const double *)(base + o_T); E.part = (double *)(base + o_part);
    E.out = out; E.corr = corr; E.corr_off = (const int64_t *)(base + o_co);
    k_eval_clouds<<<dim3(chunks, n_pairs), EVAL_NT, 0, st>>>(E);
    k_eval_finish<<<n_pairs, 32, 0, st>>>(E, chunks);
    h->launches += 13;
    CK(cudaGetLastError());
    h->eval_jobs = jobs_dev; h->eval_n = n_clouds;
    return MGICP_OK;
}

// The ICP grid is built by mgicp_preprocess, before the search radii are known.  With a radius of a few voxels (script 2: <= 3)
// a cell of 3 voxels puts every candidate into the 27 cells around the query.  With the ALL_FUNCTIONS schedule (radius = the
// cloud's size, ~110 voxels) the nearest neighbour of a point in a sparse region lies many such cells away and every re-search
// walks thousands of (mostly empty) cells: measured on 64 NCLT pairs, cell factor 3 / 6 / 10 / 16 / 25: ICP 714 / 357 / 153 /
// 56 / 78 ms.  Rule: the largest radius-to-voxel ratio of the batch, clamped to [3.5, 16].
extern "C" double mgicp_auto_icp_cell_factor(int32_t n_scales, const double *voxel_sizes, int32_t n_pairs, const double *max_dists) {
    double ratio = 0.0;
    if (!voxel_sizes || !max_dists) return 3.5;
    for (int p = 0; p < n_pairs; ++p)
        for (int s = 0; s < n_scales; ++s)
            if (voxel_sizes[s] > 0.0 && max_dists[(size_t)p * n_scales + s] > 0.0) ratio = std::max(ratio, max_dists[(size_t)p * n_scales + s] / voxel_sizes[s]);
    return std::min(16.0, std::max(3.5, ratio));
}

extern "C" int mgicp_run_batch(mgicp_handle h, void *stream, int32_t n_clouds, const void *xyz, const int64_t *cloud_off,
                               int32_t xyz_dtype, int32_t n_scales, const double *voxel_sizes, int32_t n_pairs,
                               const int32_t *pair_src, const int32_t *pair_tgt, const double *max_dists, const int32_t *max_iters,
                               const mgicp_opts *opts, const double *T_init, double *T_out, double *fitness, double *rmse,
                               int32_t *iters, int32_t *ncorr, double *stats) {
    // icp_cell_factor left at 0: the ICP grid's cell edge follows the search radius of the schedule, in voxels (see
    // mgicp_auto_icp_cell_factor): 3.5 for the script-2 schedule, 16 for the ALL_FUNCTIONS one (radius ~ the cloud's size)
    mgicp_opts o2;
    if (opts && opts->icp_cell_factor == 0.0 && max_dists && voxel_sizes && n_pairs > 0 && n_scales > 0) {
        o2 = *opts;
        o2.icp_cell_factor = mgicp_auto_icp_cell_factor(n_scales, voxel_sizes, n_pairs, max_dists);
        opts = &o2;
    }
    int rc = mgicp_preprocess(h, stream, n_clouds, xyz, cloud_off, xyz_dtype, n_scales, voxel_sizes, opts);
    if (rc) return rc;
    return mgicp_register_batch(h, stream, n_pairs, pair_src, pair_tgt, max_dists, max_iters, opts, T_init, T_out, fitness, rmse, iters,
                                ncorr, stats);
}

// one block: first non-zero error flag of the jobs -> *out (same codes as mgicp_status)
__global__ void k_job_errors(const Job *jobs, int n_jobs, int32_t *out) {
    __shared__ int s_err;
    if (threadIdx.x == 0) s_err = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < n_jobs; j += blockDim.x)
        if (jobs[j].err) atomicMax(&s_err, jobs[j].err);
    __syncthreads();
    if (threadIdx.x == 0) *out = s_err;
}

extern "C" int mgicp_job_errors(mgicp_handle h, void *stream, int32_t *err_out) {
    if (!h) return MGICP_E_INVALID;
    if (!err_out || (!h->preprocessed && !h->eval_jobs)) { h->err = "mgicp_job_errors: nothing to report on"; return MGICP_E_STATE; }
    CK(cudaSetDevice(h->device));
    const int J = h->preprocessed ? h->n_clouds * h->n_scales : h->eval_n;
    k_job_errors<<<1, 256, 0, (cudaStream_t)stream>>>(h->preprocessed ? h->jobs_dev : h->eval_jobs, J, err_out);
    h->launches += 1;
    CK(cudaGetLastError());
    return MGICP_OK;
}

extern "C" int mgicp_get_correspondences(mgicp_handle h, int32_t pair, int32_t *dst, int64_t cap, int64_t *count) {
    if (!h) return MGICP_E_INVALID;
    if (!h->preprocessed || !h->last_prev) { h->err = "mgicp_get_correspondences: no registration on this handle"; return MGICP_E_STATE; }
    if (pair < 0 || pair >= (int32_t)h->last_src.size() || !dst || !count) { h->err = "mgicp_get_correspondences: bad arguments"; return MGICP_E_INVALID; }
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    const int S = h->n_scales;
    Job js, jt;
    CK(cudaMemcpy(&js, h->jobs_dev + h->last_src[pair] * S + (S - 1), sizeof(Job), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&jt, h->jobs_dev + h->last_tgt[pair] * S + (S - 1), sizeof(Job), cudaMemcpyDeviceToHost));
    *count = 0;
    if (js.Mf <= 0 || jt.Mf <= 0) return MGICP_OK;        // the last scale did not run: empty correspondence set
    std::vector<int2> prev(js.Mf);
    std::vector<int32_t> si(js.Mf), ti(jt.Mf);
    CK(cudaMemcpy(prev.data(), h->last_prev + h->last_soff[pair], sizeof(int2) * js.Mf, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(si.data(), js.i2a, sizeof(int32_t) * js.Mf, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ti.data(), jt.i2a, sizeof(int32_t) * jt.Mf, cudaMemcpyDeviceToHost));
    // rows ordered by source index, like Open3D's correspondence_set (built in a loop over the source points)
    std::vector<std::pair<int32_t, int32_t>> rows;
    for (int i = 0; i < js.Mf; ++i)
        if (prev[i].x >= 0 && prev[i].x < jt.Mf) rows.emplace_back(si[i], ti[prev[i].x]);
    std::sort(rows.begin(), rows.end());
    if ((int64_t)rows.size() > cap) { h->err = "mgicp_get_correspondences: dst too small"; return MGICP_E_INVALID; }
    for (size_t r = 0; r < rows.size(); ++r) { dst[2 * r] = rows[r].first; dst[2 * r + 1] = rows[r].second; }
    *count = (int64_t)rows.size();
    return MGICP_OK;
}

extern "C" int mgicp_set_timing(mgicp_handle h, int32_t on) {
    if (!h) return MGICP_E_INVALID;
    h->timing = on != 0;
    h->tev_n = 0;
    return MGICP_OK;
}

extern "C" int mgicp_get_timing(mgicp_handle h, double *ms_out) {
    if (!h || !ms_out) return MGICP_E_INVALID;
    for (int i = 0; i < 16; ++i) ms_out[i] = 0.0;
    if (!h->timing || h->tev_n < 6) { h->err = "mgicp_get_timing: enable timing, then preprocess (and register) first"; return MGICP_E_STATE; }
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->tev[h->tev_n - 1]));
    for (int i = 0; i < 5; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->tev[i], h->tev[i + 1]));
        ms_out[i] = ms;
    }
    if (h->tev_n == 8) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->tev[6], h->tev[7]));
        ms_out[5] = ms;
        const int S = h->n_scales, P = h->scale_t_pairs;
        std::vector<unsigned long long> t((size_t)2 * P * S);
        CK(cudaMemcpy(t.data(), h->scale_t, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
        for (int s2 = 0; s2 < S && s2 < 8; ++s2) {
            double acc = 0.0; int cnt = 0;
            for (int p2 = 0; p2 < P; ++p2) {
                const unsigned long long a = t[((size_t)p2 * S + s2) * 2], b2 = t[((size_t)p2 * S + s2) * 2 + 1];
                if (a && b2 >= a) { acc += (double)(b2 - a) * 1e-6; ++cnt; }
            }
            ms_out[8 + s2] = cnt ? acc / cnt : 0.0;
        }
    }
    return MGICP_OK;
}

extern "C" int mgicp_check(mgicp_handle h) {
    // synchronous: first device-side error flag of the last preprocess (MGICP_E_RANGE / MGICP_E_OVERFLOW)
    if (!h) return MGICP_E_INVALID;
    if (!h->preprocessed && !h->eval_jobs) return MGICP_OK;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    const int J = h->preprocessed ? h->n_clouds * h->n_scales : h->eval_n;
    std::vector<Job> jobs(J);
    CK(cudaMemcpy(jobs.data(), h->preprocessed ? h->jobs_dev : h->eval_jobs, sizeof(Job) * J, cudaMemcpyDeviceToHost));
    for (int j = 0; j < J; ++j)
        if (jobs[j].err) {
            h->err = jobs[j].err == ERR_RANGE ? "extent / voxel_size exceeds 2^21 cells per axis" : "internal hash table overflow";
            return jobs[j].err == ERR_RANGE ? MGICP_E_RANGE : MGICP_E_OVERFLOW;
        }
    return MGICP_OK;
}

extern "C" int mgicp_get_stage(mgicp_handle h, int32_t cloud, int32_t scale, int32_t what, void *dst, int64_t cap, int64_t *count) {
    if (!h) return MGICP_E_INVALID;
    if (!h->preprocessed) { h->err = "mgicp_get_stage: nothing preprocessed"; return MGICP_E_STATE; }
    if (cloud < 0 || cloud >= h->n_clouds || scale < 0 || scale >= h->n_scales || !dst || !count) {
        h->err = "mgicp_get_stage: bad arguments"; return MGICP_E_INVALID;
    }
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    Job j;
    CK(cudaMemcpy(&j, h->jobs_dev + cloud * h->n_scales + scale, sizeof(Job), cudaMemcpyDeviceToHost));
    if (j.err) { h->err = "job has an error flag"; return j.err == ERR_RANGE ? MGICP_E_RANGE : MGICP_E_OVERFLOW; }
    auto copy_d4 = [&](const double4 *src, int n) -> int {
        if (cap < (int64_t)n * 3) { h->err = "mgicp_get_stage: dst too small"; return MGICP_E_INVALID; }
        std::vector<double4> tmp(std::max(n, 1));
        CK(cudaMemcpy(tmp.data(), src, sizeof(double4) * n, cudaMemcpyDeviceToHost));
        double *d = (double *)dst;
        for (int i = 0; i < n; ++i) { d[3 * i] = tmp[i].x; d[3 * i + 1] = tmp[i].y; d[3 * i + 2] = tmp[i].z; }
        *count = n;
        return MGICP_OK;
    };
    auto copy_raw = [&](const void *src, int rows, size_t row_bytes, int64_t row_elems) -> int {
        if (cap < (int64_t)rows * row_elems) { h->err = "mgicp_get_stage: dst too small"; return MGICP_E_INVALID; }
        CK(cudaMemcpy(dst, src, row_bytes * rows, cudaMemcpyDeviceToHost));
        *count = rows;
        return MGICP_OK;
    };
    switch (what) {
        case MGICP_STAGE_DOWNSAMPLED: return copy_raw(j.ds, j.M, sizeof(double) * 3, 3);
        case MGICP_STAGE_GRID_POINTS: return copy_d4(j.gpts, j.M);
        case MGICP_STAGE_SOR_AVG: return copy_raw(j.avg, j.M, sizeof(double), 1);
        case MGICP_STAGE_SOR_KEEP: return copy_raw(j.keep, j.M, 1, 1);
        case MGICP_STAGE_POINTS: return copy_d4(j.pts, j.Mf);
        case MGICP_STAGE_NORMALS: return copy_d4(j.nrm, j.Mf);
        case MGICP_STAGE_ICP_POINTS: return copy_d4(j.ipts, j.Mf);
        case MGICP_STAGE_ICP_NORMALS: return copy_d4(j.inrm, j.Mf);
        case MGICP_STAGE_KNN_SOR:
            if (!j.knn_sor) { h->err = "neighbour lists need opts.debug != 0"; return MGICP_E_STATE; }
            return copy_raw(j.knn_sor, j.M, sizeof(int32_t) * h->opts.sor_k, h->opts.sor_k);
        case MGICP_STAGE_KNN_NORMAL:
            if (!j.knn_nrm) { h->err = "neighbour lists need opts.debug != 0"; return MGICP_E_STATE; }
            return copy_raw(j.knn_nrm, j.Mf, sizeof(int32_t) * h->opts.normal_k, h->opts.normal_k);
        case MGICP_STAGE_BOUNDS: {
            if (cap < 6) { h->err = "mgicp_get_stage: dst too small"; return MGICP_E_INVALID; }
            std::vector<double> tmp(6);
            double *dev = nullptr;
            CK(cudaMalloc((void **)&dev, sizeof(double) * 6));
            k_bounds_decode<<<1, 32>>>(h->benc + cloud * BENC_W, dev, 6);
            CK(cudaMemcpy(dst, dev, sizeof(double) * 6, cudaMemcpyDeviceToHost));
            CK(cudaFree(dev));
            *count = 1;
            return MGICP_OK;
        }
        default: h->err = "mgicp_get_stage: unknown stage"; return MGICP_E_INVALID;
    }
}

// =============================================================================================
// FGR front end, feature stage (hybrid-radius normals + FPFH on the clouds as given): first CUDA path, see the header
// =============================================================================================
#include "mgicp_fgr.cuh"
