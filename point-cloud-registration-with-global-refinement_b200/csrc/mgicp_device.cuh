// mgicp_device.cuh -- device-side data structures shared by the kernels of mgicp.cu:
// job descriptors, the ordered (insertion-order independent) spatial hash, block scans/reductions,
// the exact grid kNN / radius-bounded NN searches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mgicp_math.cuh"

namespace mg {

typedef unsigned long long u64;

constexpr u64 EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
constexpr int TAB_PAD = 2048;          // slots past the power-of-two part (no wrap-around probing)
constexpr int COORD_BITS = 21;         // per-axis voxel / cell index range
constexpr int COORD_LIMIT = 1 << COORD_BITS;
constexpr double CELL_SLACK = 1e-7;    // relative (to the cell edge) widening of cell boxes in pruning tests
constexpr double RAD_SLACK = 1e-9;     // relative widening of search radii when enumerating cells

enum { ERR_NONE = 0, ERR_RANGE = 4, ERR_OVERFLOW = 5 };

// spatial-hash slot: key then the [start, start+count) range of the cell's points in grid order
struct __align__(16) CellSlot {
    u64 key;
    int32_t start;
    int32_t count;
};

// One (cloud, scale) job.  Static members are written by the host, the rest by kernels.
struct Job {
    // ---- static (host) ----
    const void *xyz;      // raw cloud
    int32_t dtype;        // MGICP_F32 / MGICP_F64
    int32_t cloud;        // cloud index (bounds lookup)
    int64_t n;            // raw points
    double voxel;         // voxel size of this scale
    double cell;          // kNN spatial-hash cell edge (outlier filter + normals)
    double cell_i;        // ICP spatial-hash cell edge (radius-bounded nearest neighbour)
    int32_t vbits;        // voxel table: 1<<vbits slots + TAB_PAD
    int32_t cbits_max;    // allocated cell-table bits
    u64 *vkeys;           // [ (1<<vbits) + TAB_PAD ]
    int32_t *vrank;       // same length: dense voxel id of an occupied slot
    double *vsum;         // [n*3] per-voxel coordinate sums
    int32_t *vcnt;        // [n]
    double *ds;           // [n*3] voxel centroids, canonical order (= slot order of vkeys)
    CellSlot *ctab;       // [ (1<<cbits_max) + TAB_PAD ] cells of the down-sampled cloud
    CellSlot *ftab;       // same shape: cells of the final (outlier-filtered) cloud
    CellSlot *itab;       // same shape: ICP grid (cell edge cell_i) over the final cloud
    double4 *ipts;        // [n] final points in ICP-grid order (w = index into pts)
    double4 *inrm;        // [n] final normals in ICP-grid order
    int32_t *ccursor;     // per-slot scatter cursor
    int32_t *pslot;       // [n] slot of each down-sampled point
    int32_t *order;       // [n] grid order -> canonical id
    double4 *gpts;        // [n] down-sampled points in grid order (w = canonical id)
    double *avg;          // [n] mean kNN distance (SOR)
    uint8_t *keep;        // [n]
    int32_t *newidx;      // [n+1] exclusive prefix of keep
    double4 *pts;         // [n] final points (w = grid-order index before compaction)
    double4 *nrm;         // [n] final normals
    int32_t *fb_list;     // [n] (unused scratch)
    int32_t *nbrA;        // [n*8] the 8 nearest other final points of every final point (indices in pts order, -1 padded)
    int32_t *a2i;         // [n] pts order -> ICP-grid order
    int32_t *i2a;         // [n] ICP-grid order -> pts order
    int32_t *inbr;        // [n*8] nbrA re-indexed to ICP-grid order (row = ICP-grid index)
    int32_t *knn_sor;     // debug: [n*sor_k]
    int32_t *knn_nrm;     // debug: [n*normal_k]
    // ---- dynamic (device) ----
    double org[3];        // voxel-grid origin = min_bound - voxel/2 ; also the cell-grid origin
    int32_t gdim[3];      // number of kNN cells per axis
    int32_t idim[3];      // number of ICP cells per axis
    int32_t cbits;        // kNN cell-table bits in use
    int32_t ibits;        // ICP cell-table bits in use
    int32_t M;            // voxels = down-sampled points
    int32_t Mf;           // points after outlier removal
    int32_t fb_count;     // pending brute-force queries
    int32_t ordered;      // float64 cloud with coordinates beyond float32: voxel sums in input order (k_vox_ord_*)
    int32_t err;
    double sor_thresh;
};

// read-only 32-byte load through the non-coherent path (the pointer comes out of a Job in global memory, so without
// this the compiler has to emit generic loads).  One 256-bit request (sm_100: LDG.E.ENL2.256) instead of two 128-bit
// ones; every double4 array of the workspace starts on a 256-byte boundary.
__device__ __forceinline__ double4 ldg4(const double4 *p) {
    double4 v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
// the same through L2 only (state that other SMs rewrite between passes), and its store
__device__ __forceinline__ double4 ldcg4(const double4 *p) {
    double4 v;
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stcg4(double4 *p, const double4 v) {
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

__device__ __forceinline__ u64 pack_key(int x, int y, int z) {
    return (u64)(uint32_t)x | ((u64)(uint32_t)y << COORD_BITS) | ((u64)(uint32_t)z << (2 * COORD_BITS));
}
// Block-coherent hash: the 4x4x4 block of voxels / cells that contains the key is hashed to a bucket of 64 consecutive
// slots and the key sits at its local (z,y,x) offset inside the bucket.  Neighbouring cells share cache lines, and the slot
// order (which becomes the point order) keeps the points of a block contiguous, so the lanes of a warp work on the
// same neighbourhood.  Collisions between blocks are resolved by the ordered linear probing below.  bits >= 10.
__device__ __forceinline__ uint32_t hash_key(u64 key, int bits) {
    const u64 lowmask = 3ull | (3ull << COORD_BITS) | (3ull << (2 * COORD_BITS));
    const u64 block = key & ~lowmask;
    const uint32_t local = (uint32_t)(key & 3ull) | ((uint32_t)((key >> COORD_BITS) & 3ull) << 2) | ((uint32_t)((key >> (2 * COORD_BITS)) & 3ull) << 4);
    const uint32_t bucket = (uint32_t)((block * 0x9E3779B97F4A7C15ull) >> (64 - (bits - 6)));
    return (bucket << 6) | local;
}

// Ordered linear-probing insertion (Amble & Knuth): a slot always ends up holding the smallest key
// that probes through it, so the final table layout depends only on the SET of keys, not on the
// order in which concurrent threads insert them.  That makes every later "in slot order"
// enumeration (down-sampled point order, cell order) deterministic without sorting.
// `stride` is the slot size in u64 units (1 for bare key tables, 2 for CellSlot tables).
template <int STRIDE>
__device__ __forceinline__ bool ordered_insert(u64 *tab, int bits, u64 key) {
    uint32_t i = hash_key(key, bits);
    const uint32_t cap = (1u << bits) + TAB_PAD;
    while (i < cap) {
        u64 *slot = tab + (size_t)i * STRIDE;
        u64 cur = *((volatile u64 *)slot);
        if (cur == key) return true;
        if (cur < key) { ++i; continue; }           // slot values only ever decrease: it stays < key
        u64 old = atomicMin(slot, key);
        if (old == EMPTY_KEY || old == key) return true;
        if (old > key) key = old;                   // we displaced `old`: carry it on
        ++i;
    }
    return false;
}

// lookup in a finished ordered table; -1 if absent (a larger key on the probe path proves absence)
template <int STRIDE>
__device__ __forceinline__ int ordered_find(const u64 *tab, int bits, u64 key) {
    uint32_t i = hash_key(key, bits);
    const uint32_t cap = (1u << bits) + TAB_PAD;
    while (i < cap) {
        u64 cur = __ldg(tab + (size_t)i * STRIDE);
        if (cur == key) return (int)i;
        if (cur > key) return -1;
        ++i;
    }
    return -1;
}

__device__ __forceinline__ bool cell_find(const CellSlot *tab, int bits, u64 key, int &start, int &count) {
    uint32_t i = hash_key(key, bits);
    const uint32_t cap = (1u << bits) + TAB_PAD;
    while (i < cap) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(tab + i));
        u64 cur = (u64)raw.x | ((u64)raw.y << 32);
        if (cur == key) { start = (int)raw.z; count = (int)raw.w; return true; }
        if (cur > key) return false;
        ++i;
    }
    return false;
}

// ---- block-wide helpers (blockDim.x multiple of 32, <= 1024) ---------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one int per thread; returns exclusive prefix, *total gets the block sum.
// smem: int[33]
__device__ __forceinline__ int block_excl_scan(int v, int *smem, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_incl_scan(v);
    __syncthreads();   // protect smem reuse across calls
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < nw ? smem[lane] : 0;
        int si = warp_incl_scan(s);
        smem[lane] = si - s;
        if (lane == 31) smem[32] = si;
    }
    __syncthreads();
    *total = smem[32];
    return smem[w] + inc - v;
}


// Exclusive scan over n items by one 1024-thread block with no barrier inside the loops: every warp owns a contiguous
// region; pass 1 sums it (coalesced), one block scan of the 32 warp totals, pass 2 re-walks the region with warp scans.
// load(i) -> int value (called twice per item), store(i, exclusive_prefix, value).  Returns the total.
template <class Load, class Store>
__device__ __forceinline__ int block_region_scan(const int n, int *smem /* [33] */, Load load, Store store) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int per = ((n + nw * 32 - 1) / (nw * 32)) * 32;
    const int lo = min(w * per, n), hi = min(lo + per, n);
    int sum = 0;
#pragma unroll 8
    for (int i = lo + lane; i < hi; i += 32) sum += load(i);      // independent loads: unrolled so that 8 are in flight
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) smem[w] = sum;
    __syncthreads();
    if (w == 0) {
        const int v = lane < nw ? smem[lane] : 0;
        const int inc = warp_incl_scan(v);
        smem[lane] = inc - v;
        if (lane == 31) smem[32] = inc;
    }
    __syncthreads();
    int running = smem[w];
    const int total = smem[32];
    for (int i0 = lo; i0 < hi; i0 += 8 * 32) {
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * 32 + lane; v[u] = i < hi ? load(i) : 0; }   // 8 loads in flight
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 32 + lane;
            const int inc = warp_incl_scan(v[u]);
            if (i < hi) store(i, running + inc - v[u], v[u]);
            running += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    return total;
}

// deterministic block sum of one double per thread (fixed shuffle tree, then warp partials in order)
__device__ __forceinline__ double block_sum(double v, double *smem /* [32] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += smem[i];
    return s;
}

// read-only view of one cloud's spatial hash + points
struct GridView {
    const CellSlot *tab;
    const double4 *pts;
    int bits;
    int n;
    double org[3];
    double cell;
    int dim[3];
};

// which: 0 = kNN grid over the down-sampled cloud, 1 = kNN grid over the final cloud, 2 = ICP grid over the final cloud
__device__ __forceinline__ GridView make_view(const Job &j, int which) {
    GridView g;
    g.tab = which == 0 ? j.ctab : (which == 1 ? j.ftab : j.itab);
    g.pts = which == 0 ? j.gpts : (which == 1 ? j.pts : j.ipts);
    g.bits = which == 2 ? j.ibits : j.cbits;
    g.n = which == 0 ? j.M : j.Mf;
    g.org[0] = j.org[0]; g.org[1] = j.org[1]; g.org[2] = j.org[2];
    g.cell = which == 2 ? j.cell_i : j.cell;
    const int32_t *d = which == 2 ? j.idim : j.gdim;
    g.dim[0] = d[0]; g.dim[1] = d[1]; g.dim[2] = d[2];
    return g;
}

__device__ __forceinline__ int cell_coord(double p, double org, double cell) { return (int)floor((p - org) / cell); }

// squared distance from p to the (slightly widened) box of cell k along one axis
__device__ __forceinline__ double axis_gap(double p, int k, double org, double cell, double slack) {
    double lo = org + (double)k * cell - slack;
    double hi = org + (double)(k + 1) * cell + slack;
    double d = fmax(fmax(lo - p, p - hi), 0.0);
    return d;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative exact kNN (k <= 32).  One warp answers one query.  The running top-k list lives in
// registers, one element per lane, sorted ascending by (d2, idx) across lanes; candidates are evaluated 32
// at a time and only those that beat the current k-th element are inserted (ballot + shuffle-up).
// ---------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;
constexpr int KNN_RMAX = 3;            // rings of cells tried before falling back to a scan of the whole cloud

// lexicographic (d2, idx) "a before b"
__device__ __forceinline__ bool before(double ad, int ai, double bd, int bi) { return ad < bd || (ad == bd && ai < bi); }

// compare-exchange step of a bitonic network across lanes: keep the smaller of (own, partner) when keep_min
__device__ __forceinline__ void cmpx(double &d, int &t, const int partner_xor, const bool keep_min) {
    const double od = __shfl_xor_sync(FULL, d, partner_xor);
    const int ot = __shfl_xor_sync(FULL, t, partner_xor);
    const bool other_first = before(od, ot, d, t);
    if (other_first == keep_min) { d = od; t = ot; }
}

// Merge one candidate per lane (cd, ct; invalid lanes carry +inf) into the warp-sorted top-k list.
// Few accepted candidates: serial insertion (ballot + shuffle-up).  Many: bitonic sort of the batch, then a bitonic
// merge with the list that keeps the 32 smallest of the 64.
__device__ __forceinline__ void warp_insert(double &ld2, int &lidx, int &cnt, const int k, double cd, int ct, const bool valid,
                                            const int lane) {
    double wd = __shfl_sync(FULL, ld2, k - 1);
    int wi = __shfl_sync(FULL, lidx, k - 1);
    unsigned acc = __ballot_sync(FULL, valid && before(cd, ct, wd, wi));
    if (acc == 0u) return;
    const int nacc = __popc(acc);
    if (nacc > 6) {
        if (!((acc >> lane) & 1u)) { cd = INFINITY; ct = 0x7fffffff; }
        // bitonic sort of the batch, ascending over lanes
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int j = size >> 1; j > 0; j >>= 1) {
                const bool up = (lane & size) == 0;            // ascending block
                const bool lower = (lane & j) == 0;
                cmpx(cd, ct, j, lower == up);
            }
        }
        // smallest 32 of (list ++ batch): element-wise min of the list with the reversed batch is bitonic
        const double rd = __shfl_sync(FULL, cd, 31 - lane);
        const int rt = __shfl_sync(FULL, ct, 31 - lane);
        if (before(rd, rt, ld2, lidx)) { ld2 = rd; lidx = rt; }
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) cmpx(ld2, lidx, j, (lane & j) == 0);
        cnt = min(k, cnt + nacc);
        return;
    }
    while (acc) {
        const int src = __ffs(acc) - 1;
        acc &= acc - 1;
        const double d = __shfl_sync(FULL, cd, src);
        const int t = __shfl_sync(FULL, ct, src);
        if (!before(d, t, wd, wi)) continue;                // the list tightened since the ballot (uniform branch)
        const int pos = __popc(__ballot_sync(FULL, before(ld2, lidx, d, t)));
        const double ud = __shfl_up_sync(FULL, ld2, 1);
        const int ui = __shfl_up_sync(FULL, lidx, 1);
        if (lane == pos) { ld2 = d; lidx = t; }
        else if (lane > pos) { ld2 = ud; lidx = ui; }
        if (cnt < k) ++cnt;
        wd = __shfl_sync(FULL, ld2, k - 1);
        wi = __shfl_sync(FULL, lidx, k - 1);
    }
}

// All lanes pass the same query.  On return lane t < cnt holds the t-th nearest neighbour (ascending (d2, idx)).
__device__ void knn_warp(const GridView &g, const double px, const double py, const double pz, const int k, double &ld2, int &lidx,
                         int &cnt) {
    const int lane = threadIdx.x & 31;
    const int cx = cell_coord(px, g.org[0], g.cell), cy = cell_coord(py, g.org[1], g.cell), cz = cell_coord(pz, g.org[2], g.cell);
    const double cell = g.cell, slack = CELL_SLACK * g.cell;
    // distances from the query to the lower / upper faces of its own cell: every cell-box and ring-face distance below is
    // one of these plus a whole number of cells (the few ulps this differs from recomputing each face are far inside `slack`)
    const double bx = g.org[0] + (double)cx * cell, by = g.org[1] + (double)cy * cell, bz = g.org[2] + (double)cz * cell;
    const double flx = px - bx - slack, fhx = (bx + cell) - px - slack;
    const double fly = py - by - slack, fhy = (by + cell) - py - slack;
    const double flz = pz - bz - slack, fhz = (bz + cell) - pz - slack;
    ld2 = INFINITY; lidx = 0x7fffffff; cnt = 0;
    bool done = false;
    for (int R = 0; R <= KNN_RMAX && !done; ++R) {
        const int side = 2 * R + 1, vol = side * side * side;
        for (int base = 0; base < vol; base += 32) {
            const double wprune = __shfl_sync(FULL, ld2, k - 1);
            const int e = base + lane;
            int s = 0, c = 0;
            if (e < vol) {
                int dx, dy, dz;             // e = (dz * side + dy) * side + dx, divisions by compile-time constants
                switch (R) {
                    case 0: dx = 0; dy = 0; dz = 0; break;
                    case 1: { dz = e / 9; const int r = e - 9 * dz; dy = r / 3; dx = r - 3 * dy; break; }
                    case 2: { dz = e / 25; const int r = e - 25 * dz; dy = r / 5; dx = r - 5 * dy; break; }
                    default: { dz = e / 49; const int r = e - 49 * dz; dy = r / 7; dx = r - 7 * dy; break; }
                }
                dx -= R; dy -= R; dz -= R;
                const int x = cx + dx, y = cy + dy, z = cz + dz;
                const bool shell = max(max(abs(dx), abs(dy)), abs(dz)) == R;
                if (shell && x >= 0 && x < g.dim[0] && y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
                    // once the list is full, a cell farther than the current k-th neighbour cannot contribute
                    const double gx = dx == 0 ? 0.0 : fmax((dx < 0 ? flx : fhx) + (double)(abs(dx) - 1) * cell, 0.0);
                    const double gy = dy == 0 ? 0.0 : fmax((dy < 0 ? fly : fhy) + (double)(abs(dy) - 1) * cell, 0.0);
                    const double gz = dz == 0 ? 0.0 : fmax((dz < 0 ? flz : fhz) + (double)(abs(dz) - 1) * cell, 0.0);
                    if (!(cnt == k && gx * gx + gy * gy + gz * gz > wprune))
                        if (!cell_find(g.tab, g.bits, pack_key(x, y, z), s, c)) c = 0;
                }
            }
            // pack the candidates of up to 32 cells densely over the lanes: inclusive scan of the counts, then every lane
            // binary-searches the cell its candidate slot falls into
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(FULL, incl, 31);
            for (int r0 = 0; r0 < total; r0 += 32) {
                const int gidx = r0 + lane;
                int lo = 0;                                  // smallest lane b with incl_b > gidx
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int probe = lo + step - 1;
                    const int v = __shfl_sync(FULL, incl, probe & 31);
                    if (probe < 32 && v <= gidx) lo += step;
                }
                const int bs = __shfl_sync(FULL, s, lo & 31), bc = __shfl_sync(FULL, c, lo & 31), bi = __shfl_sync(FULL, incl, lo & 31);
                const bool valid = gidx < total;
                int t = 0;
                double cd = INFINITY;
                if (valid) {
                    t = bs + (gidx - (bi - bc));
                    const double4 q = ldg4(g.pts + t);
                    cd = dist2(px, py, pz, q.x, q.y, q.z);
                }
                warp_insert(ld2, lidx, cnt, k, cd, t, valid, lane);
            }
        }
        // distance to the nearest face beyond which cells are still unexamined
        double gmin = INFINITY;
        const double Rc = (double)R * cell;
        if (cx - R > 0) gmin = fmin(gmin, flx + Rc);
        if (cx + R < g.dim[0] - 1) gmin = fmin(gmin, fhx + Rc);
        if (cy - R > 0) gmin = fmin(gmin, fly + Rc);
        if (cy + R < g.dim[1] - 1) gmin = fmin(gmin, fhy + Rc);
        if (cz - R > 0) gmin = fmin(gmin, flz + Rc);
        if (cz + R < g.dim[2] - 1) gmin = fmin(gmin, fhz + Rc);
        const double wd = __shfl_sync(FULL, ld2, k - 1);
        if (gmin == INFINITY) done = true;                                    // the whole grid has been examined
        else if (cnt == k && gmin > 0.0 && wd < gmin * gmin) done = true;     // the k-th neighbour is closer than anything unexamined
    }
    if (!done) {
        // isolated query: scan the whole cloud (coalesced, 32 candidates per step)
        ld2 = INFINITY; lidx = 0x7fffffff; cnt = 0;
        for (int o = 0; o < g.n; o += 32) {
            const int t = o + lane;
            const bool valid = t < g.n;
            double cd = INFINITY;
            if (valid) {
                const double4 q = ldg4(g.pts + t);
                cd = dist2(px, py, pz, q.x, q.y, q.z);
            }
            warp_insert(ld2, lidx, cnt, k, cd, t, valid, lane);
        }
    }
}

// Radius-bounded nearest neighbour (Open3D SearchHybrid(p, r, 1)): the nearest point, accepted iff
// d2 < r2 (strict).  `seed` is an optional candidate index (last iteration's correspondence) that
// only tightens the initial bound; the result is the exact nearest neighbour either way.
__device__ void nn_search(const GridView &g, double px, double py, double pz, double r2, int seed, int &best_j, double &best_d2) {
    double bd2 = r2;
    int bj = -1;
    if (seed >= 0) {
        const double4 q = ldg4(g.pts + seed);
        double d = dist2(px, py, pz, q.x, q.y, q.z);
        if (d < bd2) { bd2 = d; bj = seed; }
    }
    const double slack = CELL_SLACK * g.cell;
    const int cx = cell_coord(px, g.org[0], g.cell), cy = cell_coord(py, g.org[1], g.cell), cz = cell_coord(pz, g.org[2], g.cell);
    const bool inside = cx >= 0 && cx < g.dim[0] && cy >= 0 && cy < g.dim[1] && cz >= 0 && cz < g.dim[2];
    if (inside) {
        int s, c;
        if (cell_find(g.tab, g.bits, pack_key(cx, cy, cz), s, c)) {
            for (int t = s; t < s + c; ++t) {
                const double4 q = ldg4(g.pts + t);
                double d = dist2(px, py, pz, q.x, q.y, q.z);
                if (d < bd2 || (d == bd2 && bj >= 0 && t < bj)) { bd2 = d; bj = t; }
            }
        }
    }
    const double rad = sqrt(bd2) * (1.0 + RAD_SLACK) + 1e-300;
    const int x0 = max(cell_coord(px - rad, g.org[0], g.cell), 0), x1 = min(cell_coord(px + rad, g.org[0], g.cell), g.dim[0] - 1);
    const int y0 = max(cell_coord(py - rad, g.org[1], g.cell), 0), y1 = min(cell_coord(py + rad, g.org[1], g.cell), g.dim[1] - 1);
    const int z0 = max(cell_coord(pz - rad, g.org[2], g.cell), 0), z1 = min(cell_coord(pz + rad, g.org[2], g.cell), g.dim[2] - 1);
    if (x0 <= x1 && y0 <= y1 && z0 <= z1) {
        const long long ncell = (long long)(x1 - x0 + 1) * (y1 - y0 + 1) * (z1 - z0 + 1);
        if (ncell > 4096 && ncell > (long long)g.n) {
            // the box holds more cells than the cloud has points: scanning the points is cheaper
            for (int t = 0; t < g.n; ++t) {
                const double4 q = ldg4(g.pts + t);
                double d = dist2(px, py, pz, q.x, q.y, q.z);
                if (d < bd2 || (d == bd2 && bj >= 0 && t < bj)) { bd2 = d; bj = t; }
            }
        } else {
            for (int z = z0; z <= z1; ++z) {
                const double gz = axis_gap(pz, z, g.org[2], g.cell, slack);
                for (int y = y0; y <= y1; ++y) {
                    const double gy = axis_gap(py, y, g.org[1], g.cell, slack);
                    for (int x = x0; x <= x1; ++x) {
                        if (inside && x == cx && y == cy && z == cz) continue;
                        const double gx = axis_gap(px, x, g.org[0], g.cell, slack);
                        if (gx * gx + gy * gy + gz * gz > bd2) continue;
                        int s, c;
                        if (!cell_find(g.tab, g.bits, pack_key(x, y, z), s, c)) continue;
                        for (int t = s; t < s + c; ++t) {
                            const double4 q = ldg4(g.pts + t);
                            double d = dist2(px, py, pz, q.x, q.y, q.z);
                            if (d < bd2 || (d == bd2 && bj >= 0 && t < bj)) { bd2 = d; bj = t; }
                        }
                    }
                }
            }
        }
    }
    best_j = bj;
    best_d2 = bd2;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative radius-bounded nearest neighbour for 32 queries at once (one query per lane).
// The (query, cell) probes of all lanes are flattened over the warp, then the candidate points of the probed cells
// are flattened again, so every lane does useful work regardless of how uneven the per-query boxes are.  The running
// best of each query lives in shared memory (atomicMin on the bit pattern of the non-negative fp64 distance, then the
// smallest index among the candidates that reach it), which also lets later probes prune against the freshest bound.
// The result equals nn_search's: the exact nearest neighbour, accepted iff d2 < r2, ties to the smaller index.
// ---------------------------------------------------------------------------------------------
struct WarpSearch {
    unsigned long long d2bits[32];
    double px[32], py[32], pz[32];
    int idx[32];
};
constexpr int COOP_MAX_CELLS = 125;     // per query; larger boxes take the per-thread path

__device__ __forceinline__ int lane_of_slot(const int incl, const int gidx) {   // smallest lane whose inclusive prefix exceeds gidx
    int lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int probe = lo + step - 1;
        const int v = __shfl_sync(FULL, incl, probe & 31);
        if (probe < 32 && v <= gidx) lo += step;
    }
    return lo & 31;
}

__device__ void nn_search_coop(const GridView &g, WarpSearch &W, const bool need, const double px, const double py, const double pz,
                               const double r2, double &bd2, int &bj) {
    if (!__any_sync(FULL, need)) return;        // the common case late in a scale: every lane was certified without a search
    const int lane = threadIdx.x & 31;
    W.px[lane] = px; W.py[lane] = py; W.pz[lane] = pz;
    W.d2bits[lane] = (unsigned long long)__double_as_longlong(bd2);
    W.idx[lane] = bj;
    unsigned long long mybits = W.d2bits[lane];
    int x0 = 0, y0 = 0, z0 = 0, nx = 0, ny = 0, ncell = 0;
    bool big = false;
    if (need) {
        // the box only has to COVER the ball: the radius is widened by 1e-9 relative plus 1e-9 of a cell, which dwarfs the
        // rounding difference between (x - org) * (1/cell) used here and (x - org) / cell used when the points were binned
        const double rad = sqrt(bd2) * (1.0 + RAD_SLACK) + 1e-9 * g.cell;
        const double inv = 1.0 / g.cell;
        x0 = max((int)floor((px - rad - g.org[0]) * inv), 0);
        y0 = max((int)floor((py - rad - g.org[1]) * inv), 0);
        z0 = max((int)floor((pz - rad - g.org[2]) * inv), 0);
        const int x1 = min((int)floor((px + rad - g.org[0]) * inv), g.dim[0] - 1);
        const int y1 = min((int)floor((py + rad - g.org[1]) * inv), g.dim[1] - 1);
        const int z1 = min((int)floor((pz + rad - g.org[2]) * inv), g.dim[2] - 1);
        if (x0 <= x1 && y0 <= y1 && z0 <= z1) {
            nx = x1 - x0 + 1; ny = y1 - y0 + 1;
            const long long nc = (long long)nx * ny * (z1 - z0 + 1);
            if (nc > COOP_MAX_CELLS) big = true; else ncell = (int)nc;
        }
    }
    __syncwarp();
    int incl = ncell;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    for (int r0 = 0; r0 < total; r0 += 32) {
        const int gi = r0 + lane;
        const int owner = lane_of_slot(incl, gi);
        const int ox0 = __shfl_sync(FULL, x0, owner), oy0 = __shfl_sync(FULL, y0, owner), oz0 = __shfl_sync(FULL, z0, owner);
        const int onx = __shfl_sync(FULL, nx, owner), ony = __shfl_sync(FULL, ny, owner);
        const int oexcl = __shfl_sync(FULL, incl - ncell, owner);
        int s = 0, c = 0;
        if (gi < total) {
            const int e = gi - oexcl;                        // e < 125: float division is exact enough for floor
            const int ez = (int)(((float)e + 0.5f) / (float)(onx * ony));
            const int er = e - ez * onx * ony;
            const int ey = (int)(((float)er + 0.5f) / (float)onx);
            const int cx = ox0 + (er - ey * onx), cy = oy0 + ey, cz = oz0 + ez;
            // conservative fp32 prune: skip the probe only if the cell is clearly farther than the query's current best
            const float cur = (float)__longlong_as_double((long long)W.d2bits[owner]);
            const float cellf = (float)g.cell;
            const float fx = (float)(W.px[owner] - g.org[0]), fy = (float)(W.py[owner] - g.org[1]), fz = (float)(W.pz[owner] - g.org[2]);
            const float gx = fmaxf(fmaxf((float)cx * cellf - fx, fx - (float)(cx + 1) * cellf), 0.0f);
            const float gy = fmaxf(fmaxf((float)cy * cellf - fy, fy - (float)(cy + 1) * cellf), 0.0f);
            const float gz = fmaxf(fmaxf((float)cz * cellf - fz, fz - (float)(cz + 1) * cellf), 0.0f);
            const float gm = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz) - 1e-3f * cellf - 1e-5f * (fabsf(fx) + fabsf(fy) + fabsf(fz)), 0.0f);
            if (!(gm * gm > cur * 1.0001f))
                if (!cell_find(g.tab, g.bits, pack_key(cx, cy, cz), s, c)) c = 0;
        }
        int incl2 = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl2, o);
            if (lane >= o) incl2 += t;
        }
        const int total2 = __shfl_sync(FULL, incl2, 31);
        for (int r1 = 0; r1 < total2; r1 += 32) {
            const int hi = r1 + lane;
            const int b = lane_of_slot(incl2, hi);
            const int bs = __shfl_sync(FULL, s, b), bexcl = __shfl_sync(FULL, incl2 - c, b), bo = __shfl_sync(FULL, owner, b);
            bool cand = false;
            double d = 0.0;
            int t = 0;
            if (hi < total2) {
                t = bs + (hi - bexcl);
                const double4 q = ldg4(g.pts + t);
                d = dist2(W.px[bo], W.py[bo], W.pz[bo], q.x, q.y, q.z);
                cand = d < r2;
                if (cand) atomicMin(&W.d2bits[bo], (unsigned long long)__double_as_longlong(d));
            }
            __syncwarp();
            {   // a query whose best distance just improved forgets the index that belonged to the old distance
                const unsigned long long now = W.d2bits[lane];
                if (now != mybits) { mybits = now; W.idx[lane] = 0x7fffffff; }
            }
            __syncwarp();
            if (cand && (unsigned long long)__double_as_longlong(d) == W.d2bits[bo]) atomicMin(&W.idx[bo], t);
            __syncwarp();
        }
    }
    bd2 = __longlong_as_double((long long)W.d2bits[lane]);
    bj = W.idx[lane];
    __syncwarp();
    if (big) {   // very large search box (e.g. the radius-based schedule's 40 m): exact per-thread search
        int j2; double d2;
        nn_search(g, px, py, pz, r2, bj, j2, d2);
        bj = j2; bd2 = d2;
    }
}

}  // namespace mg
