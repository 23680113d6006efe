// mgicp_fgr_tc.cuh -- FGR feature matching on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// registration_fgr_based_on_feature_matching (ALL_FUNCTIONS.py:189-202) starts with the nearest neighbour of every 33-bin
// FPFH descriptor of one cloud among the descriptors of the other, both ways: 18k x 18k x 33 multiply-adds per direction and
// pair, the one dense contraction of the whole scope (85 % of the CPU oracle's FGR time).  k_fgr_nn evaluates it in fp64 on
// the CUDA cores; this file does the bulk of it on the tensor cores and keeps the result EXACT:
//
//   |a - b|^2 = |a|^2 + |b|^2 - 2 a.b.   a.b is computed as a_h.b_h + a_h.b_l + a_l.b_h with every descriptor split into two
//   fp16 numbers (a = a_h + a_l to 2^-22 relative), i.e. ONE fp16 GEMM with the three products concatenated along K
//   (3 x 33 = 99, padded to 112 = 7 x UMMA_K), fp32 accumulation in TMEM.  Worst-case error of the distance so obtained:
//   eps = 2^-16 (|a|^2 + max|b|^2)  (dropped a_l.b_l and split residuals 3 x 2^-22, fp32 accumulation of 112 terms 112 x 2^-23).
//   The epilogue (one thread per query row, reading its 256 accumulator columns with tcgen05.ld) keeps the four smallest
//   approximate distances per row.  If the fourth is more than 2 eps above the first, the exact nearest neighbour is one of
//   the first three..four: they are re-evaluated in fp64 with k_fgr_nn's own arithmetic (bin order, ties to the lower index).
//   Otherwise (1-2 % of the rows: near-duplicate descriptors) the row is queued for the brute-force fp64 search.
//   Result: bit-identical to k_fgr_nn.
//
// Kernel anatomy (one CTA per 128 query rows, 320 threads, one CTA per SM):
//   warp 0      producer: cp.async.bulk of the query tile (once) and of the database tiles (256 rows x 112 halves = 56 KB per
//               stage, 3 stages) into shared memory, completion on mbarriers.  The operands are pre-packed in global memory
//               in the canonical K-major / no-swizzle UMMA layout (k_fgr_pack), so a tile is one contiguous bulk copy;
//   warp 1      allocates TMEM (512 columns = two 128 x 256 fp32 accumulators) and issues tcgen05.mma (M128 N256 K16, 7 per
//               tile) from one elected lane; tcgen05.commit releases the smem stage and publishes the accumulator;
//   warps 2-9   epilogue: warp w owns TMEM lanes 32 (w mod 4) ..+31 = query rows and one half of a tile's columns;
//               tcgen05.ld 32x32b.x32 (the next chunk requested before the current one is examined), v = |b|^2 - 2 acc (|b|^2 by
//               ld.shared.v4), chunk minimum through a min tree, running top-4 (only groups of 8 values whose minimum beats the
//               row's bound are scanned; the two halves of a row share their bounds through shared memory); frees the
//               accumulator through an mbarrier so the MMA of tile t+2 can start.  (MGICP_FGR_EPI_WARPS=16: four warps per
//               quadrant, a quarter of the columns each, no prefetch -- measured slower.)
#pragma once
#include <cuda_fp16.h>

namespace fgrtc {

// The database rows are packed in a scrambled order (position p holds row (p * PERM_A) mod n_padded, PERM_A prime > any n):
// descriptors of neighbouring points are similar, so in file order the distances to a query fall in long smooth runs and every
// step of a run displaces the running top-4 (measured: a third of the kernel's stall samples sat in the insertion).  Scrambled,
// a row sees its candidates in an order that is as good as random: ~4 ln(n) insertions.
constexpr unsigned long long PERM_A = 2147483647ull;
constexpr int KP = 112;                 // packed K: [a_h | a_h | a_l] . [b_h | b_l | b_h], 3 x 33 = 99 padded to 7 x 16
constexpr int TM = 128;                 // query rows per CTA (UMMA M)
constexpr int TN = 256;                 // database rows per tile (UMMA N)
constexpr int STAGES = 3;
constexpr int A_BYTES = TM * KP * 2;    // 28672
constexpr int B_BYTES = TN * KP * 2;    // 57344
constexpr int NB_BYTES = TN * 4;        // |b|^2 of the tile's rows, fp32
#ifndef MGICP_FGR_SPIN_NS
#define MGICP_FGR_SPIN_NS 64
#endif
#ifndef MGICP_FGR_EPI_WARPS
#define MGICP_FGR_EPI_WARPS 8      // measured (64 NCLT pairs): 8 warps 13.0 ms, 16 warps 14.9 ms -- the epilogue is bound by issued instructions
#endif
constexpr int EPI_WARPS = MGICP_FGR_EPI_WARPS;   // PARTS per TMEM lane quadrant: each takes 1 / PARTS of a tile's columns
constexpr int PARTS = EPI_WARPS / 4;
constexpr int PCOLS = TN / PARTS;
constexpr bool EPI_PREFETCH = EPI_WARPS <= 8;    // 8 warps: the next chunk's tcgen05.ld is in flight while this one is examined (64
                                                 // registers of buffers); 16 warps: 112 registers per thread, the other warps hide it
constexpr int NT = 64 + 32 * EPI_WARPS;
constexpr int MERGE_BYTES = TM * 4 * 8 * (PARTS - 1);  // the warps of parts 1.. hand their four candidates per row to part 0's
constexpr int THR_BYTES = TM * PARTS * 4;               // the parts of a row tell each other their fourth-best value
constexpr size_t SMEM = 1024 + A_BYTES + STAGES * (B_BYTES + NB_BYTES) + MERGE_BYTES + THR_BYTES + 256;
static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "two or four epilogue warps per TMEM lane quadrant");
static_assert(SMEM <= 232448, "shared memory of k_fgr_match_tc");

struct CloudPack {                      // per cloud, device pointers
    const __half *PA;                   // [ceil(n / TM)][KP / 8][TM / 8][8][8]  query form  (a_h | a_h | a_l)
    const __half *PB;                   // [ceil(n / TN)][KP / 8][TN / 8][8][8]  database form (b_h | b_l | b_h), rows scrambled (PERM_A)
    const float *nbf;                   // [ceil(n / TN) * TN] |b|^2 in fp32 of the row at each (scrambled) position, +inf for padding
    const float *nmax;                  // [1] max |b|^2 (as fp32, rounded up)
    int32_t n;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same for the two service lanes (producer, MMA issuer): they share their schedulers with epilogue warps, and a bare
// try_wait loop took 30 % of the kernel's issued instructions; a short sleep between polls gives the slots back
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITR_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONER_%=;\n\t"
        "nanosleep.u32 %2;\n\t"
        "bra WAITR_%=;\n\t"
        "DONER_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)MGICP_FGR_SPIN_NS) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: core matrices of 8 rows x 16 bytes; SBO = distance between 8-row groups, LBO = distance between the two
// 16-byte K chunks of one UMMA_K = 16 step (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
           (1ull << 46);
}
// kind::f16, A = B = fp16 (format 0), D = fp32 (c_format 1), both K-major, M = 128, N = 256 (cute::UMMA::InstrDescriptor)
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
        "%28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
        "%28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- operand packing: fp64 descriptors -> split fp16, canonical UMMA tiles ------------------------------------------------
// grid (chunks, clouds); one thread per (row, packed-K chunk of 8)
struct PackArgs {
    const double *feat;                 // [points][33]
    const int64_t *cloud_off;           // device
    __half *const *PA;                  // per cloud
    __half *const *PB;
    float *const *nbf;
    float *const *nmax;
};
__global__ void __launch_bounds__(256) k_fgr_pack(PackArgs P) {
    const int c = blockIdx.y;
    const int64_t off = P.cloud_off[c];
    const int n = (int)(P.cloud_off[c + 1] - off);
    const int npa = (n + TM - 1) / TM * TM, npb = (n + TN - 1) / TN * TN, nrow = max(npa, npb);
    const double *F = P.feat + 33 * off;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nrow * (KP / 8); e += gridDim.x * blockDim.x) {
        const int pos = e / (KP / 8), kc = e % (KP / 8);
        // query form: position = row; database form: position pos holds row (pos * PERM_A) mod npb
        const int rowb = pos < npb ? (int)(((unsigned long long)pos * PERM_A) % (unsigned long long)npb) : n;
        __align__(16) __half a8[8];
        __align__(16) __half b8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = kc * 8 + u, seg = k / 33, bin = k - 33 * seg;
            __half ah = __float2half(0.f), al = ah, bh = ah, bl = ah;
            if (seg < 3) {
                if (pos < n) {
                    const double x = F[33 * (size_t)pos + bin];
                    ah = __double2half(x);
                    al = __double2half(x - (double)__half2float(ah));
                }
                if (rowb < n) {
                    const double x = F[33 * (size_t)rowb + bin];
                    bh = __double2half(x);
                    bl = __double2half(x - (double)__half2float(bh));
                }
            }
            a8[u] = seg == 2 ? al : ah;               // a_h | a_h | a_l
            b8[u] = seg == 1 ? bl : bh;               // b_h | b_l | b_h
        }
        if (pos < npa) {
            __half *dst = P.PA[c] + (size_t)(pos / TM) * (TM * KP) + (size_t)kc * (TM * 8) + (size_t)((pos % TM) / 8) * 64 + (pos % 8) * 8;
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(a8);
        }
        if (pos < npb) {
            __half *dst = P.PB[c] + (size_t)(pos / TN) * (TN * KP) + (size_t)kc * (TN * 8) + (size_t)((pos % TN) / 8) * 64 + (pos % 8) * 8;
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(b8);
            if (kc == 0) {
                float nf = INFINITY;
                if (rowb < n) {
                    double nb = 0.0;
                    for (int k = 0; k < 33; ++k) { const double x = F[33 * (size_t)rowb + k]; nb += x * x; }
                    nf = (float)nb;
                    atomicMax(reinterpret_cast<int *>(P.nmax[c]), __float_as_int(__double2float_ru(nb)));      // nb >= 0: int order = float order
                }
                P.nbf[c][pos] = nf;
            }
        }
    }
}

struct FbPart { double d; int32_t j; int32_t pad; };
constexpr int FB_SLICES = 128, FB_NT = 128, FB_TILE = 32, FB_CAP = 1024;

// ---- the match -------------------------------------------------------------------------------------------------------------
struct MatchArgs {
    const CloudPack *packs;             // per cloud
    const FgrPair *pairs;
    const double *feat;                 // fp64 descriptors (exact re-check)
    const int64_t *cloud_off;
    int32_t *fb_list;                   // [2 * pairs][max rows] rows that need the brute-force search
    int32_t *fb_count;                  // [2 * pairs]
    int64_t fb_stride;
    struct FbPart *fb_seed;             // [2 * pairs][FB_CAP] best exact (distance, row) among a queued row's tensor-core candidates
};

// grid (query tiles, 2 * pairs): direction 0 fills j2i (queries = descriptors of cloud j, searched among cloud i's), 1 fills i2j
__global__ void __launch_bounds__(NT, 1) k_fgr_match_tc(MatchArgs A) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const FgrPair &pr = A.pairs[blockIdx.y >> 1];
    const int dir = blockIdx.y & 1;
    const int cq = dir == 0 ? pr.fj : pr.fi, ct = dir == 0 ? pr.fi : pr.fj;
    const int cloud_q = cq == 0 ? pr.src : pr.tgt, cloud_t = ct == 0 ? pr.src : pr.tgt;
    const CloudPack Q = A.packs[cloud_q], T = A.packs[cloud_t];
    const int nq = Q.n, nt = T.n;
    if ((int)blockIdx.x * TM >= nq) return;
    int32_t *out = dir == 0 ? pr.j2i : pr.i2j;
    const int n_tiles = (nt + TN - 1) / TN;

    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = base;
    unsigned char *sB = base + A_BYTES;
    float *sNB = reinterpret_cast<float *>(sB + STAGES * B_BYTES);
    float2 *sMerge = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(sNB) + STAGES * NB_BYTES);      // [TM][4] (value, index bits)
    float *sThr = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sMerge) + MERGE_BYTES);             // [TM][PARTS]
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(sThr) + THR_BYTES);
    uint64_t *full = bars, *empty = bars + STAGES, *a_full = bars + 2 * STAGES, *acc_full = bars + 2 * STAGES + 1, *acc_empty = bars + 2 * STAGES + 3;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1 + EPI_WARPS); }     // the MMAs' commit + the epilogue warps (they read the stage's |b|^2)
        mbar_init(a_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TM * PARTS; i += NT) sThr[i] = INFINITY;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            mbar_expect_tx(a_full, A_BYTES);
            bulk_g2s(sA, Q.PA + (size_t)blockIdx.x * (TM * KP), A_BYTES, a_full);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % STAGES;
                if (t >= STAGES) mbar_wait_relaxed(&empty[s], ((t / STAGES) - 1) & 1);
                mbar_expect_tx(&full[s], B_BYTES + NB_BYTES);
                bulk_g2s(sB + (size_t)s * B_BYTES, T.PB + (size_t)t * (TN * KP), B_BYTES, &full[s]);
                bulk_g2s(sNB + s * TN, T.nbf + (size_t)t * TN, NB_BYTES, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            mbar_wait(a_full, 0);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % STAGES, acc = t & 1;
                if (t >= 2) mbar_wait_relaxed(&acc_empty[acc], ((t >> 1) - 1) & 1);  // the epilogue has drained this accumulator
                mbar_wait_relaxed(&full[s], (t / STAGES) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + (size_t)s * B_BYTES);
#ifndef FGR_TC_NO_MMA
#pragma unroll
                for (int k = 0; k < KP / 16; ++k) {
                    const uint64_t ad = umma_desc(a0 + k * 2 * (TM * 16), TM * 16, 128);
                    const uint64_t bd = umma_desc(b0 + k * 2 * (TN * 16), TN * 16, 128);
                    umma_f16(tmem + acc * TN, ad, bd, k > 0 ? 1u : 0u);
                }
#endif
                umma_commit(&empty[s]);            // the stage can be refilled once these MMAs have read it
                umma_commit(&acc_full[acc]);       // ... and the accumulator is complete
            }
        }
    } else {
        // ===== epilogue: thread = query row, PARTS warps per row quadrant (each takes 1 / PARTS of the tile's columns) =====
        const int q4 = warp & 3;                   // TMEM lane quadrant this warp may read
        const int part = (warp - 2) >> 2;          // columns [part * PCOLS, +PCOLS) of every tile
        const int row = blockIdx.x * TM + q4 * 32 + lane;
        float t0 = INFINITY, t1 = INFINITY, t2 = INFINITY, t3 = INFINITY;
        int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
        auto insert = [&](const float x, const int j) {
            if (x < t2) {
                t3 = t2; i3 = i2;
                if (x < t1) {
                    t2 = t1; i2 = i1;
                    if (x < t0) { t1 = t0; i1 = i0; t0 = x; i0 = j; } else { t1 = x; i1 = j; }
                } else { t2 = x; i2 = j; }
            } else { t3 = x; i3 = j; }
        };
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        volatile float *thr = sThr + (q4 * 32 + lane) * PARTS;
        // tf: what a value has to beat to matter.  The fourth-best of ANY part of the row bounds the row's final fourth-best from
        // above, so the smallest of them filters for all parts (read once per tile; a stale value is still a valid bound).  A value
        // of the row's final top four is below every such bound at every time: it enters its part's list and stays there.
        float tf = INFINITY;
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t % STAGES, acc = t & 1;
            mbar_wait(&acc_full[acc], (t >> 1) & 1);
            // the stage's |b|^2 came in with the bulk copy: acquire ITS barrier too (completed long ago: the MMAs waited for it),
            // rather than relying on the chain copy -> MMA thread -> tcgen05.commit -> this thread
            mbar_wait(&full[s], (t / STAGES) & 1);
            tc_fence_after();
            const uint32_t nb_s = smem_u32(sNB + s * TN + part * PCOLS);        // stays valid until this warp arrives on empty[s] below
            const uint32_t col0 = (uint32_t)(acc * TN + part * PCOLS);
#ifdef FGR_TC_NO_EPI
            if (t >= 0) { tc_fence_before(); __syncwarp(); if (lane == 0) { mbar_arrive(&acc_empty[acc]); mbar_arrive(&empty[s]); } continue; }
#endif
#pragma unroll
            for (int p = 0; p < PARTS; ++p) tf = fminf(tf, thr[p]);
            uint32_t ra[32], rb[EPI_PREFETCH ? 32 : 1];
            tmem_ld32_issue(tmem + lane_base + col0, ra);
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < PCOLS; cc += 32) {
                // 8 warps: the next chunk's accumulators are requested before this chunk is looked at
                uint32_t (&cur)[32] = (EPI_PREFETCH && ((cc >> 5) & 1)) ? reinterpret_cast<uint32_t (&)[32]>(rb) : ra;
                if (EPI_PREFETCH && cc + 32 < PCOLS) tmem_ld32_issue(tmem + lane_base + col0 + cc + 32, ((cc >> 5) & 1) ? ra : reinterpret_cast<uint32_t (&)[32]>(rb));
                // branch-free first: the 32 values and their minimum (independent FMAs, a min tree).  Only a chunk that holds
                // something below the bound goes through the insertion; comparing and branching per element made the epilogue
                // 10x slower than the MMAs feeding it
                float v[32];
#pragma unroll
                for (int u = 0; u < 32; u += 4) {
                    float n0, n1, n2, n3;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(n0), "=f"(n1), "=f"(n2), "=f"(n3) : "r"(nb_s + 4u * (uint32_t)(cc + u)));
                    v[u] = fmaf(-2.0f, __uint_as_float(cur[u]), n0);         v[u + 1] = fmaf(-2.0f, __uint_as_float(cur[u + 1]), n1);
                    v[u + 2] = fmaf(-2.0f, __uint_as_float(cur[u + 2]), n2); v[u + 3] = fmaf(-2.0f, __uint_as_float(cur[u + 3]), n3);
                }
                if (!EPI_PREFETCH && cc + 32 < PCOLS) tmem_ld32_issue(tmem + lane_base + col0 + cc + 32, ra);      // `ra` is consumed
                float m16[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) m16[u] = fminf(v[u], v[u + 16]);
#pragma unroll
                for (int u = 0; u < 8; ++u) m16[u] = fminf(m16[u], m16[u + 8]);
#pragma unroll
                for (int u = 0; u < 4; ++u) m16[u] = fminf(m16[u], m16[u + 4]);
                const float mn = fminf(fminf(m16[0], m16[1]), fminf(m16[2], m16[3]));
                if (mn < tf) {
                    // Not rare per WARP (one of 32 lanes qualifies in most chunks), so it is kept short: the min tree's last level
                    // holds the minima of the four groups u = g (mod 4); only a group whose minimum qualifies is scanned (8 values
                    // parked in local memory, a bit mask, one insertion site per group -- with the insertion inlined 32 times the
                    // epilogue's code did not fit the instruction cache).  The order in which equal approximate values arrive
                    // differs from column order; the exact re-check and the certificate below do not depend on it.
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (m16[g] < tf) {
                            float vl[8];
                            unsigned mask = 0u;
#pragma unroll
                            for (int e = 0; e < 8; ++e) { vl[e] = v[g + 4 * e]; mask |= (v[g + 4 * e] < tf ? 1u : 0u) << e; }
                            while (mask) {
                                const int e = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const float x = vl[e];
                                if (x < tf) { insert(x, t * TN + part * PCOLS + cc + g + 4 * e); tf = fminf(tf, t3); }
                            }
                        }
                    }
                }
                if (cc + 32 < PCOLS) tmem_ld_wait();
            }
            thr[part] = t3;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&acc_empty[acc]); mbar_arrive(&empty[s]); }      // one arrival per warp (128 serialised arrivals cost more than the tile)
        }
        // ---- the parts of a row meet: the threads of parts 1.. hand their four candidates over ----
        if (part > 0) {
            float2 *m = sMerge + ((part - 1) * TM + q4 * 32 + lane) * 4;
            m[0] = make_float2(t0, __int_as_float(i0)); m[1] = make_float2(t1, __int_as_float(i1));
            m[2] = make_float2(t2, __int_as_float(i2)); m[3] = make_float2(t3, __int_as_float(i3));
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * EPI_WARPS) : "memory");
        if (part > 0) goto done;
        for (int pp = 0; pp < PARTS - 1; ++pp) {
            const float2 *m = sMerge + (pp * TM + q4 * 32 + lane) * 4;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float x = m[u].x;
                const int j = __float_as_int(m[u].y);
                // equal approximate values: both stay candidates as long as there is room (the exact re-check decides)
                if (x < t3) insert(x, j);
            }
        }
        // ---- exact re-check of the candidates (k_fgr_nn's arithmetic), or the brute-force queue ----
        if (row < nq) {
            const double *Fq = A.feat + 33 * A.cloud_off[cloud_q], *Ft = A.feat + 33 * A.cloud_off[cloud_t];
            double qf[33];
            double na = 0.0;
#pragma unroll
            for (int k = 0; k < 33; ++k) { qf[k] = Fq[33 * (size_t)row + k]; na += qf[k] * qf[k]; }
            const float eps = __fmul_ru(1.52587890625e-5f, __fadd_ru(__double2float_ru(na), *T.nmax));      // 2^-16 (|a|^2 + max |b|^2)
            double best = INFINITY;
            int bj = -1;
            const int cand[4] = {i0, i1, i2, i3};
            const unsigned long long npb = (unsigned long long)n_tiles * TN;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = cand[u] < 0 ? -1 : (int)(((unsigned long long)cand[u] * PERM_A) % npb);      // packed position -> database row
                if (j >= 0 && j < nt) {
                    const double sdist = fgr_feat_dist2(qf, Ft + 33 * (size_t)j);
                    if (sdist < best || (sdist == best && j < bj)) { best = sdist; bj = j; }
                }
            }
            // every database row that is not in the list has an approximate v >= t3, i.e. an exact distance >= |a|^2 + t3 - eps:
            // if that is above the best exact distance in the list, the list's best is the nearest neighbour (t3 = +inf: the list
            // holds every row).  `best` is exact, so only one eps is spent.
            const bool certain = (double)t3 - (double)eps > (best - na) * (1.0 + 1e-12) + 1e-9;
            if (certain) {
                out[row] = bj;
            } else {
                out[row] = -1;
                const int slot = atomicAdd(&A.fb_count[blockIdx.y], 1);
                A.fb_list[(size_t)blockIdx.y * A.fb_stride + slot] = row;
                if (slot < FB_CAP) { FbPart sd; sd.d = best; sd.j = bj; sd.pad = 0; A.fb_seed[(size_t)blockIdx.y * FB_CAP + slot] = sd; }
            }
        }
    done:;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---- brute-force fp64 search for the rows the tensor-core pass could not settle ---------------------------------------------
// grid (FB_SLICES, 2 * pairs): block (slice, pair-direction) scans database rows [slice range] for every queued row (one thread
// per queued row, database rows staged through shared memory like k_fgr_nn) and writes its (distance, index) per row;
// k_fgr_match_fb2 takes the minimum over the slices (ties to the lower index).
__global__ void __launch_bounds__(FB_NT) k_fgr_match_fb(MatchArgs A, FbPart *part) {
    __shared__ double tile[FB_TILE][33];
    const FgrPair &pr = A.pairs[blockIdx.y >> 1];
    const int dir = blockIdx.y & 1;
    const int cq = dir == 0 ? pr.fj : pr.fi, ct = dir == 0 ? pr.fi : pr.fj;
    const int cloud_q = cq == 0 ? pr.src : pr.tgt, cloud_t = ct == 0 ? pr.src : pr.tgt;
    const int nt = A.packs[cloud_t].n;
    const double *Fq = A.feat + 33 * A.cloud_off[cloud_q], *Ft = A.feat + 33 * A.cloud_off[cloud_t];
    const int nfb = min(A.fb_count[blockIdx.y], FB_CAP);
    const int per = ((nt + FB_SLICES - 1) / FB_SLICES + FB_TILE - 1) / FB_TILE * FB_TILE;
    const int lo = min((int)blockIdx.x * per, nt), hi = min(lo + per, nt);
    for (int e0 = 0; e0 < nfb; e0 += FB_NT) {                       // block-uniform
        const int e = e0 + threadIdx.x;
        const int row = e < nfb ? A.fb_list[(size_t)blockIdx.y * A.fb_stride + e] : 0;
        double qf[33];
#pragma unroll
        for (int k = 0; k < 33; ++k) qf[k] = Fq[33 * (size_t)row + k];
        // The scan starts from the best exact distance among the row's tensor-core candidates (usually already the answer: the row
        // is here because several rows are nearly as close, not because the candidates are bad).  A database row is first
        // summed with FMAs, a third of the bins at a time: all terms are >= 0, so once the partial sum -- within 33 * 2^-53 of the
        // plainly summed one, the margin below is 1e-13 -- is not below `best` for ANY lane, the row cannot win (nor tie) and
        // the warp moves on, nearly always after the first 11 bins.  Only a row that passes is evaluated with k_fgr_nn's own
        // arithmetic, and ties go to the lower index.
        const bool live = e < nfb;
        FbPart sd; sd.d = INFINITY; sd.j = -1;
        if (live) sd = A.fb_seed[(size_t)blockIdx.y * FB_CAP + e];
        double best = live ? sd.d : -INFINITY;
        int32_t bj = live ? sd.j : -1;
        if (live && bj < 0) best = INFINITY;
        const double SAFE = 1.0 - 1e-13;
        for (int t0 = lo; t0 < hi; t0 += FB_TILE) {
            __syncthreads();
            for (int x = threadIdx.x; x < FB_TILE * 33; x += FB_NT) {
                const int t = t0 + x / 33;
                tile[x / 33][x % 33] = t < hi ? Ft[33 * (size_t)t + x % 33] : 0.0;
            }
            __syncthreads();
            const int m = min(FB_TILE, hi - t0);
            for (int b = 0; b < m; ++b) {
                const double *tb = tile[b];
                double sf = 0.0;
#pragma unroll
                for (int k = 0; k < 11; ++k) { const double d = qf[k] - tb[k]; sf = fma(d, d, sf); }
                if (!__any_sync(0xffffffffu, sf * SAFE <= best)) continue;
#pragma unroll
                for (int k = 11; k < 22; ++k) { const double d = qf[k] - tb[k]; sf = fma(d, d, sf); }
                if (!__any_sync(0xffffffffu, sf * SAFE <= best)) continue;
#pragma unroll
                for (int k = 22; k < 33; ++k) { const double d = qf[k] - tb[k]; sf = fma(d, d, sf); }
                if (sf * SAFE <= best) {                      // '<=': identical descriptors (distance 0) tie, the lower index wins
                    const double sdist = fgr_feat_dist2(qf, tb);
                    const int t = t0 + b;
                    if (sdist < best || (sdist == best && t < bj)) { best = sdist; bj = t; }
                }
            }
        }
        if (e < nfb) {
            FbPart o; o.d = best; o.j = bj; o.pad = 0;
            part[((size_t)blockIdx.y * FB_CAP + e) * FB_SLICES + blockIdx.x] = o;
        }
    }
}
// grid (chunks, 2 * pairs), one thread per queued row
__global__ void __launch_bounds__(128) k_fgr_match_fb2(MatchArgs A, const FbPart *part) {
    const FgrPair &pr = A.pairs[blockIdx.y >> 1];
    const int dir = blockIdx.y & 1;
    int32_t *out = dir == 0 ? pr.j2i : pr.i2j;
    const int nfb = min(A.fb_count[blockIdx.y], FB_CAP);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nfb; e += gridDim.x * blockDim.x) {
        double best = INFINITY;
        int32_t bj = -1;
        for (int sl = 0; sl < FB_SLICES; ++sl) {
            const FbPart o = part[((size_t)blockIdx.y * FB_CAP + e) * FB_SLICES + sl];
            if (o.j >= 0 && (o.d < best || (o.d == best && o.j < bj))) { best = o.d; bj = o.j; }      // every slice starts from the same seed
        }
        out[A.fb_list[(size_t)blockIdx.y * A.fb_stride + e]] = bj;
    }
}

}  // namespace fgrtc
