// Convolution forward / data-gradient on tcgen05 (TF32), TMA operands, operand reuse in shared memory, persistent CTAs
// with double-buffered accumulators (math mode 3).
//
// What the measurements on tc_conv_tma.cu (mode 2) said (tests/diag_conv_timeline.py, profiles/r1_conv_timeline_*.txt):
//   * its issue loops run inside `if (lane == 0)`: ptxas then treats every operand as lane-varying and wraps each
//     UTCHMMA / UTMALDG in an ELECT + 5 x R2UR + BRA.U.ANY waterfall -- ~200 clk of issue per 64-clk MMA;
//   * one tile per CTA: the co-resident CTAs of an SM, and in fact the whole chip, run in lockstep -- a main-loop phase with
//     idle HBM, then an epilogue phase (every CTA storing) with an idle tensor pipe, each about half of the time;
//   * the epilogue stored from the tcgen05.ld layout: 16 bytes to each of 32 different lines per instruction, and the
//     address arithmetic (runtime division by the patch width) dominated its instruction count.
// This kernel:
//   * whole-warp issue loops, `elect.sync` only around the instructions that must come from one thread; all operands are
//     warp-uniform (uniform registers, no waterfall); descriptors are built from a constant high word + a 32-bit low word;
//   * the issuing warps have the HIGHEST warp ids (the SM sub-partition arbiter prefers high warp ids): the instruction-heavy
//     epilogue warps never delay an MMA / TMA issue;
//   * persistent: one CTA per SM walks tiles round-robin; the accumulators of tile i (MT*BN TMEM columns) are drained by the
//     epilogue warps while the MMA warp fills the other set with tile i+1; the TMA rings run ahead across tiles;
//   * y-tap reuse: GEMM rows of a sub-tile are a (bh x bw) patch of output pixels ordered (py, px) with bw % 8 == 0, so the
//     A operands of the vertical taps of one kernel column are the SAME box shifted by whole patch rows = multiples of the
//     1024-byte swizzle atom: ONE box of (bh + taps - 1) patch rows per (channel chunk, tx, y-phase), each tap a different
//     start address in the matrix descriptor (3x3: 3 boxes of 18 rows instead of 9 of 16);
//   * weight reuse: MT sub-tiles (MT accumulators) per CTA share every weight box;
//   * image-spanning patches: on the small maps of the deep layers (10 x 53, 20 x 106) a 128-pixel patch of ONE image wastes up
//     to a third of the GEMM rows on padding.  A sub-tile may therefore take its (bh x bw) patch from nb consecutive images
//     (bw * bh * nb == 128).  The box is loaded through a tensor map whose dimensions are ordered (C, W, B, H), so shared
//     memory holds rows ordered (py, image, px): a vertical tap is still ONE start-address shift (of nb * bw rows);
//   * epilogue: TMEM -> registers -> padded shared-memory chunk -> full 128-byte row segments, 4 rows per store instruction,
//     row pointers computed once per sub-tile; masked column statistics from the staged chunk.
//
//   * two MMA-issuing warps, one per accumulator (sub-tile): the tensor pipe accepts about one MMA ahead of the one it is
//     executing (measured: per tap, time = issue-loop overhead + 8 x 64 clk, not the maximum of the two), so a single
//     issuing thread's loop overhead (~50 clk per MMA) is exposed; two threads hide each other's.
//
//   warps 0-3: epilogue (TMEM lane quadrant = warp % 4)   warp 4: TMA producer   warps 5,6: MMA issuers (5 owns the TMEM allocation)
//   (EPI_WARPS = 8 shifts the producer / issuers to warps 8 / 9,10)
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "tc_api.h"
#include "tc_common.cuh"

namespace {

using namespace sdt_tc;

constexpr int BKF = 32;
// Epilogue warps (4 or 8): warp w reads TMEM lane quadrant w % 4 (the hardware's restriction) and takes the 32-column chunks with
// chunk % (EPI_WARPS / 4) == w / 4.  Eight warps were measured and rejected (profiles/r2_ablation_epilogue_warps.txt): the small-K
// launches that look epilogue-bound in ncu (64 -> 64 stride 2: tensor pipe 31-40 %, DRAM 30-40 %) did not get faster, and the 18 KB
// of extra staging cost the deep layers a weight stage: 3.00 instead of 2.97 ms per step.
constexpr int EPI_WARPS = 4;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int W_TMA = EPI_WARPS;       // warp ids: epilogue 0 .. EPI_WARPS-1, TMA producer, two MMA issuers (highest ids)
constexpr int W_MMA = EPI_WARPS + 1;
constexpr int THREADS = EPI_THREADS + 96;
constexpr int MAX_TH = 8;
constexpr int MAX_GROUPS = 4;
constexpr int B_RING_MAX = 6, A_RING_MAX = 4;
constexpr int SMEM_MAX = 227 * 1024;          // one persistent CTA per SM
constexpr int N_BARS = 2 * A_RING_MAX + 2 * B_RING_MAX + 4;
constexpr int TAIL_BYTES = N_BARS * 8 + 16;   // barriers, TMEM slot
constexpr int STG_PITCH = 36;                 // floats per staged row: 32 columns + 4 (conflict-free 128-bit rows)
// epilogue scratch: 4 warps x 32 rows x STG_PITCH staging + statistics partials [2][slots][BN] per sub-tile; a slot is a run of
// GEMM rows of one image: 32 rows (4 slots, one per epilogue warp) for single-image patches, min(bw, 32) rows for image-spanning
// ones (up to 16).  With 4 slots all MT sub-tiles keep their own partials and the epilogue warps meet once per tile; with more
// slots one set is reused (two barriers per sub-tile; those are the deep, main-loop-bound layers).
__host__ __device__ constexpr int epi_bytes(int bn, int mt, int slots) {
    return EPI_WARPS * 32 * STG_PITCH * 4 + (slots > 4 ? 1 : mt) * 2 * slots * bn * 4;
}

struct YGeom {
    int bw, bh, nb;          // sub-tile patch: bw x bh pixels of nb consecutive images (bw * bh * nb == 128, bw % 8 == 0)
    int lbw, lnb;            // log2(bw), log2(nb)
    int lrun;                // log2 of the statistics run length (rows of one image that are adjacent in the GEMM tile)
    int tiles_x, tiles_y;    // patches per image
    int subtiles;            // ceil(B / nb) * tiles_x * tiles_y
    int box_rows;            // bh + (taps per group - 1)
    int a_box_bytes;         // box_rows * nb * bw * 128
    int a_stages, b_stages;
    int n_groups;            // tap groups = boxes per (channel chunk, tx)
    // per group, 16 bits: [3:0] y_add + 8 (box origin = y0*y_mul + y_off + y_add), [7:4] taps, [11:8] first ty, [15:12] ty step + 8;
    // tap j of a group reads the box shifted by j patch rows
    unsigned long long groups;
};
// Up to four problems that share source, destination, weights' shape and patch geometry and differ in their grid and offsets
// only -- the stride-parity classes of one data gradient -- run as ONE launch: four 140-tile launches of a 2 x 2-tap
// problem each paid their own pipeline fill and an epilogue that nothing overlapped; as one 560-tile launch the persistent
// CTAs drain the accumulators of one class while the tensor pipe works on the next.  A single problem is the n == 1 case.
constexpr int MAX_CLASSES = 4;
struct YClasses {
    int n;
    int u_min;                            // min over classes of their unit count: units [0, n * u_min) are walked class-minor
    int unit_start[MAX_CLASSES + 1];      // prefix sums of ceil(subtiles / MT): a CTA's MT sub-tiles share the weight boxes
    int subtiles[MAX_CLASSES], tiles_x[MAX_CLASSES], tpi[MAX_CLASSES];
    int GH[MAX_CLASSES], GW[MAX_CLASSES];
    int y_off[MAX_CLASSES], x_off[MAX_CLASSES], dy_off[MAX_CLASSES], dx_off[MAX_CLASSES];
};
struct YMaps {
    CUtensorMap b[MAX_CLASSES];           // the weight operand of each class
};
// (kernel parameters: select without dynamic indexing, which would copy the struct to local memory)
__device__ __forceinline__ int sel4(const int (&a)[MAX_CLASSES], int c) { return c == 0 ? a[0] : c == 1 ? a[1] : c == 2 ? a[2] : a[3]; }

__host__ __device__ __forceinline__ int grp_y_add(unsigned long long p, int gi) { return (int)((p >> (16 * gi)) & 15u) - 8; }
__host__ __device__ __forceinline__ int grp_taps(unsigned long long p, int gi) { return (int)((p >> (16 * gi + 4)) & 15u); }
__host__ __device__ __forceinline__ int grp_ty0(unsigned long long p, int gi) { return (int)((p >> (16 * gi + 8)) & 15u); }
__host__ __device__ __forceinline__ int grp_step(unsigned long long p, int gi) { return (int)((p >> (16 * gi + 12)) & 15u) - 8; }

constexpr uint32_t DESC_HI = desc_hi(1024, kSwizzle128B);        // K-major SWIZZLE_128B: 8-row groups 1024 B apart
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return sdt_tc::desc_lo(smem_addr, 16); }
__device__ __forceinline__ void mma_tf32_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    mma_tf32_lohi(tmem_d, a_lo, b_lo, DESC_HI, idesc, accumulate);
}

// profiling aids (tests/diag_conv_timeline.py); only the DBG instantiations read them
__device__ long long* g_timeline = nullptr;      // 8 x int64 per CTA
__device__ int g_timeline_ctas = 0;
__device__ int g_dbg_flags = 0;                  // 1 no TMA traffic, 2 epilogue drains TMEM only, 4 no MMAs (results are wrong)
__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int BN, int MT, bool DBG>
__global__ void __launch_bounds__(THREADS, 1) tc_conv_ytap_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ YMaps tmBs,
                                                                const sdt_conv_desc d, const YGeom g, const YClasses cl) {
    constexpr int B_BYTES = BN * 128;
    constexpr int ACC_COLS = MT * BN;                 // one accumulator set; two sets (double buffer) are allocated
    constexpr int N_ISS = MT >= 2 ? 2 : 1;            // MMA-issuing warps; issuer i owns sub-tiles i, i + N_ISS, ...
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smA = smem_u32(smem_raw);
    const int a_stage_bytes = MT * g.a_box_bytes;
    const int ring_bytes = g.a_stages * a_stage_bytes + g.b_stages * B_BYTES;
    const int slots = 128 >> g.lrun;
    const int epi = epi_bytes(BN, MT, slots);
    {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if ((smA & 1023u) != 0 || ring_bytes + epi + TAIL_BYTES > (int)dyn) {
            if (threadIdx.x == 0) printf("tc_conv_ytap_kernel: shared-memory window misaligned or too small\n");
            __trap();
        }
    }
    const uint32_t smB = smA + g.a_stages * a_stage_bytes;
    float* stg_all = reinterpret_cast<float*>(smem_raw + ring_bytes);              // 4 warps x 32 rows x STG_PITCH floats
    float* s_red = stg_all + EPI_WARPS * 32 * STG_PITCH;                                   // [MT or 1][2][slots][BN]
    const uint32_t bars = smA + ring_bytes + epi;
    // barriers: fullA[A_RING_MAX], emptyA[..], fullB[B_RING_MAX], emptyB[..], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + ring_bytes + epi + N_BARS * 8);
    auto fullA = [&](int s) { return bars + 8u * s; };
    auto emptyA = [&](int s) { return bars + 8u * (A_RING_MAX + s); };
    auto fullB = [&](int s) { return bars + 8u * (2 * A_RING_MAX + s); };
    auto emptyB = [&](int s) { return bars + 8u * (2 * A_RING_MAX + B_RING_MAX + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (2 * A_RING_MAX + 2 * B_RING_MAX + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (2 * A_RING_MAX + 2 * B_RING_MAX + 2 + a); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.N;
    const int n_ntiles = N / BN;
    const int n_tiles = (cl.n == 1 ? cl.unit_start[1] : cl.n == 2 ? cl.unit_start[2] : cl.n == 3 ? cl.unit_start[3] : cl.unit_start[4]) * n_ntiles;
    // tile -> class, first sub-tile within the class
    auto class_of = [&](int unit) {
        int c = 0;
#pragma unroll
        for (int k = 1; k < MAX_CLASSES; ++k)
            if (k < cl.n && unit >= cl.unit_start[k]) c = k;
        return c;
    };
    auto unit_base = [&](int c) { return c == 0 ? 0 : c == 1 ? cl.unit_start[1] : c == 2 ? cl.unit_start[2] : cl.unit_start[3]; };
    // Walk order of the units of a multi-class launch: position-major, class-minor.  The parity classes of a strided data gradient
    // read the SAME dy region for the same position; walked class after class (the first version) every class streamed dy from
    // DRAM again (ncu: 267 MB read for 70 MB of dy on the 64 -> 64 stride-2 layer); with the classes of one position on
    // co-running CTAs three of the four reads hit L2.  Units beyond n * u_min (classes of unequal size) follow in class order.
    auto unit_of = [&](int v) {
        if (cl.n == 1) return v;
        const int lim = cl.n * cl.u_min;
        if (v < lim) {
            const int q = v / cl.n;
            return unit_base(v - q * cl.n) + q;
        }
        int r = v - lim;
#pragma unroll
        for (int k = 0; k < MAX_CLASSES; ++k) {
            if (k < cl.n) {
                const int extra = (k == 0 ? cl.unit_start[1] : k == 1 ? cl.unit_start[2] - cl.unit_start[1]
                                   : k == 2 ? cl.unit_start[3] - cl.unit_start[2] : cl.unit_start[4] - cl.unit_start[3]) - cl.u_min;
                if (r < extra) return unit_base(k) + cl.u_min + r;
                r -= extra;
            }
        }
        return v;
    };
    const int chunks = d.C / BKF;
    const int dbg = DBG ? g_dbg_flags : 0;
    long long* tl = nullptr;
    if (DBG) {
        tl = (g_timeline != nullptr && (int)blockIdx.x < g_timeline_ctas) ? g_timeline + 8 * blockIdx.x : nullptr;
        if (tid == 0 && tl != nullptr) {
            tl[0] = -clock64();
            tl[1] = gtimer();
        }
    }

    if (tid == 0) {
        for (int s = 0; s < A_RING_MAX; ++s) {
            mbar_init(fullA(s), 1);
            mbar_init(emptyA(s), N_ISS);
        }
        for (int s = 0; s < B_RING_MAX; ++s) {
            mbar_init(fullB(s), 1);
            mbar_init(emptyB(s), N_ISS);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full(a), N_ISS);
            mbar_init(tmem_empty(a), EPI_WARPS);    // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), 2 * ACC_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation) overlapped the tail of the preceding kernel; global memory from here on
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    if (DBG && tl && tid == 0) tl[2] = gtimer();

    if (warp == W_TMA) {
        // ================= TMA producer (whole warp walks the loop, one elected lane issues) =================
        int sa = 0, pa = 1, sbi = 0, pb = 1;      // ring slot + parity to wait on the empty barriers (first lap passes)
        long long wait_empty = 0;
        for (int tile = blockIdx.x; tile < n_tiles && !(dbg & 1); tile += gridDim.x) {
            const int vunit = tile / n_ntiles;
            const int n0 = (tile - vunit * n_ntiles) * BN;
            const int grp_idx = unit_of(vunit);
            const int cls = class_of(grp_idx);
            const int t0 = (grp_idx - unit_base(cls)) * MT;
            const int c_sub = sel4(cl.subtiles, cls), tpi = sel4(cl.tpi, cls), c_tx = sel4(cl.tiles_x, cls);
            const int c_yoff = sel4(cl.y_off, cls), c_xoff = sel4(cl.x_off, cls);
            const CUtensorMap* tmB = &tmBs.b[0] + cls;
            const int nvalid = min(MT, c_sub - t0);
            int sb[MT], sy[MT], sx[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int t = min(t0 + m, c_sub - 1);
                sb[m] = t / tpi;
                const int rem = t - sb[m] * tpi;
                sb[m] <<= g.lnb;                               // first image of the patch
                sy[m] = (rem / c_tx) * g.bh * d.y_mul + c_yoff;
                sx[m] = (rem % c_tx) * g.bw * d.x_mul + c_xoff;
            }
            for (int ch = 0; ch < chunks; ++ch) {
                const int c0 = ch * BKF;
                for (int tx = 0; tx < d.TW; ++tx) {
                    for (int gi = 0; gi < g.n_groups; ++gi) {
                        const int y_add = grp_y_add(g.groups, gi), taps = grp_taps(g.groups, gi);
                        const int step = grp_step(g.groups, gi);
                        int ty = grp_ty0(g.groups, gi);
                        long long tw = 0;
                        if (DBG && tl) tw = clock64();
                        mbar_wait_spin(emptyA(sa), (uint32_t)pa);
                        if (DBG && tl) wait_empty += clock64() - tw;
                        if (elect_one()) {
                            mbar_expect_tx(fullA(sa), (uint32_t)(nvalid * g.a_box_bytes));
#pragma unroll
                            for (int m = 0; m < MT; ++m)
                                if (m < nvalid) {
                                    // tensor-map dimensions: (C, W, H, B), or (C, W, B, H) for image-spanning patches
                                    const int yy = sy[m] + y_add;
                                    tma_load_4d(smA + sa * a_stage_bytes + m * g.a_box_bytes, &tmA, c0, sx[m] + tx * d.tx_mul,
                                                g.nb > 1 ? sb[m] : yy, g.nb > 1 ? yy : sb[m], fullA(sa));
                                }
                        }
                        __syncwarp();
                        if (++sa == g.a_stages) { sa = 0; pa ^= 1; }
                        for (int j = 0; j < taps; ++j, ty += step) {
                            if (DBG && tl) tw = clock64();
                            mbar_wait_spin(emptyB(sbi), (uint32_t)pb);
                            if (DBG && tl) wait_empty += clock64() - tw;
                            if (elect_one()) {
                                mbar_expect_tx(fullB(sbi), B_BYTES);
                                tma_load_2d(smB + sbi * B_BYTES, tmB, (ty * d.TW + tx) * d.C + c0, n0, fullB(sbi));
                            }
                            __syncwarp();
                            if (++sbi == g.b_stages) { sbi = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
        if (DBG && tl && lane == 0) tl[6] = wait_empty;
    } else if (warp >= W_MMA) {
        // ================= MMA issuers (whole warp walks the loop, one elected lane issues) =================
        const int iss = warp - W_MMA;
        if (iss < N_ISS) {
        const uint32_t idesc = make_idesc_tf32(BN, 0, 0);
        const uint32_t shift_lo = (uint32_t)(g.bw * g.nb * 128) >> 4;    // one patch row (of all nb images), in descriptor address units
        const uint32_t box_lo = (uint32_t)g.a_box_bytes >> 4;
        int sa = 0, pa = 0, sbi = 0, pb = 0;
        long long wait_full = 0, wait_acc = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int grp_idx = unit_of(tile / n_ntiles);
            const int cls = class_of(grp_idx);
            const int t0i = (grp_idx - unit_base(cls)) * MT;
            const int c_subi = sel4(cl.subtiles, cls);
            const int nvalid = min(MT, c_subi - t0i);
            // Vertical taps that fall outside the source for EVERY row of a sub-tile contribute zeros (TMA's out-of-bounds fill): skip
            // their MMAs.  The 6 x 3 / no-padding layer's data gradient walks a 10-row grid over a 5-row source: half of its taps.
            // yb = source row of patch row 0 at tap shift 0 (before the group's y_add); rows_m = patch rows inside the grid.
            int yb[MT / N_ISS], rows_m[MT / N_ISS];
            {
                const int tpi_i = sel4(cl.tpi, cls), c_txi = sel4(cl.tiles_x, cls), c_GHi = sel4(cl.GH, cls), c_yoffi = sel4(cl.y_off, cls);
#pragma unroll
                for (int mm = 0; mm < MT / N_ISS; ++mm) {
                    const int t = min(t0i + iss + mm * N_ISS, c_subi - 1);
                    const int rem = t - (t / tpi_i) * tpi_i;
                    const int y0 = (rem / c_txi) * g.bh;
                    yb[mm] = y0 * d.y_mul + c_yoffi;
                    rows_m[mm] = min(g.bh, c_GHi - y0);
                }
            }
            uint32_t started_m = 0;                  // bit mm: accumulator of sub-tile iss + mm * N_ISS has received an MMA
            const int acc = it & 1;
            const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * ACC_COLS);
            long long tw = 0;
            if (DBG && tl) tw = clock64();
            mbar_wait_spin(tmem_empty(acc), (uint32_t)(((it >> 1) & 1) ^ 1));     // the epilogue has drained this accumulator set
            if (DBG && tl) wait_acc += clock64() - tw;
            tc_fence_after();
            for (int ch = 0; ch < chunks; ++ch) {
                for (int tx = 0; tx < d.TW; ++tx) {
                    for (int gi = 0; gi < g.n_groups; ++gi) {
                        const int taps = grp_taps(g.groups, gi);
                        const int y_add = grp_y_add(g.groups, gi);
                        const bool first_box = (ch | tx | gi) == 0;
                        if (DBG && tl) tw = clock64();
                        if (!(dbg & 1)) mbar_wait_spin(fullA(sa), (uint32_t)pa);
                        if (DBG && tl) wait_full += clock64() - tw;
                        uint32_t a_lo = desc_lo(smA + sa * a_stage_bytes);
                        for (int j = 0; j < taps; ++j, a_lo += shift_lo) {
                            if (DBG && tl) tw = clock64();
                            if (!(dbg & 1)) mbar_wait_spin(fullB(sbi), (uint32_t)pb);
                            if (DBG && tl) wait_full += clock64() - tw;
                            if (!(dbg & 16)) tc_fence_after();
                            const uint32_t b_lo = desc_lo(smB + sbi * B_BYTES);
                            if (!(dbg & 64) && elect_one()) {
#pragma unroll
                                for (int mm = 0; mm < MT / N_ISS; ++mm) {
                                    const int m = iss + mm * N_ISS;
                                    // source rows of this tap: ys + py * y_mul, py < rows_m; the very first tap is always issued (it
                                    // initialises the accumulator, zeros are harmless)
                                    const int ys = yb[mm] + y_add + j * d.y_mul;
                                    const bool live = (first_box && j == 0) || (ys + (rows_m[mm] - 1) * d.y_mul >= 0 && ys < d.SH);
                                    if (m < nvalid && live && !(dbg & 4)) {
                                        const uint32_t st = (started_m >> mm) & 1u;
#pragma unroll
                                        for (int k4 = 0; k4 < 4; ++k4)
                                            mma_tf32_lo(tmem_acc + (uint32_t)(m * BN), a_lo + m * box_lo + 2u * k4, b_lo + 2u * k4, idesc,
                                                        st | (uint32_t)k4);
                                        started_m |= 1u << mm;
                                    }
                                }
                                if (!(dbg & 8)) mma_commit(emptyB(sbi));
                            }
                            if (!(dbg & 64)) __syncwarp();
                            if (++sbi == g.b_stages) { sbi = 0; pb ^= 1; }
                        }
                        if (!(dbg & 32) && elect_one()) mma_commit(emptyA(sa));
                        __syncwarp();
                        if (++sa == g.a_stages) { sa = 0; pa ^= 1; }
                    }
                }
            }
            if (elect_one()) mma_commit(tmem_full(acc));
            __syncwarp();
            if (DBG && tl && lane == 0 && iss == 0 && it == 0) tl[3] = gtimer();
        }
        if (DBG && tl && lane == 0 && iss == 0) {
            tl[7] = wait_full;
            tl[4] = gtimer();
        }
        }
    } else {
        // ================= epilogue (TMEM lane quadrant = warp % 4, column-chunk parity = warp / 4), one tile behind the MMA issuer
        const int q = warp & 3, half = warp >> 2;
        constexpr int CH_STEP = EPI_WARPS / 4;                // chunks handled by this warp: half, half + CH_STEP, ...
        const int r = q * 32 + lane;
        // GEMM row -> (py, image, px)
        const int py = r >> (g.lbw + g.lnb), pn = (r >> g.lbw) & (g.nb - 1), px = r & (g.bw - 1);
        const bool per_m = slots > 4;                        // one set of statistics partials, reduced after every sub-tile
        // column sums of one sub-tile: the runs of each image, in row order -> partial rows [B][tiles_x * tiles_y][2][N]
        auto reduce_stats = [&](const float* red, int b0, int rem, int n0, int tpi) {
            for (int i = tid; i < BN << g.lnb; i += EPI_THREADS) {
                const int c = i % BN, n = i / BN;
                if (b0 + n >= d.B) continue;
                float s1 = 0.f, s2 = 0.f;
                for (int sl = 0; sl < slots; ++sl) {
                    if ((((sl << g.lrun) >> g.lbw) & (g.nb - 1)) != n) continue;
                    s1 += red[(0 * slots + sl) * BN + c];
                    s2 += red[(1 * slots + sl) * BN + c];
                }
                const size_t prow = (size_t)(b0 + n) * tpi + rem;
                d.stat_partial[(prow * 2 + 0) * N + n0 + c] = s1;
                d.stat_partial[(prow * 2 + 1) * N + n0 + c] = s2;
            }
        };
        float* stg = stg_all + (size_t)(warp * 32) * STG_PITCH;
        const int sub = lane >> 3, col4 = lane & 7;          // write-back: 8 lanes per 128-byte row segment
        long long wait_tmem = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int vunit = tile / n_ntiles;
            const int n0 = (tile - vunit * n_ntiles) * BN;
            const int grp_idx = unit_of(vunit);
            const int cls = class_of(grp_idx);
            const int t0 = (grp_idx - unit_base(cls)) * MT;
            const int tpi = sel4(cl.tpi, cls), c_tx = sel4(cl.tiles_x, cls);
            const int c_GH = sel4(cl.GH, cls), c_GW = sel4(cl.GW, cls), c_dyoff = sel4(cl.dy_off, cls), c_dxoff = sel4(cl.dx_off, cls);
            const int nvalid = min(MT, sel4(cl.subtiles, cls) - t0);
            const int acc = it & 1;
            long long tw = 0;
            if (DBG && tl) tw = clock64();
            mbar_wait(tmem_full(acc), (uint32_t)((it >> 1) & 1));
            if (DBG && tl) wait_tmem += clock64() - tw;
            tc_fence_after();
            for (int m = 0; m < nvalid; ++m) {
                const int t = t0 + m;
                const int ib = t / tpi;
                const int rem = t - ib * tpi;
                const int b0 = ib << g.lnb;
                const int ty0 = (rem / c_tx) * g.bh, tx0 = (rem % c_tx) * g.bw;
                const bool ok = (ty0 + py) < c_GH && (tx0 + px) < c_GW && (b0 + pn) < d.B;
                float* red_m = s_red + (per_m ? 0 : m * 2 * slots * BN);
                const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
                // destination of the 8 rows this lane writes back (rows sub, sub+4, ... of the warp's 32), column n0
                float* rowp[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int rr = q * 32 + 4 * k + sub;
                    const int gy = ty0 + (rr >> (g.lbw + g.lnb)), gx = tx0 + (rr & (g.bw - 1));
                    const int b = b0 + ((rr >> g.lbw) & (g.nb - 1));
                    rowp[k] = d.dst + (((long long)b * d.DH + (gy * d.dy_mul + c_dyoff)) * d.DW + (gx * d.dx_mul + c_dxoff)) * N + n0 + col4 * 4;
                }
                for (int c = half; c < BN / 32; c += CH_STEP) {
                    float v[32];
                    if (!(DBG && (dbg & 128)))
                        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + m * BN + c * 32), v);
                    if (m == nvalid - 1 && c + CH_STEP >= BN / 32) {  // this warp's last read of the accumulator set: hand it back to the MMA issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tmem_empty(acc));
                    }
                    if (DBG && (dbg & 2)) continue;
                    float4* row = reinterpret_cast<float4*>(stg + (size_t)lane * STG_PITCH);
#pragma unroll
                    for (int j = 0; j < 8; ++j) row[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    if (d.stat_partial != nullptr) {
                        // masked column sums of the staged chunk in four blocks of 8 rows (all loads before any store: the
                        // partials live in the same shared memory), combined into runs of 8, 16 or 32 rows of one image
                        float b1[4], b2[4];
#pragma unroll
                        for (int blk = 0; blk < 4; ++blk) {
                            float s1 = 0.f, s2 = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int i = blk * 8 + j;
                                const float x = stg[(size_t)i * STG_PITCH + lane];
                                if ((okmask >> i) & 1u) {
                                    s1 += x;
                                    s2 += x * x;
                                }
                            }
                            b1[blk] = s1;
                            b2[blk] = s2;
                        }
                        float* r1 = red_m + c * 32 + lane;
                        float* r2 = r1 + slots * BN;
                        if (g.lrun == 3) {
#pragma unroll
                            for (int blk = 0; blk < 4; ++blk) {
                                r1[(q * 4 + blk) * BN] = b1[blk];
                                r2[(q * 4 + blk) * BN] = b2[blk];
                            }
                        } else if (g.lrun == 4) {
                            r1[(q * 2 + 0) * BN] = b1[0] + b1[1];
                            r2[(q * 2 + 0) * BN] = b2[0] + b2[1];
                            r1[(q * 2 + 1) * BN] = b1[2] + b1[3];
                            r2[(q * 2 + 1) * BN] = b2[2] + b2[3];
                        } else {
                            r1[q * BN] = (b1[0] + b1[1]) + (b1[2] + b1[3]);
                            r2[q * BN] = (b2[0] + b2[1]) + (b2[2] + b2[3]);
                        }
                    }
                    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (d.bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(d.bias + n0 + c * 32) + col4);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int i = 4 * k + sub;
                        if ((okmask >> i) & 1u) {
                            float4* p = reinterpret_cast<float4*>(rowp[k] + c * 32);
                            float4 o = *reinterpret_cast<const float4*>(stg + (size_t)i * STG_PITCH + col4 * 4);
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                            if (d.accumulate) {
                                const float4 old = *p;
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            *p = o;
                        }
                    }
                    __syncwarp();
                }
                if (d.stat_partial != nullptr && per_m) {
                    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");          // the epilogue warps
                    reduce_stats(s_red, b0, rem, n0, tpi);
                    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");          // s_red is reused by the next sub-tile
                }
            }
            if (d.stat_partial != nullptr && !per_m) {
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                for (int m = 0; m < nvalid; ++m) {
                    const int ib = (t0 + m) / tpi;
                    reduce_stats(s_red + m * 2 * slots * BN, ib << g.lnb, t0 + m - ib * tpi, n0, tpi);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * ACC_COLS);
    }
    if (DBG && tl && tid == 0) {
        tl[5] = gtimer();
        tl[0] += clock64();
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct Plan {
    bool ok;
    int bn, mt, tiles;
    YGeom g;
    int smem;
    double cost;
};

// Vertical tap structure: the groups of taps that share one box.  Returns the taps per group (max), 0 if unsupported.
int tap_structure(const sdt_conv_desc* d, YGeom* g) {
    if (d->TH > MAX_TH) return 0;
    g->groups = 0;
    if (d->ty_mul == 1) {                       // y = gy*y_mul + y_off + ty: taps ty = p, p + y_mul, ... share the box of phase p
        const int n = d->TH < d->y_mul ? d->TH : d->y_mul;
        if (n > MAX_GROUPS || d->y_mul > 7) return 0;
        g->n_groups = n;
        int max_taps = 0;
        for (int p = 0; p < n; ++p) {
            const int taps = (d->TH - p + d->y_mul - 1) / d->y_mul;
            if (taps > max_taps) max_taps = taps;
            g->groups |= (unsigned long long)((p + 8) | (taps << 4) | (p << 8) | ((d->y_mul + 8) << 12)) << (16 * p);
        }
        return max_taps;
    }
    if (d->ty_mul == -1 && d->y_mul == 1) {     // data gradient: y = gy + y_off - ty: one box from ty = TH-1 (shift 0) down to 0
        g->n_groups = 1;
        g->groups = (unsigned long long)((-(d->TH - 1) + 8) | (d->TH << 4) | ((d->TH - 1) << 8) | ((-1 + 8) << 12));
        return d->TH;
    }
    return 0;
}

// Image-spanning patches need a tensor map whose dimension order is (C, W, B, H): strides not ascending.  Probe the driver
// once (encoding dereferences nothing); without it the planner keeps to single-image patches.
bool spanning_maps_ok() {
    static int ok = -1;
    if (ok < 0) {
        ok = 0;
        EncodeTiledFn enc = get_encode();
        if (enc != nullptr) {
            alignas(64) CUtensorMap tm;
            const cuuint64_t dims[4] = {64, 53, 32, 10};
            const cuuint64_t strides[3] = {64 * 4, 10 * 53 * 64 * 4, 53 * 64 * 4};
            const cuuint32_t box[4] = {32, 8, 8, 4};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            ok = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, reinterpret_cast<void*>(0x10000), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        } else {
            ok = 1;          // no driver (host-only planning, e.g. sdt_conv_plan in the CPU tests): plan as the GPU box would
        }
    }
    return ok == 1;
}

// tuning / test overrides of the planner: N tile, accumulators per CTA, patch width, images per patch (0 = planner's choice);
// initialised from SDT_YTAP_BN / _MT / _BW / _NB, changed by sdt_debug_conv_force (tests force the image-spanning geometries)
int env_int(const char* name) { return getenv(name) ? atoi(getenv(name)) : 0; }
int g_force[4] = {env_int("SDT_YTAP_BN"), env_int("SDT_YTAP_MT"), env_int("SDT_YTAP_BW"), env_int("SDT_YTAP_NB")};

// Fraction of the (tile row, vertical tap) pairs whose source rows are not all outside the source: the MMA issuers skip the rest.
double live_tap_fraction(const sdt_conv_desc* d, const YGeom& g) {
    long long live = 0, all = 0;
    for (int ty = 0; ty < g.tiles_y; ++ty) {
        const int y0 = ty * g.bh, rows = std::min(g.bh, d->GH - y0);
        for (int gi = 0; gi < g.n_groups; ++gi)
            for (int j = 0; j < grp_taps(g.groups, gi); ++j) {
                const int ys = y0 * d->y_mul + d->y_off + grp_y_add(g.groups, gi) + j * d->y_mul;
                live += (ys + (rows - 1) * d->y_mul >= 0 && ys < d->SH) ? 1 : 0;
                ++all;
            }
    }
    return all ? (double)live / (double)all : 1.0;
}

// n_classes problems of (nearly) d's size share the launch: the tile count that fills the machine is the sum
Plan make_plan(const sdt_conv_desc* d, int n_classes = 1) {
    Plan best{};
    best.ok = false;
    if (d->N % 64 != 0 || d->C % 32 != 0) return best;
    const int K = d->TH * d->TW * d->C;
    const int force_bn = g_force[0], force_mt = g_force[1], force_bw = g_force[2], force_nb = g_force[3];
    const int nb_max = spanning_maps_ok() ? 16 : 1;
    for (int bn = 128; bn >= 64; bn /= 2) {
        if (d->N % bn != 0 || (force_bn && bn != force_bn && d->N % force_bn == 0)) continue;
        for (int mt = 4; mt >= 1; mt /= 2) {
            if (mt * bn > 256 || (force_mt && mt != force_mt && force_mt * bn <= 256)) continue;
            for (int bw = 8; bw <= 128; bw *= 2)
            for (int nb = 1; nb * bw <= 128 && nb <= nb_max; nb *= 2) {
                if ((force_bw && bw != force_bw) || (force_nb && nb != force_nb && force_nb * bw <= 128)) continue;
                if (nb > 1 && nb / 2 >= d->B) continue;               // never more than one half-empty patch of images
                YGeom g{};
                const int max_taps = tap_structure(d, &g);
                if (max_taps == 0) return best;
                g.bw = bw;
                g.nb = nb;
                for (g.lbw = 0; (1 << g.lbw) < bw; ++g.lbw) {}
                for (g.lnb = 0; (1 << g.lnb) < nb; ++g.lnb) {}
                g.lrun = (nb == 1 || bw >= 32) ? 5 : g.lbw;
                const int slots = 128 >> g.lrun;
                g.bh = 128 / (bw * nb);
                g.box_rows = g.bh + max_taps - 1;
                if (bw * d->x_mul > 256 || g.box_rows * d->y_mul > 256) continue;
                g.a_box_bytes = g.box_rows * nb * bw * 128;
                g.tiles_x = (d->GW + bw - 1) / bw;
                g.tiles_y = (d->GH + g.bh - 1) / g.bh;
                g.subtiles = ((d->B + nb - 1) / nb) * g.tiles_x * g.tiles_y;
                const long long tiles = (long long)((g.subtiles + mt - 1) / mt) * (d->N / bn) * n_classes;
                const int budget = SMEM_MAX - epi_bytes(bn, mt, slots) - TAIL_BYTES;
                const int a_stage = mt * g.a_box_bytes, b_stage = bn * 128;
                int as = 2, bs = 2;
                if (as * a_stage + bs * b_stage > budget) continue;
                for (bool grew = true; grew;) {      // spend what is left: weight boxes first (finer grained), then A stages
                    grew = false;
                    if (bs < B_RING_MAX && bs < 2 * as + 1 && as * a_stage + (bs + 1) * b_stage <= budget) { ++bs; grew = true; }
                    else if (as < A_RING_MAX && (as + 1) * a_stage + bs * b_stage <= budget) { ++as; grew = true; }
                }
                g.a_stages = as;
                g.b_stages = bs;
                // cost model (clocks per tile): shared-memory traffic (TMA fill + tensor-core operand reads, 128 B/clk) against
                // tensor-pipe clocks; the persistent CTAs (one per SM) run ceil(tiles / 148) tiles each
                const double fill = (double)(d->C / BKF) * d->TW * g.n_groups * g.a_box_bytes * mt + (double)K * bn * 4.0;
                const double live = live_tap_fraction(d, g);         // skipped taps neither read their operands nor use the tensor pipe
                const double reads = (double)mt * (K / 8) * (128 + bn) * 32.0 * live;
                const double mma_clk = (double)mt * (K / 8) * (bn / 2.0) * live;
                const double smem_clk = (fill + reads) / 128.0;
                const double rounds = (double)((tiles + 147) / 148);
                // a single issuing warp (MT == 1) exposes its loop overhead: measured ~1.5x slower per MMA than two issuers
                // (image-spanning patches only where they save tiles: at equal cost the single-image geometry wins)
                const double cost = (smem_clk > mma_clk ? smem_clk : mma_clk) * rounds * (mt == 1 ? 1.5 : 1.0) * (nb > 1 ? 1.03 : 1.0);
                if (!best.ok || cost < best.cost) {
                    best.ok = true;
                    best.bn = bn;
                    best.mt = mt;
                    best.g = g;
                    best.tiles = (int)tiles;
                    best.smem = as * a_stage + bs * b_stage + epi_bytes(bn, mt, slots) + TAIL_BYTES;
                    best.cost = cost;
                }
            }
        }
    }
    return best;
}

bool g_host_debug = false;       // set by sdt_debug_conv_timeline / sdt_debug_conv_flags: launch the DBG instantiation

template <int BN, int MT, bool DBG>
int launch_ytap2(const sdt_conv_desc* d, const Plan& pl, const CUtensorMap& tmA, const YMaps& tmB, const YClasses& cl, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        SDT_CUDA_OK(cudaFuncSetAttribute(tc_conv_ytap_kernel<BN, MT, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        attr_set = true;
    }
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        SDT_CUDA_OK(cudaGetDevice(&dev));
        SDT_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    // SDT_YTAP_SM_RESERVE=n: the persistent grid leaves n SMs free (multi-GPU: room for the NCCL kernels of the gradient buckets
    // that run beside the backward pass -- a 148-CTA persistent kernel holding every SM makes them queue behind it)
    static const int sm_reserve = getenv("SDT_YTAP_SM_RESERVE") ? atoi(getenv("SDT_YTAP_SM_RESERVE")) : 0;
    const int sm_use = sm_count - sm_reserve > 0 ? sm_count - sm_reserve : sm_count;
    const int tiles = cl.unit_start[cl.n] * (d->N / BN);
    const int grid = tiles < sm_use ? tiles : sm_use;          // persistent: one CTA per SM, static round-robin tiles
    sdt::launch(tc_conv_ytap_kernel<BN, MT, DBG>, dim3(grid), dim3(THREADS), pl.smem, st, tmA, tmB, *d, pl.g, cl);
    SDT_LAUNCH_OK("tc_conv_ytap_kernel");
    sdt_note_tc_launch();
    return SDT_OK;
}

// ds[0..n): the problems of one launch (n == 1: a plain convolution; n > 1: compatible parity classes, see multi_ok)
template <int BN, int MT>
int launch_ytap(const sdt_conv_desc* ds, int n, const Plan& pl, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    SDT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    const sdt_conv_desc* d = ds;
    const YGeom& g = pl.g;
    alignas(64) CUtensorMap tmA;
    alignas(64) YMaps tmB;
    YClasses cl{};
    {
        // (C, W, H, B) with one image per box, or (C, W, B, H) with nb images per box row: shared-memory rows (py, image, px)
        const bool span = g.nb > 1;
        const cuuint64_t row_stride = (cuuint64_t)d->SW * d->C * 4, img_stride = (cuuint64_t)d->SH * d->SW * d->C * 4;
        const cuuint32_t box_y = (cuuint32_t)(g.box_rows * d->y_mul);
        const cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->SW, (cuuint64_t)(span ? d->B : d->SH), (cuuint64_t)(span ? d->SH : d->B)};
        const cuuint64_t strides[3] = {(cuuint64_t)d->C * 4, span ? img_stride : row_stride, span ? row_stride : img_stride};
        const cuuint32_t box[4] = {32, (cuuint32_t)(g.bw * d->x_mul), span ? (cuuint32_t)g.nb : box_y, span ? box_y : 1u};
        const cuuint32_t estr[4] = {1, (cuuint32_t)d->x_mul, span ? 1u : (cuuint32_t)d->y_mul, span ? (cuuint32_t)d->y_mul : 1u};
        const CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->src), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A, y-tap) failed with %d", (int)r);
    }
    cl.n = n;
    for (int c = 0; c < MAX_CLASSES; ++c) {
        const sdt_conv_desc* dc = ds + (c < n ? c : n - 1);       // unused entries repeat the last class
        const int K = dc->TH * dc->TW * dc->C;
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)dc->N};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)BN};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmB.b[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dc->wt_nk), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B, y-tap) failed with %d", (int)r);
        cl.tiles_x[c] = (dc->GW + g.bw - 1) / g.bw;
        cl.tpi[c] = cl.tiles_x[c] * ((dc->GH + g.bh - 1) / g.bh);
        cl.subtiles[c] = ((dc->B + g.nb - 1) / g.nb) * cl.tpi[c];
        cl.GH[c] = dc->GH; cl.GW[c] = dc->GW;
        cl.y_off[c] = dc->y_off; cl.x_off[c] = dc->x_off; cl.dy_off[c] = dc->dy_off; cl.dx_off[c] = dc->dx_off;
        cl.unit_start[c + 1] = cl.unit_start[c] + (c < n ? (cl.subtiles[c] + MT - 1) / MT : 0);
    }
    cl.u_min = cl.unit_start[1];
    for (int c = 1; c < n; ++c) cl.u_min = std::min(cl.u_min, cl.unit_start[c + 1] - cl.unit_start[c]);
    static const bool class_major = getenv("SDT_YTAP_CLASS_MAJOR") != nullptr;      // tuning aid: the first version's class-after-class walk
    if (class_major) cl.u_min = 0;
    if (g_host_debug) return launch_ytap2<BN, MT, true>(d, pl, tmA, tmB, cl, st);
    return launch_ytap2<BN, MT, false>(d, pl, tmA, tmB, cl, st);
}

int launch_planned(const sdt_conv_desc* ds, int n, const Plan& pl, cudaStream_t st) {
    if (pl.bn == 128) {
        if (pl.mt == 2) return launch_ytap<128, 2>(ds, n, pl, st);
        return launch_ytap<128, 1>(ds, n, pl, st);
    }
    if (pl.mt == 4) return launch_ytap<64, 4>(ds, n, pl, st);
    if (pl.mt == 2) return launch_ytap<64, 2>(ds, n, pl, st);
    return launch_ytap<64, 1>(ds, n, pl, st);
}

}  // namespace

bool sdt_tc_conv_ytap_eligible(const sdt_conv_desc* d) { return sdt_tc_conv_ytap_shape_ok(d) && get_encode() != nullptr; }

bool sdt_tc_conv_ytap_shape_ok(const sdt_conv_desc* d) {
    if (d->wt_nk == nullptr || d->xf_scale != nullptr) return false;        // plain (already activated) source only
    if (d->C % 32 != 0 || d->N % 64 != 0) return false;
    if (d->GH < 2) return false;                                             // single-row maps: tc_conv_tma.cu (1-D layers are remapped by the dispatcher)
    if (d->x_mul < 1 || d->x_mul > 8 || d->y_mul < 1 || d->y_mul > 8) return false;
    if ((((uintptr_t)d->src | (uintptr_t)d->wt_nk | (uintptr_t)d->dst | (uintptr_t)d->bias) & 15) != 0) return false;
    return make_plan(d).ok;
}

// statistics rows: one per (image, patch position), image-major
int sdt_tc_conv_ytap_row_tiles(const sdt_conv_desc* d) {
    const Plan pl = make_plan(d);
    return d->B * pl.g.tiles_x * pl.g.tiles_y;
}

// profiling aids (not part of the public header)
extern "C" int sdt_debug_conv_flags(int flags) {
    SDT_CUDA_OK(cudaMemcpyToSymbol(g_dbg_flags, &flags, sizeof(flags)));
    g_host_debug = flags != 0;
    return SDT_OK;
}
extern "C" int sdt_debug_conv_force(int bn, int mt, int bw, int nb) {
    g_force[0] = bn; g_force[1] = mt; g_force[2] = bw; g_force[3] = nb;
    return SDT_OK;
}
// per-CTA timeline buffer of 8 x int64 records, or NULL to switch off
extern "C" int sdt_debug_conv_timeline(void* buf, int ctas) {
    long long* p = static_cast<long long*>(buf);
    SDT_CUDA_OK(cudaMemcpyToSymbol(g_timeline, &p, sizeof(p)));
    SDT_CUDA_OK(cudaMemcpyToSymbol(g_timeline_ctas, &ctas, sizeof(ctas)));
    g_host_debug = p != nullptr;
    return SDT_OK;
}

int sdt_tc_conv_ytap_describe(const sdt_conv_desc* d, int32_t* out10) {
    const Plan pl = make_plan(d);
    if (!pl.ok) return 0;
    out10[1] = pl.bn; out10[2] = pl.mt; out10[3] = pl.g.bh | (pl.g.nb << 8); out10[4] = pl.g.bw; out10[5] = pl.g.box_rows;
    out10[6] = pl.g.a_stages; out10[7] = pl.g.b_stages; out10[8] = pl.smem;
    out10[9] = pl.tiles;
    return 1;
}

int sdt_tc_conv_ytap_launch(const sdt_conv_desc* d, cudaStream_t st) {
    const Plan pl = make_plan(d);
    SDT_REQUIRE(pl.ok, "sdt_tc_conv_ytap_launch: no plan for this descriptor");
    return launch_planned(d, 1, pl, st);
}

// Can ds[0..n) run as one launch?  Same source / destination / shapes / tap structure; only grids and offsets differ, and
// no statistics epilogue (its partial rows are per problem).
bool sdt_tc_conv_ytap_multi_ok(const sdt_conv_desc* ds, int n) {
    if (n < 2 || n > MAX_CLASSES) return false;
    for (int c = 0; c < n; ++c) {
        const sdt_conv_desc& a = ds[0];
        const sdt_conv_desc& b = ds[c];
        if (!sdt_tc_conv_ytap_eligible(&b) || b.stat_partial != nullptr) return false;
        if (b.src != a.src || b.dst != a.dst || b.bias != a.bias || b.accumulate != a.accumulate) return false;
        if (b.B != a.B || b.SH != a.SH || b.SW != a.SW || b.C != a.C || b.N != a.N || b.TH != a.TH || b.TW != a.TW) return false;
        if (b.y_mul != a.y_mul || b.ty_mul != a.ty_mul || b.x_mul != a.x_mul || b.tx_mul != a.tx_mul) return false;
        if (b.DH != a.DH || b.DW != a.DW || b.dy_mul != a.dy_mul || b.dx_mul != a.dx_mul) return false;
        if (b.GH > a.GH || b.GW > a.GW || ((uintptr_t)b.wt_nk & 15) != 0) return false;      // planned on the largest (first) class
    }
    return make_plan(ds, n).ok;
}

int sdt_tc_conv_ytap_launch_multi(const sdt_conv_desc* ds, int n, cudaStream_t st) {
    SDT_REQUIRE(sdt_tc_conv_ytap_multi_ok(ds, n), "sdt_tc_conv_ytap_launch_multi: the problems cannot share a launch");
    return launch_planned(ds, n, make_plan(ds, n), st);
}
