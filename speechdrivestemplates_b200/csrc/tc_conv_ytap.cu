// Convolution forward / data-gradient on tcgen05 (TF32), TMA operands, with operand REUSE in shared memory (math mode 3).
//
// Why: tc_conv_tma.cu fetches one 128-pixel x 32-channel A box (16 KB) and one weight box per (tap, channel chunk), i.e.
// 24-32 KB of L2 -> shared-memory traffic per 128x{64,128}x32 MMA block.  Every encoder layer measured at the same
// ~12 TB/s of L2 -> SM traffic (profiles/r2_l2_bound_analysis.md): the kernel sits on the chip's L2 bandwidth cap
// (~6300 B/clk), not on the tensor pipe (22-33 %).  This kernel moves fewer bytes per FLOP:
//
//   * y-tap reuse.  GEMM rows of a sub-tile are a (bh x bw) patch of output pixels ordered (py, px) with bw % 8 == 0.
//     For a fixed horizontal tap tx, the A operands of the vertical taps ty are the SAME pixels shifted by whole patch
//     rows, i.e. by bw rows of 128 B = a multiple of the 1024-byte swizzle atom.  So ONE box of (bh + max_shift) patch
//     rows is loaded per (channel chunk, tx, y-phase) and each ty is just a different start address in the shared-memory
//     matrix descriptor (3x3: 3 boxes of 18 rows instead of 9 boxes of 16; 4x4 stride 2: two y-phases x 4 tx boxes of
//     17 rows instead of 16 boxes of 16).
//   * weight reuse.  A CTA owns MT sub-tiles (MT accumulators in TMEM, MT*BN columns); every weight box is used by
//     MT * 4 MMAs instead of 4.
//
//   warp 0  : TMA producer (one thread), two rings: A stages (MT boxes each) and B stages (one weight box per tap).
//   warp 1  : MMA issuer (one thread), owns the TMEM allocation.
//   warps 2-5: epilogue per sub-tile: tcgen05.ld, bias / accumulate, 128-byte row stores, masked column statistics.
// Two CTAs per SM (<= ~110 KB shared memory, <= 256 TMEM columns each): one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <stdlib.h>

#include "tc_api.h"
#include "tc_common.cuh"

namespace {

using namespace sdt_tc;

constexpr int BKF = 32;
constexpr int THREADS = 192;
constexpr int MAX_TH = 8;
constexpr int B_RING_MAX = 6, A_RING_MAX = 4;
constexpr int SMEM_TWO_PER_SM = 113 * 1024;   // 2 x (113 + 1 reserved) KB == 228 KB
constexpr int SMEM_ONE_PER_SM = 200 * 1024;
constexpr int TAIL_BYTES = (2 * A_RING_MAX + 2 * B_RING_MAX + 1) * 8 + 16;   // barriers, TMEM slot

struct YGeom {
    int bw, bh;              // sub-tile patch (bw * bh == 128, bw % 8 == 0)
    int tiles_x, tiles_y;    // patches per image
    int subtiles;            // B * tiles_x * tiles_y
    int box_rows;            // bh + max vertical shift
    int a_box_bytes;         // box_rows * bw * 128
    int n_phase;             // distinct y phases (boxes per (chunk, tx))
    int a_stages, b_stages;
    int y_org;               // added to y0*y_mul + y_off + phase for the box origin
    int bar_off;             // byte offset of the barrier block = max(ring bytes, epilogue staging bytes)
    int phase[MAX_TH];       // per ty
    int shift[MAX_TH];       // per ty, in patch rows
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// optional per-CTA timeline (profiling aid, tests/diag_conv_timeline.py): 8 x int64 per CTA
__device__ long long* g_timeline = nullptr;
__device__ int g_timeline_ctas = 0;
__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TL(slot, val)                                                                                         \
    do {                                                                                                      \
        if (tl != nullptr) tl[slot] = (val);                                                                  \
    } while (0)

template <int BN, int MT>
__global__ void __launch_bounds__(THREADS, 2) tc_conv_ytap_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const sdt_conv_desc d, const YGeom g) {
    constexpr int B_BYTES = BN * 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;   // 0: the dynamic window is 1024-byte aligned (checked below)
    uint8_t* sm = smem_raw + pad;
    const int a_stage_bytes = MT * g.a_box_bytes;
    {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if (pad + g.bar_off + TAIL_BYTES > dyn || g.a_stages * a_stage_bytes + g.b_stages * B_BYTES > g.bar_off) {
            if (threadIdx.x == 0) printf("tc_conv_ytap_kernel: shared-memory window misaligned (pad %u)\n", pad);
            __trap();
        }
    }
    const uint32_t smA = raw_addr + pad;
    const uint32_t smB = smA + g.a_stages * a_stage_bytes;
    const uint32_t bars = smA + g.bar_off;              // fullA[A_RING_MAX], emptyA[..], fullB[B_RING_MAX], emptyB[..], tmem_full
    uint8_t* after = sm + g.bar_off;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(after + (2 * A_RING_MAX + 2 * B_RING_MAX + 1) * 8);
    constexpr int STG_PITCH = BN + 4;                  // floats; the epilogue's staging tile and s_red alias the rings
    auto fullA = [&](int s) { return bars + 8u * s; };
    auto emptyA = [&](int s) { return bars + 8u * (A_RING_MAX + s); };
    auto fullB = [&](int s) { return bars + 8u * (2 * A_RING_MAX + s); };
    auto emptyB = [&](int s) { return bars + 8u * (2 * A_RING_MAX + B_RING_MAX + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * A_RING_MAX + 2 * B_RING_MAX);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.N;
    const int n0 = blockIdx.y * BN;
    const int t0 = blockIdx.x * MT;
    const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
    long long* tl = (g_timeline != nullptr && cta_lin < g_timeline_ctas) ? g_timeline + 8 * cta_lin : nullptr;
    if (tid == 0 && tl != nullptr) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tl[0] = smid;
        tl[1] = gtimer();
    }
    const int nvalid = min(MT, g.subtiles - t0);
    const int tpi = g.tiles_x * g.tiles_y;
    const int chunks = d.C / BKF;

    if (tid == 0) {
        for (int s = 0; s < A_RING_MAX; ++s) {
            mbar_init(fullA(s), 1);
            mbar_init(emptyA(s), 1);
        }
        for (int s = 0; s < B_RING_MAX; ++s) {
            mbar_init(fullB(s), 1);
            mbar_init(emptyB(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), MT * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int sb[MT], sy[MT], sx[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int t = min(t0 + m, g.subtiles - 1);
                sb[m] = t / tpi;
                const int rem = t - sb[m] * tpi;
                sy[m] = (rem / g.tiles_x) * g.bh * d.y_mul + d.y_off + g.y_org;
                sx[m] = (rem % g.tiles_x) * g.bw * d.x_mul + d.x_off;
            }
            int sa = 0, pa = 1, sbi = 0, pb = 1;      // ring slot + parity to wait on the empty barriers (first lap passes)
            long long wait_empty = 0;
            for (int ch = 0; ch < chunks; ++ch) {
                const int c0 = ch * BKF;
                for (int tx = 0; tx < d.TW; ++tx) {
                    for (int ph = 0; ph < g.n_phase; ++ph) {
                        long long tw = tl ? clock64() : 0;
                        mbar_wait(emptyA(sa), (uint32_t)pa);
                        if (tl) wait_empty += clock64() - tw;
                        mbar_expect_tx(fullA(sa), (uint32_t)(nvalid * g.a_box_bytes));
#pragma unroll
                        for (int m = 0; m < MT; ++m)
                            if (m < nvalid)
                                tma_load_4d(smA + sa * a_stage_bytes + m * g.a_box_bytes, &tmA, c0, sx[m] + tx * d.tx_mul, sy[m] + ph,
                                            sb[m], fullA(sa));
                        if (++sa == g.a_stages) { sa = 0; pa ^= 1; }
                        for (int ty = 0; ty < d.TH; ++ty) {
                            if (g.phase[ty] != ph) continue;
                            tw = tl ? clock64() : 0;
                            mbar_wait(emptyB(sbi), (uint32_t)pb);
                            if (tl) wait_empty += clock64() - tw;
                            mbar_expect_tx(fullB(sbi), B_BYTES);
                            tma_load_2d(smB + sbi * B_BYTES, &tmB, (ty * d.TW + tx) * d.C + c0, n0, fullB(sbi));
                            if (++sbi == g.b_stages) { sbi = 0; pb ^= 1; }
                        }
                    }
                }
            }
            TL(6, wait_empty);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BN, 0, 0);
            int sa = 0, pa = 0, sbi = 0, pb = 0;
            uint32_t started = 0;
            long long wait_full = 0;
            TL(2, gtimer());
            for (int ch = 0; ch < chunks; ++ch) {
                for (int tx = 0; tx < d.TW; ++tx) {
                    for (int ph = 0; ph < g.n_phase; ++ph) {
                        long long tw = tl ? clock64() : 0;
                        mbar_wait(fullA(sa), (uint32_t)pa);
                        if (tl) wait_full += clock64() - tw;
                        for (int ty = 0; ty < d.TH; ++ty) {
                            if (g.phase[ty] != ph) continue;
                            tw = tl ? clock64() : 0;
                            mbar_wait(fullB(sbi), (uint32_t)pb);
                            if (tl) wait_full += clock64() - tw;
                            if (tl && !started) tl[3] = gtimer();
                            tc_fence_after();
                            const uint64_t db = make_smem_desc(smB + sbi * B_BYTES, 16, 1024);
                            const uint32_t a_off = (uint32_t)(g.shift[ty] * g.bw * 128);
#pragma unroll
                            for (int m = 0; m < MT; ++m) {
                                if (m < nvalid) {
                                    const uint64_t da = make_smem_desc(smA + sa * a_stage_bytes + m * g.a_box_bytes + a_off, 16, 1024);
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4)
                                        mma_tf32(tmem_base + (uint32_t)(m * BN), da + 2u * k4, db + 2u * k4, idesc, started | (uint32_t)k4);
                                }
                            }
                            started = 1;
                            mma_commit(emptyB(sbi));
                            if (++sbi == g.b_stages) { sbi = 0; pb ^= 1; }
                        }
                        mma_commit(emptyA(sa));
                        if (++sa == g.a_stages) { sa = 0; pa ^= 1; }
                    }
                }
            }
            mma_commit(tmem_full_bar);
            TL(7, wait_full);
        }
        __syncwarp();
    } else {
        // ================= epilogue (warps 2..5; TMEM lane quadrant = warp % 4) =================
        // TMEM -> registers (thread = row) -> padded shared-memory tile -> full-row coalesced global stores.  A direct
        // store from the tcgen05.ld layout writes 16 bytes to each of 32 different lines per instruction and cost as much
        // as the whole main loop (tests/diag_conv_timeline.py).  Each warp stages and writes back only its own 32 rows.
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int py = r / g.bw, px = r - py * g.bw;
        float* stg = reinterpret_cast<float*>(sm) + (size_t)(q * 32) * STG_PITCH;      // this warp's 32 rows
        float* s_red = reinterpret_cast<float*>(sm) + (size_t)128 * STG_PITCH;           // [MT][2][4][BN]
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        if (tid == 64) TL(4, gtimer());
        for (int m = 0; m < nvalid; ++m) {
            const int t = t0 + m;
            const int b = t / tpi;
            const int rem = t - b * tpi;
            const int ty0 = (rem / g.tiles_x) * g.bh, tx0 = (rem % g.tiles_x) * g.bw;
            const bool ok = (ty0 + py) < d.GH && (tx0 + px) < d.GW;
            const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * BN + c * 32), v);
                float4* row = reinterpret_cast<float4*>(stg + (size_t)lane * STG_PITCH + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) row[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            __syncwarp();
            if (d.stat_partial != nullptr) {
                // column sums over this warp's valid rows (rows of the patch outside the output grid are masked out)
#pragma unroll
                for (int c = 0; c < BN / 32; ++c) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                    for (int i = 0; i < 32; ++i) {
                        const float x = stg[(size_t)i * STG_PITCH + c * 32 + lane];
                        if ((okmask >> i) & 1u) {
                            s1 += x;
                            s2 += x * x;
                        }
                    }
                    s_red[((m * 2 + 0) * 4 + q) * BN + c * 32 + lane] = s1;
                    s_red[((m * 2 + 1) * 4 + q) * BN + c * 32 + lane] = s2;
                }
            }
            // write back: ROWS_PER_STORE rows per instruction, each row BN*4 contiguous bytes
            constexpr int LANES_PER_ROW = BN / 4, ROWS_PER_STORE = 32 / LANES_PER_ROW;
            const int sub = lane / LANES_PER_ROW, col4 = lane % LANES_PER_ROW;
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d.bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(d.bias + n0) + col4);
#pragma unroll 4
            for (int i0 = 0; i0 < 32; i0 += ROWS_PER_STORE) {
                const int i = i0 + sub;
                if ((okmask >> i) & 1u) {
                    const int rr = q * 32 + i;
                    const int gy = ty0 + rr / g.bw, gx = tx0 + rr % g.bw;
                    float4* p = reinterpret_cast<float4*>(
                                    d.dst + (((long long)b * d.DH + (gy * d.dy_mul + d.dy_off)) * d.DW + (gx * d.dx_mul + d.dx_off)) * N + n0) + col4;
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
    }
    tc_fence_before();
    __syncthreads();
    if (d.stat_partial != nullptr) {
        const float* s_red = reinterpret_cast<const float*>(sm) + (size_t)128 * STG_PITCH;
        for (int i = tid; i < nvalid * BN; i += THREADS) {
            const int m = i / BN, c = i - m * BN;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                s1 += s_red[((m * 2 + 0) * 4 + q) * BN + c];
                s2 += s_red[((m * 2 + 1) * 4 + q) * BN + c];
            }
            d.stat_partial[((size_t)(t0 + m) * 2 + 0) * N + n0 + c] = s1;
            d.stat_partial[((size_t)(t0 + m) * 2 + 1) * N + n0 + c] = s2;
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, MT * BN);
    }
    if (tid == 0) TL(5, gtimer());
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
    int bn, mt;
    YGeom g;
    int smem;
    double cost;
};


// Vertical tap structure: which box (phase) and which patch-row shift each ty reads.
bool tap_structure(const sdt_conv_desc* d, YGeom* g, int* max_shift) {
    if (d->TH > MAX_TH) return false;
    if (d->ty_mul == 1) {                       // y = gy*y_mul + y_off + ty  ->  phase ty % y_mul, shift ty / y_mul
        g->n_phase = d->TH < d->y_mul ? d->TH : d->y_mul;
        g->y_org = 0;
        *max_shift = 0;
        for (int ty = 0; ty < d->TH; ++ty) {
            g->phase[ty] = ty % d->y_mul;
            g->shift[ty] = ty / d->y_mul;
            if (g->shift[ty] > *max_shift) *max_shift = g->shift[ty];
        }
        return true;
    }
    if (d->ty_mul == -1 && d->y_mul == 1) {     // data gradient: y = gy + y_off - ty
        g->n_phase = 1;
        g->y_org = -(d->TH - 1);
        *max_shift = d->TH - 1;
        for (int ty = 0; ty < d->TH; ++ty) {
            g->phase[ty] = 0;
            g->shift[ty] = d->TH - 1 - ty;
        }
        return true;
    }
    return false;
}

Plan make_plan(const sdt_conv_desc* d) {
    Plan best{};
    best.ok = false;
    if (d->N % 64 != 0 || d->C % 32 != 0) return best;
    const int K = d->TH * d->TW * d->C;
    for (int bn = 128; bn >= 64; bn /= 2) {
        if (d->N % bn != 0) continue;
        for (int mt = 4; mt >= 1; mt /= 2) {
            if (mt * bn > 256) continue;
            for (int bw = 8; bw <= 128; bw *= 2) {
                YGeom g{};
                int max_shift = 0;
                if (!tap_structure(d, &g, &max_shift)) return best;
                g.bw = bw;
                g.bh = 128 / bw;
                g.box_rows = g.bh + max_shift;
                if (bw * d->x_mul > 256 || g.box_rows * d->y_mul > 256) continue;
                g.a_box_bytes = g.box_rows * bw * 128;
                g.tiles_x = (d->GW + bw - 1) / bw;
                g.tiles_y = (d->GH + g.bh - 1) / g.bh;
                g.subtiles = d->B * g.tiles_x * g.tiles_y;
                const long long ctas = (long long)((g.subtiles + mt - 1) / mt) * (d->N / bn);
                const int budget = ctas > 148 ? SMEM_TWO_PER_SM : SMEM_ONE_PER_SM;
                const int a_stage = mt * g.a_box_bytes, b_stage = bn * 128;
                int as = 2, bs = 2;
                if (as * a_stage + bs * b_stage + TAIL_BYTES > budget) continue;
                for (bool grew = true; grew;) {      // spend what is left: weight boxes first (finer grained), then A stages
                    grew = false;
                    if (bs < B_RING_MAX && bs < 2 * as + 1 && as * a_stage + (bs + 1) * b_stage + TAIL_BYTES <= budget) { ++bs; grew = true; }
                    else if (as < A_RING_MAX && (as + 1) * a_stage + bs * b_stage + TAIL_BYTES <= budget) { ++as; grew = true; }
                }
                g.a_stages = as;
                g.b_stages = bs;
                // cost model (clocks per SM): L2 -> shared-memory bytes at ~42 B/clk/SM against tensor-pipe clocks
                const double a_bytes = (double)(d->C / BKF) * d->TW * g.n_phase * g.a_box_bytes * g.subtiles * (d->N / bn);
                const double b_bytes = (double)ctas * K * bn * 4.0;
                const double mma_clk = (double)g.subtiles * (d->N / bn) * (K / 8) * (bn / 2.0);
                const double waves = (double)((ctas + 295) / 296) * 296.0 / (double)ctas;   // quantisation on 148 x 2 slots
                const double mem_clk = (a_bytes + b_bytes) / 42.0;
                const double cost = ((mem_clk > mma_clk ? mem_clk : mma_clk) + 0.15 * (mem_clk + mma_clk)) * (ctas > 296 ? waves : 1.0);
                if (!best.ok || cost < best.cost) {
                    best.ok = true;
                    best.bn = bn;
                    best.mt = mt;
                    best.g = g;
                    const int ring = as * a_stage + bs * b_stage;
                    const int staging = 128 * (bn + 4) * 4 + mt * 8 * bn * 4;        // epilogue tile + statistics scratch
                    best.g.bar_off = ring > staging ? ring : staging;
                    best.smem = best.g.bar_off + TAIL_BYTES;
                    best.cost = cost;
                }
            }
        }
    }
    return best;
}

template <int BN, int MT>
int launch_ytap(const sdt_conv_desc* d, const Plan& pl, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    SDT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    static int attr_smem = 0;
    if (pl.smem > attr_smem) {
        SDT_CUDA_OK(cudaFuncSetAttribute(tc_conv_ytap_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ONE_PER_SM));
        attr_smem = SMEM_ONE_PER_SM;
    }
    const YGeom& g = pl.g;
    alignas(64) CUtensorMap tmA, tmB;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->SW, (cuuint64_t)d->SH, (cuuint64_t)d->B};
        const cuuint64_t strides[3] = {(cuuint64_t)d->C * 4, (cuuint64_t)d->SW * d->C * 4, (cuuint64_t)d->SH * d->SW * d->C * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)(g.bw * d->x_mul), (cuuint32_t)(g.box_rows * d->y_mul), 1};
        const cuuint32_t estr[4] = {1, (cuuint32_t)d->x_mul, (cuuint32_t)d->y_mul, 1};
        const CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->src), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A, y-tap) failed with %d", (int)r);
    }
    {
        const int K = d->TH * d->TW * d->C;
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)d->N};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)BN};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d->wt_nk), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B, y-tap) failed with %d", (int)r);
    }
    dim3 grid((g.subtiles + MT - 1) / MT, d->N / BN);
    tc_conv_ytap_kernel<BN, MT><<<grid, THREADS, pl.smem, st>>>(tmA, tmB, *d, g);
    SDT_LAUNCH_OK("tc_conv_ytap_kernel");
    sdt_note_tc_launch();
    return SDT_OK;
}

}  // namespace

bool sdt_tc_conv_ytap_eligible(const sdt_conv_desc* d) { return sdt_tc_conv_ytap_shape_ok(d) && get_encode() != nullptr; }

bool sdt_tc_conv_ytap_shape_ok(const sdt_conv_desc* d) {
    if (d->wt_nk == nullptr || d->xf_scale != nullptr) return false;        // plain (already activated) source only
    if (d->C % 32 != 0 || d->N % 64 != 0) return false;
    if (d->GH < 2 || d->TH < 2) return false;                                // nothing to reuse vertically: tc_conv_tma.cu
    if (d->x_mul < 1 || d->x_mul > 8 || d->y_mul < 1 || d->y_mul > 8) return false;
    if ((((uintptr_t)d->src | (uintptr_t)d->wt_nk | (uintptr_t)d->dst | (uintptr_t)d->bias) & 15) != 0) return false;
    return make_plan(d).ok;
}

int sdt_tc_conv_ytap_row_tiles(const sdt_conv_desc* d) { return make_plan(d).g.subtiles; }


// profiling aid (not part of the public header): per-CTA timeline buffer of 8 x int64 records, or NULL to switch off
extern "C" int sdt_debug_conv_timeline(void* buf, int ctas) {
    long long* p = static_cast<long long*>(buf);
    SDT_CUDA_OK(cudaMemcpyToSymbol(g_timeline, &p, sizeof(p)));
    SDT_CUDA_OK(cudaMemcpyToSymbol(g_timeline_ctas, &ctas, sizeof(ctas)));
    return SDT_OK;
}

int sdt_tc_conv_ytap_describe(const sdt_conv_desc* d, int32_t* out10) {
    const Plan pl = make_plan(d);
    if (!pl.ok) return 0;
    out10[1] = pl.bn; out10[2] = pl.mt; out10[3] = pl.g.bh; out10[4] = pl.g.bw; out10[5] = pl.g.box_rows;
    out10[6] = pl.g.a_stages; out10[7] = pl.g.b_stages; out10[8] = pl.smem;
    out10[9] = ((pl.g.subtiles + pl.mt - 1) / pl.mt) * (d->N / pl.bn);
    return 1;
}

int sdt_tc_conv_ytap_launch(const sdt_conv_desc* d, cudaStream_t st) {
    const Plan pl = make_plan(d);
    SDT_REQUIRE(pl.ok, "sdt_tc_conv_ytap_launch: no plan for this descriptor");
    static const bool debug = getenv("SDT_YTAP_DEBUG") != nullptr;
    if (debug)
        fprintf(stderr, "ytap: B%d C%d N%d G%dx%d T%dx%d ymul%d tymul%d -> BN%d MT%d patch %dx%d box_rows %d A%d B%d smem %d ctas %d\n", d->B, d->C,
                d->N, d->GH, d->GW, d->TH, d->TW, d->y_mul, d->ty_mul, pl.bn, pl.mt, pl.g.bh, pl.g.bw, pl.g.box_rows, pl.g.a_stages,
                pl.g.b_stages, pl.smem, ((pl.g.subtiles + pl.mt - 1) / pl.mt) * (d->N / pl.bn));
    if (pl.bn == 128) {
        if (pl.mt == 2) return launch_ytap<128, 2>(d, pl, st);
        return launch_ytap<128, 1>(d, pl, st);
    }
    if (pl.mt == 4) return launch_ytap<64, 4>(d, pl, st);
    if (pl.mt == 2) return launch_ytap<64, 2>(d, pl, st);
    return launch_ytap<64, 1>(d, pl, st);
}
