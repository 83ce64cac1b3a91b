// Convolution forward / data-gradient on tcgen05 with BOTH operands delivered by TMA (cp.async.bulk.tensor).
//
// Why: the register-staged producers of tc_conv.cu push every operand byte through the SM's 128 B/clk L1/shared pipe
// three times (L1 fill, st.shared, tensor-core fetch) and the ncu captures show that pipe, not DRAM / L2 / the tensor
// pipe, at 50-57 % (profiles/r1_ncu_tc_conv_v1_register_producers.txt).  TMA writes the SWIZZLE_128B operand tiles
// straight from L2 into shared memory: no LSU instructions, no registers, one pass.
//
// The A operand has no loader transform here, so the caller feeds an ACTIVATED channels-last tensor (the 2-D encoder
// materialises LeakyReLU(scale*raw+shift); the 1-D stacks and all data-gradients already are plain tensors).
//   A tile  : 128 GEMM rows = a (bh x bw) patch of the output grid of ONE image.  For tap (ty,tx) and 32-channel chunk c0 the
//             tile is the 4-D TMA box {32, bw*xs, bh*ys, 1} of the (C,W,H,B) tensor at coordinates
//             (c0, x0*x_mul + x_off + tx*tx_mul, y0*y_mul + y_off + ty*ty_mul, b) with element strides (1, x_mul, y_mul, 1):
//             strided convolutions and the stride-parity classes of the data gradient are plain boxes, the zero padding
//             is TMA's out-of-bounds fill (negative / too large coordinates).  Row r of the box = (r / bw, r % bw).
//   B tile  : 2-D box {32, BN} of the K-major (N, K) weight copy at (k, n0).
//   warps 0-3: epilogue: tcgen05.ld, bias / accumulate, 128-byte row stores, masked column statistics.
//   warp 4  : TMA producer: wait empty[s] -> arrive.expect_tx(full[s]) -> two bulk tensor copies.
//   warp 5  : MMA issuer: 4 x tcgen05.mma per stage, tcgen05.commit -> empty[s]; owns the TMEM allocation.
// 3 stages of 32 KB (BN = 128) -> two CTAs per SM.
#include <cuda.h>
#include <stdlib.h>

#include "tc_api.h"
#include "tc_common.cuh"

namespace {

using namespace sdt_tc;

constexpr int BM = 128;
constexpr int BKF = 32;
constexpr int THREADS = 192;

struct TileGeom {
    int bw, bh;            // patch width / height, bw * bh == 128
    int tiles_x, tiles_y;  // patches per image
};

// profiling aid (tests/diag_conv1d_timeline.py): 8 x int64 per CTA -- globaltimer (ns) at CTA start, after the prologue, when the first
// operand stage has landed, after the last MMA was issued, when the accumulator is complete, at the end of the epilogue, at CTA end
__device__ long long* g_tma_timeline = nullptr;
__device__ int g_tma_timeline_ctas = 0;
__device__ int g_tma_dbg = 0;            // with a timeline only (results are wrong): 1 = one MMA per k-block instead of four, 2 = no MMAs, 3 = no MMAs + plain arrive
                                         // instead of tcgen05.commit, 4 = as 3 and weight boxes only (one TMA operation per stage)
__device__ __forceinline__ long long gtimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// DEEP: grids of at most one CTA per SM (the 1-D stacks) are latency-bound on the k loop, not on occupancy: give the
// single resident CTA the whole shared memory as a 6-8 stage ring instead of leaving room for a second CTA.
constexpr int RN_CL = 4;               // CTAs per cluster of the fused row-norm epilogue: the four 64-column quarters of N = 256
template <int BN, bool DEEP>
struct TmaCfg {
    static constexpr int STAGES = DEEP ? (BN == 64 ? 8 : 6) : (BN == 64 ? 4 : 3);      // ring slots of one k-block each
    // k-blocks per barrier hand-over.  The k loop of the latency-bound (DEEP) launches is paced by the mbarrier hand-shake itself:
    // ~150 ns per iteration whatever the box sizes, TMA operations, ring depth, producers, MMAs or commits
    // (profiles/r2_ablation_fused_rownorm.txt).  Two k-blocks per full / empty barrier halve the iterations on both sides.
    static constexpr int KPS = DEEP ? 2 : 1;
    static constexpr int GROUPS = STAGES / KPS;
    static constexpr int A_BYTES = BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int BAR_BYTES = 256;
    static constexpr int RED_BYTES = 2 * 4 * BN * 4;              // statistics partials [2][4][BN]; the row-norm epilogue keeps its
    static constexpr int XCH_BYTES = 2 * RN_CL * BM * 4;          // exchange buffer [2 passes][rank][row] behind them (4 KB)
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + RED_BYTES + XCH_BYTES + 1024;
};

// ---- thread-block cluster helpers (fused row-norm epilogue) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {          // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t local_smem_addr, uint32_t rank, float v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_smem_addr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// RN: fused channel-LayerNorm + activation epilogue (sdt_conv_desc.rn_*).  The 256 output channels of a GEMM row are spread over the
// four CTAs of a cluster (blockIdx.y = 64-column quarter = cluster rank); each epilogue thread owns one row, keeps its 64 accumulators
// in registers, and the row sums travel through distributed shared memory: pass 1 the sums (-> mean), pass 2 the centred squares
// (-> rstd), each CTA writing its partial into every member's exchange buffer, a cluster barrier, a fixed-order sum (identical in all
// four CTAs).  Replaces the separate sdt_rownorm_act_fwd launch behind every 1-D convolution of the generator.
template <int BN, bool DEEP, bool RN>
__global__ void __launch_bounds__(THREADS, DEEP ? 1 : 2) tc_conv_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const sdt_conv_desc d, const TileGeom tg) {
    using Cfg = TmaCfg<BN, DEEP>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    uint8_t* sm = smem_raw + pad;
    const uint32_t smA = raw_addr + pad;
    const uint32_t smB = smA + STAGES * Cfg::A_BYTES;
    uint8_t* after = sm + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES);
    const uint32_t bars = smB + STAGES * Cfg::B_BYTES;          // full[STAGES], empty[STAGES], tmem_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(after + (2 * STAGES + 1) * 8);
    float* s_red = reinterpret_cast<float*>(after + Cfg::BAR_BYTES);   // [2][4][BN]
    float* s_xch = s_red + 2 * 4 * BN;                                  // [2][RN_CL][BM] (RN only)
    float rnv[RN ? 64 : 1];                                             // RN: this thread's row of the tile (epilogue warps)
    long long rn_dst_off = -1;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = d.TH * d.TW * d.C, KB = K / BKF;
    const int N = d.N;
    const int n0 = blockIdx.y * BN;
    const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
    long long* tl = (g_tma_timeline != nullptr && cta_lin < g_tma_timeline_ctas) ? g_tma_timeline + 8 * cta_lin : nullptr;
    if (tl && tid == 0) tl[0] = gtimer_ns();
    // patch of this CTA
    const int tpi = tg.tiles_x * tg.tiles_y;
    const int b = blockIdx.x / tpi;
    const int trem = blockIdx.x - b * tpi;
    const int y0 = (trem / tg.tiles_x) * tg.bh, x0 = (trem % tg.tiles_x) * tg.bw;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation) overlapped the tail of the preceding kernel; global memory from here on
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    if (tl && tid == 0) tl[1] = gtimer_ns();

    // warps 0-3: epilogue (TMEM lane quadrant = warp id); warp 4: TMA producer; warp 5: MMA issuer.  The issuing warps walk
    // their loops as whole warps and elect one lane per instruction (tc_common.cuh: elect_one), and have the highest warp ids.
    if (warp == 4) {
        // ================= TMA producer =================
        const int pdbg = tl ? g_tma_dbg : 0;
        constexpr int KPS = Cfg::KPS, GROUPS = Cfg::GROUPS;
        int g = 0, par = 1;
        int txx = 0, tyy = 0, c0 = 0;               // (tap, channel chunk) of the next k-block; order: chunk major, tap minor
        for (int kb = 0; kb < KB; kb += KPS) {
            const int nkb = KB - kb < KPS ? KB - kb : KPS;
            mbar_wait_spin(empty_bar(g), (uint32_t)par);
            const bool leader = elect_one();
            if (leader) mbar_expect_tx(full_bar(g), (uint32_t)(nkb * (pdbg == 4 ? Cfg::B_BYTES : Cfg::A_BYTES + Cfg::B_BYTES)));
#pragma unroll
            for (int j = 0; j < KPS; ++j) {
                if (j < nkb) {
                    const int s = g * KPS + j;
                    const int wy = y0 * d.y_mul + d.y_off + tyy * d.ty_mul;
                    const int wx = x0 * d.x_mul + d.x_off + txx * d.tx_mul;
                    if (leader) {
                        if (pdbg != 4) tma_load_4d(smA + s * Cfg::A_BYTES, &tmA, c0, wx, wy, b, full_bar(g));
                        tma_load_2d(smB + s * Cfg::B_BYTES, &tmB, (tyy * d.TW + txx) * d.C + c0, n0, full_bar(g));
                    }
                    if (++txx == d.TW) {
                        txx = 0;
                        if (++tyy == d.TH) { tyy = 0; c0 += BKF; }
                    }
                }
            }
            __syncwarp();
            if (++g == GROUPS) { g = 0; par ^= 1; }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_tf32(BN, 0, 0);
        constexpr uint32_t HI = desc_hi(1024, kSwizzle128B);
        constexpr int KPS = Cfg::KPS, GROUPS = Cfg::GROUPS;
        int g = 0, par = 0;
        uint32_t started = 0;
        const int dbg = tl ? g_tma_dbg : 0;
        for (int kb = 0; kb < KB; kb += KPS) {
            const int nkb = KB - kb < KPS ? KB - kb : KPS;
            mbar_wait_spin(full_bar(g), (uint32_t)par);
            if (tl && kb == 0 && lane == 0) tl[2] = gtimer_ns();
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < KPS; ++j) {
                    if (j < nkb) {
                        const int s = g * KPS + j;
                        const uint32_t a_lo = desc_lo(smA + s * Cfg::A_BYTES, 16), b_lo = desc_lo(smB + s * Cfg::B_BYTES, 16);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            if (dbg == 0 || (dbg == 1 && k4 == 0))
                                mma_tf32_lohi(tmem_base, a_lo + 2u * k4, b_lo + 2u * k4, HI, idesc, (started | (uint32_t)(j | k4)) ? 1u : 0u);
                    }
                }
                if (dbg >= 3) mbar_arrive(empty_bar(g));          // no MMAs and the stage handed back WITHOUT tcgen05.commit
                else mma_commit(empty_bar(g));
            }
            __syncwarp();
            started = 1;
            if (++g == GROUPS) { g = 0; par ^= 1; }
        }
        if (elect_one()) mma_commit(tmem_full_bar);
        __syncwarp();
        if (tl && lane == 0) tl[3] = gtimer_ns();
    } else {
        // ================= epilogue (warps 2..5; TMEM lane quadrant = warp % 4) =================
        const int q = warp;
        const int r = q * 32 + lane;
        const int py = r / tg.bw, px = r - py * tg.bw;
        const int gy = y0 + py, gx = x0 + px;
        const bool ok = gy < d.GH && gx < d.GW;
        const long long dst_off = ok ? (((long long)b * d.DH + (gy * d.dy_mul + d.dy_off)) * d.DW + (gx * d.dx_mul + d.dx_off)) * N : -1;
        mbar_wait(tmem_full_bar, 0);
        if (tl && tid == 0) tl[4] = gtimer_ns();
        tc_fence_after();
        if constexpr (RN) {
            // pass 1: the row's 64 accumulators into registers, their sum to every CTA of the cluster
            static_assert(!RN || BN == 64, "row-norm epilogue: 64-column quarters");
            float v0[32], v1[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16), v0);
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + 32u, v1);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                rnv[i] = v0[i];
                rnv[32 + i] = v1[i];
            }
#pragma unroll
            for (int i = 0; i < 64; ++i) sum += rnv[i];
            rn_dst_off = dst_off;
            const uint32_t rank = cluster_ctarank();
            const uint32_t slot = smem_u32(s_xch + (0 * RN_CL + rank) * BM + r);
#pragma unroll
            for (uint32_t t = 0; t < RN_CL; ++t) st_cluster_f32(slot, t, sum);
        } else
        for (int c = 0; c < BN / 32; ++c) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            const int ncol = n0 + c * 32;
            if (ok) {
                float* p = d.dst + dst_off + ncol;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    if (d.bias != nullptr) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(d.bias + ncol) + j);
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    }
                    if (d.accumulate) {
                        const float4 old = reinterpret_cast<const float4*>(p)[j];
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    reinterpret_cast<float4*>(p)[j] = o;
                }
            }
            if (d.stat_partial != nullptr) {
                // rows of the patch that fall outside the output grid computed on real neighbours: mask them out
                float w[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v[i] = ok ? v[i] : 0.f;
                    w[i] = v[i] * v[i];
                }
                const float s1 = warp_transpose_sum(v, lane);
                const float s2 = warp_transpose_sum(w, lane);
                s_red[(0 * 4 + q) * BN + c * 32 + lane] = s1;
                s_red[(1 * 4 + q) * BN + c * 32 + lane] = s2;
            }
        }
    }
    if constexpr (RN) {
        cluster_sync_all();
        const int r = (warp & 3) * 32 + lane;
        float mu = 0.f;
        if (warp < 4) {
            mu = ((s_xch[0 * BM + r] + s_xch[1 * BM + r]) + (s_xch[2 * BM + r] + s_xch[3 * BM + r])) * (1.0f / (float)(RN_CL * BN));
            float qs = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const float dl = rnv[i] - mu;
                qs = fmaf(dl, dl, qs);
            }
            const uint32_t rank = cluster_ctarank();
            const uint32_t slot = smem_u32(s_xch + (1 * RN_CL + rank) * BM + r);
#pragma unroll
            for (uint32_t t = 0; t < RN_CL; ++t) st_cluster_f32(slot, t, qs);
        }
        cluster_sync_all();
        if (warp < 4 && rn_dst_off >= 0) {
            const float* x2 = s_xch + RN_CL * BM;
            const float var = ((x2[0 * BM + r] + x2[1 * BM + r]) + (x2[2 * BM + r] + x2[3 * BM + r])) * (1.0f / (float)(RN_CL * BN));   // biased
            const float rs = 1.0f / sqrtf(var + d.rn_eps);
            float4* praw = reinterpret_cast<float4*>(d.dst + rn_dst_off + n0);
            float4* pact = reinterpret_cast<float4*>(d.rn_act + rn_dst_off + n0);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                praw[j] = make_float4(rnv[4 * j], rnv[4 * j + 1], rnv[4 * j + 2], rnv[4 * j + 3]);
                float4 o;
                o.x = sdt::out_round(sdt::leaky((rnv[4 * j] - mu) * rs, d.rn_slope), d.rn_out_tf32);
                o.y = sdt::out_round(sdt::leaky((rnv[4 * j + 1] - mu) * rs, d.rn_slope), d.rn_out_tf32);
                o.z = sdt::out_round(sdt::leaky((rnv[4 * j + 2] - mu) * rs, d.rn_slope), d.rn_out_tf32);
                o.w = sdt::out_round(sdt::leaky((rnv[4 * j + 3] - mu) * rs, d.rn_slope), d.rn_out_tf32);
                pact[j] = o;
            }
            if (n0 == 0) {
                d.rn_mean[rn_dst_off / N] = mu;
                d.rn_rstd[rn_dst_off / N] = rs;
            }
        }
    }
    if (tl && tid == 0) tl[5] = gtimer_ns();
    tc_fence_before();
    __syncthreads();
    if (d.stat_partial != nullptr) {
        for (int c = tid; c < BN; c += THREADS) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                s1 += s_red[(0 * 4 + q) * BN + c];
                s2 += s_red[(1 * 4 + q) * BN + c];
            }
            d.stat_partial[((size_t)blockIdx.x * 2 + 0) * N + n0 + c] = s1;
            d.stat_partial[((size_t)blockIdx.x * 2 + 1) * N + n0 + c] = s2;
        }
    }
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
        if (tl && lane == 0) tl[6] = gtimer_ns();
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

TileGeom pick_geom(const sdt_conv_desc* d) {
    TileGeom best{128, 1, 0, 0};
    long long best_rows = -1;
    for (int bw = 128; bw >= 8; bw /= 2) {
        const int bh = 128 / bw;
        if (bw * d->x_mul > 256 || bh * d->y_mul > 256) continue;          // TMA box extent limit
        const int tx = (d->GW + bw - 1) / bw, ty = (d->GH + bh - 1) / bh;
        const long long rows = (long long)tx * ty * 128;
        if (best_rows < 0 || rows < best_rows) {
            best_rows = rows;
            best = TileGeom{bw, bh, tx, ty};
        }
    }
    return best;
}

template <int BN, bool DEEP, bool RN = false>
int launch_tma(const sdt_conv_desc* d, const TileGeom& tg, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    SDT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    static bool attr_set = false;
    if (!attr_set) {
        SDT_CUDA_OK(cudaFuncSetAttribute(tc_conv_tma_kernel<BN, DEEP, RN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TmaCfg<BN, DEEP>::SMEM));
        attr_set = true;
    }
    alignas(64) CUtensorMap tmA, tmB;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->SW, (cuuint64_t)d->SH, (cuuint64_t)d->B};
        const cuuint64_t strides[3] = {(cuuint64_t)d->C * 4, (cuuint64_t)d->SW * d->C * 4, (cuuint64_t)d->SH * d->SW * d->C * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)(tg.bw * d->x_mul), (cuuint32_t)(tg.bh * d->y_mul), 1};
        const cuuint32_t estr[4] = {1, (cuuint32_t)d->x_mul, (cuuint32_t)d->y_mul, 1};
        const CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->src), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed with %d", (int)r);
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
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }
    dim3 grid(d->B * tg.tiles_x * tg.tiles_y, d->N / BN);
    if (RN)
        sdt::launch_cluster(tc_conv_tma_kernel<BN, DEEP, RN>, dim3(grid), dim3(THREADS), TmaCfg<BN, DEEP>::SMEM, st, 1, RN_CL, 1, tmA, tmB, *d, tg);
    else
        sdt::launch(tc_conv_tma_kernel<BN, DEEP, RN>, dim3(grid), dim3(THREADS), TmaCfg<BN, DEEP>::SMEM, st, tmA, tmB, *d, tg);
    SDT_LAUNCH_OK("tc_conv_tma_kernel");
    sdt_note_tc_launch();
    return SDT_OK;
}

}  // namespace

// per-CTA timeline buffer of 8 x int64 records, or NULL to switch off (diagnostics only)
extern "C" int sdt_debug_tma_flags(int flags) {
    SDT_CUDA_OK(cudaMemcpyToSymbol(g_tma_dbg, &flags, sizeof(flags)));
    return SDT_OK;
}
extern "C" int sdt_debug_tma_timeline(void* buf, int ctas) {
    long long* p = static_cast<long long*>(buf);
    SDT_CUDA_OK(cudaMemcpyToSymbol(g_tma_timeline, &p, sizeof(p)));
    SDT_CUDA_OK(cudaMemcpyToSymbol(g_tma_timeline_ctas, &ctas, sizeof(ctas)));
    return SDT_OK;
}

bool sdt_tc_conv_tma_eligible(const sdt_conv_desc* d) {
    if (d->wt_nk == nullptr || d->xf_scale != nullptr) return false;        // plain (already activated) source only
    if (d->C % 32 != 0 || d->N % 64 != 0) return false;
    if (d->x_mul < 1 || d->x_mul > 8 || d->y_mul < 1 || d->y_mul > 8) return false;
    if ((((uintptr_t)d->src | (uintptr_t)d->wt_nk | (uintptr_t)d->dst | (uintptr_t)d->bias) & 15) != 0) return false;
    return get_encode() != nullptr;
}

int sdt_tc_conv_tma_row_tiles(const sdt_conv_desc* d) {
    const TileGeom tg = pick_geom(d);
    return d->B * tg.tiles_x * tg.tiles_y;
}

bool sdt_tc_conv_tma_rownorm_ok(const sdt_conv_desc* d) {
    if (!sdt_tc_conv_tma_eligible(d)) return false;
    if (d->N != RN_CL * 64 || d->bias != nullptr || d->accumulate || d->stat_partial != nullptr) return false;
    if (d->dy_mul != 1 || d->dx_mul != 1 || d->dy_off != 0 || d->dx_off != 0 || d->DH != d->GH || d->DW != d->GW) return false;   // plain forward
    return true;
}

int sdt_tc_conv_tma_launch(const sdt_conv_desc* d, cudaStream_t st) {
    const TileGeom tg = pick_geom(d);
    const long long tiles = (long long)d->B * tg.tiles_x * tg.tiles_y;
    if (d->rn_act != nullptr) {
        SDT_REQUIRE(sdt_tc_conv_tma_rownorm_ok(d) && d->rn_mean && d->rn_rstd, "sdt_conv_gemm: the fused row-norm epilogue is not available for this problem");
        SDT_REQUIRE(((((uintptr_t)d->rn_act) & 15) == 0), "sdt_conv_gemm: rn_act must be 16-byte aligned");
        return tiles * RN_CL <= 148 ? launch_tma<64, true, true>(d, tg, st) : launch_tma<64, false, true>(d, tg, st);
    }
    int bn = d->N % 128 == 0 ? 128 : 64;
    if (bn == 128 && tiles * (d->N / 128) < 2 * 148) bn = 64;               // small problems: more CTAs
    if (getenv("SDT_TMA_BN128") != nullptr && d->N % 128 == 0) bn = 128;    // tuning aid (tests/diag_conv1d_timeline.py)
    const bool deep = tiles * (d->N / bn) <= 148;
    if (bn == 128) return deep ? launch_tma<128, true>(d, tg, st) : launch_tma<128, false>(d, tg, st);
    return deep ? launch_tma<64, true>(d, tg, st) : launch_tma<64, false>(d, tg, st);
}
