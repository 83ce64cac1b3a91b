// Convolution forward / data-gradient on tcgen05 with CTA PAIRS (cta_group::2), math mode 4 -- the next step after
// tc_conv_ytap.cu (mode 3), whose ncu captures show the tensor pipe 57-85 % "active" but at 40-50 % of the TF32 rate: with both
// operands in shared memory a 128x128x8 TF32 MMA reads 8 KB per 64 clk, the whole 128 B/clk port of ONE SM, and the TMA fill and
// the epilogue staging come on top.  A CTA pair executes one 256-row MMA: each CTA supplies its own 128 pixel rows (A) but only
// HALF of the weight tile (B), so per CTA an MMA step reads 4 KB + BN/2 * 32 B instead of 4 KB + BN * 32 B, and N = 256 tiles
// (128-clk MMAs, half the issue rate) fit the 512 TMEM columns double-buffered.
//
// Same structure as mode 3 otherwise (persistent, y-tap reuse, double-buffered accumulators, staged epilogue), per CTA of a pair:
//   warps 0-3: epilogue of this CTA's sub-tiles (its own TMEM lanes)       warp 4: TMA producer: its A boxes + its half of B,
//   warps 5,6: MMA issuers -- LEADER CTA (cluster rank 0) only; tcgen05.mma.cta_group::2, tcgen05.commit multicast to both CTAs
//              completing on the LEADER's full barriers (cp.async.bulk.tensor .cta_group::2, peer bit of the barrier address cleared)
// Barriers: full{A,B} live on the leader (one expect_tx covering both CTAs' bytes); empty{A,B} and tmem_full exist in both CTAs and
// receive the multicast commits; tmem_empty lives on the leader and counts the 8 epilogue warps of both CTAs (the peer's arrive
// through the cluster-shared window).
#include <cuda.h>
#include <stdlib.h>

#include "tc_api.h"
#include "tc_common.cuh"

namespace {

using namespace sdt_tc;

constexpr int BKF = 32;
constexpr int THREADS = 224;
constexpr int EPI_THREADS = 128;
constexpr int MAX_TH = 8;
constexpr int MAX_GROUPS = 4;
constexpr int B_RING_MAX = 6, A_RING_MAX = 4;
constexpr int SMEM_MAX = 227 * 1024;
constexpr int N_BARS = 2 * A_RING_MAX + 2 * B_RING_MAX + 4;
constexpr int TAIL_BYTES = N_BARS * 8 + 16;
constexpr int STG_PITCH = 36;
#define EPI_BYTES(BN_, MT_) (4 * 32 * STG_PITCH * 4 + (MT_) * 2 * 4 * (BN_) * 4)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the EVEN CTA of a pair

struct PGeom {
    int bw, bh, lbw;
    int tiles_x, tiles_y, subtiles;
    int box_rows, a_box_bytes;
    int a_stages, b_stages;
    int n_groups;
    unsigned long long groups;     // as YGeom::groups in tc_conv_ytap.cu
};
__host__ __device__ __forceinline__ int grp_y_add(unsigned long long p, int gi) { return (int)((p >> (16 * gi)) & 15u) - 8; }
__host__ __device__ __forceinline__ int grp_taps(unsigned long long p, int gi) { return (int)((p >> (16 * gi + 4)) & 15u); }
__host__ __device__ __forceinline__ int grp_ty0(unsigned long long p, int gi) { return (int)((p >> (16 * gi + 8)) & 15u); }
__host__ __device__ __forceinline__ int grp_step(unsigned long long p, int gi) { return (int)((p >> (16 * gi + 12)) & 15u) - 8; }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem_addr, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// TMA loads of a CTA pair: the transaction bytes complete on the LEADER's barrier
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const void* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar & PEER_BIT_MASK)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const void* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar & PEER_BIT_MASK)
        : "memory");
}
__device__ __forceinline__ void mma2_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of this thread's MMAs -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma2_commit_mc(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
// arrive on the barrier at this offset in the LEADER CTA (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_tf32_m256(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}
constexpr uint32_t DESC_HI = desc_hi(1024, kSwizzle128B);

template <int BN, int MT>
__global__ void __launch_bounds__(THREADS, 1) tc_conv_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const sdt_conv_desc d, const PGeom g) {
    constexpr int B_HALF_BYTES = (BN / 2) * 128;      // this CTA's half of the weight tile
    constexpr int ACC_COLS = MT * BN;
    constexpr int N_ISS = MT >= 2 ? 2 : 1;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smA = smem_u32(smem_raw);
    const int a_stage_bytes = MT * g.a_box_bytes;
    const int ring_bytes = g.a_stages * a_stage_bytes + g.b_stages * B_HALF_BYTES;
    {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if ((smA & 1023u) != 0 || ring_bytes + EPI_BYTES(BN, MT) + TAIL_BYTES > (int)dyn) {
            if (threadIdx.x == 0) printf("tc_conv_pair_kernel: shared-memory window misaligned or too small\n");
            __trap();
        }
    }
    const uint32_t smB = smA + g.a_stages * a_stage_bytes;
    float* stg_all = reinterpret_cast<float*>(smem_raw + ring_bytes);
    float* s_red = stg_all + 4 * 32 * STG_PITCH;
    const uint32_t bars = smA + ring_bytes + EPI_BYTES(BN, MT);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + ring_bytes + EPI_BYTES(BN, MT) + N_BARS * 8);
    auto fullA = [&](int s) { return bars + 8u * s; };
    auto emptyA = [&](int s) { return bars + 8u * (A_RING_MAX + s); };
    auto fullB = [&](int s) { return bars + 8u * (2 * A_RING_MAX + s); };
    auto emptyB = [&](int s) { return bars + 8u * (2 * A_RING_MAX + B_RING_MAX + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (2 * A_RING_MAX + 2 * B_RING_MAX + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (2 * A_RING_MAX + 2 * B_RING_MAX + 2 + a); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();           // 0 = leader
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int N = d.N;
    const int n_ntiles = N / BN;
    const int n_groups = (g.subtiles + MT - 1) / MT;    // groups of MT sub-tiles; a cluster tile = two consecutive groups x one N tile
    const int n_ctiles = ((n_groups + 1) / 2) * n_ntiles;
    const int tpi = g.tiles_x * g.tiles_y;
    const int chunks = d.C / BKF;

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
            mbar_init(tmem_empty(a), 8);            // the epilogue warps of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc2(smem_u32(tmem_slot), 2 * ACC_COLS);
    tc_fence_before();
    cluster_sync_all();                             // barriers of both CTAs initialised, TMEM allocated in both
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();

    // this CTA's group of a cluster tile
    auto my_group = [&](int ct) { return (ct / n_ntiles) * 2 + rank; };

    if (warp == 4) {
        // ================= TMA producer (both CTAs) =================
        int sa = 0, pa = 1, sbi = 0, pb = 1;
        for (int ct = pair; ct < n_ctiles; ct += n_pairs) {
            const int grp = my_group(ct);
            const int n0 = (ct % n_ntiles) * BN;
            const int t0 = grp * MT;
            const int nvalid = max(0, min(MT, g.subtiles - t0));
            const int nvalid_peer = max(0, min(MT, g.subtiles - (grp ^ 1) * MT));
            int sb[MT], sy[MT], sx[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int t = max(0, min(t0 + m, g.subtiles - 1));
                sb[m] = t / tpi;
                const int rem = t - sb[m] * tpi;
                sy[m] = (rem / g.tiles_x) * g.bh * d.y_mul + d.y_off;
                sx[m] = (rem % g.tiles_x) * g.bw * d.x_mul + d.x_off;
            }
            for (int ch = 0; ch < chunks; ++ch) {
                const int c0 = ch * BKF;
                for (int tx = 0; tx < d.TW; ++tx) {
                    for (int gi = 0; gi < g.n_groups; ++gi) {
                        const int y_add = grp_y_add(g.groups, gi), taps = grp_taps(g.groups, gi);
                        const int step = grp_step(g.groups, gi);
                        int ty = grp_ty0(g.groups, gi);
                        mbar_wait(emptyA(sa), (uint32_t)pa);
                        if (elect_one()) {
                            if (rank == 0) mbar_expect_tx(fullA(sa), (uint32_t)((nvalid + nvalid_peer) * g.a_box_bytes));
#pragma unroll
                            for (int m = 0; m < MT; ++m)
                                if (m < nvalid)
                                    tma2_load_4d(smA + sa * a_stage_bytes + m * g.a_box_bytes, &tmA, c0, sx[m] + tx * d.tx_mul, sy[m] + y_add,
                                                 sb[m], fullA(sa));
                        }
                        __syncwarp();
                        if (++sa == g.a_stages) { sa = 0; pa ^= 1; }
                        for (int j = 0; j < taps; ++j, ty += step) {
                            mbar_wait(emptyB(sbi), (uint32_t)pb);
                            if (elect_one()) {
                                if (rank == 0) mbar_expect_tx(fullB(sbi), 2 * B_HALF_BYTES);
                                tma2_load_2d(smB + sbi * B_HALF_BYTES, &tmB, (ty * d.TW + tx) * d.C + c0, n0 + rank * (BN / 2), fullB(sbi));
                            }
                            __syncwarp();
                            if (++sbi == g.b_stages) { sbi = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp >= 5) {
        // ================= MMA issuers: leader CTA only =================
        const int iss = warp - 5;
        if (rank == 0 && iss < N_ISS) {
            const uint32_t idesc = make_idesc_tf32_m256(BN);
            const uint32_t shift_lo = (uint32_t)(g.bw * 128) >> 4;
            const uint32_t box_lo = (uint32_t)g.a_box_bytes >> 4;
            int sa = 0, pa = 0, sbi = 0, pb = 0;
            int it = 0;
            for (int ct = pair; ct < n_ctiles; ct += n_pairs, ++it) {
                const int acc = it & 1;
                const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * ACC_COLS);
                mbar_wait(tmem_empty(acc), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                uint32_t started = 0;
                for (int ch = 0; ch < chunks; ++ch) {
                    for (int tx = 0; tx < d.TW; ++tx) {
                        for (int gi = 0; gi < g.n_groups; ++gi) {
                            const int taps = grp_taps(g.groups, gi);
                            mbar_wait(fullA(sa), (uint32_t)pa);
                            uint32_t a_lo = desc_lo(smA + sa * a_stage_bytes, 16);
                            for (int j = 0; j < taps; ++j, a_lo += shift_lo) {
                                mbar_wait(fullB(sbi), (uint32_t)pb);
                                tc_fence_after();
                                const uint32_t b_lo = desc_lo(smB + sbi * B_HALF_BYTES, 16);
                                if (elect_one()) {
#pragma unroll
                                    for (int mm = 0; mm < MT / N_ISS; ++mm) {
                                        const int m = iss + mm * N_ISS;
#pragma unroll
                                        for (int k4 = 0; k4 < 4; ++k4)
                                            mma2_tf32_lohi(tmem_acc + (uint32_t)(m * BN), a_lo + m * box_lo + 2u * k4, b_lo + 2u * k4, DESC_HI, idesc,
                                                           started | (uint32_t)k4);
                                    }
                                    mma2_commit_mc(emptyB(sbi));
                                }
                                __syncwarp();
                                started = 1;
                                if (++sbi == g.b_stages) { sbi = 0; pb ^= 1; }
                            }
                            if (elect_one()) mma2_commit_mc(emptyA(sa));
                            __syncwarp();
                            if (++sa == g.a_stages) { sa = 0; pa ^= 1; }
                        }
                    }
                }
                if (elect_one()) mma2_commit_mc(tmem_full(acc));
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue (both CTAs, own sub-tiles / own TMEM lanes) =================
        const int q = warp;
        const int r = q * 32 + lane;
        const int py = r >> g.lbw, px = r & (g.bw - 1);
        float* stg = stg_all + (size_t)(q * 32) * STG_PITCH;
        const int sub = lane >> 3, col4 = lane & 7;
        int it = 0;
        for (int ct = pair; ct < n_ctiles; ct += n_pairs, ++it) {
            const int grp = my_group(ct);
            const int n0 = (ct % n_ntiles) * BN;
            const int t0 = grp * MT;
            const int nvalid = max(0, min(MT, g.subtiles - t0));
            const int acc = it & 1;
            mbar_wait(tmem_full(acc), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            if (nvalid == 0) {                          // odd group count: the peer of the last pair has nothing to store
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(tmem_empty(acc));
            }
            for (int m = 0; m < nvalid; ++m) {
                const int t = t0 + m;
                const int b = t / tpi;
                const int rem = t - b * tpi;
                const int ty0 = (rem / g.tiles_x) * g.bh, tx0 = (rem % g.tiles_x) * g.bw;
                const bool ok = (ty0 + py) < d.GH && (tx0 + px) < d.GW;
                const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
                float* rowp[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int rr = q * 32 + 4 * k + sub;
                    const int gy = ty0 + (rr >> g.lbw), gx = tx0 + (rr & (g.bw - 1));
                    rowp[k] = d.dst + (((long long)b * d.DH + (gy * d.dy_mul + d.dy_off)) * d.DW + (gx * d.dx_mul + d.dx_off)) * N + n0 + col4 * 4;
                }
                for (int c = 0; c < BN / 32; ++c) {
                    float v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + m * BN + c * 32), v);
                    if (m == nvalid - 1 && c == BN / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(tmem_empty(acc));
                    }
                    float4* row = reinterpret_cast<float4*>(stg + (size_t)lane * STG_PITCH);
#pragma unroll
                    for (int j = 0; j < 8; ++j) row[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    if (d.stat_partial != nullptr) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                        for (int i = 0; i < 32; ++i) {
                            const float x = stg[(size_t)i * STG_PITCH + lane];
                            if ((okmask >> i) & 1u) {
                                s1 += x;
                                s2 += x * x;
                            }
                        }
                        s_red[((m * 2 + 0) * 4 + q) * BN + c * 32 + lane] = s1;
                        s_red[((m * 2 + 1) * 4 + q) * BN + c * 32 + lane] = s2;
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
            }
            if (d.stat_partial != nullptr) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = tid; i < nvalid * BN; i += EPI_THREADS) {
                    const int m = i / BN, c = i - m * BN;
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        s1 += s_red[((m * 2 + 0) * 4 + qq) * BN + c];
                        s2 += s_red[((m * 2 + 1) * 4 + qq) * BN + c];
                    }
                    d.stat_partial[((size_t)(t0 + m) * 2 + 0) * N + n0 + c] = s1;
                    d.stat_partial[((size_t)(t0 + m) * 2 + 1) * N + n0 + c] = s2;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                             // nobody of the pair still uses the other's shared memory / TMEM
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 2 * ACC_COLS);
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

struct PPlan {
    bool ok;
    int bn, mt, ctiles;
    PGeom g;
    int smem;
    double cost;
};

int tap_structure(const sdt_conv_desc* d, PGeom* g) {
    if (d->TH > MAX_TH) return 0;
    g->groups = 0;
    if (d->ty_mul == 1) {
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
    if (d->ty_mul == -1 && d->y_mul == 1) {
        g->n_groups = 1;
        g->groups = (unsigned long long)((-(d->TH - 1) + 8) | (d->TH << 4) | ((d->TH - 1) << 8) | ((-1 + 8) << 12));
        return d->TH;
    }
    return 0;
}

PPlan make_plan(const sdt_conv_desc* d) {
    PPlan best{};
    best.ok = false;
    if (d->N % 128 != 0 || d->C % 32 != 0) return best;
    const int K = d->TH * d->TW * d->C;
    for (int bn = 256; bn >= 128; bn /= 2) {
        if (d->N % bn != 0) continue;
        const int mt = 256 / bn;                       // 2 accumulator sets x MT x BN == 512 TMEM columns
        for (int bw = 8; bw <= 128; bw *= 2) {
            PGeom g{};
            const int max_taps = tap_structure(d, &g);
            if (max_taps == 0) return best;
            g.bw = bw;
            for (g.lbw = 0; (1 << g.lbw) < bw; ++g.lbw) {}
            g.bh = 128 / bw;
            g.box_rows = g.bh + max_taps - 1;
            if (bw * d->x_mul > 256 || g.box_rows * d->y_mul > 256) continue;
            g.a_box_bytes = g.box_rows * bw * 128;
            g.tiles_x = (d->GW + bw - 1) / bw;
            g.tiles_y = (d->GH + g.bh - 1) / g.bh;
            g.subtiles = d->B * g.tiles_x * g.tiles_y;
            const int groups = (g.subtiles + mt - 1) / mt;
            const long long ctiles = (long long)((groups + 1) / 2) * (d->N / bn);
            const int budget = SMEM_MAX - EPI_BYTES(bn, mt) - TAIL_BYTES;
            const int a_stage = mt * g.a_box_bytes, b_stage = (bn / 2) * 128;
            int as = 2, bs = 2;
            if (as * a_stage + bs * b_stage > budget) continue;
            for (bool grew = true; grew;) {
                grew = false;
                if (bs < B_RING_MAX && bs < 2 * as + 1 && as * a_stage + (bs + 1) * b_stage <= budget) { ++bs; grew = true; }
                else if (as < A_RING_MAX && (as + 1) * a_stage + bs * b_stage <= budget) { ++as; grew = true; }
            }
            g.a_stages = as;
            g.b_stages = bs;
            const double fill = (double)(d->C / BKF) * d->TW * g.n_groups * g.a_box_bytes * mt + (double)K * (bn / 2) * 4.0;
            const double reads = (double)mt * (K / 8) * (128 + bn / 2) * 32.0;
            const double mma_clk = (double)mt * (K / 8) * (bn / 2.0);
            const double smem_clk = (fill + reads) / 128.0;
            const double rounds = (double)((ctiles + 73) / 74);
            const double cost = (smem_clk > mma_clk ? smem_clk : mma_clk) * rounds;
            if (!best.ok || cost < best.cost) {
                best.ok = true;
                best.bn = bn;
                best.mt = mt;
                best.g = g;
                best.ctiles = (int)ctiles;
                best.smem = as * a_stage + bs * b_stage + EPI_BYTES(bn, mt) + TAIL_BYTES;
                best.cost = cost;
            }
        }
        if (best.ok) break;                            // prefer the widest N tile that fits
    }
    return best;
}

template <int BN, int MT>
int launch_pair(const sdt_conv_desc* d, const PPlan& pl, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    SDT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    static bool attr_set = false;
    if (!attr_set) {
        SDT_CUDA_OK(cudaFuncSetAttribute(tc_conv_pair_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        attr_set = true;
    }
    const PGeom& g = pl.g;
    alignas(64) CUtensorMap tmA, tmB;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->SW, (cuuint64_t)d->SH, (cuuint64_t)d->B};
        const cuuint64_t strides[3] = {(cuuint64_t)d->C * 4, (cuuint64_t)d->SW * d->C * 4, (cuuint64_t)d->SH * d->SW * d->C * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)(g.bw * d->x_mul), (cuuint32_t)(g.box_rows * d->y_mul), 1};
        const cuuint32_t estr[4] = {1, (cuuint32_t)d->x_mul, (cuuint32_t)d->y_mul, 1};
        const CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->src), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A, pair) failed with %d", (int)r);
    }
    {
        const int K = d->TH * d->TW * d->C;
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)d->N};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)(BN / 2)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d->wt_nk), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B, pair) failed with %d", (int)r);
    }
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        SDT_CUDA_OK(cudaGetDevice(&dev));
        SDT_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    int pairs = sm_count / 2;
    if (pl.ctiles < pairs) pairs = pl.ctiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = sdt::pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, tc_conv_pair_kernel<BN, MT>, tmA, tmB, *d, g);
    SDT_LAUNCH_OK("tc_conv_pair_kernel");
    sdt_note_tc_launch();
    return SDT_OK;
}

}  // namespace

bool sdt_tc_conv_pair_shape_ok(const sdt_conv_desc* d) {
    if (d->wt_nk == nullptr || d->xf_scale != nullptr) return false;
    if (d->C % 32 != 0 || d->N % 128 != 0) return false;
    if (d->GH < 2 || d->TH < 2) return false;
    if (d->x_mul < 1 || d->x_mul > 8 || d->y_mul < 1 || d->y_mul > 8) return false;
    if ((((uintptr_t)d->src | (uintptr_t)d->wt_nk | (uintptr_t)d->dst | (uintptr_t)d->bias) & 15) != 0) return false;
    return make_plan(d).ok;
}
bool sdt_tc_conv_pair_eligible(const sdt_conv_desc* d) { return sdt_tc_conv_pair_shape_ok(d) && get_encode() != nullptr; }
int sdt_tc_conv_pair_row_tiles(const sdt_conv_desc* d) { return make_plan(d).g.subtiles; }
int sdt_tc_conv_pair_describe(const sdt_conv_desc* d, int32_t* out10) {
    const PPlan pl = make_plan(d);
    if (!pl.ok) return 0;
    out10[1] = pl.bn; out10[2] = pl.mt; out10[3] = pl.g.bh; out10[4] = pl.g.bw; out10[5] = pl.g.box_rows;
    out10[6] = pl.g.a_stages; out10[7] = pl.g.b_stages; out10[8] = pl.smem;
    out10[9] = pl.ctiles;
    return 1;
}
int sdt_tc_conv_pair_launch(const sdt_conv_desc* d, cudaStream_t st) {
    const PPlan pl = make_plan(d);
    SDT_REQUIRE(pl.ok, "sdt_tc_conv_pair_launch: no plan for this descriptor");
    if (pl.bn == 256) return launch_pair<256, 1>(d, pl, st);
    return launch_pair<128, 2>(d, pl, st);
}
