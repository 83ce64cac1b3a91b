// Convolution forward / data-gradient as implicit GEMM on the 5th-generation tensor cores (tcgen05, kind::tf32,
// fp32 accumulation in TMEM) -- "math mode 1" of sdt_conv_gemm.  Same descriptor, same fusion contract as the FFMA
// kernel in conv_gemm.cu: the previous layer's normalisation + LeakyReLU is applied by the operand loader, the
// per-channel sum / sum-of-squares for the next normalisation come out of the epilogue.
//
// CTA = 128 output rows x BN output channels, K consumed in blocks of 32 floats (one 128-byte swizzle row).
//   warps 0..11 : three producer groups of 128 threads; group g builds k-blocks g, g+3, ...: thread r gathers the
//                 128 contiguous bytes of im2col row r (channels-last => one tap, 32 channels), applies scale/shift +
//                 LeakyReLU, and stores them into the K-major SWIZZLE_128B operand tile (16-byte chunk j of row r lands
//                 at chunk j ^ (r & 7)); weight rows ([N][K] K-major copy of the parameter) are copied the same way.
//                 fence.proxy.async + mbarrier.arrive hand the stage to the tensor core.
//   warp 12     : one elected thread issues 4 x tcgen05.mma (M128 x BN x K8) per stage and tcgen05.commit's the stage
//                 back to the producers; a final commit signals the epilogue.
//   epilogue    : the 12 producer warps read the accumulator with tcgen05.ld (warp w owns TMEM lanes 32*(w%4)..+31),
//                 add bias, store 128-byte row segments, and reduce the column statistics with a shuffle transpose.
#include <limits.h>

#include "tc_api.h"
#include "tc_common.cuh"

namespace {

using namespace sdt_tc;

constexpr int BM = 128;
constexpr int BKF = 32;                  // floats per k-block
constexpr int GROUPS = 3;
constexpr int PRODUCERS = GROUPS * 128;
constexpr int THREADS = PRODUCERS + 32;

struct __align__(16) RowInfo {
    long long base;     // element offset of src[b, sy0, sx0, 0] (may point outside the tensor; guarded by sy0/sx0)
    int sy0, sx0;       // sy0 = INT_MIN/2 marks a row outside the problem
};

template <int BN>
struct TcCfg {
    // 3 stages keep ~1 stage per producer group in flight and leave 60-150 KB of the SM's 228 KB to L1, which is what
    // absorbs the im2col tap re-reads
    static constexpr int STAGES = 3;
    static constexpr int A_BYTES = BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int BAR_BYTES = 256;                       // (2*STAGES+1) mbarriers + TMEM slot, keeps what follows 16-byte aligned
    static_assert((2 * STAGES + 1) * 8 + 16 <= BAR_BYTES, "barrier block");
    static constexpr int XF_BYTES = 2 * 256 * 4;
    static constexpr int RED_BYTES = 2 * 4 * BN * 4;
    static constexpr int ROW_BYTES = BM * 16 + BM * 8;           // {base, sy0, sx0} per row + dst offset per row
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + XF_BYTES + RED_BYTES + ROW_BYTES + 1024;
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) tc_conv_kernel(const sdt_conv_desc d) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    using Cfg = TcCfg<BN>;
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
    float* s_scale = reinterpret_cast<float*>(after + Cfg::BAR_BYTES);
    float* s_shift = s_scale + 256;
    float* s_red = s_shift + 256;                                 // [2][4][BN]
    RowInfo* s_rows = reinterpret_cast<RowInfo*>(s_red + 2 * 4 * BN);
    long long* s_dst = reinterpret_cast<long long*>(s_rows + BM);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = d.TH * d.TW * d.C, KB = K / BKF;
    const int N = d.N, P = d.GH * d.GW;
    const int n0 = blockIdx.y * BN;
    const bool has_xf = d.xf_scale != nullptr;

    // image of this tile (per-image tiling) -- only needed for the loader transform
    int tile_b = 0;
    if (d.per_image_tiles) tile_b = blockIdx.x / ((P + BM - 1) / BM);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == PRODUCERS / 32) tmem_alloc(smem_u32(tmem_slot), BN);
    if (has_xf) {
        for (int c = tid; c < d.C; c += THREADS) {
            s_scale[c] = d.xf_scale[(size_t)tile_b * d.xf_bstride + c];
            s_shift[c] = d.xf_shift[(size_t)tile_b * d.xf_bstride + c];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- decode the 128 rows of the tile once into shared memory
    if (tid < BM) {
        const int r = tid;
        int row_b, rem;
        bool ok;
        if (d.per_image_tiles) {
            const int tpi = (P + BM - 1) / BM;
            row_b = blockIdx.x / tpi;
            rem = (blockIdx.x % tpi) * BM + r;
            ok = rem < P;
        } else {
            const long long gm = (long long)blockIdx.x * BM + r;
            ok = gm < (long long)d.B * P;
            row_b = ok ? (int)(gm / P) : 0;
            rem = ok ? (int)(gm % P) : 0;
        }
        const int gy = rem / d.GW, gx = rem % d.GW;
        RowInfo ri;
        ri.sy0 = ok ? gy * d.y_mul + d.y_off : INT_MIN / 2;
        ri.sx0 = gx * d.x_mul + d.x_off;
        ri.base = (((long long)row_b * d.SH + ri.sy0) * d.SW + ri.sx0) * d.C;
        s_rows[r] = ri;
        s_dst[r] = ok ? (((long long)row_b * d.DH + (gy * d.dy_mul + d.dy_off)) * d.DW + (gx * d.dx_mul + d.dx_off)) * N : -1;
    }
    __syncthreads();
    const long long dst_off = tid < PRODUCERS ? s_dst[tid & 127] : -1;

    if (tid < PRODUCERS) {
        // ================= producers =================
        // piece map: 16-byte chunk `ch` (fixed per thread) of rows rsub, rsub+16, ...: the 8 lanes of a row read its
        // 128 bytes contiguously (4 cache lines per warp instruction) and write one swizzled 128-byte smem row.
        const int g = tid >> 7, t = tid & 127;
        const int ch = t & 7, rsub = t >> 3;
        const uint32_t sw_off = (uint32_t)((ch ^ (rsub & 7)) << 4);      // (r & 7) == (rsub & 7) for r = rsub + 16*i
        constexpr int B_PIECES = BN / 16;
        const int taps = d.TH * d.TW;
        for (int kb = g; kb < KB; kb += GROUPS) {
            const int s = kb % STAGES, round = kb / STAGES;
            // k-block order: channel chunk major, tap minor -- the taps of one 32-channel chunk re-read the same input
            // lines back to back, so they hit L1 instead of L2 (the GEMM K index itself stays (tap, channel))
            const int cchunk = kb / taps, tap = kb - cchunk * taps;
            const int c0 = cchunk * BKF;
            const int k = tap * d.C + c0;
            const int tyy = tap / d.TW, txx = tap - tyy * d.TW;
            const int dy_ = tyy * d.ty_mul, dx_ = txx * d.tx_mul;
            const long long delta = ((long long)dy_ * d.SW + dx_) * d.C + c0 + ch * 4;
            float4 a[8];
            unsigned vmask = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const RowInfo ri = s_rows[rsub + 16 * i];
                const int sy = ri.sy0 + dy_, sx = ri.sx0 + dx_;
                if (sy >= 0 && sy < d.SH && sx >= 0 && sx < d.SW) {
                    a[i] = __ldg(reinterpret_cast<const float4*>(d.src + ri.base + delta));
                    vmask |= 1u << i;
                } else {
                    a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float4 bw[B_PIECES];
            {
                const float* wp = d.wt_nk + (size_t)(n0 + rsub) * K + k + ch * 4;
#pragma unroll
                for (int i = 0; i < B_PIECES; ++i) bw[i] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(16 * i) * K));
            }
            if (has_xf) {
                const float4 c = *reinterpret_cast<const float4*>(s_scale + c0 + ch * 4);
                const float4 h = *reinterpret_cast<const float4*>(s_shift + c0 + ch * 4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (vmask & (1u << i)) {
                        a[i].x = sdt::leaky(fmaf(a[i].x, c.x, h.x), d.xf_slope);
                        a[i].y = sdt::leaky(fmaf(a[i].y, c.y, h.y), d.xf_slope);
                        a[i].z = sdt::leaky(fmaf(a[i].z, c.z, h.z), d.xf_slope);
                        a[i].w = sdt::leaky(fmaf(a[i].w, c.w, h.w), d.xf_slope);
                    }
                }
            }
            mbar_wait(empty_bar(s), (uint32_t)((round & 1) ^ 1));
            const uint32_t abase = smA + s * Cfg::A_BYTES + rsub * 128 + sw_off;
#pragma unroll
            for (int i = 0; i < 8; ++i) st_shared_v4(abase + i * 16 * 128, a[i]);
            const uint32_t bbase = smB + s * Cfg::B_BYTES + rsub * 128 + sw_off;
#pragma unroll
            for (int i = 0; i < B_PIECES; ++i) st_shared_v4(bbase + i * 16 * 128, bw[i]);
            fence_proxy_async_smem();
            mbar_arrive(full_bar(s));
        }
    } else {
        // ================= MMA issuer: the whole warp walks the loop, lane 0 issues =================
        const uint32_t idesc = make_idesc_tf32(BN, 0, 0);
        for (int kb = 0; kb < KB; ++kb) {
            if (lane == 0) {
                const int s = kb % STAGES, round = kb / STAGES;
                mbar_wait(full_bar(s), (uint32_t)(round & 1));
                tc_fence_after();
                const uint64_t da = make_smem_desc(smA + s * Cfg::A_BYTES, 16, 1024);
                const uint64_t db = make_smem_desc(smB + s * Cfg::B_BYTES, 16, 1024);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    // advance 32 bytes (8 tf32) along K inside the 128-byte swizzle row: +2 in the 16-byte address field
                    mma_tf32(tmem_base, da + 2u * k4, db + 2u * k4, idesc, (uint32_t)((kb | k4) != 0));
                }
                mma_commit(empty_bar(s));
            }
            __syncwarp();
        }
        if (lane == 0) mma_commit(tmem_full_bar);
        __syncwarp();
    }

    // ================= epilogue =================
    if (tid < PRODUCERS) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3, g = warp >> 2;
        for (int c = g; c < BN / 32; c += GROUPS) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            const int ncol = n0 + c * 32;
            if (dst_off >= 0) {
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
                // rows outside the tile range hold exact zeros (their operand rows were zero)
                float w[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) w[i] = v[i] * v[i];
                const float s1 = warp_transpose_sum(v, lane);
                const float s2 = warp_transpose_sum(w, lane);
                s_red[(0 * 4 + q) * BN + c * 32 + lane] = s1;
                s_red[(1 * 4 + q) * BN + c * 32 + lane] = s2;
            }
        }
    }
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
    if (warp == PRODUCERS / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

template <int BN>
int launch_tc(const sdt_conv_desc* d, int row_tiles, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        SDT_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM));
        attr_set = true;
    }
    dim3 grid(row_tiles, d->N / BN);
    sdt::launch(tc_conv_kernel<BN>, dim3(grid), dim3(THREADS), TcCfg<BN>::SMEM, st, *d);
    SDT_LAUNCH_OK("tc_conv_kernel");
    sdt_note_tc_launch();
    return SDT_OK;
}

}  // namespace

bool sdt_tc_conv_eligible(const sdt_conv_desc* d) {
    if (d->wt_nk == nullptr) return false;
    if (d->C % 32 != 0 || (d->C > 256 && d->xf_scale != nullptr)) return false;
    if (!(d->N == 64 || d->N == 128 || d->N == 256)) return false;
    // the loader transform's scale/shift are staged once per CTA: either one image per CTA (per-(b,c) statistics) or
    // statistics that do not depend on the image at all (BatchNorm, xf_bstride == 0)
    if (d->xf_scale != nullptr && !d->per_image_tiles && d->xf_bstride != 0) return false;
    if ((((uintptr_t)d->src | (uintptr_t)d->wt_nk | (uintptr_t)d->dst | (uintptr_t)d->bias) & 15) != 0) return false;
    return true;
}

int sdt_tc_conv_launch(const sdt_conv_desc* d, int row_tiles, cudaStream_t st) {
    // widest N tile that still gives every SM a CTA; small problems (the 1-D stacks) split N to fill the chip
    int bn = d->N;
    while (bn > 64 && (long long)row_tiles * (d->N / bn) < 148) bn /= 2;
    if (bn == 256) return launch_tc<256>(d, row_tiles, st);
    if (bn == 128) return launch_tc<128>(d, row_tiles, st);
    return launch_tc<64>(d, row_tiles, st);
}
