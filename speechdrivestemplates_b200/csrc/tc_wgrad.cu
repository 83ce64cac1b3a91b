// Convolution weight-gradient on the tcgen05 tensor cores (kind::tf32, fp32 accumulation in TMEM) -- "math mode 1"
// of sdt_conv_wgrad.
//
//   D[m = k-index (tap, ci)][n = output channel] = sum over pixels  A(pix, k-index) * dy[pix, n]
//
// Both operands are contiguous along their M/N dimension in HBM (channels-last), i.e. MN-major: a stage holds 32
// pixels; each pixel contributes 128-byte segments (32 channels) that are copied -- A through the previous layer's
// scale/shift + LeakyReLU -- into MN-major SWIZZLE_128B_BASE32B atoms (the only MN-major swizzle tf32 supports:
// 4 pixels x 128 B; 32-byte unit u of pixel p lands at unit u ^ (p & 3)).  One tcgen05.mma (K = 8) consumes two
// atoms along K.  CTA tile = 128 k-indices x Cout,
// split-K over pixel ranges (gridDim.z), partials reduced in fixed order by sdt_conv_wgrad_reduce.
#include "tc_api.h"
#include "tc_common.cuh"

namespace {

using namespace sdt_tc;

constexpr int BM = 128;                  // k-indices per CTA (GEMM M)
constexpr int PIX = 32;                  // pixels per stage
constexpr int GROUPS = 3;
constexpr int PRODUCERS = GROUPS * 128;
constexpr int THREADS = PRODUCERS + 32;

template <int BN>
struct WgCfg {
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int A_BYTES = PIX * BM * 4;        // 16 KB: 4 k-groups x 4 MN atoms x 1 KB
    static constexpr int B_BYTES = PIX * BN * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + 1024;
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) tc_wgrad_kernel(const sdt_conv_desc d) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    using Cfg = WgCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NCH = BN / 32;                         // 32-channel chunks of dy per pixel
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    uint8_t* sm = smem_raw + pad;
    const uint32_t smA = raw_addr + pad;
    const uint32_t smB = smA + STAGES * Cfg::A_BYTES;
    uint8_t* after = sm + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES);
    const uint32_t bars = smB + STAGES * Cfg::B_BYTES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(after + (2 * STAGES + 1) * 8);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Kc = d.TH * d.TW * d.C;
    const int P = d.GH * d.GW;
    const long long Mtot = (long long)d.B * P;
    const int kidx0 = blockIdx.x * BM;
    long long chunk = (Mtot + d.splits - 1) / d.splits;
    chunk = (chunk + PIX - 1) / PIX * PIX;
    const long long p_begin = (long long)blockIdx.z * chunk;
    const long long p_end = p_begin + chunk < Mtot ? p_begin + chunk : Mtot;
    const int KB = p_end > p_begin ? (int)((p_end - p_begin + PIX - 1) / PIX) : 0;
    float* out = d.wpart + (size_t)blockIdx.z * BN * Kc;

    if (KB == 0) {   // empty split: the partial is all zeros (uniform over the CTA, no tensor-core work)
        for (int e = tid; e < BN * BM; e += THREADS) {
            const int n = e / BM, m = e % BM;
            if (kidx0 + m < Kc) out[(size_t)n * Kc + kidx0 + m] = 0.f;
        }
        return;
    }

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == PRODUCERS / 32) tmem_alloc(smem_u32(tmem_slot), BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool has_xf = d.xf_scale != nullptr;

    if (tid < PRODUCERS) {
        // ================= producers =================
        // piece map: 16-byte piece `w` (fixed per thread) of (pixel, 32-channel chunk) pairs pr, pr+16, ...: the 8 lanes
        // of a pair read its 128 bytes contiguously (4 cache lines per warp instruction).
        const int g = tid >> 7, t = tid & 127;
        const int w = t & 7, pr0 = t >> 3;                  // pair index = pr0 + 16*i
        // A pairs: (pixel kk = pair >> 2, chunk ja = pair & 3); ja = pr0 & 3 is loop invariant
        const int ja = pr0 & 3;
        const int kidx = kidx0 + ja * 32;
        const bool chunk_ok = kidx < Kc;
        int c0 = 0, tyy = 0, txx = 0;
        if (chunk_ok) {
            const int tap = kidx / d.C;
            c0 = kidx - tap * d.C;
            tyy = tap / d.TW;
            txx = tap - tyy * d.TW;
        }
        const int dy_ = tyy * d.ty_mul, dx_ = txx * d.tx_mul;
        constexpr int B_PIECES = 2 * NCH;                   // 32 pixels x NCH chunks x 8 pieces / 128 threads
        for (int kb = g; kb < KB; kb += GROUPS) {
            const int s = kb % STAGES, round = kb / STAGES;
            const long long pbase = p_begin + (long long)kb * PIX;
            // ---- A: im2col segments through the loader transform
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = (pr0 + 16 * i) >> 2;
                const long long pix = pbase + kk;
                a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (chunk_ok && pix < p_end) {
                    const int b = (int)(pix / P);
                    const int rem = (int)(pix - (long long)b * P);
                    const int gy = rem / d.GW, gx = rem - gy * d.GW;
                    const int sy = gy * d.y_mul + d.y_off + dy_;
                    const int sx = gx * d.x_mul + d.x_off + dx_;
                    if (sy >= 0 && sy < d.SH && sx >= 0 && sx < d.SW) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(d.src + (((size_t)b * d.SH + sy) * d.SW + sx) * d.C + c0 + w * 4));
                        if (has_xf) {
                            const size_t o = (size_t)b * d.xf_bstride + c0 + w * 4;
                            const float4 c = __ldg(reinterpret_cast<const float4*>(d.xf_scale + o));
                            const float4 h = __ldg(reinterpret_cast<const float4*>(d.xf_shift + o));
                            v.x = sdt::leaky(fmaf(v.x, c.x, h.x), d.xf_slope);
                            v.y = sdt::leaky(fmaf(v.y, c.y, h.y), d.xf_slope);
                            v.z = sdt::leaky(fmaf(v.z, c.z, h.z), d.xf_slope);
                            v.w = sdt::leaky(fmaf(v.w, c.w, h.w), d.xf_slope);
                        }
                        a[i] = v;
                    }
                }
            }
            // ---- B: dy rows, pairs (pixel = pair / NCH, chunk jb = pair % NCH)
            float4 bw[B_PIECES];
#pragma unroll
            for (int i = 0; i < B_PIECES; ++i) {
                const int pair = pr0 + 16 * i;
                const int kp = pair / NCH, jb = pair % NCH;
                const long long pb = pbase + kp;
                bw[i] = pb < p_end ? __ldg(reinterpret_cast<const float4*>(d.dy + pb * BN + jb * 32 + w * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(empty_bar(s), (uint32_t)((round & 1) ^ 1));
            const uint32_t abase = smA + s * Cfg::A_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = (pr0 + 16 * i) >> 2;
                const uint32_t off = (uint32_t)(((kk >> 2) * 4 + ja) * 512 + (kk & 3) * 128 + ((((w >> 1) ^ (kk & 3)) << 5) | ((w & 1) << 4)));
                st_shared_v4(abase + off, a[i]);
            }
            const uint32_t bbase = smB + s * Cfg::B_BYTES;
#pragma unroll
            for (int i = 0; i < B_PIECES; ++i) {
                const int pair = pr0 + 16 * i;
                const int kp = pair / NCH, jb = pair % NCH;
                const uint32_t off = (uint32_t)(((kp >> 2) * NCH + jb) * 512 + (kp & 3) * 128 + ((((w >> 1) ^ (kp & 3)) << 5) | ((w & 1) << 4)));
                st_shared_v4(bbase + off, bw[i]);
            }
            fence_proxy_async_smem();
            mbar_arrive(full_bar(s));
        }
    } else {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_tf32(BN, 1, 1);          // both operands MN-major
        for (int kb = 0; kb < KB; ++kb) {
            if (lane == 0) {
                const int s = kb % STAGES, round = kb / STAGES;
                mbar_wait(full_bar(s), (uint32_t)(round & 1));
                tc_fence_after();
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    // 8 pixels = two K-atoms of 4 pixels; atoms are 512 B, MN atoms adjacent (LBO 512), K atoms a row of
                    // MN atoms apart (SBO = 4*512 for A, NCH*512 for B)
                    const uint64_t da = make_smem_desc(smA + s * Cfg::A_BYTES + k4 * 4096, 512, 2048, kSwizzle128B_Base32B);
                    const uint64_t db = make_smem_desc(smB + s * Cfg::B_BYTES + k4 * NCH * 1024, 512, NCH * 512, kSwizzle128B_Base32B);
                    mma_tf32(tmem_base, da, db, idesc, (uint32_t)((kb | k4) != 0));
                }
                mma_commit(empty_bar(s));
            }
            __syncwarp();
        }
        if (lane == 0) mma_commit(tmem_full_bar);
        __syncwarp();
    }

    // ================= epilogue: D[m][n] -> wpart[z][n][kidx0 + m] (lanes = consecutive m: coalesced) =================
    if (tid < PRODUCERS) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3, g = warp >> 2;
        const int m = q * 32 + lane;
        const bool m_ok = kidx0 + m < Kc;
        for (int c = g; c < NCH; c += GROUPS) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            if (m_ok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) out[(size_t)(c * 32 + i) * Kc + kidx0 + m] = v[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PRODUCERS / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

template <int BN>
int launch_wg(const sdt_conv_desc* d, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        SDT_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<BN>::SMEM));
        attr_set = true;
    }
    const int Kc = d->TH * d->TW * d->C;
    dim3 grid((Kc + BM - 1) / BM, 1, d->splits);
    sdt::launch(tc_wgrad_kernel<BN>, dim3(grid), dim3(THREADS), WgCfg<BN>::SMEM, st, *d);
    SDT_LAUNCH_OK("tc_wgrad_kernel");
    sdt_note_tc_launch();
    return SDT_OK;
}

}  // namespace

bool sdt_tc_wgrad_eligible(const sdt_conv_desc* d) {
    if (d->C % 32 != 0) return false;
    if (!(d->N == 64 || d->N == 128 || d->N == 256)) return false;
    if ((((uintptr_t)d->src | (uintptr_t)d->dy | (uintptr_t)d->wpart | (uintptr_t)d->xf_scale | (uintptr_t)d->xf_shift) & 15) != 0) return false;
    return true;
}

int sdt_tc_wgrad_launch(const sdt_conv_desc* d, cudaStream_t st) {
    if (d->N == 256) return launch_wg<256>(d, st);
    if (d->N == 128) return launch_wg<128>(d, st);
    return launch_wg<64>(d, st);
}
