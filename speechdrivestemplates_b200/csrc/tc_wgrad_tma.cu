// Convolution weight-gradient on tcgen05 with both operands delivered by TMA (math mode 2).
//
//   D[m = k-index (tap, ci)][n = output channel] = sum over pixels  A(pix, k-index) * dy[pix, n]
//
// A k-block is a (pb x ph x pw) = 32-pixel box of the output grid (it may span images for small maps).  Both operands are
// MN-major: per 32-channel chunk one 4-D TMA box {32 ch, pw*xs, ph*ys, pb} lands as [32 pixels][128 B] with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, which is tcgen05's SWIZZLE_128B_BASE32B layout (the only MN-major swizzle for
// tf32): 8 K-atoms of 4 pixels x 128 B, K-atoms 512 B apart (SBO), chunks (MN atoms) 4 KB apart (LBO).  Pixels outside
// the output grid are TMA zero-fill on the dy side, so they contribute nothing; the im2col taps use the same signed
// coordinates / element strides as the forward kernel.  The source must be a plain (activated) tensor.
// CTA tile = 128 k-indices x BN channels, split-K over the list of pixel boxes (gridDim.z).
#include <cuda.h>

#include "tc_api.h"
#include "tc_common.cuh"

namespace {

using namespace sdt_tc;

constexpr int BM = 128;
constexpr int PIX = 32;
constexpr int THREADS = 192;

struct BoxGeom {
    int pw, ph, pb;        // pixel box, pw * ph * pb == 32
    int nx, ny, nb;        // boxes along x, y, batch
};

template <int BN>
struct WgTmaCfg {
    static constexpr int STAGES = BN == 64 ? 4 : 3;
    static constexpr int A_BYTES = PIX * BM * 4;
    static constexpr int B_BYTES = PIX * BN * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + 1024;
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 2) tc_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const sdt_conv_desc d, const BoxGeom bg) {
    using Cfg = WgTmaCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NCH = BN / 32;
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
    const int Ntot = d.N;
    const int n0 = blockIdx.y * BN;
    const int kidx0 = blockIdx.x * BM;
    const int NP = bg.nx * bg.ny * bg.nb;
    const int chunk = (NP + d.splits - 1) / d.splits;
    const int p_begin = blockIdx.z * chunk;
    const int p_end = p_begin + chunk < NP ? p_begin + chunk : NP;
    const int KB = p_end > p_begin ? p_end - p_begin : 0;
    float* out = d.wpart + (size_t)blockIdx.z * Ntot * Kc;

    if (KB == 0) {
        sdt::pdl_wait();
        for (int e = tid; e < BN * BM; e += THREADS) {
            const int n = e / BM, m = e % BM;
            if (kidx0 + m < Kc) out[(size_t)(n0 + n) * Kc + kidx0 + m] = 0.f;
        }
        return;
    }

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

    // warps 0-3: epilogue (TMEM lane quadrant = warp id); warp 4: TMA producer; warp 5: MMA issuer.  The issuing warps walk
    // their loops as whole warps and elect one lane per instruction (tc_common.cuh: elect_one), and have the highest warp ids.
    if (warp == 4) {
        // ================= TMA producer =================
        // the four 32-channel chunks of this CTA's 128 k-indices: tap and channel offset are loop invariant
        int c0[4], dyo[4], dxo[4];
        int n_chunks = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kidx = kidx0 + j * 32;
            c0[j] = dyo[j] = dxo[j] = 0;
            if (kidx < Kc) {
                const int tap = kidx / d.C;
                c0[j] = kidx - tap * d.C;
                const int tyy = tap / d.TW, txx = tap - tyy * d.TW;
                dyo[j] = d.y_off + tyy * d.ty_mul;
                dxo[j] = d.x_off + txx * d.tx_mul;
                n_chunks = j + 1;
            }
        }
        const uint32_t tx_bytes = (uint32_t)(n_chunks * 4096 + Cfg::B_BYTES);
        const int per_image = bg.ny * bg.nx;
        int s = 0, par = 1;
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait_spin(empty_bar(s), (uint32_t)par);
            const int pi = p_begin + kb;
            const int bb = pi / per_image;
            const int rem = pi - bb * per_image;
            const int yy = rem / bg.nx, xx = rem - yy * bg.nx;
            const int b0 = bb * bg.pb, y0 = yy * bg.ph, x0 = xx * bg.pw;
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), tx_bytes);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < n_chunks)
                        tma_load_4d(smA + s * Cfg::A_BYTES + j * 4096, &tmA, c0[j], x0 * d.x_mul + dxo[j], y0 * d.y_mul + dyo[j], b0, full_bar(s));
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                    tma_load_4d(smB + s * Cfg::B_BYTES + j * 4096, &tmB, n0 + j * 32, x0, y0, b0, full_bar(s));
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; par ^= 1; }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_tf32(BN, 1, 1);
        // MN-major SWIZZLE_128B_BASE32B: 8 pixels = two K-atoms 512 B apart (SBO); 32-channel chunks (MN atoms) 4 KB apart (LBO)
        constexpr uint32_t HI = desc_hi(512, kSwizzle128B_Base32B);
        int s = 0, par = 0;
        uint32_t started = 0;
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait_spin(full_bar(s), (uint32_t)par);
            tc_fence_after();
            const uint32_t a_lo = desc_lo(smA + s * Cfg::A_BYTES, 4096), b_lo = desc_lo(smB + s * Cfg::B_BYTES, 4096);
            if (elect_one()) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    mma_tf32_lohi(tmem_base, a_lo + (uint32_t)(k4 * 64), b_lo + (uint32_t)(k4 * 64), HI, idesc, started | (uint32_t)k4);
                mma_commit(empty_bar(s));
            }
            __syncwarp();
            started = 1;
            if (++s == STAGES) { s = 0; par ^= 1; }
        }
        if (elect_one()) mma_commit(tmem_full_bar);
        __syncwarp();
    } else {
        // ================= epilogue: D[m][n] -> wpart[z][n0 + n][kidx0 + m] =================
        const int q = warp;
        const int m = q * 32 + lane;
        const bool m_ok = kidx0 + m < Kc;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        for (int c = 0; c < NCH; ++c) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            if (m_ok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) out[(size_t)(n0 + c * 32 + i) * Kc + kidx0 + m] = v[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
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

BoxGeom pick_box(const sdt_conv_desc* d) {
    BoxGeom best{32, 1, 1, 0, 0, 0};
    long long best_n = -1;
    for (int pw = 32; pw >= 1; pw /= 2) {
        for (int ph = 32 / pw; ph >= 1; ph /= 2) {
            const int pb = 32 / (pw * ph);
            if (pw * d->x_mul > 256 || ph * d->y_mul > 256) continue;
            const int nx = (d->GW + pw - 1) / pw, ny = (d->GH + ph - 1) / ph, nb = (d->B + pb - 1) / pb;
            const long long n = (long long)nx * ny * nb;
            if (best_n < 0 || n < best_n) {
                best_n = n;
                best = BoxGeom{pw, ph, pb, nx, ny, nb};
            }
        }
    }
    return best;
}

template <int BN>
int launch_wg_tma(const sdt_conv_desc* d, const BoxGeom& bg, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    SDT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    static bool attr_set = false;
    if (!attr_set) {
        SDT_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgTmaCfg<BN>::SMEM));
        attr_set = true;
    }
    alignas(64) CUtensorMap tmA, tmB;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->SW, (cuuint64_t)d->SH, (cuuint64_t)d->B};
        const cuuint64_t strides[3] = {(cuuint64_t)d->C * 4, (cuuint64_t)d->SW * d->C * 4, (cuuint64_t)d->SH * d->SW * d->C * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)(bg.pw * d->x_mul), (cuuint32_t)(bg.ph * d->y_mul), (cuuint32_t)bg.pb};
        const cuuint32_t estr[4] = {1, (cuuint32_t)d->x_mul, (cuuint32_t)d->y_mul, 1};
        const CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->src), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad A) failed with %d", (int)r);
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)d->N, (cuuint64_t)d->GW, (cuuint64_t)d->GH, (cuuint64_t)d->B};
        const cuuint64_t strides[3] = {(cuuint64_t)d->N * 4, (cuuint64_t)d->GW * d->N * 4, (cuuint64_t)d->GH * d->GW * d->N * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)bg.pw, (cuuint32_t)bg.ph, (cuuint32_t)bg.pb};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->dy), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad dy) failed with %d", (int)r);
    }
    const int Kc = d->TH * d->TW * d->C;
    dim3 grid((Kc + BM - 1) / BM, d->N / BN, d->splits);
    sdt::launch(tc_wgrad_tma_kernel<BN>, dim3(grid), dim3(THREADS), WgTmaCfg<BN>::SMEM, st, tmA, tmB, *d, bg);
    SDT_LAUNCH_OK("tc_wgrad_tma_kernel");
    sdt_note_tc_launch();
    return SDT_OK;
}

}  // namespace

bool sdt_tc_wgrad_tma_eligible(const sdt_conv_desc* d) {
    if (d->xf_scale != nullptr) return false;
    if (d->C % 32 != 0 || d->N % 64 != 0) return false;
    if (d->x_mul < 1 || d->x_mul > 8 || d->y_mul < 1 || d->y_mul > 8) return false;
    if ((((uintptr_t)d->src | (uintptr_t)d->dy | (uintptr_t)d->wpart) & 15) != 0) return false;
    return get_encode() != nullptr;
}

int sdt_tc_wgrad_tma_launch(const sdt_conv_desc* d, cudaStream_t st) {
    const BoxGeom bg = pick_box(d);
    if (d->N % 128 == 0) return launch_wg_tma<128>(d, bg, st);
    return launch_wg_tma<64>(d, bg, st);
}
