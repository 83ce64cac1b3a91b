// Convolution weight gradient on tcgen05 with operand reuse (math modes 3, 4) for 2-D layers with C % 128 == 0:
//
//   D_tap[m = input channel][n = output channel] = sum over pixels  x(pix shifted by tap, m) * dy[pix, n]
//
// tc_wgrad_tma.cu computes one 128 (k-index) x 128 tile per CTA and streams a 16 KB x box + a 16 KB dy box per 32-pixel k-block:
// 32 FLOP per byte of L2 -> shared-memory traffic, which is what bounds it (~46 B/clk/SM measured).  Here a CTA owns up to three
// VERTICAL taps of one kernel column for a block of 128 input channels: their x operands are the same pixels shifted by whole rows
// of the (pw = 8)-wide pixel box, i.e. by multiples of 1024 B = two K-atoms of the MN-major SWIZZLE_128B_BASE32B layout, so ONE box
// of (ph + taps - 1) rows serves all of them (a different start address in the matrix descriptor per tap), and the dy box is
// shared by the taps' accumulators: 40 KB per 12 MMAs instead of 96 KB (3x3), 36 KB per 8 instead of 64 KB (4x4 stride 2).
//
//   warps 0-3: epilogue (TMEM lane quadrant = warp id)   warp 4: TMA producer   warp 5: MMA issuer (whole-warp loops, elect.sync)
// One CTA per SM (up to 384 TMEM columns, 5 stages of <= 40 KB); split-K over the list of pixel boxes (gridDim.z = d.splits), the
// partials are reduced by sdt_conv_wgrad_reduce exactly as for the other weight-gradient kernels.
#include <cuda.h>

#include "tc_api.h"
#include "tc_common.cuh"

namespace {

using namespace sdt_tc;

constexpr int THREADS = 192;
constexpr int PW = 8, PH = 4;              // pixel box of a k-block: 4 rows x 8 columns of the output grid
constexpr int NA_MAX = 3;                  // accumulators (vertical taps) per CTA
constexpr int BN = 128;
constexpr int MAX_GROUPS = 8;
constexpr int SMEM_MAX = 227 * 1024;
constexpr int STAGES_MAX = 6;

struct WGeom {
    int nx, ny;                // pixel boxes per image
    int box_rows;              // PH + NA - 1
    int a_chunk_bytes;         // box_rows * PW * 128: one 32-channel chunk of the x box
    int stages;
    int n_groups;
    // per group, 16 bits: [3:0] taps, [7:4] first ty, [11:8] ty step
    unsigned long long groups;
};
__host__ __device__ __forceinline__ int wg_taps(unsigned long long p, int gi) { return (int)((p >> (16 * gi)) & 15u); }
__host__ __device__ __forceinline__ int wg_ty0(unsigned long long p, int gi) { return (int)((p >> (16 * gi + 4)) & 15u); }
__host__ __device__ __forceinline__ int wg_step(unsigned long long p, int gi) { return (int)((p >> (16 * gi + 8)) & 15u); }

__global__ void __launch_bounds__(THREADS, 1) tc_wgrad_ytap_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const sdt_conv_desc d, const WGeom g) {
    constexpr int B_BYTES = 32 * BN * 4;               // dy box: 32 pixels x 128 channels
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smA = smem_u32(smem_raw);
    const int a_bytes = 4 * g.a_chunk_bytes;           // four 32-channel chunks
    const uint32_t smB = smA + g.stages * a_bytes;
    const int ring_bytes = g.stages * (a_bytes + B_BYTES);
    const uint32_t bars = smA + ring_bytes;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + ring_bytes + (2 * STAGES_MAX + 1) * 8);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES_MAX + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES_MAX);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Kc = d.TH * d.TW * d.C;
    const int N = d.N;
    // blockIdx.x = ((channel block * TW) + tx) * n_groups + group ; blockIdx.y = n tile ; blockIdx.z = split
    const int gi = blockIdx.x % g.n_groups;
    const int tx = (blockIdx.x / g.n_groups) % d.TW;
    const int cb = blockIdx.x / (g.n_groups * d.TW);
    const int na = wg_taps(g.groups, gi), ty0 = wg_ty0(g.groups, gi), tstep = wg_step(g.groups, gi);
    const int n0 = blockIdx.y * BN;
    const int NP = g.nx * g.ny * d.B;
    const int chunk = (NP + d.splits - 1) / d.splits;
    const int p_begin = blockIdx.z * chunk;
    const int p_end = p_begin + chunk < NP ? p_begin + chunk : NP;
    const int KB = p_end > p_begin ? p_end - p_begin : 0;
    float* out = d.wpart + (size_t)blockIdx.z * N * Kc;

    if (KB == 0) {                                      // empty split: its partial is zero
        sdt::pdl_wait();
        for (int j = 0; j < na; ++j) {
            const int kbase = ((ty0 + j * tstep) * d.TW + tx) * d.C + cb * 128;
            for (int e = tid; e < BN * 128; e += THREADS) out[(size_t)(n0 + e / 128) * Kc + kbase + (e % 128)] = 0.f;
        }
        return;
    }
    {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if ((smA & 1023u) != 0 || ring_bytes + (2 * STAGES_MAX + 1) * 8 + 16 > (int)dyn) {
            if (tid == 0) printf("tc_wgrad_ytap_kernel: shared-memory window misaligned or too small\n");
            __trap();
        }
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES_MAX; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();

    if (warp == 4) {
        // ================= TMA producer =================
        const int per_image = g.nx * g.ny;
        const uint32_t tx_bytes = (uint32_t)(a_bytes + B_BYTES);
        const int x_add = d.x_off + tx * d.tx_mul, y_add = d.y_off + ty0 * d.ty_mul;
        int s = 0, par = 1;
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait_spin(empty_bar(s), (uint32_t)par);
            const int pi = p_begin + kb;
            const int bb = pi / per_image;
            const int rem = pi - bb * per_image;
            const int yy = rem / g.nx, xx = rem - yy * g.nx;
            const int y0 = yy * PH, x0 = xx * PW;
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), tx_bytes);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    tma_load_4d(smA + s * a_bytes + j * g.a_chunk_bytes, &tmA, cb * 128 + j * 32, x0 * d.x_mul + x_add, y0 * d.y_mul + y_add, bb,
                                full_bar(s));
                    tma_load_4d(smB + s * B_BYTES + j * 4096, &tmB, n0 + j * 32, x0, y0, bb, full_bar(s));
                }
            }
            __syncwarp();
            if (++s == g.stages) { s = 0; par ^= 1; }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_tf32(BN, 1, 1);
        // MN-major SWIZZLE_128B_BASE32B: K-atoms (4 pixels x 128 B) 512 B apart (SBO); 32-channel chunks (MN atoms) one chunk apart (LBO)
        constexpr uint32_t HI = desc_hi(512, kSwizzle128B_Base32B);
        int s = 0, par = 0;
        uint32_t started = 0;
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait_spin(full_bar(s), (uint32_t)par);
            tc_fence_after();
            const uint32_t a_lo = desc_lo(smA + s * a_bytes, (uint32_t)g.a_chunk_bytes), b_lo = desc_lo(smB + s * B_BYTES, 4096);
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < NA_MAX; ++j) {
                    if (j < na) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)      // tap j = the box shifted by j rows of 8 pixels (1024 B); k4 = 8 pixels further
                            mma_tf32_lohi(tmem_base + (uint32_t)(j * BN), a_lo + (uint32_t)((j + k4) * 64), b_lo + (uint32_t)(k4 * 64), HI, idesc,
                                          started | (uint32_t)k4);
                    }
                }
                mma_commit(empty_bar(s));
            }
            __syncwarp();
            started = 1;
            if (++s == g.stages) { s = 0; par ^= 1; }
        }
        if (elect_one()) mma_commit(tmem_full_bar);
        __syncwarp();
    } else {
        // ================= epilogue: D_tap[m][n] -> wpart[z][n0 + n][(tap * C) + cb*128 + m] =================
        const int q = warp;
        const int m = q * 32 + lane;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        for (int j = 0; j < na; ++j) {
            const int kbase = ((ty0 + j * tstep) * d.TW + tx) * d.C + cb * 128;
            for (int c = 0; c < BN / 32; ++c) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * BN + c * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) out[(size_t)(n0 + c * 32 + i) * Kc + kbase + m] = v[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
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

// groups of <= NA_MAX vertical taps that share a box: taps of one y-phase (ty = p, p + y_mul, ...) in runs of consecutive shifts
bool make_geom(const sdt_conv_desc* d, WGeom* g, int* na_max) {
    g->groups = 0;
    g->n_groups = 0;
    *na_max = 0;
    if (d->ty_mul != 1 || d->TH > 15 || d->y_mul > 15) return false;
    const int phases = d->TH < d->y_mul ? d->TH : d->y_mul;
    for (int p = 0; p < phases; ++p) {
        const int cnt = (d->TH - p + d->y_mul - 1) / d->y_mul;
        for (int first = 0; first < cnt; first += NA_MAX) {
            const int n = cnt - first < NA_MAX ? cnt - first : NA_MAX;
            if (g->n_groups >= MAX_GROUPS / 2) return false;         // 4 groups of 16 bits
            g->groups |= (unsigned long long)(n | ((p + first * d->y_mul) << 4) | (d->y_mul << 8)) << (16 * g->n_groups);
            ++g->n_groups;
            if (n > *na_max) *na_max = n;
        }
    }
    g->box_rows = PH + *na_max - 1;
    g->a_chunk_bytes = g->box_rows * PW * 128;
    g->nx = (d->GW + PW - 1) / PW;
    g->ny = (d->GH + PH - 1) / PH;
    const int stage = 4 * g->a_chunk_bytes + 32 * BN * 4;
    int st = (SMEM_MAX - (2 * STAGES_MAX + 1) * 8 - 16) / stage;
    if (st > STAGES_MAX) st = STAGES_MAX;
    g->stages = st;
    return st >= 3 && PW * d->x_mul <= 256 && g->box_rows * d->y_mul <= 256;
}

}  // namespace

bool sdt_tc_wgrad_ytap_eligible(const sdt_conv_desc* d) {
    if (d->xf_scale != nullptr) return false;
    if (d->C % 128 != 0 || d->N % 128 != 0) return false;
    if (d->GH < 2 || d->TH < 2) return false;                               // 1-D layers / single taps: nothing to share
    if (d->x_mul < 1 || d->x_mul > 8 || d->y_mul < 1 || d->y_mul > 8) return false;
    if ((((uintptr_t)d->src | (uintptr_t)d->dy | (uintptr_t)d->wpart) & 15) != 0) return false;
    WGeom g;
    int na;
    if (!make_geom(d, &g, &na) || na < 2) return false;
    return get_encode() != nullptr;
}

int sdt_tc_wgrad_ytap_launch(const sdt_conv_desc* d, cudaStream_t st) {
    WGeom g;
    int na;
    SDT_REQUIRE(make_geom(d, &g, &na), "sdt_tc_wgrad_ytap_launch: unsupported geometry");
    EncodeTiledFn enc = get_encode();
    SDT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    static bool attr_set = false;
    if (!attr_set) {
        SDT_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_ytap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        attr_set = true;
    }
    alignas(64) CUtensorMap tmA, tmB;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->SW, (cuuint64_t)d->SH, (cuuint64_t)d->B};
        const cuuint64_t strides[3] = {(cuuint64_t)d->C * 4, (cuuint64_t)d->SW * d->C * 4, (cuuint64_t)d->SH * d->SW * d->C * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)(PW * d->x_mul), (cuuint32_t)(g.box_rows * d->y_mul), 1};
        const cuuint32_t estr[4] = {1, (cuuint32_t)d->x_mul, (cuuint32_t)d->y_mul, 1};
        const CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->src), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad x, y-tap) failed with %d", (int)r);
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)d->N, (cuuint64_t)d->GW, (cuuint64_t)d->GH, (cuuint64_t)d->B};
        const cuuint64_t strides[3] = {(cuuint64_t)d->N * 4, (cuuint64_t)d->GW * d->N * 4, (cuuint64_t)d->GH * d->GW * d->N * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)PW, (cuuint32_t)PH, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->dy), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SDT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad dy, y-tap) failed with %d", (int)r);
    }
    const int smem = g.stages * (4 * g.a_chunk_bytes + 32 * BN * 4) + (2 * STAGES_MAX + 1) * 8 + 16;
    dim3 grid((d->C / 128) * d->TW * g.n_groups, d->N / BN, d->splits);
    sdt::launch(tc_wgrad_ytap_kernel, grid, dim3(THREADS), smem, st, tmA, tmB, *d, g);
    SDT_LAUNCH_OK("tc_wgrad_ytap_kernel");
    sdt_note_tc_launch();
    return SDT_OK;
}
