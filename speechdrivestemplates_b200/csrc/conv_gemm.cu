// Convolution (1-D / 2-D, forward / data-gradient / weight-gradient) as implicit GEMM on channels-last fp32
// tensors -- the fp32 FFMA ("math mode 0") path.  Replaces nn.Conv1d/nn.Conv2d inside ConvNormRelu
// (core/networks/building_blocks.py:15-22,31-36,49) and their autograd backward.
//
// The previous layer's normalisation + activation is applied by the A-operand loader (x*scale+shift, LeakyReLU),
// and the per-channel sum / sum-of-squares the NEXT normalisation needs are produced by the epilogue, so the
// normalised activation tensor is never written to HBM (building_blocks.py:50-54 fused away).
//
// Tiling: BMxBNx16 CTA tiles, 256 threads, 8x8 / 8x4 / 4x4 register tiles, register-prefetch double buffering.
#include "common.cuh"
#include <stdlib.h>

#include "tc_api.h"

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

struct RowSmem {
    int b, sy0, sx0;
    long long dst;  // element offset of the output row, -1 = row out of range
};

__device__ __forceinline__ float xf_apply(float v, float sc, float sh, float slope) {
    return sdt::leaky(fmaf(v, sc, sh), slope);
}

// ------------------------------------------------------------------------------------------------
// forward / dgrad:  dst[row(m), n] = sum_k A(m,k) * wt[k, n]
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(NTHREADS, 2) conv_gemm_kernel(const sdt_conv_desc d) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    static_assert((BM / TM) * (BN / TN) == NTHREADS, "thread tile");
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int GM = TM / 4, GN = TN / 4;
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];
    __shared__ RowSmem rows[BM];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int K = d.TH * d.TW * d.C;
    const int N = d.N;
    const int P = d.GH * d.GW;
    const int n0 = blockIdx.y * BN;

    // ---- decode the BM rows of this tile once
    for (int r = tid; r < BM; r += NTHREADS) {
        int b, rem;
        bool ok;
        if (d.per_image_tiles) {
            const int tpi = (P + BM - 1) / BM;
            b = blockIdx.x / tpi;
            rem = (blockIdx.x % tpi) * BM + r;
            ok = rem < P;
        } else {
            const long long gm = (long long)blockIdx.x * BM + r;
            ok = gm < (long long)d.B * P;
            b = ok ? (int)(gm / P) : 0;
            rem = ok ? (int)(gm % P) : 0;
        }
        RowSmem ri;
        const int gy = rem / d.GW, gx = rem % d.GW;
        ri.b = b;
        ri.sy0 = gy * d.y_mul + d.y_off;
        ri.sx0 = gx * d.x_mul + d.x_off;
        ri.dst = ok ? (((long long)b * d.DH + (gy * d.dy_mul + d.dy_off)) * d.DW + (gx * d.dx_mul + d.dx_off)) * N : -1;
        rows[r] = ri;
    }
    __syncthreads();

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // ---- loaders.  A: row-fastest mapping (conflict-free transposed smem stores)
    constexpr int A_PER_THREAD = BM * BK / NTHREADS;                 // scalars
    constexpr int A_ITERS = VEC ? A_PER_THREAD / 4 : A_PER_THREAD;
    constexpr int B_PER_THREAD = BK * BN / NTHREADS;
    constexpr int B_ITERS = B_PER_THREAD / 4;                        // float4 slots (scalar fallback inside)
    float a_reg[A_PER_THREAD];
    float b_reg[B_PER_THREAD];
    const int a_row = tid % BM;
    const int a_k0 = tid / BM;  // + (NTHREADS/BM) * i
    const bool has_xf = d.xf_scale != nullptr;
    const bool n_vec = (N % 4) == 0;

    auto load_tiles = [&](int kt) {
        const int kbase = kt * BK;
        const RowSmem ri = rows[a_row];
        const bool row_ok = ri.dst >= 0;
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            const int kk = (a_k0 + (NTHREADS / BM) * i) * (VEC ? 4 : 1);
            const int k = kbase + kk;
            float v[VEC ? 4 : 1];
#pragma unroll
            for (int q = 0; q < (VEC ? 4 : 1); ++q) v[q] = 0.f;
            if (row_ok && k < K) {
                const int tap = k / d.C, c = k - tap * d.C;
                const int tyy = tap / d.TW, txx = tap - tyy * d.TW;
                const int sy = ri.sy0 + tyy * d.ty_mul, sx = ri.sx0 + txx * d.tx_mul;
                if (sy >= 0 && sy < d.SH && sx >= 0 && sx < d.SW) {
                    const float* p = d.src + (((size_t)ri.b * d.SH + sy) * d.SW + sx) * d.C + c;
                    if (VEC) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        if (has_xf) {
                            const size_t o = (size_t)ri.b * d.xf_bstride + c;
                            const float4 sc = __ldg(reinterpret_cast<const float4*>(d.xf_scale + o));
                            const float4 sh = __ldg(reinterpret_cast<const float4*>(d.xf_shift + o));
                            v[0] = xf_apply(v[0], sc.x, sh.x, d.xf_slope);
                            v[1] = xf_apply(v[1], sc.y, sh.y, d.xf_slope);
                            v[2] = xf_apply(v[2], sc.z, sh.z, d.xf_slope);
                            v[3] = xf_apply(v[3], sc.w, sh.w, d.xf_slope);
                        }
                    } else {
                        v[0] = __ldg(p);
                        if (has_xf) {
                            const size_t o = (size_t)ri.b * d.xf_bstride + c;
                            v[0] = xf_apply(v[0], __ldg(d.xf_scale + o), __ldg(d.xf_shift + o), d.xf_slope);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < (VEC ? 4 : 1); ++q) a_reg[i * (VEC ? 4 : 1) + q] = v[q];
        }
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            const int f = tid + i * NTHREADS;
            const int kk = f / (BN / 4), nv = f % (BN / 4);
            const int k = kbase + kk, n = n0 + nv * 4;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < K) {
                const float* p = d.wt + (size_t)k * N + n;
                if (n_vec) {
                    if (n < N) t = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (n + 0 < N) t.x = __ldg(p + 0);
                    if (n + 1 < N) t.y = __ldg(p + 1);
                    if (n + 2 < N) t.z = __ldg(p + 2);
                    if (n + 3 < N) t.w = __ldg(p + 3);
                }
            }
            b_reg[i * 4 + 0] = t.x; b_reg[i * 4 + 1] = t.y; b_reg[i * 4 + 2] = t.z; b_reg[i * 4 + 3] = t.w;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            const int kk = (a_k0 + (NTHREADS / BM) * i) * (VEC ? 4 : 1);
#pragma unroll
            for (int q = 0; q < (VEC ? 4 : 1); ++q) As[buf][kk + q][a_row] = a_reg[i * (VEC ? 4 : 1) + q];
        }
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            const int f = tid + i * NTHREADS;
            const int kk = f / (BN / 4), nv = f % (BN / 4);
            *reinterpret_cast<float4*>(&Bs[buf][kk][nv * 4]) =
                make_float4(b_reg[i * 4 + 0], b_reg[i * 4 + 1], b_reg[i * 4 + 2], b_reg[i * 4 + 3]);
        }
    };

    const int KT = (K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) load_tiles(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float af[TM], bf[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][g * (BM / GM) + ty * 4]);
                af[g * 4 + 0] = t.x; af[g * 4 + 1] = t.y; af[g * 4 + 2] = t.z; af[g * 4 + 3] = t.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * (BN / GN) + tx * 4]);
                bf[g * 4 + 0] = t.x; bf[g * 4 + 1] = t.y; bf[g * 4 + 2] = t.z; bf[g * 4 + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
        }
        if (kt + 1 < KT) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue: bias, store (optionally accumulate)
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = (i / 4) * (BM / GM) + ty * 4 + (i % 4);
        const long long off = rows[r].dst;
        if (off < 0) continue;
#pragma unroll
        for (int g = 0; g < GN; ++g) {
            const int n = n0 + g * (BN / GN) + tx * 4;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q] = acc[i][g * 4 + q];
                if (d.bias != nullptr && n + q < N) v[q] += __ldg(d.bias + n + q);
            }
            float* p = d.dst + off + n;
            if (n_vec) {
                if (n < N) {
                    float4 o = make_float4(v[0], v[1], v[2], v[3]);
                    if (d.accumulate) {
                        const float4 old = *reinterpret_cast<const float4*>(p);
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    *reinterpret_cast<float4*>(p) = o;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (n + q < N) p[q] = d.accumulate ? p[q] + v[q] : v[q];
            }
        }
    }

    // ---- epilogue: column sums / sums of squares of this row tile (input of sdt_norm_finalize).
    // Rows outside the range hold exact zeros (their A rows were zero), so they do not disturb the sums.
    if (d.stat_partial != nullptr) {
        __syncthreads();  // everyone is done with As/Bs
        float* red_s = &As[0][0][0];   // [BM/TM][BN]
        float* red_q = &Bs[0][0][0];
        static_assert((BM / TM) * BN <= 2 * BK * LDA && (BM / TM) * BN <= 2 * BK * LDB, "reduction scratch");
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                s += acc[i][j];
                q = fmaf(acc[i][j], acc[i][j], q);
            }
            const int c = (j / 4) * (BN / GN) + tx * 4 + (j % 4);
            red_s[ty * BN + c] = s;
            red_q[ty * BN + c] = q;
        }
        __syncthreads();
        for (int c = tid; c < BN; c += NTHREADS) {
            if (n0 + c < N) {
                float s = 0.f, q = 0.f;
                for (int t = 0; t < BM / TM; ++t) {
                    s += red_s[t * BN + c];
                    q += red_q[t * BN + c];
                }
                d.stat_partial[((size_t)blockIdx.x * 2 + 0) * N + n0 + c] = s;
                d.stat_partial[((size_t)blockIdx.x * 2 + 1) * N + n0 + c] = q;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// wgrad:  wpart[z][n][k] = sum_{m in split z} dy[m, n] * A(m, k)
//   GEMM rows = output channels n (d.N), GEMM cols = k = (ty,tx,c), contraction over pixels m.
// ------------------------------------------------------------------------------------------------
template <int BM, int TM, bool VECA, bool VECB>
__global__ void __launch_bounds__(NTHREADS, 2) conv_wgrad_kernel(const sdt_conv_desc d) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    constexpr int BN = 128, TN = 8;
    static_assert((BM / TM) * (BN / TN) == NTHREADS, "thread tile");
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int GM = TM / 4, GN = TN / 4;
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int Kc = d.TH * d.TW * d.C;       // GEMM N extent
    const int Nout = d.N;                   // GEMM M extent
    const int P = d.GH * d.GW;
    const long long Mtot = (long long)d.B * P;
    const int m0 = blockIdx.y * BM;         // output-channel offset
    const int c0 = blockIdx.x * BN;         // k offset
    long long chunk = (Mtot + d.splits - 1) / d.splits;
    chunk = (chunk + BK - 1) / BK * BK;
    const long long p_begin = (long long)blockIdx.z * chunk;
    const long long p_end = p_begin + chunk < Mtot ? p_begin + chunk : Mtot;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    constexpr int A_PER_THREAD = BK * BM / NTHREADS;
    constexpr int B_PER_THREAD = BK * BN / NTHREADS;  // 8
    float a_reg[A_PER_THREAD], b_reg[B_PER_THREAD];
    const bool has_xf = d.xf_scale != nullptr;

    // B-operand column decode is loop invariant.
    // VECB: thread owns 4 consecutive k (same tap) at nv = tid % 32; rows kk = tid/32 + 8*i
    // scalar: thread owns k = c0 + tid % 128; rows kk = tid/128 + 2*i
    const int bcol = VECB ? (tid % (BN / 4)) * 4 : tid % BN;
    const int kcol = c0 + bcol;
    const bool col_ok = kcol < Kc;
    int tap = 0, cc = 0, tyy = 0, txx = 0;
    if (col_ok) {
        tap = kcol / d.C; cc = kcol - tap * d.C;
        tyy = tap / d.TW; txx = tap - tyy * d.TW;
    }

    auto load_tiles = [&](long long pbase) {
        // A: dy[pix][m0 + m]
        if (VECA) {
#pragma unroll
            for (int i = 0; i < A_PER_THREAD / 4; ++i) {
                const int f = tid + i * NTHREADS;
                const int kk = f / (BM / 4), mv = f % (BM / 4);
                const long long pix = pbase + kk;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pix < p_end && m0 + mv * 4 < Nout) t = __ldg(reinterpret_cast<const float4*>(d.dy + pix * Nout + m0 + mv * 4));
                a_reg[i * 4 + 0] = t.x; a_reg[i * 4 + 1] = t.y; a_reg[i * 4 + 2] = t.z; a_reg[i * 4 + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < A_PER_THREAD; ++i) {
                const int f = tid + i * NTHREADS;
                const int kk = f / BM, m = f % BM;
                const long long pix = pbase + kk;
                a_reg[i] = (pix < p_end && m0 + m < Nout) ? __ldg(d.dy + pix * Nout + m0 + m) : 0.f;
            }
        }
        // B: xf(src[b, sy, sx, c])
#pragma unroll
        for (int i = 0; i < (VECB ? B_PER_THREAD / 4 : B_PER_THREAD); ++i) {
            const int kk = VECB ? tid / (BN / 4) + (NTHREADS / (BN / 4)) * i : tid / BN + (NTHREADS / BN) * i;
            const long long pix = pbase + kk;
            float v[VECB ? 4 : 1];
#pragma unroll
            for (int q = 0; q < (VECB ? 4 : 1); ++q) v[q] = 0.f;
            if (col_ok && pix < p_end) {
                const int b = (int)(pix / P);
                const int rem = (int)(pix - (long long)b * P);
                const int gy = rem / d.GW, gx = rem - gy * d.GW;
                const int sy = gy * d.y_mul + d.y_off + tyy * d.ty_mul;
                const int sx = gx * d.x_mul + d.x_off + txx * d.tx_mul;
                if (sy >= 0 && sy < d.SH && sx >= 0 && sx < d.SW) {
                    const float* p = d.src + (((size_t)b * d.SH + sy) * d.SW + sx) * d.C + cc;
                    if (VECB) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        if (has_xf) {
                            const size_t o = (size_t)b * d.xf_bstride + cc;
                            const float4 sc = __ldg(reinterpret_cast<const float4*>(d.xf_scale + o));
                            const float4 sh = __ldg(reinterpret_cast<const float4*>(d.xf_shift + o));
                            v[0] = xf_apply(v[0], sc.x, sh.x, d.xf_slope);
                            v[1] = xf_apply(v[1], sc.y, sh.y, d.xf_slope);
                            v[2] = xf_apply(v[2], sc.z, sh.z, d.xf_slope);
                            v[3] = xf_apply(v[3], sc.w, sh.w, d.xf_slope);
                        }
                    } else {
                        v[0] = __ldg(p);
                        if (has_xf) {
                            const size_t o = (size_t)b * d.xf_bstride + cc;
                            v[0] = xf_apply(v[0], __ldg(d.xf_scale + o), __ldg(d.xf_shift + o), d.xf_slope);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < (VECB ? 4 : 1); ++q) b_reg[i * (VECB ? 4 : 1) + q] = v[q];
        }
    };
    auto store_tiles = [&](int buf) {
        if (VECA) {
#pragma unroll
            for (int i = 0; i < A_PER_THREAD / 4; ++i) {
                const int f = tid + i * NTHREADS;
                const int kk = f / (BM / 4), mv = f % (BM / 4);
                *reinterpret_cast<float4*>(&As[buf][kk][mv * 4]) =
                    make_float4(a_reg[i * 4 + 0], a_reg[i * 4 + 1], a_reg[i * 4 + 2], a_reg[i * 4 + 3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < A_PER_THREAD; ++i) {
                const int f = tid + i * NTHREADS;
                As[buf][f / BM][f % BM] = a_reg[i];
            }
        }
#pragma unroll
        for (int i = 0; i < (VECB ? B_PER_THREAD / 4 : B_PER_THREAD); ++i) {
            const int kk = VECB ? tid / (BN / 4) + (NTHREADS / (BN / 4)) * i : tid / BN + (NTHREADS / BN) * i;
            if (VECB) {
                *reinterpret_cast<float4*>(&Bs[buf][kk][bcol]) =
                    make_float4(b_reg[i * 4 + 0], b_reg[i * 4 + 1], b_reg[i * 4 + 2], b_reg[i * 4 + 3]);
            } else {
                Bs[buf][kk][bcol] = b_reg[i];
            }
        }
    };

    const int KT = p_end > p_begin ? (int)((p_end - p_begin + BK - 1) / BK) : 0;
    if (KT > 0) {
        load_tiles(p_begin);
        store_tiles(0);
    }
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) load_tiles(p_begin + (long long)(kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float af[TM], bf[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][g * (BM / GM) + ty * 4]);
                af[g * 4 + 0] = t.x; af[g * 4 + 1] = t.y; af[g * 4 + 2] = t.z; af[g * 4 + 3] = t.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * (BN / GN) + tx * 4]);
                bf[g * 4 + 0] = t.x; bf[g * 4 + 1] = t.y; bf[g * 4 + 2] = t.z; bf[g * 4 + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
        }
        if (kt + 1 < KT) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    float* out = d.wpart + (size_t)blockIdx.z * Nout * Kc;
    const bool k_vec = (Kc % 4) == 0;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * (BM / GM) + ty * 4 + (i % 4);
        if (m >= Nout) continue;
#pragma unroll
        for (int g = 0; g < GN; ++g) {
            const int k = c0 + g * (BN / GN) + tx * 4;
            float* p = out + (size_t)m * Kc + k;
            if (k_vec) {
                if (k < Kc) *reinterpret_cast<float4*>(p) = make_float4(acc[i][g * 4], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (k + q < Kc) p[q] = acc[i][g * 4 + q];
            }
        }
    }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ wpart, int splits, int N, int C, int T,
                                    float* __restrict__ grad, int accumulate) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    // one thread per element of the reference-layout gradient (n, c, t); reads are strided but the tensor is small
    const long long total = (long long)N * C * T;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int t = (int)(e % T);
    const int c = (int)((e / T) % C);
    const int n = (int)(e / ((long long)T * C));
    const size_t src = (size_t)n * T * C + (size_t)t * C + c;
    const size_t stride = (size_t)N * T * C;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += wpart[z * stride + src];
    grad[e] = accumulate ? grad[e] + s : s;
}

// C % 32 == 0: CTA = (output channel n, 32 input channels); warp w sums taps w, w+8, .. over the splits with 128-byte
// coalesced reads of the (T, C) partials, the (c, t) transpose into the reference layout goes through shared memory.
__global__ void __launch_bounds__(256) wgrad_reduce_tiled_kernel(const float* __restrict__ wpart, int splits, int N, int C, int T,
                                                                 float* __restrict__ grad, int accumulate) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    extern __shared__ float s_t[];          // [T][33]
    const int cchunks = C >> 5;
    const int n = blockIdx.x / cchunks, c0 = (blockIdx.x - n * cchunks) << 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t stride = (size_t)N * T * C;
    for (int t = warp; t < T; t += 8) {
        const float* p = wpart + ((size_t)n * T + t) * C + c0 + lane;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int z = 0;
        for (; z + 4 <= splits; z += 4) {
            s0 += __ldg(p + (size_t)z * stride);
            s1 += __ldg(p + (size_t)(z + 1) * stride);
            s2 += __ldg(p + (size_t)(z + 2) * stride);
            s3 += __ldg(p + (size_t)(z + 3) * stride);
        }
        for (; z < splits; ++z) s0 += __ldg(p + (size_t)z * stride);
        s_t[t * 33 + lane] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    float* out = grad + ((size_t)n * C + c0) * T;
    for (int i = threadIdx.x; i < 32 * T; i += 256) {
        const int c = i / T, t = i - c * T;
        const float v = s_t[t * 33 + c];
        out[i] = accumulate ? out[i] + v : v;
    }
}

// All split-K reductions of a group of layers in ONE launch (the fused trainers defer them to the end of a gradient bucket):
// blockIdx.y = item, blockIdx.x = (output channel, 32-channel chunk) of that item, same tiling as wgrad_reduce_tiled_kernel.
struct ReduceItem {     // mirrors sdt_reduce_item
    const float* wpart;
    float* grad;
    int32_t splits, N, C, T, accumulate, pad0;
};

__global__ void __launch_bounds__(256) wgrad_reduce_batch_kernel(const ReduceItem* __restrict__ items) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    extern __shared__ float s_t[];          // [T][33]
    const ReduceItem it = items[blockIdx.y];
    const int cchunks = it.C >> 5;
    if ((int)blockIdx.x >= it.N * cchunks) return;
    const int C = it.C, T = it.T, splits = it.splits;
    const int n = blockIdx.x / cchunks, c0 = (blockIdx.x - n * cchunks) << 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t stride = (size_t)it.N * T * C;
    for (int t = warp; t < T; t += 8) {
        const float* p = it.wpart + ((size_t)n * T + t) * C + c0 + lane;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int z = 0;
        for (; z + 4 <= splits; z += 4) {
            s0 += __ldg(p + (size_t)z * stride);
            s1 += __ldg(p + (size_t)(z + 1) * stride);
            s2 += __ldg(p + (size_t)(z + 2) * stride);
            s3 += __ldg(p + (size_t)(z + 3) * stride);
        }
        for (; z < splits; ++z) s0 += __ldg(p + (size_t)z * stride);
        s_t[t * 33 + lane] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    float* out = it.grad + ((size_t)n * C + c0) * T;
    for (int i = threadIdx.x; i < 32 * T; i += 256) {
        const int c = i / T, t = i - c * T;
        const float v = s_t[t * 33 + c];
        out[i] = it.accumulate ? out[i] + v : v;
    }
}

__global__ void weight_prep_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int mode, int ky0,
                                   int kx0, int kstep, int TH, int TW, float* __restrict__ out) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long total = (long long)TH * TW * Cin * Cout;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    int co, ci, jy, jx;
    if (mode == 2) {  // out[co*K + (jy*TW+jx)*Cin + ci]
        ci = (int)(e % Cin);
        const int tap = (int)((e / Cin) % (TH * TW));
        co = (int)(e / ((long long)Cin * TH * TW));
        jy = tap / TW; jx = tap % TW;
    } else if (mode == 3) {  // out[ci*K' + (jy*TW+jx)*Cout + co]
        co = (int)(e % Cout);
        const int tap = (int)((e / Cout) % (TH * TW));
        ci = (int)(e / ((long long)Cout * TH * TW));
        jy = tap / TW; jx = tap % TW;
    } else if (mode == 0) {  // out[((jy*TW+jx)*Cin + ci)*Cout + co]
        co = (int)(e % Cout);
        ci = (int)((e / Cout) % Cin);
        const int tap = (int)(e / ((long long)Cout * Cin));
        jy = tap / TW; jx = tap % TW;
    } else {          // out[((jy*TW+jx)*Cout + co)*Cin + ci]
        ci = (int)(e % Cin);
        co = (int)((e / Cin) % Cout);
        const int tap = (int)(e / ((long long)Cout * Cin));
        jy = tap / TW; jx = tap % TW;
    }
    const int ky = ky0 + kstep * jy, kx = kx0 + kstep * jx;
    // modes 2 / 3 are tensor-core operands: stored RN-rounded to TF32 (the MMA would truncate, common.cuh: tf32_rna)
    out[e] = sdt::out_round(w[(((size_t)co * Cin + ci) * KH + ky) * KW + kx], mode >= 2);
}

// ------------------------------------------------------------------------------------------------
// wgrad for tiny contractions (K = TH*TW*C <= 16, e.g. the 1->64 3x3 first encoder layer, K = 9): a 128-wide GEMM tile
// would be >90 % padding, so this is a streaming kernel instead: each CTA walks a contiguous pixel range, stages 64
// pixels of dy (64 x N) and of the im2col patch (64 x 16) in shared memory, thread (q, n) accumulates the 16 taps of
// output channel n over a quarter of the pixels; one partial (N x K) per CTA, reduced by wgrad_reduce_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int SK_PIX = 64;
__global__ void __launch_bounds__(256) conv_wgrad_smallk_kernel(const sdt_conv_desc d) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ __align__(16) float dy_s[SK_PIX][64 + 1];
    __shared__ __align__(16) float a_s[SK_PIX][16];
    __shared__ float red[4][64][16 + 1];
    const int tid = threadIdx.x;
    const int Kc = d.TH * d.TW * d.C, N = d.N, P = d.GH * d.GW;
    const long long Mtot = (long long)d.B * P;
    long long chunk = (Mtot + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + SK_PIX - 1) / SK_PIX * SK_PIX;
    const long long p_begin = (long long)blockIdx.x * chunk;
    const long long p_end = p_begin + chunk < Mtot ? p_begin + chunk : Mtot;
    const int n = tid & 63, q = tid >> 6;
    const bool has_xf = d.xf_scale != nullptr;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    for (long long pb = p_begin; pb < p_end; pb += SK_PIX) {
        // stage dy: 64 pixels x N (N <= 64), coalesced
        for (int e = tid; e < SK_PIX * 64; e += 256) {
            const int pp = e >> 6, nn = e & 63;
            const long long pix = pb + pp;
            dy_s[pp][nn] = (pix < p_end && nn < N) ? __ldg(d.dy + pix * N + nn) : 0.f;
        }
        // stage the patches: 64 pixels x 16 (k >= Kc -> 0)
        for (int e = tid; e < SK_PIX * 16; e += 256) {
            const int pp = e >> 4, k = e & 15;
            const long long pix = pb + pp;
            float v = 0.f;
            if (pix < p_end && k < Kc) {
                const int b = (int)(pix / P);
                const int rem = (int)(pix - (long long)b * P);
                const int gy = rem / d.GW, gx = rem - gy * d.GW;
                const int tap = k / d.C, c = k - tap * d.C;
                const int tyy = tap / d.TW, txx = tap - tyy * d.TW;
                const int sy = gy * d.y_mul + d.y_off + tyy * d.ty_mul, sx = gx * d.x_mul + d.x_off + txx * d.tx_mul;
                if (sy >= 0 && sy < d.SH && sx >= 0 && sx < d.SW) {
                    v = __ldg(d.src + (((size_t)b * d.SH + sy) * d.SW + sx) * d.C + c);
                    if (has_xf) {
                        const size_t o = (size_t)b * d.xf_bstride + c;
                        v = xf_apply(v, __ldg(d.xf_scale + o), __ldg(d.xf_shift + o), d.xf_slope);
                    }
                }
            }
            a_s[pp][k] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int pp = q; pp < SK_PIX; pp += 4) {
            const float g = dy_s[pp][n];
            const float4 a0 = *reinterpret_cast<const float4*>(&a_s[pp][0]);
            const float4 a1 = *reinterpret_cast<const float4*>(&a_s[pp][4]);
            const float4 a2 = *reinterpret_cast<const float4*>(&a_s[pp][8]);
            const float4 a3 = *reinterpret_cast<const float4*>(&a_s[pp][12]);
            acc[0] = fmaf(g, a0.x, acc[0]); acc[1] = fmaf(g, a0.y, acc[1]); acc[2] = fmaf(g, a0.z, acc[2]); acc[3] = fmaf(g, a0.w, acc[3]);
            acc[4] = fmaf(g, a1.x, acc[4]); acc[5] = fmaf(g, a1.y, acc[5]); acc[6] = fmaf(g, a1.z, acc[6]); acc[7] = fmaf(g, a1.w, acc[7]);
            acc[8] = fmaf(g, a2.x, acc[8]); acc[9] = fmaf(g, a2.y, acc[9]); acc[10] = fmaf(g, a2.z, acc[10]); acc[11] = fmaf(g, a2.w, acc[11]);
            acc[12] = fmaf(g, a3.x, acc[12]); acc[13] = fmaf(g, a3.y, acc[13]); acc[14] = fmaf(g, a3.z, acc[14]); acc[15] = fmaf(g, a3.w, acc[15]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) red[q][n][k] = acc[k];
    __syncthreads();
    float* out = d.wpart + (size_t)blockIdx.x * N * Kc;
    for (int e = tid; e < N * Kc; e += 256) {
        const int nn = e / Kc, k = e - nn * Kc;
        out[e] = ((red[0][nn][k] + red[1][nn][k]) + red[2][nn][k]) + red[3][nn][k];
    }
}

struct PrepItem {     // mirrors sdt_prep_item
    const float* w;
    float* out;
    int32_t Cout, Cin, KH, KW, mode, ky0, kx0, kstep, TH, TW, pad0, pad1;
};

__global__ void weight_prep_batch_kernel(const PrepItem* __restrict__ items) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const PrepItem it = items[blockIdx.y];
    // 32-bit index arithmetic (an operand has < 2^31 elements; the 64-bit divisions of the first version were most of its time)
    const unsigned Cin = it.Cin, Cout = it.Cout, T = it.TH * it.TW, TW = it.TW;
    const unsigned cin_p = (it.mode == 2 && it.pad0 > it.Cin) ? it.pad0 : Cin;      // mode 2: input channels zero-padded to pad0
    const unsigned total = T * cin_p * Cout;
    const unsigned KHW = it.KH * it.KW;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        unsigned co, ci, tap;
        if (it.mode == 2) {
            const unsigned q = e / cin_p;
            ci = e - q * cin_p; co = q / T; tap = q - co * T;
            if (ci >= Cin) { it.out[e] = 0.f; continue; }
        }
        else if (it.mode == 3) { const unsigned q = e / Cout; co = e - q * Cout; ci = q / T; tap = q - ci * T; }
        else if (it.mode == 0) { const unsigned q = e / Cout; co = e - q * Cout; tap = q / Cin; ci = q - tap * Cin; }
        else { const unsigned q = e / Cin; ci = e - q * Cin; tap = q / Cout; co = q - tap * Cout; }
        const unsigned jy = tap / TW, jx = tap - jy * TW;
        const unsigned ky = it.ky0 + it.kstep * jy, kx = it.kx0 + it.kstep * jx;
        it.out[e] = sdt::out_round(__ldg(it.w + ((size_t)co * Cin + ci) * KHW + ky * it.KW + kx), it.mode >= 2);
    }
}

int check_desc(const sdt_conv_desc* d, const char* who) {
    SDT_REQUIRE(d != nullptr, "%s: null descriptor", who);
    SDT_REQUIRE(d->src != nullptr, "%s: null src", who);
    SDT_REQUIRE(d->B > 0 && d->SH > 0 && d->SW > 0 && d->C > 0 && d->GH > 0 && d->GW > 0 && d->TH > 0 && d->TW > 0 && d->N > 0,
                "%s: non-positive extent (B=%d SH=%d SW=%d C=%d GH=%d GW=%d TH=%d TW=%d N=%d)", who, d->B, d->SH, d->SW,
                d->C, d->GH, d->GW, d->TH, d->TW, d->N);
    SDT_REQUIRE((d->xf_scale == nullptr) == (d->xf_shift == nullptr), "%s: xf_scale and xf_shift must come together", who);
    SDT_REQUIRE(d->xf_scale == nullptr || d->xf_bstride == 0 || d->xf_bstride == d->C, "%s: xf_bstride must be 0 or C", who);
    SDT_REQUIRE(d->math >= 0 && d->math <= 5, "%s: math must be 0 (process default) or mode + 1 (1..5), got %d", who, d->math);
    return SDT_OK;
}

inline int row_tiles_for(const sdt_conv_desc* d, int BM) {
    const long long P = (long long)d->GH * d->GW;
    if (d->per_image_tiles) return (int)(d->B * ((P + BM - 1) / BM));
    return (int)((d->B * P + BM - 1) / BM);
}

// tile choice shared by sdt_conv_row_tiles and sdt_conv_gemm
inline int pick_bm(const sdt_conv_desc* d) {
    const long long M = (long long)d->B * d->GH * d->GW;
    // small problems (1-D stacks): 64x64 tiles to get more CTAs in flight
    if (!d->per_image_tiles && M * d->N <= (long long)128 * 128 * 148) return 64;
    return 128;
}

}  // namespace

// 1-D layers arrive as B images of one row (B,1,L,C).  The persistent kernel tiles (rows x columns) patches inside ONE image, so
// for it the batch becomes the image height: (B,1,L,C) and (1,B,L,C) are the same memory, a patch of bh clips x bw positions
// fills all 128 GEMM rows (a per-clip patch of a 64-frame sequence fills half), and zero padding along the sequence is still
// TMA's out-of-bounds fill in x.  There are no vertical taps (TH == 1), so rows never mix clips.
// Measured (B = 32): the 36 small 1-D launches of a step are 0.1 ms SLOWER this way than on tc_conv_tma.cu (a persistent CTA with
// the whole shared memory and 512 TMEM columns per SM is a heavy vehicle for 16-64 tiles), so it is off unless SDT_REMAP_1D=1.
// math mode of one problem: its own (`math` = mode + 1) or the process default
inline int mode_of(const sdt_conv_desc* d) { return d->math > 0 ? d->math - 1 : sdt_get_conv_math(); }

inline bool remap_1d(const sdt_conv_desc* d, sdt_conv_desc* out) {
    static const bool on = getenv("SDT_REMAP_1D") != nullptr && getenv("SDT_REMAP_1D")[0] == '1';
    if (!on || mode_of(d) < 3) return false;
    if (!(d->SH == 1 && d->GH == 1 && d->DH == 1 && d->TH == 1 && d->B > 1 && d->per_image_tiles == 0)) return false;
    *out = *d;
    out->SH = out->GH = out->DH = d->B;
    out->B = 1;
    out->y_mul = 1;
    out->ty_mul = 1;
    out->y_off = 0;
    out->dy_mul = 1;
    out->dy_off = 0;
    return true;
}

inline bool use_pair(const sdt_conv_desc* d) { return mode_of(d) == 4 && sdt_tc_conv_pair_eligible(d); }
inline bool use_ytap(const sdt_conv_desc* d) { return mode_of(d) >= 3 && sdt_tc_conv_ytap_eligible(d); }
inline bool use_tma(const sdt_conv_desc* d) { return mode_of(d) >= 2 && sdt_tc_conv_tma_eligible(d); }
inline bool use_tc(const sdt_conv_desc* d) { return mode_of(d) >= 1 && sdt_tc_conv_eligible(d); }

extern "C" int sdt_conv_row_tiles(const sdt_conv_desc* d) {
    if (check_desc(d, "sdt_conv_row_tiles") != SDT_OK) return -1;
    sdt_conv_desc r;
    if (remap_1d(d, &r) && use_ytap(&r)) return sdt_tc_conv_ytap_row_tiles(&r);
    if (use_pair(d)) return sdt_tc_conv_pair_row_tiles(d);
    if (use_ytap(d)) return sdt_tc_conv_ytap_row_tiles(d);
    if (use_tma(d)) return sdt_tc_conv_tma_row_tiles(d);
    return row_tiles_for(d, use_tc(d) ? 128 : pick_bm(d));
}

extern "C" int sdt_conv_rownorm_ok(const sdt_conv_desc* d) {
    if (check_desc(d, "sdt_conv_rownorm_ok") != SDT_OK) return 0;
    if (d->GH != 1 || d->SH != 1) return 0;                 // 1-D layers (the 2-D encoder normalises per image, not per row)
    return (use_tma(d) && sdt_tc_conv_tma_rownorm_ok(d)) ? 1 : 0;
}

extern "C" int sdt_conv_plan(const sdt_conv_desc* d, int32_t* out10) {
    if (int rc = check_desc(d, "sdt_conv_plan")) return rc;
    SDT_REQUIRE(out10 != nullptr, "sdt_conv_plan: null output");
    for (int i = 0; i < 10; ++i) out10[i] = 0;
    sdt_conv_desc r1;
    if (remap_1d(d, &r1) && sdt_tc_conv_ytap_shape_ok(&r1)) {
        out10[0] = 3;
        sdt_tc_conv_ytap_describe(&r1, out10);
    } else if (mode_of(d) == 4 && sdt_tc_conv_pair_shape_ok(d)) {
        out10[0] = 4;
        sdt_tc_conv_pair_describe(d, out10);
    } else if (mode_of(d) >= 3 && sdt_tc_conv_ytap_shape_ok(d)) {
        out10[0] = 3;
        sdt_tc_conv_ytap_describe(d, out10);
    } else if (use_tma(d)) {
        out10[0] = 2;
    } else if (use_tc(d)) {
        out10[0] = 1;
    }
    return SDT_OK;
}

extern "C" int sdt_conv_gemm(const sdt_conv_desc* d, void* stream) {
    if (int rc = check_desc(d, "sdt_conv_gemm")) return rc;
    SDT_REQUIRE(d->dst && (d->wt || d->wt_nk), "sdt_conv_gemm: null dst or no weight operand");
    SDT_REQUIRE(!(d->stat_partial && d->bias), "sdt_conv_gemm: statistics epilogue excludes bias");
    SDT_REQUIRE(!(d->stat_partial && d->accumulate), "sdt_conv_gemm: statistics epilogue excludes accumulate");
    cudaStream_t st = sdt::as_stream(stream);
    if (d->rn_act != nullptr) {                                                 // fused row-norm epilogue: TMA kernel only
        SDT_REQUIRE(sdt_conv_rownorm_ok(d), "sdt_conv_gemm: rn_act set but sdt_conv_rownorm_ok() is 0 for this problem");
        return sdt_tc_conv_tma_launch(d, st);
    }
    sdt_conv_desc r1;
    if (remap_1d(d, &r1) && use_ytap(&r1)) return sdt_tc_conv_ytap_launch(&r1, st);   // 1-D layers on the persistent kernel
    if (use_pair(d)) return sdt_tc_conv_pair_launch(d, st);                     // math mode 4: CTA pairs (cta_group::2), experimental
    if (use_ytap(d)) return sdt_tc_conv_ytap_launch(d, st);                     // math mode 3: + operand reuse in shared memory
    if (use_tma(d)) return sdt_tc_conv_tma_launch(d, st);                       // math mode 2: tcgen05 TF32, TMA operands
    if (use_tc(d)) return sdt_tc_conv_launch(d, row_tiles_for(d, 128), st);   // math mode 1/2: tcgen05 TF32
    SDT_REQUIRE(d->wt != nullptr, "sdt_conv_gemm: the FFMA path needs the (K,N) operand `wt` (tcgen05 path not eligible here)");
    const bool vec = (d->C % 4) == 0;
    const int bm = pick_bm(d);
    const sdt_conv_desc dd = *d;
    if (bm == 64) {
        dim3 grid(row_tiles_for(d, 64), sdt::ceil_div(d->N, 64));
        if (vec) sdt::launch(conv_gemm_kernel<64, 64, 4, 4, true>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else sdt::launch(conv_gemm_kernel<64, 64, 4, 4, false>, dim3(grid), dim3(NTHREADS), 0, st, dd);
    } else if (d->N <= 64) {
        dim3 grid(row_tiles_for(d, 128), sdt::ceil_div(d->N, 64));
        if (vec) sdt::launch(conv_gemm_kernel<128, 64, 8, 4, true>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else sdt::launch(conv_gemm_kernel<128, 64, 8, 4, false>, dim3(grid), dim3(NTHREADS), 0, st, dd);
    } else {
        dim3 grid(row_tiles_for(d, 128), sdt::ceil_div(d->N, 128));
        if (vec) sdt::launch(conv_gemm_kernel<128, 128, 8, 8, true>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else sdt::launch(conv_gemm_kernel<128, 128, 8, 8, false>, dim3(grid), dim3(NTHREADS), 0, st, dd);
    }
    SDT_LAUNCH_OK("conv_gemm_kernel");
    return SDT_OK;
}

// n independent convolution problems in stream order.  In math mode >= 3, problems that differ in their grids and offsets
// only (the stride-parity classes of one data gradient) run as ONE persistent launch; anything else is n launches.
// returns in *launched (may be NULL) the number of kernels launched
extern "C" int sdt_conv_gemm_multi(const sdt_conv_desc* descs, int n, void* stream, int* launched) {
    SDT_REQUIRE(descs != nullptr && n >= 1 && n <= 16, "sdt_conv_gemm_multi: %d problems (1..16)", n);
    if (launched) *launched = 0;
    if (mode_of(descs) == 3 && n >= 2 && n <= 4) {
        bool ok = true;
        for (int c = 0; c < n && ok; ++c) {
            if (int rc = check_desc(descs + c, "sdt_conv_gemm_multi")) return rc;
            ok = descs[c].dst && descs[c].wt_nk && descs[c].math == descs[0].math && use_ytap(descs + c);
        }
        if (ok && sdt_tc_conv_ytap_multi_ok(descs, n)) {
            if (launched) *launched = 1;
            return sdt_tc_conv_ytap_launch_multi(descs, n, sdt::as_stream(stream));
        }
    }
    for (int c = 0; c < n; ++c) {
        if (int rc = sdt_conv_gemm(descs + c, stream)) return rc;
        if (launched) ++*launched;
    }
    return SDT_OK;
}

extern "C" int sdt_conv_wgrad(const sdt_conv_desc* d, void* stream) {
    if (int rc = check_desc(d, "sdt_conv_wgrad")) return rc;
    SDT_REQUIRE(d->dy && d->wpart, "sdt_conv_wgrad: null dy/wpart");
    SDT_REQUIRE(d->splits >= 1 && d->splits <= 65535, "sdt_conv_wgrad: splits=%d out of range", d->splits);
    const int Kc = d->TH * d->TW * d->C;
    const bool veca = (d->N % 4) == 0, vecb = (d->C % 4) == 0;
    cudaStream_t st = sdt::as_stream(stream);
    if (mode_of(d) >= 3 && sdt_tc_wgrad_ytap_eligible(d)) return sdt_tc_wgrad_ytap_launch(d, st);   // + vertical-tap reuse
    if (mode_of(d) >= 2 && sdt_tc_wgrad_tma_eligible(d)) return sdt_tc_wgrad_tma_launch(d, st);   // tcgen05 + TMA
    if (mode_of(d) >= 1 && sdt_tc_wgrad_eligible(d)) return sdt_tc_wgrad_launch(d, st);   // tcgen05 TF32
    if (Kc <= 16 && d->N <= 64) {       // tiny contraction: streaming kernel, one partial per CTA (gridDim.x == splits)
        sdt::launch(conv_wgrad_smallk_kernel, dim3(d->splits), dim3(256), 0, st, *d);
        SDT_LAUNCH_OK("conv_wgrad_smallk_kernel");
        return SDT_OK;
    }
    const sdt_conv_desc dd = *d;
    if (d->N <= 64) {
        dim3 grid(sdt::ceil_div(Kc, 128), sdt::ceil_div(d->N, 64), d->splits);
        if (veca && vecb) sdt::launch(conv_wgrad_kernel<64, 4, true, true>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else if (veca) sdt::launch(conv_wgrad_kernel<64, 4, true, false>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else if (vecb) sdt::launch(conv_wgrad_kernel<64, 4, false, true>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else sdt::launch(conv_wgrad_kernel<64, 4, false, false>, dim3(grid), dim3(NTHREADS), 0, st, dd);
    } else {
        dim3 grid(sdt::ceil_div(Kc, 128), sdt::ceil_div(d->N, 128), d->splits);
        if (veca && vecb) sdt::launch(conv_wgrad_kernel<128, 8, true, true>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else if (veca) sdt::launch(conv_wgrad_kernel<128, 8, true, false>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else if (vecb) sdt::launch(conv_wgrad_kernel<128, 8, false, true>, dim3(grid), dim3(NTHREADS), 0, st, dd);
        else sdt::launch(conv_wgrad_kernel<128, 8, false, false>, dim3(grid), dim3(NTHREADS), 0, st, dd);
    }
    SDT_LAUNCH_OK("conv_wgrad_kernel");
    return SDT_OK;
}

extern "C" int sdt_conv_wgrad_reduce(const float* wpart, int splits, int N, int C, int T, float* grad, int accumulate,
                                     void* stream) {
    SDT_REQUIRE(wpart && grad && splits >= 1 && N > 0 && C > 0 && T > 0, "sdt_conv_wgrad_reduce: bad arguments");
    const long long total = (long long)N * C * T;
    if (C % 32 == 0 && T <= 256)
        sdt::launch(wgrad_reduce_tiled_kernel, dim3(N * (C / 32)), dim3(256), (size_t)T * 33 * sizeof(float), sdt::as_stream(stream), wpart, splits, N, C, T, grad, accumulate);
    else
        sdt::launch(wgrad_reduce_kernel, dim3(sdt::ceil_div(total, 256)), dim3(256), 0, sdt::as_stream(stream), wpart, splits, N, C, T, grad, accumulate);
    SDT_LAUNCH_OK("wgrad_reduce_kernel");
    return SDT_OK;
}

extern "C" int sdt_conv_wgrad_reduce_batch(const sdt_reduce_item* items_device, int n_items, int max_ctas, int max_T, void* stream) {
    SDT_REQUIRE(items_device && n_items > 0 && max_ctas > 0, "sdt_conv_wgrad_reduce_batch: bad arguments");
    SDT_REQUIRE(max_T > 0 && max_T <= 256, "sdt_conv_wgrad_reduce_batch: taps per item must be in [1, 256] (max_T=%d)", max_T);
    static_assert(sizeof(ReduceItem) == sizeof(sdt_reduce_item), "sdt_reduce_item layout");
    sdt::launch(wgrad_reduce_batch_kernel, dim3(max_ctas, n_items), dim3(256), (size_t)max_T * 33 * sizeof(float), sdt::as_stream(stream),
                reinterpret_cast<const ReduceItem*>(items_device));
    SDT_LAUNCH_OK("wgrad_reduce_batch_kernel");
    return SDT_OK;
}

extern "C" int sdt_weight_prep(const float* w, int Cout, int Cin, int KH, int KW, int mode, int ky0, int kx0, int kstep,
                               int TH, int TW, float* out, void* stream) {
    SDT_REQUIRE(w && out, "sdt_weight_prep: null pointer");
    SDT_REQUIRE(Cout > 0 && Cin > 0 && KH > 0 && KW > 0 && TH > 0 && TW > 0 && kstep > 0, "sdt_weight_prep: bad extents");
    SDT_REQUIRE(mode >= 0 && mode <= 3, "sdt_weight_prep: mode must be 0..3");
    SDT_REQUIRE(ky0 >= 0 && kx0 >= 0 && ky0 + kstep * (TH - 1) < KH && kx0 + kstep * (TW - 1) < KW,
                "sdt_weight_prep: tap selection outside the %dx%d kernel", KH, KW);
    const long long total = (long long)TH * TW * Cin * Cout;
    sdt::launch(weight_prep_kernel, dim3(sdt::ceil_div(total, 256)), dim3(256), 0, sdt::as_stream(stream), w, Cout, Cin, KH, KW, mode, ky0, kx0,
                                                                                       kstep, TH, TW, out);
    SDT_LAUNCH_OK("weight_prep_kernel");
    return SDT_OK;
}

extern "C" int sdt_weight_prep_batch(const sdt_prep_item* items_device, int n_items, long long max_elems, void* stream) {
    SDT_REQUIRE(items_device && n_items > 0 && max_elems > 0, "sdt_weight_prep_batch: bad arguments");
    static_assert(sizeof(PrepItem) == sizeof(sdt_prep_item), "sdt_prep_item layout");
    int gx = sdt::ceil_div(max_elems, 256 * 4);
    if (gx > 1024) gx = 1024;
    sdt::launch(weight_prep_batch_kernel, dim3(gx, n_items), dim3(256), 0, sdt::as_stream(stream), reinterpret_cast<const PrepItem*>(items_device));
    SDT_LAUNCH_OK("weight_prep_batch_kernel");
    return SDT_OK;
}
