// First block of the audio encoder: Conv2d(1 -> 64, 3x3, stride 1, pad 1, no bias) + InstanceNorm2d + LeakyReLU
// (reference: core/networks/keypoints_generation/generator.py:17 through building_blocks.py ConvNormRelu).
//
// It is the largest map of the network (B x 80 x T x 64 fp32 = 280 MB at B = 32) and has ONE input channel, so it is
// pure HBM traffic.  With a single input channel the convolution is a 9-tap linear map of the mel image, which lets
// both directions run in a single pass over the big map:
//   forward : the per-(image, channel) InstanceNorm statistics follow from the 9 tap means and the 9x9 tap second
//             moments of the image (54 numbers per image, accumulated in double):  mean_c = w_c . mu,
//             var_c = w_c^T (M - mu mu^T) w_c.  One kernel then writes LeakyReLU((conv - mean) * rstd) directly; the raw
//             convolution output is never stored.
//   backward: with g_hat = g * LeakyReLU'(act) and xhat recovered from act (LeakyReLU with slope > 0 is invertible),
//             dW[c,t] = sum_b rstd_bc * ( G_t - S1 * mu_t - S2 * (rstd_bc * (M w_c)_t + shift_bc * mu_t) ),
//             S1 = sum g_hat, S2 = sum g_hat * xhat, G_t = sum g_hat * tap_t  -- 11 sums per (image, channel) from one
//             pass over (g, act); the input gradient of the normalisation is never materialised (mel needs no gradient).
#include "common.cuh"

namespace {

constexpr int kTaps = 9;
constexpr int kMom = kTaps + kTaps * (kTaps + 1) / 2;   // 9 means + 45 upper-triangle second moments
constexpr int kC = 64;                                  // output channels of the block
constexpr int kBwdQ = 2 + kTaps;                        // S1, S2, G_t

__host__ __device__ inline int tri(int a, int b) {      // index of (a <= b) in the packed upper triangle
    return a * kTaps - a * (a - 1) / 2 + (b - a);
}

constexpr int kChunk = 512;                             // pixels of one image row handled by a block
constexpr int kPitch = kChunk + 2;

// A block works on the unit (row y, columns [x0, x0 + kChunk)) of image b; blockIdx.x = y * chunks + chunk.
struct Unit {
    int y, x0, n;
};
__device__ __forceinline__ Unit unit_of_block(int W) {
    const int chunks = (W + kChunk - 1) / kChunk;
    Unit u;
    u.y = blockIdx.x / chunks;
    u.x0 = (blockIdx.x - u.y * chunks) * kChunk;
    u.n = min(kChunk, W - u.x0);
    return u;
}

// rows y-1, y, y+1 of image b, columns x0-1 .. x0+kChunk, zero outside the image: s[r * kPitch + 1 + (x - x0)]
__device__ __forceinline__ void stage_rows(const float* __restrict__ x, int b, const Unit& u, int H, int W, float* s) {
    for (int i = threadIdx.x; i < 3 * kPitch; i += blockDim.x) {
        const int r = i / kPitch, xx = u.x0 + (i - r * kPitch) - 1;
        const int yy = u.y + r - 1;
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((size_t)b * H + yy) * W + xx];
        s[i] = v;
    }
}

__global__ void __launch_bounds__(128) fl_moments_kernel(const float* __restrict__ x, int H, int W, double* __restrict__ partial) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ float s_rows[3 * kPitch];
    __shared__ double s_red[4][kMom];
    const int b = blockIdx.y;
    const Unit u = unit_of_block(W);
    stage_rows(x, b, u, H, W, s_rows);
    __syncthreads();
    const int pitch = kPitch;
    double acc[kMom];
#pragma unroll
    for (int i = 0; i < kMom; ++i) acc[i] = 0.0;
    for (int px = threadIdx.x; px < u.n; px += blockDim.x) {
        float t[kTaps];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) t[ky * 3 + kx] = s_rows[ky * pitch + px + kx];
#pragma unroll
        for (int a = 0; a < kTaps; ++a) {
            acc[a] += (double)t[a];
#pragma unroll
            for (int c = a; c < kTaps; ++c) acc[kTaps + tri(a, c)] += (double)t[a] * (double)t[c];
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < kMom; ++i) {
        const double v = sdt::warp_sum_d(acc[i]);
        if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kMom)
        partial[((size_t)b * gridDim.x + blockIdx.x) * kMom + threadIdx.x] =
            s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x];
}

__global__ void __launch_bounds__(128) fl_stats_kernel(const double* __restrict__ partial, const float* __restrict__ w, int units,
                                                       double count, float eps, double* __restrict__ moments,
                                                       float* __restrict__ scale, float* __restrict__ shift) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ double s_m[kMom];
    const int b = blockIdx.x;
    if (threadIdx.x < kMom) {
        double s = 0.0;
        for (int i = 0; i < units; ++i) s += partial[((size_t)b * units + i) * kMom + threadIdx.x];
        s /= count;
        s_m[threadIdx.x] = s;
        moments[(size_t)b * kMom + threadIdx.x] = s;
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c >= kC) return;
    double wc[kTaps], mean = 0.0;
#pragma unroll
    for (int t = 0; t < kTaps; ++t) {
        wc[t] = (double)w[c * kTaps + t];
        mean += wc[t] * s_m[t];
    }
    double var = 0.0;
#pragma unroll
    for (int a = 0; a < kTaps; ++a)
#pragma unroll
        for (int d = 0; d < kTaps; ++d) {
            const double cov = s_m[kTaps + (a <= d ? tri(a, d) : tri(d, a))] - s_m[a] * s_m[d];
            var += wc[a] * wc[d] * cov;
        }
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);       // biased variance, InstanceNorm2d
    scale[(size_t)b * kC + c] = (float)rstd;
    shift[(size_t)b * kC + c] = (float)(-mean * rstd);
}

// 256 threads: 16 channel quads x 16 pixel lanes; a warp stores 2 adjacent pixels = 512 contiguous bytes
__global__ void __launch_bounds__(256) fl_act_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ scale, const float* __restrict__ shift, int H,
                                                     int W, float slope, float* __restrict__ act, int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ float s_rows[3 * kPitch];
    const int b = blockIdx.y;
    const Unit u = unit_of_block(W);
    stage_rows(x, b, u, H, W, s_rows);
    const int cg = threadIdx.x & 15, pl = threadIdx.x >> 4;
    float wq[4][kTaps], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = cg * 4 + j;
        const float sc = scale[(size_t)b * kC + c];
        sh[j] = shift[(size_t)b * kC + c];
#pragma unroll
        for (int t = 0; t < kTaps; ++t) wq[j][t] = w[c * kTaps + t] * sc;
    }
    __syncthreads();
    const int pitch = kPitch;
    float4* out = reinterpret_cast<float4*>(act + (((size_t)b * H + u.y) * W + u.x0) * kC) + cg;
    for (int px = pl; px < u.n; px += 16) {
        float t[kTaps];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) t[ky * 3 + kx] = s_rows[ky * pitch + px + kx];
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = sh[j];
#pragma unroll
            for (int k = 0; k < kTaps; ++k) a = fmaf(wq[j][k], t[k], a);
            v[j] = sdt::out_round(sdt::leaky(a, slope), out_tf32);
        }
        out[(size_t)px * (kC / 4)] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// 256 threads: 32 channel pairs x 8 pixel lanes; one image row per block.  The pre-activation xhat = conv(x) * rstd - mean * rstd
// is RECOMPUTED from the staged input rows with exactly the forward's arithmetic (fl_act_kernel: shift + sum_k (w_k * rstd) * x_k,
// same order) instead of being recovered from the stored activation: the pass reads g only (280 MB at B = 32) instead of g + act
// (560 MB), and xhat is the unrounded fp32 value rather than the TF32-rounded one the activation map holds.
__global__ void __launch_bounds__(256) fl_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                     const float* __restrict__ w, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, int H, int W, float slope,
                                                     float* __restrict__ partial) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ float s_rows[3 * kPitch];
    __shared__ float s_red[kBwdQ][8][kC];
    const int b = blockIdx.y;
    const Unit u = unit_of_block(W);
    stage_rows(x, b, u, H, W, s_rows);
    const int cp = threadIdx.x & 31, pl = threadIdx.x >> 5;
    float wq[2][kTaps], sh[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int c = cp * 2 + j;
        const float sc = scale[(size_t)b * kC + c];
        sh[j] = shift[(size_t)b * kC + c];
#pragma unroll
        for (int t = 0; t < kTaps; ++t) wq[j][t] = w[c * kTaps + t] * sc;
    }
    __syncthreads();
    const int pitch = kPitch;
    float acc[2][kBwdQ];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < kBwdQ; ++q) acc[j][q] = 0.f;
    const float2* gp = reinterpret_cast<const float2*>(g + (((size_t)b * H + u.y) * W + u.x0) * kC) + cp;
#pragma unroll 4
    for (int px = pl; px < u.n; px += 8) {
        const float2 gv = __ldg(gp + (size_t)px * (kC / 2));
        float t[kTaps];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) t[ky * 3 + kx] = s_rows[ky * pitch + px + kx];
        const float gg[2] = {gv.x, gv.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float xh = sh[j];
#pragma unroll
            for (int k = 0; k < kTaps; ++k) xh = fmaf(wq[j][k], t[k], xh);
            const float gh = xh > 0.f ? gg[j] : gg[j] * slope;
            acc[j][0] += gh;
            acc[j][1] = fmaf(gh, xh, acc[j][1]);
#pragma unroll
            for (int k = 0; k < kTaps; ++k) acc[j][2 + k] = fmaf(gh, t[k], acc[j][2 + k]);
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < kBwdQ; ++q) s_red[q][pl][cp * 2 + j] = acc[j][q];
    __syncthreads();
    for (int o = threadIdx.x; o < kBwdQ * kC; o += blockDim.x) {
        const int q = o / kC, c = o - q * kC;
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l) s += s_red[q][l][c];
        partial[(((size_t)b * gridDim.x + blockIdx.x) * kBwdQ + q) * kC + c] = s;
    }
}

// one block per image, one thread per (q, c): the sum over the image's units in double (coalesced: consecutive threads read
// consecutive floats), stored IN PLACE as a float-float pair in the image's unit rows 0 (high part) and 1 (low part).
// (The finalize kernel used to walk the units itself, one 4-byte read per 2.8 KB stride and thread: 85 us.)
__global__ void __launch_bounds__(kBwdQ * kC) fl_bwd_colsum_kernel(float* __restrict__ partial, int units) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int b = blockIdx.x, t = threadIdx.x;
    float* base = partial + (size_t)b * units * kBwdQ * kC;
    double s = 0.0;
#pragma unroll 8
    for (int i = 0; i < units; ++i) s += (double)base[(size_t)i * kBwdQ * kC + t];
    __syncthreads();                            // rows 0 and 1 have been read by everybody
    const float hi = (float)s;
    base[t] = hi;
    base[kBwdQ * kC + t] = (float)(s - (double)hi);
}

// one block per output channel: the closed-form combination over the batch
__global__ void __launch_bounds__(256) fl_bwd_finalize_kernel(const float* __restrict__ partial, const double* __restrict__ moments,
                                                              const float* __restrict__ w, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int B, int units,
                                                              float* __restrict__ dw) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    extern __shared__ double s_sum[];          // [B][kBwdQ]
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < B * kBwdQ; i += blockDim.x) {
        const int b = i / kBwdQ, q = i - b * kBwdQ;
        const float* row = partial + ((size_t)b * units * kBwdQ + q) * kC + c;
        double s;
        if (units >= 2) {                       // reduced by fl_bwd_colsum_kernel
            s = (double)row[0] + (double)row[(size_t)kBwdQ * kC];
        } else {
            s = (double)row[0];
        }
        s_sum[i] = s;
    }
    __syncthreads();
    // thread (b, t): the closed form of image b for tap t; then a fixed-order sum over the batch (deterministic)
    double wc[kTaps];
#pragma unroll
    for (int k = 0; k < kTaps; ++k) wc[k] = (double)w[c * kTaps + k];
    double* s_out = s_sum + (size_t)B * kBwdQ;          // [B][kTaps]
    for (int i = threadIdx.x; i < B * kTaps; i += blockDim.x) {
        const int b = i / kTaps, t = i - b * kTaps;
        const double* m = moments + (size_t)b * kMom;
        const double sc = (double)scale[(size_t)b * kC + c], sh = (double)shift[(size_t)b * kC + c];
        const double S1 = s_sum[b * kBwdQ + 0], S2 = s_sum[b * kBwdQ + 1], G = s_sum[b * kBwdQ + 2 + t];
        double mw = 0.0;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) mw += m[kTaps + (t <= k ? tri(t, k) : tri(k, t))] * wc[k];
        s_out[i] = sc * (G - S1 * m[t] - S2 * (sc * mw + sh * m[t]));
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= kTaps) return;
    double out = 0.0;
    for (int b = 0; b < B; ++b) out += s_out[b * kTaps + t];
    dw[c * kTaps + t] = (float)out;
}

}  // namespace

extern "C" int sdt_first_layer_units(int H, int W) { return H * ((W + kChunk - 1) / kChunk); }

extern "C" int sdt_first_layer_fwd(const float* x, const float* w, int B, int H, int W, int C, float eps, float slope,
                                   double* mom_partial, double* moments, float* scale, float* shift, float* act,
                                   int out_tf32, void* stream) {
    SDT_REQUIRE(x && w && mom_partial && moments && scale && shift, "sdt_first_layer_fwd: null pointer");        // act == NULL: statistics only
    SDT_REQUIRE(C == kC, "sdt_first_layer_fwd: the block has %d output channels (got %d)", kC, C);
    SDT_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "sdt_first_layer_fwd: bad extents");
    const int units = sdt_first_layer_units(H, W);
    cudaStream_t st = sdt::as_stream(stream);
    sdt::launch(fl_moments_kernel, dim3(units, B), dim3(128), 0, st, x, H, W, mom_partial);
    SDT_LAUNCH_OK("fl_moments_kernel");
    sdt::launch(fl_stats_kernel, dim3(B), dim3(128), 0, st, mom_partial, w, units, (double)H * W, eps, moments, scale, shift);
    SDT_LAUNCH_OK("fl_stats_kernel");
    if (act == nullptr) return SDT_OK;
    sdt::launch(fl_act_kernel, dim3(units, B), dim3(256), 0, st, x, w, scale, shift, H, W, slope, act, out_tf32);
    SDT_LAUNCH_OK("fl_act_kernel");
    return SDT_OK;
}

extern "C" int sdt_first_layer_act(const float* x, const float* w, const float* scale, const float* shift, int B, int H, int W, int C,
                                   float slope, float* act, int out_tf32, void* stream) {
    SDT_REQUIRE(x && w && scale && shift && act, "sdt_first_layer_act: null pointer");
    SDT_REQUIRE(C == kC, "sdt_first_layer_act: the block has %d output channels (got %d)", kC, C);
    SDT_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "sdt_first_layer_act: bad extents");
    sdt::launch(fl_act_kernel, dim3(sdt_first_layer_units(H, W), B), dim3(256), 0, sdt::as_stream(stream), x, w, scale, shift, H, W, slope, act,
                out_tf32);
    SDT_LAUNCH_OK("fl_act_kernel");
    return SDT_OK;
}

extern "C" int sdt_first_layer_bwd(const float* g_act, const float* act, const float* x, const float* w, const double* moments,
                                   const float* scale, const float* shift, int B, int H, int W, int C, float slope,
                                   float* partial, float* dw, void* stream) {
    (void)act;      // not read any more: the pre-activation is recomputed from x (kept in the signature for ABI stability; may be NULL)
    SDT_REQUIRE(g_act && x && w && moments && scale && shift && partial && dw, "sdt_first_layer_bwd: null pointer");
    SDT_REQUIRE(C == kC, "sdt_first_layer_bwd: the block has %d output channels (got %d)", kC, C);
    SDT_REQUIRE(slope > 0.f, "sdt_first_layer_bwd: needs an invertible activation (slope > 0), got %g", (double)slope);
    SDT_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "sdt_first_layer_bwd: bad extents");
    const int units = sdt_first_layer_units(H, W);
    const size_t fsmem = (size_t)B * (kBwdQ + kTaps) * sizeof(double);
    SDT_REQUIRE(fsmem <= 40 * 1024, "sdt_first_layer_bwd: batch %d too large for the finalize kernel", B);
    cudaStream_t st = sdt::as_stream(stream);
    sdt::launch(fl_bwd_kernel, dim3(units, B), dim3(256), 0, st, g_act, x, w, scale, shift, H, W, slope, partial);
    SDT_LAUNCH_OK("fl_bwd_kernel");
    if (units >= 2) {
        sdt::launch(fl_bwd_colsum_kernel, dim3(B), dim3(kBwdQ * kC), 0, st, partial, units);
        SDT_LAUNCH_OK("fl_bwd_colsum_kernel");
    }
    sdt::launch(fl_bwd_finalize_kernel, dim3(kC), dim3(256), fsmem, st, partial, moments, w, scale, shift, B, units, dw);
    SDT_LAUNCH_OK("fl_bwd_finalize_kernel");
    return SDT_OK;
}
