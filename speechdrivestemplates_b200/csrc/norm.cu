// Normalisation kernels: InstanceNorm2d / BatchNorm{1,2}d statistics finalisation and backward, and the
// channel-LayerNorm that the reference's "InstanceNorm1d on the permuted tensor" amounts to
// (core/networks/building_blocks.py:23-27,38-43,50-54).  All reductions are fixed-order (deterministic).
#include <algorithm>

#include "common.cuh"

namespace {

// ---- statistics finalisation ---------------------------------------------------------------------
// sum of the (tiles, 2, C) partials of one (group, channel): the 8 warps of the CTA split the tiles, lane = channel
// (128-byte coalesced reads), doubles combined through shared memory.  CTA = (group, 32 channels).
__device__ __forceinline__ void sum_partials_32ch(const float* __restrict__ p, int tiles, int C, int c, bool valid,
                                                  double (*s_red)[2][32], double& s, double& q) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double a = 0.0, b = 0.0;
    if (valid)
        for (int t = warp; t < tiles; t += 8) {
            a += (double)__ldg(p + ((size_t)t * 2 + 0) * C + c);
            b += (double)__ldg(p + ((size_t)t * 2 + 1) * C + c);
        }
    s_red[warp][0][lane] = a;
    s_red[warp][1][lane] = b;
    __syncthreads();
    s = 0.0; q = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        s += s_red[w][0][lane];
        q += s_red[w][1][lane];
    }
}

// ---- statistics of a column window of a channels-last map (time-tiled inference) ------------------------------------
// x (B, H, W, C); positions p = h * (w1 - w0) + (w - w0) over the window; part `blockIdx.x` sums rows_per_part positions.
// lane = channel (coalesced), the row groups of the CTA are combined in shared memory in fixed order.
__global__ void chan_stats_kernel(const float* __restrict__ x, int H, int W, int C, int w0, int w1, int rows_per_part,
                                  float* __restrict__ partial, int n_parts) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    extern __shared__ float s_cs[];                     // (rgroups, 2, C)
    const int part = blockIdx.x, b = blockIdx.y;
    const int rgroups = blockDim.x / C, rg = threadIdx.x / C, c = threadIdx.x % C;
    const int ww = w1 - w0;
    const long long P = (long long)H * ww;
    const long long p0 = (long long)part * rows_per_part;
    const long long p1 = p0 + rows_per_part < P ? p0 + rows_per_part : P;
    float s = 0.f, q = 0.f;
    if (rg < rgroups)
        for (long long p = p0 + rg; p < p1; p += rgroups) {
            const int h = (int)(p / ww), w = w0 + (int)(p % ww);
            const float v = __ldg(x + (((size_t)b * H + h) * W + w) * C + c);
            s += v;
            q += v * v;
        }
    if (rg < rgroups) {
        s_cs[(rg * 2 + 0) * C + c] = s;
        s_cs[(rg * 2 + 1) * C + c] = q;
    }
    __syncthreads();
    if (rg == 0) {
        for (int r = 1; r < rgroups; ++r) {
            s += s_cs[(r * 2 + 0) * C + c];
            q += s_cs[(r * 2 + 1) * C + c];
        }
        float* o = partial + ((size_t)b * n_parts + part) * 2 * C;
        o[c] = s;
        o[C + c] = q;
    }
}

__global__ void norm_finalize_kernel(const float* __restrict__ partial, int groups, int tiles_per_group, int C,
                                     double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float eps, float* __restrict__ scale, float* __restrict__ shift,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                     float* __restrict__ running_mean, float* __restrict__ running_var,
                                     int64_t* __restrict__ nbt, float momentum) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ double s_red[8][2][32];
    const int cchunks = (C + 31) / 32;
    const int g = blockIdx.x / cchunks, c = (blockIdx.x % cchunks) * 32 + (threadIdx.x & 31);
    const int e = g * C + c;
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
    double s, q;
    sum_partials_32ch(partial + (size_t)g * tiles_per_group * 2 * C, tiles_per_group, C, c, c < C, s_red, s, q);
    if (threadIdx.x >= 32 || c >= C) return;
    const double mean = s / count;
    double var = q / count - mean * mean;   // biased
    if (var < 0.0) var = 0.0;
    const float meanf = (float)mean;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const float sc = rstd * ga;
    scale[e] = sc;
    shift[e] = be - meanf * sc;
    if (mean_out) mean_out[e] = meanf;
    if (rstd_out) rstd_out[e] = rstd;
    if (running_mean != nullptr && g == 0) {
        const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * meanf;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

__global__ void bn_eval_kernel(const float* __restrict__ rm, const float* __restrict__ rv, const float* __restrict__ gamma,
                               const float* __restrict__ beta, float eps, int C, float* __restrict__ scale,
                               float* __restrict__ shift) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float rstd = 1.0f / sqrtf(rv[c] + eps);
    const float sc = rstd * (gamma ? gamma[c] : 1.f);
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - rm[c] * sc;
}

// ---- backward of normalise -> affine -> activation over (B, P, C) ----------------------------------
// CTA = 256 threads = 8 row-lanes x 32 channel-quads (C handled in chunks of 128), tile of ROWS_PER_TILE pixels.

__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              int P, int C, int groups_is_batch, float slope,
                                                              float* __restrict__ partial, int tiles_per_image) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ float red[2][8][128];
    const int b = blockIdx.x / tiles_per_image, tile = blockIdx.x % tiles_per_image;
    const int cq = threadIdx.x % 32, rl = threadIdx.x / 32;
    const int rows = (P + tiles_per_image - 1) / tiles_per_image;
    const int p0 = tile * rows;
    const int p1 = min(P, p0 + rows);
    for (int cb = 0; cb < C; cb += 128) {
        const int c = cb + cq * 4;
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < C) {
            const int so = (groups_is_batch ? b * C : 0) + c;
            float mu[4], rs[4], ga[4], be[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                mu[q] = mean[so + q]; rs[q] = rstd[so + q];
                ga[q] = gamma ? gamma[c + q] : 1.f; be[q] = beta ? beta[c + q] : 0.f;
            }
            for (int p = p0 + rl; p < p1; p += 8) {
                const size_t o = ((size_t)b * P + p) * C + c;
                const float4 gv = *reinterpret_cast<const float4*>(g + o);
                const float4 xv = *reinterpret_cast<const float4*>(x + o);
                const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float xh = (xx[q] - mu[q]) * rs[q];
                    const float y = fmaf(xh, ga[q], be[q]);
                    const float gp = gg[q] * sdt::leaky_grad(y, slope);
                    s1[q] += gp;
                    s2[q] = fmaf(gp, xh, s2[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            red[0][rl][cq * 4 + q] = s1[q];
            red[1][rl][cq * 4 + q] = s2[q];
        }
        __syncthreads();
        if (threadIdx.x < 128 && cb + threadIdx.x < C) {
            float a = 0.f, bsum = 0.f;
            for (int r = 0; r < 8; ++r) {
                a += red[0][r][threadIdx.x];
                bsum += red[1][r][threadIdx.x];
            }
            partial[((size_t)blockIdx.x * 2 + 0) * C + cb + threadIdx.x] = a;
            partial[((size_t)blockIdx.x * 2 + 1) * C + cb + threadIdx.x] = bsum;
        }
        __syncthreads();
    }
}

__global__ void norm_bwd_finalize_kernel(const float* __restrict__ partial, int groups, int tiles_per_group, int C,
                                         double count, float* __restrict__ m1, float* __restrict__ m2,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ double s_red[8][2][32];
    const int cchunks = (C + 31) / 32;
    const int gidx = blockIdx.x / cchunks, c = (blockIdx.x % cchunks) * 32 + (threadIdx.x & 31);
    const int e = gidx * C + c;
    double s1, s2;
    sum_partials_32ch(partial + (size_t)gidx * tiles_per_group * 2 * C, tiles_per_group, C, c, c < C, s_red, s1, s2);
    if (threadIdx.x >= 32 || c >= C) return;
    m1[e] = (float)(s1 / count);
    m2[e] = (float)(s2 / count);
    if (dgamma != nullptr && gidx == 0) {   // BatchNorm affine gradients: dgamma = sum g'*xhat, dbeta = sum g'
        dgamma[c] = accumulate ? dgamma[c] + (float)s2 : (float)s2;
        dbeta[c] = accumulate ? dbeta[c] + (float)s1 : (float)s1;
    }
}

__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(float* __restrict__ g, const float* __restrict__ x,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ m1, const float* __restrict__ m2,
                                                             long long total4, int P, int C, int groups_is_batch, float slope,
                                                             int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long e4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e4 >= total4) return;
    const long long e = e4 * 4;
    const int c = (int)(e % C);
    const int b = (int)(e / ((long long)P * C));
    const int so = (groups_is_batch ? b * C : 0) + c;
    float4 gv = *reinterpret_cast<const float4*>(g + e);
    const float4 xv = *reinterpret_cast<const float4*>(x + e);
    float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float rs = rstd[so + q];
        const float ga = gamma ? gamma[c + q] : 1.f, be = beta ? beta[c + q] : 0.f;
        const float xh = (xx[q] - mean[so + q]) * rs;
        const float y = fmaf(xh, ga, be);
        const float gp = gg[q] * sdt::leaky_grad(y, slope);
        gg[q] = sdt::out_round(rs * ga * (gp - m1[so + q] - xh * m2[so + q]), out_tf32);
    }
    *reinterpret_cast<float4*>(g + e) = make_float4(gg[0], gg[1], gg[2], gg[3]);
}

// C in {64, 128, 256}: every thread owns one float4 channel group for the whole tile (all 256 threads busy for the
// 64-channel maps, which are the largest), statistics hoisted into registers, rows unrolled for loads in flight.
__global__ void __launch_bounds__(256) norm_bwd_reduce_pow2_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   int P, int C, int groups_is_batch, float slope,
                                                                   float* __restrict__ partial, int tiles_per_image) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ float red[2][1024];
    const int b = blockIdx.x / tiles_per_image, tile = blockIdx.x % tiles_per_image;
    const int c4n = C >> 2, nrl = 256 / c4n;
    const int cq = threadIdx.x % c4n, rl = threadIdx.x / c4n;
    const int c = cq * 4;
    const int rows = (P + tiles_per_image - 1) / tiles_per_image;       // tiles_per_image is the caller's choice of parallelism
    const int p0 = tile * rows, p1 = min(P, p0 + rows);
    const int so = (groups_is_batch ? b * C : 0) + c;
    const float4 mu = *reinterpret_cast<const float4*>(mean + so), rs = *reinterpret_cast<const float4*>(rstd + so);
    const float4 ga = gamma ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 be = beta ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float mu_[4] = {mu.x, mu.y, mu.z, mu.w}, rs_[4] = {rs.x, rs.y, rs.z, rs.w};
    const float ga_[4] = {ga.x, ga.y, ga.z, ga.w}, be_[4] = {be.x, be.y, be.z, be.w};
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* gp4 = reinterpret_cast<const float4*>(g + (size_t)b * P * C) + cq;
    const float4* xp4 = reinterpret_cast<const float4*>(x + (size_t)b * P * C) + cq;
#pragma unroll 4
    for (int p = p0 + rl; p < p1; p += nrl) {
        const float4 gv = __ldg(gp4 + (size_t)p * c4n);
        const float4 xv = __ldg(xp4 + (size_t)p * c4n);
        const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float xh = (xx[q] - mu_[q]) * rs_[q];
            const float y = fmaf(xh, ga_[q], be_[q]);
            const float gp = gg[q] * sdt::leaky_grad(y, slope);
            s1[q] += gp;
            s2[q] = fmaf(gp, xh, s2[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        red[0][rl * C + c + q] = s1[q];
        red[1][rl * C + c + q] = s2[q];
    }
    __syncthreads();
    if (threadIdx.x < C) {
        float a = 0.f, bsum = 0.f;
        for (int r = 0; r < nrl; ++r) {
            a += red[0][r * C + threadIdx.x];
            bsum += red[1][r * C + threadIdx.x];
        }
        partial[((size_t)blockIdx.x * 2 + 0) * C + threadIdx.x] = a;
        partial[((size_t)blockIdx.x * 2 + 1) * C + threadIdx.x] = bsum;
    }
}

// grid (x, B): image from the block, a thread keeps ONE channel group across its grid-stride loop (stride is a
// multiple of C/4), so the six statistic vectors are loaded once.
__global__ void __launch_bounds__(256) norm_bwd_apply_pow2_kernel(float* __restrict__ g, const float* __restrict__ x,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ m1, const float* __restrict__ m2,
                                                                  long long per_image4, int C, int groups_is_batch, float slope,
                                                                  int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int b = blockIdx.y;
    const int c4n = C >> 2;
    const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(e0 % c4n) * 4;
    const int so = (groups_is_batch ? b * C : 0) + c;
    const float4 mu = *reinterpret_cast<const float4*>(mean + so), rs = *reinterpret_cast<const float4*>(rstd + so);
    const float4 a1 = *reinterpret_cast<const float4*>(m1 + so), a2 = *reinterpret_cast<const float4*>(m2 + so);
    const float4 ga = gamma ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 be = beta ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float mu_[4] = {mu.x, mu.y, mu.z, mu.w}, rs_[4] = {rs.x, rs.y, rs.z, rs.w};
    const float a1_[4] = {a1.x, a1.y, a1.z, a1.w}, a2_[4] = {a2.x, a2.y, a2.z, a2.w};
    const float ga_[4] = {ga.x, ga.y, ga.z, ga.w}, be_[4] = {be.x, be.y, be.z, be.w};
    float4* gi = reinterpret_cast<float4*>(g) + (size_t)b * per_image4;
    const float4* xi = reinterpret_cast<const float4*>(x) + (size_t)b * per_image4;
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 4
    for (long long e = e0; e < per_image4; e += stride) {
        const float4 gv = gi[e];
        const float4 xv = __ldg(xi + e);
        float gg[4] = {gv.x, gv.y, gv.z, gv.w};
        const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float xh = (xx[q] - mu_[q]) * rs_[q];
            const float y = fmaf(xh, ga_[q], be_[q]);
            const float gp = gg[q] * sdt::leaky_grad(y, slope);
            gg[q] = sdt::out_round(rs_[q] * ga_[q] * (gp - a1_[q] - xh * a2_[q]), out_tf32);
        }
        gi[e] = make_float4(gg[0], gg[1], gg[2], gg[3]);
    }
}

// ---- channel LayerNorm (+ activation) over rows of (R, C): one warp per row ---------------------------
template <int MAXV>
__global__ void __launch_bounds__(256) rownorm_fwd_kernel(const float* __restrict__ x, int R, int C, float eps, float slope,
                                                          float* __restrict__ y, float* __restrict__ mean,
                                                          float* __restrict__ rstd, int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (row >= R) return;
    const float* xr = x + (size_t)row * C;
    float v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < C ? xr[c] : 0.f;
        s += v[i];
    }
    const float mu = sdt::warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        const float dlt = c < C ? v[i] - mu : 0.f;
        q = fmaf(dlt, dlt, q);
    }
    const float var = sdt::warp_sum(q) / (float)C;   // biased
    const float rs = 1.0f / sqrtf(var + eps);
    float* yr = y + (size_t)row * C;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) yr[c] = sdt::out_round(sdt::leaky((v[i] - mu) * rs, slope), out_tf32);
    }
    if (lane == 0) {
        mean[row] = mu;
        rstd[row] = rs;
    }
}

template <int MAXV>
__global__ void __launch_bounds__(256) rownorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          int R, int C, float slope, float* __restrict__ gx, int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (row >= R) return;
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + (size_t)row * C;
    const float* gr = gy + (size_t)row * C;
    float xh[MAXV], gp[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        xh[i] = 0.f; gp[i] = 0.f;
        if (c < C) {
            xh[i] = (xr[c] - mu) * rs;
            gp[i] = gr[c] * sdt::leaky_grad(xh[i], slope);
        }
        s1 += gp[i];
        s2 = fmaf(gp[i], xh[i], s2);
    }
    const float m1 = sdt::warp_sum(s1) / (float)C, m2 = sdt::warp_sum(s2) / (float)C;
    float* o = gx + (size_t)row * C;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) o[c] = sdt::out_round(rs * (gp[i] - m1 - xh[i] * m2), out_tf32);
    }
}

__global__ void scale_shift_act_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                       const float* __restrict__ shift, long long total, int P, int C, int bstride,
                                       float slope, float* __restrict__ y, int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const int b = (int)(e / ((long long)P * C));
    const int so = b * bstride + c;
    y[e] = sdt::out_round(sdt::leaky(fmaf(x[e], scale[so], shift[so]), slope), out_tf32);
}

// C % 4 == 0: one float4 per thread, image index from the block (gridDim.y = B)
__global__ void __launch_bounds__(256) scale_shift_act_v4_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, long long per_image4, int C,
                                                                 int bstride, float slope, float* __restrict__ y, int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int b = blockIdx.y;
    const float4* xi = reinterpret_cast<const float4*>(x) + (size_t)b * per_image4;
    float4* yi = reinterpret_cast<float4*>(y) + (size_t)b * per_image4;
    const float* sc = scale + (size_t)b * bstride;
    const float* sh = shift + (size_t)b * bstride;
    const int c4n = C / 4;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_image4; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % c4n) * 4;
        const float4 v = xi[e];
        const float4 s4 = *reinterpret_cast<const float4*>(sc + c), h4 = *reinterpret_cast<const float4*>(sh + c);
        float4 o;
        o.x = sdt::out_round(sdt::leaky(fmaf(v.x, s4.x, h4.x), slope), out_tf32);
        o.y = sdt::out_round(sdt::leaky(fmaf(v.y, s4.y, h4.y), slope), out_tf32);
        o.z = sdt::out_round(sdt::leaky(fmaf(v.z, s4.z, h4.z), slope), out_tf32);
        o.w = sdt::out_round(sdt::leaky(fmaf(v.w, s4.w, h4.w), slope), out_tf32);
        yi[e] = o;
    }
}

}  // namespace

extern "C" int sdt_norm_finalize(const float* partial, int groups, int tiles_per_group, int C, double count,
                                 const float* gamma, const float* beta, float eps, float* scale, float* shift, float* mean,
                                 float* rstd, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                 float momentum, void* stream) {
    SDT_REQUIRE(partial && scale && shift, "sdt_norm_finalize: null pointer");
    SDT_REQUIRE(groups > 0 && tiles_per_group > 0 && C > 0 && count > 0, "sdt_norm_finalize: bad extents");
    SDT_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "sdt_norm_finalize: running stats come together");
    SDT_REQUIRE(running_mean == nullptr || groups == 1, "sdt_norm_finalize: running statistics need groups == 1 (BatchNorm)");
    sdt::launch(norm_finalize_kernel, dim3(groups * sdt::ceil_div(C, 32)), dim3(256), 0, sdt::as_stream(stream), partial, groups, tiles_per_group, C, count, gamma, beta, eps, scale, shift, mean, rstd, running_mean, running_var,
        num_batches_tracked, momentum);
    SDT_LAUNCH_OK("norm_finalize_kernel");
    return SDT_OK;
}

extern "C" int sdt_chan_stats(const float* x, int B, int H, int W, int C, int w0, int w1, int rows_per_part, float* partial,
                              int n_parts, void* stream) {
    SDT_REQUIRE(x && partial, "sdt_chan_stats: null pointer");
    SDT_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C <= 256 && 256 % C == 0, "sdt_chan_stats: need C dividing 256 (C=%d)", C);
    SDT_REQUIRE(w0 >= 0 && w1 >= w0 && w1 <= W && rows_per_part > 0 && n_parts > 0, "sdt_chan_stats: bad window [%d, %d) of %d", w0, w1, W);
    SDT_REQUIRE((long long)n_parts * rows_per_part >= (long long)H * (w1 - w0), "sdt_chan_stats: %d parts of %d rows do not cover the window", n_parts, rows_per_part);
    sdt::launch(chan_stats_kernel, dim3(n_parts, B), dim3(256), (size_t)(256 / C) * 2 * C * sizeof(float), sdt::as_stream(stream), x, H, W, C, w0,
                w1, rows_per_part, partial, n_parts);
    SDT_LAUNCH_OK("chan_stats_kernel");
    return SDT_OK;
}

extern "C" int sdt_bn_eval_scale_shift(const float* running_mean, const float* running_var, const float* gamma,
                                       const float* beta, float eps, int C, float* scale, float* shift, void* stream) {
    SDT_REQUIRE(running_mean && running_var && scale && shift && C > 0, "sdt_bn_eval_scale_shift: bad arguments");
    sdt::launch(bn_eval_kernel, dim3(sdt::ceil_div(C, 128)), dim3(128), 0, sdt::as_stream(stream), running_mean, running_var, gamma, beta, eps, C,
                                                                             scale, shift);
    SDT_LAUNCH_OK("bn_eval_kernel");
    return SDT_OK;
}

extern "C" int sdt_norm_bwd_reduce(const float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                                   const float* beta, int B, int P, int C, int groups, float slope, float* partial,
                                   int tiles_per_image, void* stream) {
    SDT_REQUIRE(g && x && mean && rstd && partial, "sdt_norm_bwd_reduce: null pointer");
    SDT_REQUIRE(B > 0 && P > 0 && C > 0 && C % 4 == 0, "sdt_norm_bwd_reduce: need C %% 4 == 0 (C=%d)", C);
    SDT_REQUIRE(groups == B || groups == 1, "sdt_norm_bwd_reduce: groups must be B or 1");
    SDT_REQUIRE(tiles_per_image >= 1 && tiles_per_image <= P, "sdt_norm_bwd_reduce: tiles_per_image=%d outside [1, P=%d]", tiles_per_image, P);
    if (C == 64 || C == 128 || C == 256)
        sdt::launch(norm_bwd_reduce_pow2_kernel, dim3(B * tiles_per_image), dim3(256), 0, sdt::as_stream(stream), g, x, mean, rstd, gamma, beta, P, C, groups == B ? 1 : 0, slope, partial, tiles_per_image);
    else
        sdt::launch(norm_bwd_reduce_kernel, dim3(B * tiles_per_image), dim3(256), 0, sdt::as_stream(stream), g, x, mean, rstd, gamma, beta, P, C,
                                                                                        groups == B ? 1 : 0,
                                                                                        slope, partial, tiles_per_image);
    SDT_LAUNCH_OK("norm_bwd_reduce_kernel");
    return SDT_OK;
}

extern "C" int sdt_norm_bwd_finalize(const float* partial, int groups, int tiles_per_group, int C, double count, float* m1,
                                     float* m2, float* dgamma, float* dbeta, int accumulate, void* stream) {
    SDT_REQUIRE(partial && m1 && m2, "sdt_norm_bwd_finalize: null pointer");
    SDT_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "sdt_norm_bwd_finalize: dgamma/dbeta come together");
    SDT_REQUIRE(dgamma == nullptr || groups == 1, "sdt_norm_bwd_finalize: affine gradients need groups == 1");
    sdt::launch(norm_bwd_finalize_kernel, dim3(groups * sdt::ceil_div(C, 32)), dim3(256), 0, sdt::as_stream(stream), partial, groups, tiles_per_group, C, count, m1, m2, dgamma, dbeta, accumulate);
    SDT_LAUNCH_OK("norm_bwd_finalize_kernel");
    return SDT_OK;
}

extern "C" int sdt_norm_bwd_apply(float* g, const float* x, const float* mean, const float* rstd, const float* gamma,
                                  const float* beta, const float* m1, const float* m2, int B, int P, int C, int groups,
                                  float slope, int out_tf32, void* stream) {
    SDT_REQUIRE(g && x && mean && rstd && m1 && m2, "sdt_norm_bwd_apply: null pointer");
    SDT_REQUIRE(C % 4 == 0, "sdt_norm_bwd_apply: need C %% 4 == 0 (C=%d)", C);
    SDT_REQUIRE(groups == B || groups == 1, "sdt_norm_bwd_apply: groups must be B or 1");
    const long long total4 = (long long)B * P * C / 4;
    if ((C == 64 || C == 128 || C == 256) && B <= 65535) {
        const long long per_image4 = (long long)P * C / 4;
        const int gx = (int)std::min<long long>(sdt::ceil_div(per_image4, 256 * 4), 4096);
        sdt::launch(norm_bwd_apply_pow2_kernel, dim3(gx, B), dim3(256), 0, sdt::as_stream(stream), g, x, mean, rstd, gamma, beta, m1, m2,
                                                                                  per_image4, C, groups == B ? 1 : 0, slope, out_tf32);
    } else
        sdt::launch(norm_bwd_apply_kernel, dim3(sdt::ceil_div(total4, 256)), dim3(256), 0, sdt::as_stream(stream), g, x, mean, rstd, gamma, beta, m1, m2,
                                                                                             total4, P, C, groups == B ? 1 : 0, slope, out_tf32);
    SDT_LAUNCH_OK("norm_bwd_apply_kernel");
    return SDT_OK;
}

extern "C" int sdt_rownorm_act_fwd(const float* x, int R, int C, float eps, float slope, float* y, float* mean, float* rstd,
                                   int out_tf32, void* stream) {
    SDT_REQUIRE(x && y && mean && rstd && R > 0 && C > 0, "sdt_rownorm_act_fwd: bad arguments");
    SDT_REQUIRE(C <= 1024, "sdt_rownorm_act_fwd: C=%d > 1024 unsupported", C);
    cudaStream_t st = sdt::as_stream(stream);
    const int grid = sdt::ceil_div(R, 8);
    if (C <= 256) sdt::launch(rownorm_fwd_kernel<8>, dim3(grid), dim3(256), 0, st, x, R, C, eps, slope, y, mean, rstd, out_tf32);
    else sdt::launch(rownorm_fwd_kernel<32>, dim3(grid), dim3(256), 0, st, x, R, C, eps, slope, y, mean, rstd, out_tf32);
    SDT_LAUNCH_OK("rownorm_fwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_rownorm_act_bwd(const float* g_y, const float* x, const float* mean, const float* rstd, int R, int C,
                                   float slope, float* g_x, int out_tf32, void* stream) {
    SDT_REQUIRE(g_y && x && mean && rstd && g_x && R > 0 && C > 0, "sdt_rownorm_act_bwd: bad arguments");
    SDT_REQUIRE(C <= 1024, "sdt_rownorm_act_bwd: C=%d > 1024 unsupported", C);
    cudaStream_t st = sdt::as_stream(stream);
    const int grid = sdt::ceil_div(R, 8);
    if (C <= 256) sdt::launch(rownorm_bwd_kernel<8>, dim3(grid), dim3(256), 0, st, g_y, x, mean, rstd, R, C, slope, g_x, out_tf32);
    else sdt::launch(rownorm_bwd_kernel<32>, dim3(grid), dim3(256), 0, st, g_y, x, mean, rstd, R, C, slope, g_x, out_tf32);
    SDT_LAUNCH_OK("rownorm_bwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_scale_shift_act(const float* x, const float* scale, const float* shift, int B, int P, int C, int bstride,
                                   float slope, float* y, int out_tf32, void* stream) {
    SDT_REQUIRE(x && scale && shift && y && B > 0 && P > 0 && C > 0, "sdt_scale_shift_act: bad arguments");
    const long long total = (long long)B * P * C;
    if (C % 4 == 0 && B <= 65535 && ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)scale | (uintptr_t)shift) & 15) == 0)) {
        const long long per_image4 = (long long)P * C / 4;
        int gx = sdt::ceil_div(per_image4, 256 * 2);
        if (gx > 2048) gx = 2048;
        sdt::launch(scale_shift_act_v4_kernel, dim3(gx, B), dim3(256), 0, sdt::as_stream(stream), x, scale, shift, per_image4, C, bstride, slope, y, out_tf32);
        SDT_LAUNCH_OK("scale_shift_act_v4_kernel");
        return SDT_OK;
    }
    sdt::launch(scale_shift_act_kernel, dim3(sdt::ceil_div(total, 256)), dim3(256), 0, sdt::as_stream(stream), x, scale, shift, total, P, C, bstride,
                                                                                         slope, y, out_tf32);
    SDT_LAUNCH_OK("scale_shift_act_kernel");
    return SDT_OK;
}
