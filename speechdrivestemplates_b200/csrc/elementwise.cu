// Resampling / concat / loss / optimizer kernels of the Voice2Pose and Pose2Pose steps (channels-last fp32).
// Reference call sites are cited per kernel; all reductions are fixed-order two-stage sums.
#include "common.cuh"

namespace {

// PyTorch's area_pixel_compute_source_index for align_corners=False (linear/bilinear), computed in fp32 like ATen.
__device__ __forceinline__ void lerp_coeff(int j, int in_size, int out_size, int& i0, int& i1, float& w0, float& w1) {
    const float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)j + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    w1 = src - (float)i0;
    w0 = 1.f - w1;
}

// ---- F.interpolate(x,(1,F),'bilinear').squeeze(2) + cat(code) : generator.py:41-42,109-111 ----------------------------
__global__ void enc_to_seq_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                      const float* __restrict__ shift, int bstride, float slope, int B, int H, int W, int C,
                                      const float* __restrict__ code, int D, int F, float* __restrict__ out, int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long total = (long long)B * F * (C + D);
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int ch = (int)(e % (C + D));
    const int j = (int)((e / (C + D)) % F);
    const int b = (int)(e / ((long long)(C + D) * F));
    if (ch >= C) {
        out[e] = sdt::out_round(code[b * D + (ch - C)], out_tf32);
        return;
    }
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    lerp_coeff(0, H, 1, y0, y1, wy0, wy1);
    lerp_coeff(j, W, F, x0, x1, wx0, wx1);
    const float sc = scale[b * bstride + ch], sh = shift[b * bstride + ch];
    auto at = [&](int yy, int xx) { return sdt::leaky(fmaf(x[(((size_t)b * H + yy) * W + xx) * C + ch], sc, sh), slope); };
    // ATen: w_y0*(w_x0*v00 + w_x1*v01) + w_y1*(w_x0*v10 + w_x1*v11)
    out[e] = sdt::out_round(wy0 * (wx0 * at(y0, x0) + wx1 * at(y0, x1)) + wy1 * (wx0 * at(y1, x0) + wx1 * at(y1, x1)), out_tf32);
}

// adjoint in gather form (deterministic): every (b, y, x, c) sums the output frames that sampled it
__global__ void enc_to_seq_bwd_kernel(const float* __restrict__ g_out, int B, int H, int W, int C, int D, int F,
                                      float* __restrict__ g_act) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long total = (long long)B * H * W * C;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const int xx = (int)((e / C) % W);
    const int yy = (int)((e / ((long long)C * W)) % H);
    const int b = (int)(e / ((long long)C * W * H));
    int y0, y1;
    float wy0, wy1;
    lerp_coeff(0, H, 1, y0, y1, wy0, wy1);
    float wy = 0.f;
    if (yy == y0) wy += wy0;
    if (yy == y1) wy += wy1;
    float acc = 0.f;
    if (wy != 0.f) {
        // only frames whose source position (j + 0.5) * W / F - 0.5 lies within one pixel of xx can have sampled it: a window
        // of ~2 F / W + 4 frames instead of all F (same terms in the same ascending order, the others had weight 0)
        const int jlo = max(0, (int)(((long long)(xx - 1) * F) / W) - 2);
        const int jhi = min(F - 1, (int)(((long long)(xx + 2) * F + W - 1) / W) + 2);
        for (int j = jlo; j <= jhi; ++j) {
            int x0, x1;
            float wx0, wx1;
            lerp_coeff(j, W, F, x0, x1, wx0, wx1);
            float wx = 0.f;
            if (xx == x0) wx += wx0;
            if (xx == x1) wx += wx1;
            if (wx != 0.f) acc = fmaf(wy * wx, g_out[((size_t)b * F + j) * (C + D) + c], acc);
        }
    }
    g_act[e] = acc;
}

__global__ void code_grad_from_seq_kernel(const float* __restrict__ g_out, int B, int C, int D, int F,
                                          float* __restrict__ g_code) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * D) return;
    const int b = e / D, dd = e % D;
    float acc = 0.f;
    for (int j = 0; j < F; ++j) acc += g_out[((size_t)b * F + j) * (C + D) + C + dd];
    g_code[e] = acc;
}

// ---- F.interpolate(x, Lout, 'linear') (+ skip): generator.py:79-83, autoencoder.py:62-66 ----------------------------
__global__ void upsample_add_fwd_kernel(const float* __restrict__ x, const float* __restrict__ skip, int B, int Lin,
                                        int Lout, int C, float* __restrict__ out, int out_tf32) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long total = (long long)B * Lout * C;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const int j = (int)((e / C) % Lout);
    const int b = (int)(e / ((long long)C * Lout));
    int i0, i1;
    float w0, w1;
    lerp_coeff(j, Lin, Lout, i0, i1, w0, w1);
    float v = w0 * x[((size_t)b * Lin + i0) * C + c] + w1 * x[((size_t)b * Lin + i1) * C + c];
    if (skip != nullptr) v += skip[e];
    out[e] = sdt::out_round(v, out_tf32);
}

__global__ void upsample_bwd_kernel(const float* __restrict__ g_out, int B, int Lin, int Lout, int C,
                                    float* __restrict__ g_x, int accumulate) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long total = (long long)B * Lin * C;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const int i = (int)((e / C) % Lin);
    const int b = (int)(e / ((long long)C * Lin));
    // output positions that can reference input i lie in a window around i*Lout/Lin
    const float inv = (float)Lout / (float)Lin;
    int j_lo = (int)floorf(((float)i - 1.0f) * inv) - 2, j_hi = (int)ceilf(((float)i + 1.5f) * inv) + 2;
    if (j_lo < 0) j_lo = 0;
    if (j_hi > Lout - 1) j_hi = Lout - 1;
    float acc = 0.f;
    for (int j = j_lo; j <= j_hi; ++j) {
        int i0, i1;
        float w0, w1;
        lerp_coeff(j, Lin, Lout, i0, i1, w0, w1);
        float w = 0.f;
        if (i == i0) w += w0;
        if (i == i1) w += w1;
        if (w != 0.f) acc = fmaf(w, g_out[((size_t)b * Lout + j) * C + c], acc);
    }
    g_x[e] = accumulate ? g_x[e] + acc : acc;
}

// ---- L1 loss: voice2pose.py:141-142 -----------------------------------------------------------------------------------
constexpr int kLossBlocks = 256;

__global__ void __launch_bounds__(256) l1_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                         long long n, float lambda, float* __restrict__ g_pred,
                                                         float* __restrict__ partial) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ float red[8];
    float acc = 0.f;
    const float gscale = lambda / (float)n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float dlt = pred[e] - gt[e];
        acc += fabsf(dlt) * lambda;
        if (g_pred != nullptr) g_pred[e] = dlt > 0.f ? gscale : (dlt < 0.f ? -gscale : 0.f);
    }
    acc = sdt::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        partial[blockIdx.x] = s;
    }
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int count, double denom, float* __restrict__ out) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < count; ++i) s += (double)partial[i];
        out[0] = (float)(s / denom);
    }
}

// ---- clip-code gather + batch-statistics KL: voice2pose.py:94,147-157; closed-form gradient per SURVEY App. E --------
__global__ void code_gather_kl_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int B, int D,
                                      float lambda, float* __restrict__ code, float* __restrict__ out,
                                      float* __restrict__ g_code) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    // one CTA; thread dd < D owns one code dimension
    __shared__ int all_nonzero;
    __shared__ float terms[1024];
    const int dd = threadIdx.x;
    if (dd == 0) all_nonzero = 1;
    __syncthreads();
    float m = 0.f, var = 0.f;
    if (dd < D) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) {
            const float v = table[(size_t)idx[b] * D + dd];
            code[b * D + dd] = v;
            s += v;
        }
        m = s / (float)B;
        float q = 0.f;
        for (int b = 0; b < B; ++b) {
            const float dlt = code[b * D + dd] - m;
            q = fmaf(dlt, dlt, q);
        }
        var = q / (float)(B - 1);   // unbiased; B == 1 gives NaN like torch.var
        if (!(var != 0.f)) atomicAnd(&all_nonzero, 0);
    }
    __syncthreads();
    const bool apply = all_nonzero != 0;
    if (dd < D) {
        terms[dd] = apply ? (-logf(var) + m * m + var - 1.f) : 0.f;
        const float k = lambda * 0.5f / (float)D;
        for (int b = 0; b < B; ++b) {
            float gval = 0.f;
            if (apply) gval = k * (2.f * m / (float)B + (1.f - 1.f / var) * 2.f * (code[b * D + dd] - m) / (float)(B - 1));
            g_code[b * D + dd] = gval;
        }
    }
    __syncthreads();
    if (dd == 0) {
        float s = 0.f;
        for (int i = 0; i < D; ++i) s += terms[i];
        out[0] = apply ? 0.5f * (s / (float)D) * lambda : 0.f;
        out[1] = apply ? 1.f : 0.f;
    }
}

__global__ void code_scatter_grad_kernel(const float* __restrict__ ga, const float* __restrict__ gb,
                                         const int64_t* __restrict__ idx, int B, int D, float* __restrict__ g_table) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * D) return;
    const int b = e / D, dd = e % D;
    const int64_t row = idx[b];
    for (int p = 0; p < b; ++p)
        if (idx[p] == row) return;   // an earlier occurrence owns this row
    float s = 0.f;
    for (int p = b; p < B; ++p)
        if (idx[p] == row) s += (ga ? ga[p * D + dd] : 0.f) + (gb ? gb[p * D + dd] : 0.f);
    g_table[(size_t)row * D + dd] += s;
}

// buffer scatter table[idx[b]] = src[b] (pose2pose.py:135-137).  Duplicate indices: the LAST occurrence wins, which is what the
// sequential CPU index_put does; torch's CUDA index_put_ leaves the winner unspecified.
__global__ void code_store_rows_kernel(const float* __restrict__ src_a, float* __restrict__ table_a, const float* __restrict__ src_b,
                                       float* __restrict__ table_b, const int64_t* __restrict__ idx, int B, int D) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * D) return;
    const int b = e / D, dd = e % D;
    const int64_t row = idx[b];
    for (int p = b + 1; p < B; ++p)
        if (idx[p] == row) return;   // a later occurrence owns this row
    table_a[(size_t)row * D + dd] = src_a[e];
    if (src_b != nullptr) table_b[(size_t)row * D + dd] = src_b[e];
}

__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ g, int R, int C, float* __restrict__ out,
                                                      int accumulate) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    // one CTA per 32 columns; 32 row-lanes with four loads in flight each (8 lanes of one dependent load chain took 29 us
    // for 2048 rows); fixed summation order
    __shared__ float red[32][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < C) {
        int r = rl;
        for (; r + 96 < R; r += 128) {
            s0 += g[(size_t)r * C + c];
            s1 += g[(size_t)(r + 32) * C + c];
            s2 += g[(size_t)(r + 64) * C + c];
            s3 += g[(size_t)(r + 96) * C + c];
        }
        for (; r < R; r += 32) s0 += g[(size_t)r * C + c];
    }
    red[rl][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (rl == 0 && c < C) {
        float t = 0.f;
        for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x & 31];
        out[c] = accumulate ? out[c] + t : t;
    }
}

__global__ void __launch_bounds__(1024) mse_const_kernel(const float* __restrict__ s, long long n, float target, float lambda,
                                                         float* __restrict__ out, float* __restrict__ g_s) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ float red[32];
    float acc = 0.f;
    const float gscale = lambda * 2.f / (float)n;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        const float dlt = s[e] - target;
        acc = fmaf(dlt, dlt, acc);
        if (g_s != nullptr) g_s[e] = gscale * dlt;
    }
    acc = sdt::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 32; ++i) t += red[i];
        out[0] = lambda * (t / (float)n);
    }
}

__global__ void motion_diff_fwd_kernel(const float* __restrict__ x, int B, int T, int C, float* __restrict__ out) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long total = (long long)B * (T - 1) * C;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const int t = (int)((e / C) % (T - 1));
    const int b = (int)(e / ((long long)C * (T - 1)));
    const size_t o = ((size_t)b * T + t) * C + c;
    out[e] = x[o + C] - x[o];
}

__global__ void motion_diff_bwd_kernel(const float* __restrict__ g_out, int B, int T, int C, float* __restrict__ g_x,
                                       int accumulate) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long total = (long long)B * T * C;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const int t = (int)((e / C) % T);
    const int b = (int)(e / ((long long)C * T));
    float v = 0.f;
    if (t >= 1) v += g_out[((size_t)b * (T - 1) + t - 1) * C + c];
    if (t < T - 1) v -= g_out[((size_t)b * (T - 1) + t) * C + c];
    g_x[e] = accumulate ? g_x[e] + v : v;
}

// ---- pose VAE head / reparameterisation: autoencoder.py:31-35,84-87; pose2pose.py:77 ---------------------------------
__global__ void pose_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                     const float* __restrict__ shift, float slope, int B, int L, int D2,
                                     float* __restrict__ mu, float* __restrict__ logvar) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * D2) return;
    const int b = e / D2, ch = e % D2;
    const float v = sdt::leaky(fmaf(x[((size_t)b * L + 0) * D2 + ch], scale[ch], shift[ch]), slope);
    if (ch & 1) logvar[b * (D2 / 2) + ch / 2] = v;
    else mu[b * (D2 / 2) + ch / 2] = v;
}

__global__ void __launch_bounds__(1024) vae_reparam_kl_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                                              const float* __restrict__ eps, int n, float lambda,
                                                              float* __restrict__ code, float* __restrict__ out) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ float red[32];
    float acc = 0.f;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const float m = mu[e], lv = logvar[e];
        code[e] = m + expf(0.5f * lv) * eps[e];
        acc += -lv + m * m + expf(lv) - 1.f;
    }
    acc = sdt::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 32; ++i) t += red[i];
        out[0] = 0.5f * (t / (float)n) * lambda;
    }
}

// backward of the reparameterisation + KL:
//   d mu = g_code + lambda*mu/n ;  d logvar = g_code*0.5*exp(0.5 logvar)*eps + lambda*0.5*(exp(logvar) - 1)/n
__global__ void vae_reparam_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                          const float* __restrict__ eps, const float* __restrict__ g_code, int n, float lambda,
                                          float* __restrict__ g_mu, float* __restrict__ g_logvar) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float m = mu[e], lv = logvar[e], gc = g_code[e];
    g_mu[e] = gc + lambda * m / (float)n;
    g_logvar[e] = gc * 0.5f * expf(0.5f * lv) * eps[e] + lambda * 0.5f * (expf(lv) - 1.f) / (float)n;
}

// adjoint of the PoseSeqEncoder tail (autoencoder.py:31-35): g_act (B,L,2D) = 0 except t = 0, where even channels get
// g_mu and odd channels g_logvar
__global__ void pose_head_bwd_kernel(const float* __restrict__ g_mu, const float* __restrict__ g_logvar, int B, int L, int D2,
                                     float* __restrict__ g_act) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * L * D2) return;
    const int ch = e % D2, t = (e / D2) % L, b = e / (D2 * L);
    float v = 0.f;
    if (t == 0) v = (ch & 1) ? g_logvar[b * (D2 / 2) + ch / 2] : g_mu[b * (D2 / 2) + ch / 2];
    g_act[e] = v;
}

// ---- Adam: torch.optim.Adam single-tensor form, flat buffer -------------------------------------------
// scalars: [0] step_size = lr / (1 - beta1^t), [1] 1/sqrt(1 - beta2^t), [2] t (as float, informational),
//          [3] learning rate used when the lr argument is negative; 8 bytes at scalars+4 hold t as int64.
__global__ void adam_advance_kernel(float* scalars, float lr, double beta1, double beta2) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    long long* tptr = reinterpret_cast<long long*>(scalars + 4);
    const long long t = *tptr + 1;
    *tptr = t;
    if (lr < 0.f) lr = scalars[3];   // learning rate kept on the device (a captured CUDA graph can then follow MultiStepLR)
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    scalars[0] = (float)((double)lr / bc1);
    scalars[1] = (float)(1.0 / sqrt(bc2));
    scalars[2] = (float)t;
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        const float* __restrict__ scalars, float beta1, float beta2,
                                                        float omb1, float omb2, float eps, float grad_scale, float wd) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    // omb1/omb2 = (float)(1 - beta) evaluated in double on the host, as torch does for add_(alpha=1-beta1)
    const float step_size = scalars[0], inv_sqrt_bc2 = scalars[1];
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
        float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vq[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float gq = wd != 0.f ? gg[q] * grad_scale + wd * pp[q] : gg[q] * grad_scale;   // grad.add(param, alpha=wd)
            mm[q] = beta1 * mm[q] + omb1 * gq;
            vq[q] = beta2 * vq[q] + omb2 * gq * gq;
            const float denom = sqrtf(vq[q]) * inv_sqrt_bc2 + eps;
            pp[q] -= step_size * (mm[q] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vq[0], vq[1], vq[2], vq[3]);
    }
    // tail
    if (blockIdx.x == 0) {
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
            const float gq = wd != 0.f ? g[i] * grad_scale + wd * p[i] : g[i] * grad_scale;
            const float mq = beta1 * m[i] + omb1 * gq;
            const float vq = beta2 * v[i] + omb2 * gq * gq;
            m[i] = mq; v[i] = vq;
            p[i] -= step_size * (mq / (sqrtf(vq) * inv_sqrt_bc2 + eps));
        }
    }
}

}  // namespace

#define GRID1D(total) sdt::ceil_div((long long)(total), 256), 256, 0, sdt::as_stream(stream)

extern "C" int sdt_enc_to_seq_fwd(const float* x, const float* scale, const float* shift, int xf_bstride, float slope, int B,
                                  int H, int W, int C, const float* code, int D, int F, float* out, int out_tf32, void* stream) {
    SDT_REQUIRE(x && scale && shift && out, "sdt_enc_to_seq_fwd: null pointer");
    SDT_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && F > 0 && D >= 0, "sdt_enc_to_seq_fwd: bad extents");
    SDT_REQUIRE(D == 0 || code != nullptr, "sdt_enc_to_seq_fwd: D > 0 needs code");
    sdt::launch(enc_to_seq_fwd_kernel, dim3(sdt::ceil_div((long long)((long long)B * F * (C + D)), 256)), dim3(256), 0, sdt::as_stream(stream), x, scale, shift, xf_bstride, slope, B, H, W, C, code, D, F, out, out_tf32);
    SDT_LAUNCH_OK("enc_to_seq_fwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_enc_to_seq_bwd(const float* g_out, int B, int H, int W, int C, int D, int F, float* g_act, float* g_code,
                                  void* stream) {
    SDT_REQUIRE(g_out && g_act, "sdt_enc_to_seq_bwd: null pointer");
    SDT_REQUIRE(D == 0 || g_code != nullptr, "sdt_enc_to_seq_bwd: D > 0 needs g_code");
    sdt::launch(enc_to_seq_bwd_kernel, dim3(sdt::ceil_div((long long)((long long)B * H * W * C), 256)), dim3(256), 0, sdt::as_stream(stream), g_out, B, H, W, C, D, F, g_act);
    SDT_LAUNCH_OK("enc_to_seq_bwd_kernel");
    if (D > 0) {
        sdt::launch(code_grad_from_seq_kernel, dim3(sdt::ceil_div((long long)(B * D), 256)), dim3(256), 0, sdt::as_stream(stream), g_out, B, C, D, F, g_code);
        SDT_LAUNCH_OK("code_grad_from_seq_kernel");
    }
    return SDT_OK;
}

extern "C" int sdt_upsample_add_fwd(const float* x, const float* skip, int B, int Lin, int Lout, int C, float* out,
                                    int out_tf32, void* stream) {
    SDT_REQUIRE(x && out && B > 0 && Lin > 0 && Lout > 0 && C > 0, "sdt_upsample_add_fwd: bad arguments");
    sdt::launch(upsample_add_fwd_kernel, dim3(sdt::ceil_div((long long)((long long)B * Lout * C), 256)), dim3(256), 0, sdt::as_stream(stream), x, skip, B, Lin, Lout, C, out, out_tf32);
    SDT_LAUNCH_OK("upsample_add_fwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_upsample_bwd(const float* g_out, int B, int Lin, int Lout, int C, float* g_x, int accumulate, void* stream) {
    SDT_REQUIRE(g_out && g_x && B > 0 && Lin > 0 && Lout > 0 && C > 0, "sdt_upsample_bwd: bad arguments");
    sdt::launch(upsample_bwd_kernel, dim3(sdt::ceil_div((long long)((long long)B * Lin * C), 256)), dim3(256), 0, sdt::as_stream(stream), g_out, B, Lin, Lout, C, g_x, accumulate);
    SDT_LAUNCH_OK("upsample_bwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_l1_loss(const float* pred, const float* gt, int64_t n, float lambda, float* loss_out, float* g_pred,
                           float* partial, void* stream) {
    SDT_REQUIRE(pred && gt && loss_out && partial && n > 0, "sdt_l1_loss: bad arguments");
    sdt::launch(l1_partial_kernel, dim3(kLossBlocks), dim3(256), 0, sdt::as_stream(stream), pred, gt, n, lambda, g_pred, partial);
    SDT_LAUNCH_OK("l1_partial_kernel");
    sdt::launch(sum_partials_kernel, dim3(1), dim3(32), 0, sdt::as_stream(stream), partial, kLossBlocks, (double)n, loss_out);
    SDT_LAUNCH_OK("sum_partials_kernel");
    return SDT_OK;
}

extern "C" int sdt_code_gather_kl(const float* table, const int64_t* idx, int B, int D, float lambda, float* code, float* out,
                                  float* g_code, void* stream) {
    SDT_REQUIRE(table && idx && code && out && g_code, "sdt_code_gather_kl: null pointer");
    SDT_REQUIRE(B > 0 && D > 0 && D <= 1024, "sdt_code_gather_kl: need 0 < D <= 1024");
    sdt::launch(code_gather_kl_kernel, dim3(1), dim3(((D + 31) / 32) * 32), 0, sdt::as_stream(stream), table, idx, B, D, lambda, code, out, g_code);
    SDT_LAUNCH_OK("code_gather_kl_kernel");
    return SDT_OK;
}

extern "C" int sdt_code_scatter_grad(const float* g_code_a, const float* g_code_b, const int64_t* idx, int B, int D,
                                     float* g_table, void* stream) {
    SDT_REQUIRE(idx && g_table && (g_code_a || g_code_b), "sdt_code_scatter_grad: null pointer");
    sdt::launch(code_scatter_grad_kernel, dim3(sdt::ceil_div((long long)(B * D), 256)), dim3(256), 0, sdt::as_stream(stream), g_code_a, g_code_b, idx, B, D, g_table);
    SDT_LAUNCH_OK("code_scatter_grad_kernel");
    return SDT_OK;
}

extern "C" int sdt_code_store_rows(const float* src_a, float* table_a, const float* src_b, float* table_b, const int64_t* idx,
                                   int B, int D, void* stream) {
    SDT_REQUIRE(src_a && table_a && idx && B > 0 && D > 0, "sdt_code_store_rows: bad arguments");
    SDT_REQUIRE((src_b == nullptr) == (table_b == nullptr), "sdt_code_store_rows: src_b and table_b come together");
    sdt::launch(code_store_rows_kernel, dim3(sdt::ceil_div((long long)(B * D), 256)), dim3(256), 0, sdt::as_stream(stream), src_a, table_a, src_b, table_b, idx, B, D);
    SDT_LAUNCH_OK("code_store_rows_kernel");
    return SDT_OK;
}

extern "C" int sdt_colsum(const float* g, int R, int C, float* out, int accumulate, void* stream) {
    SDT_REQUIRE(g && out && R > 0 && C > 0, "sdt_colsum: bad arguments");
    sdt::launch(colsum_kernel, dim3(sdt::ceil_div(C, 32)), dim3(1024), 0, sdt::as_stream(stream), g, R, C, out, accumulate);
    SDT_LAUNCH_OK("colsum_kernel");
    return SDT_OK;
}

extern "C" int sdt_mse_const_loss(const float* s, int64_t n, float target, float lambda, float* out, float* g_s, void* stream) {
    SDT_REQUIRE(s && out && n > 0, "sdt_mse_const_loss: bad arguments");
    sdt::launch(mse_const_kernel, dim3(1), dim3(1024), 0, sdt::as_stream(stream), s, n, target, lambda, out, g_s);
    SDT_LAUNCH_OK("mse_const_kernel");
    return SDT_OK;
}

extern "C" int sdt_motion_diff_fwd(const float* x, int B, int T, int C, float* out, void* stream) {
    SDT_REQUIRE(x && out && B > 0 && T > 1 && C > 0, "sdt_motion_diff_fwd: bad arguments");
    sdt::launch(motion_diff_fwd_kernel, dim3(sdt::ceil_div((long long)((long long)B * (T - 1) * C), 256)), dim3(256), 0, sdt::as_stream(stream), x, B, T, C, out);
    SDT_LAUNCH_OK("motion_diff_fwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_motion_diff_bwd(const float* g_out, int B, int T, int C, float* g_x, int accumulate, void* stream) {
    SDT_REQUIRE(g_out && g_x && B > 0 && T > 1 && C > 0, "sdt_motion_diff_bwd: bad arguments");
    sdt::launch(motion_diff_bwd_kernel, dim3(sdt::ceil_div((long long)((long long)B * T * C), 256)), dim3(256), 0, sdt::as_stream(stream), g_out, B, T, C, g_x, accumulate);
    SDT_LAUNCH_OK("motion_diff_bwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_pose_head_fwd(const float* x, const float* scale, const float* shift, float slope, int B, int L, int D2,
                                 float* mu, float* logvar, void* stream) {
    SDT_REQUIRE(x && scale && shift && mu && logvar, "sdt_pose_head_fwd: null pointer");
    SDT_REQUIRE(B > 0 && L > 0 && D2 > 0 && D2 % 2 == 0, "sdt_pose_head_fwd: bad extents");
    sdt::launch(pose_head_fwd_kernel, dim3(sdt::ceil_div((long long)(B * D2), 256)), dim3(256), 0, sdt::as_stream(stream), x, scale, shift, slope, B, L, D2, mu, logvar);
    SDT_LAUNCH_OK("pose_head_fwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_vae_reparam_kl(const float* mu, const float* logvar, const float* eps, int n, float lambda, float* code,
                                  float* out, void* stream) {
    SDT_REQUIRE(mu && logvar && eps && code && out && n > 0, "sdt_vae_reparam_kl: bad arguments");
    sdt::launch(vae_reparam_kl_kernel, dim3(1), dim3(1024), 0, sdt::as_stream(stream), mu, logvar, eps, n, lambda, code, out);
    SDT_LAUNCH_OK("vae_reparam_kl_kernel");
    return SDT_OK;
}

extern "C" int sdt_adam_advance(float* scalars, float lr, double beta1, double beta2, void* stream) {
    SDT_REQUIRE(scalars, "sdt_adam_advance: null pointer");
    sdt::launch(adam_advance_kernel, dim3(1), dim3(1), 0, sdt::as_stream(stream), scalars, lr, beta1, beta2);
    SDT_LAUNCH_OK("adam_advance_kernel");
    return SDT_OK;
}

extern "C" int sdt_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* scalars, double beta1, double beta2, double eps, float grad_scale, float weight_decay,
                             void* stream) {
    SDT_REQUIRE(param && grad && exp_avg && exp_avg_sq && scalars && n > 0, "sdt_adam_flat: bad arguments");
    SDT_REQUIRE(weight_decay >= 0.f, "sdt_adam_flat: negative weight_decay");
    SDT_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                "sdt_adam_flat: buffers must be 16-byte aligned");
    int blocks = sdt::ceil_div(n / 4 + 1, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    sdt::launch(adam_flat_kernel, dim3(blocks), dim3(256), 0, sdt::as_stream(stream), param, grad, exp_avg, exp_avg_sq, n, scalars, (float)beta1,
                                                                 (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2),
                                                                 (float)eps, grad_scale, weight_decay);
    SDT_LAUNCH_OK("adam_flat_kernel");
    return SDT_OK;
}

extern "C" int sdt_vae_reparam_kl_bwd(const float* mu, const float* logvar, const float* eps, const float* g_code, int n,
                                      float lambda, float* g_mu, float* g_logvar, void* stream) {
    SDT_REQUIRE(mu && logvar && eps && g_code && g_mu && g_logvar && n > 0, "sdt_vae_reparam_kl_bwd: bad arguments");
    sdt::launch(vae_reparam_kl_bwd_kernel, dim3(sdt::ceil_div((long long)(n), 256)), dim3(256), 0, sdt::as_stream(stream), mu, logvar, eps, g_code, n, lambda, g_mu, g_logvar);
    SDT_LAUNCH_OK("vae_reparam_kl_bwd_kernel");
    return SDT_OK;
}

extern "C" int sdt_pose_head_bwd(const float* g_mu, const float* g_logvar, int B, int L, int D2, float* g_act, void* stream) {
    SDT_REQUIRE(g_mu && g_logvar && g_act && B > 0 && L > 0 && D2 > 0 && D2 % 2 == 0, "sdt_pose_head_bwd: bad arguments");
    sdt::launch(pose_head_bwd_kernel, dim3(sdt::ceil_div((long long)(B * L * D2), 256)), dim3(256), 0, sdt::as_stream(stream), g_mu, g_logvar, B, L, D2, g_act);
    SDT_LAUNCH_OK("pose_head_bwd_kernel");
    return SDT_OK;
}
