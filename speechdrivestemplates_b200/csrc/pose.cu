// Keypoint indexing / normalisation kernels -- the bit-exact gates.
// IEEE round-to-nearest single operations only: no FMA contraction, no reciprocal-multiply.
#include "common.cuh"

namespace {

constexpr int K137 = 137, K121 = 121;
constexpr int kRootNode122 = 1;            // gesture_dataset.py:42 (index in the 122-keypoint layout)
constexpr int kHandRootL = 6, kHandRootR = 3, kHeadRoot = 39;   // :43-45 (121 layout)

// 121-layout index -> 137-layout index: 122 = [0..7, 15, 16, 25..136] (gesture_dataset.py:134), then drop #1 (:143)
__device__ __forceinline__ int idx121_to_137(int k) {
    const int k122 = k == 0 ? 0 : k + 1;
    if (k122 < 8) return k122;
    if (k122 == 8) return 15;
    if (k122 == 9) return 16;
    return k122 + 15;
}

// part root of keypoint k in the 121 layout, or -1 (gesture_dataset.py:147-165)
__device__ __forceinline__ int part_root(int k) {
    if (k >= 9 && k < 79 && k != kHeadRoot) return kHeadRoot;
    if (k >= 79 && k < 100) return kHandRootL;
    if (k >= 100 && k < 121) return kHandRootR;
    return -1;
}

// raw (T,3,137) -> normalised (T,2,121); gesture_dataset.py:95-105,131-191
__global__ void pose_preprocess_kernel(const float* __restrict__ raw, int T, const float* __restrict__ mean,
                                       const float* __restrict__ stdv, int hierarchical, float* __restrict__ out) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * 2 * K121) return;
    const int k = e % K121, xy = (e / K121) % 2, t = e / (2 * K121);
    const float* row = raw + ((size_t)t * 3 + xy) * K137;
    const float root = row[kRootNode122];            // 122-index 1 == 137-index 1
    float v = __fsub_rn(row[idx121_to_137(k)], root);
    if (hierarchical) {
        const int pr = part_root(k);
        if (pr >= 0) v = __fsub_rn(v, __fsub_rn(row[idx121_to_137(pr)], root));
    }
    const int s = xy * K121 + k;
    out[e] = __fdiv_rn(__fsub_rn(v, mean[s]), stdv[s]);
}

// (B,T,2,121) f32 -> f64 final results; gesture_dataset.py:193-220
__global__ void pose_final_kernel(const float* __restrict__ poses, int B, int T, const double* __restrict__ mean,
                                  const double* __restrict__ stdv, const double* __restrict__ scale, int hierarchical,
                                  double* __restrict__ out) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long total = (long long)B * T * 2 * K121;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int k = (int)(e % K121);
    const int xy = (int)((e / K121) % 2);
    const int b = (int)(e / ((long long)T * 2 * K121));
    const int s = xy * K121 + k;
    const double* mb = mean + (size_t)b * 2 * K121;
    const double* sb = stdv + (size_t)b * 2 * K121;
    double v = __dadd_rn(__dmul_rn((double)poses[e], sb[s]), mb[s]);
    if (hierarchical) {
        const int pr = part_root(k);
        if (pr >= 0) {
            const int sr = xy * K121 + pr;
            const double r = __dadd_rn(__dmul_rn((double)poses[e - k + pr], sb[sr]), mb[sr]);
            v = __dadd_rn(v, r);
        }
    }
    out[e] = __dmul_rn(v, scale[b]);
}

// transform_normalized_parted2global (gesture_dataset.py:221-234), fp32: denormalise with the parted statistics,
// parted_to_global, normalise with the global statistics.  Stats are (2K) f32 arrays of ONE speaker (the reference
// assumes a single speaker per batch there).
__global__ void pose_parted2global_kernel(const float* __restrict__ poses, long long total, const float* __restrict__ mean_p,
                                          const float* __restrict__ std_p, const float* __restrict__ mean_g,
                                          const float* __restrict__ std_g, float* __restrict__ out) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int k = (int)(e % K121);
    const int xy = (int)((e / K121) % 2);
    const int s = xy * K121 + k;
    float v = __fadd_rn(__fmul_rn(poses[e], std_p[s]), mean_p[s]);
    const int pr = part_root(k);
    if (pr >= 0) {
        const int sr = xy * K121 + pr;
        v = __fadd_rn(v, __fadd_rn(__fmul_rn(poses[e - k + pr], std_p[sr]), mean_p[sr]));
    }
    out[e] = __fdiv_rn(__fsub_rn(v, mean_g[s]), std_g[s]);
}

// evaluate_step (voice2pose.py:412-430): per-clip partial sums, then a fixed-order finish.
__global__ void __launch_bounds__(256) pose_metrics_partial_kernel(const double* __restrict__ pred, const double* __restrict__ gt,
                                                                   int T, double* __restrict__ partial) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    __shared__ double red[8];
    __shared__ double lipg[1024], lipp[1024];
    __shared__ double maxg;
    const int b = blockIdx.x;
    const double* pb = pred + (size_t)b * T * 2 * K121;
    const double* gb = gt + (size_t)b * T * 2 * K121;
    double acc = 0.0;
    for (int e = threadIdx.x; e < T * K121; e += blockDim.x) {
        const int t = e / K121, k = e % K121;
        const double dx = __dsub_rn(pb[((size_t)t * 2 + 0) * K121 + k], gb[((size_t)t * 2 + 0) * K121 + k]);
        const double dy = __dsub_rn(pb[((size_t)t * 2 + 1) * K121 + k], gb[((size_t)t * 2 + 1) * K121 + k]);
        acc += sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    }
    acc = sdt::warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const double* pt = pb + (size_t)t * 2 * K121;
        const double* gtt = gb + (size_t)t * 2 * K121;
        const double px = pt[75] - pt[71], py = pt[K121 + 75] - pt[K121 + 71];
        const double gx = gtt[75] - gtt[71], gy = gtt[K121 + 75] - gtt[K121 + 71];
        lipp[t] = sqrt(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)));
        lipg[t] = sqrt(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        partial[2 * b + 0] = s;
        double m = lipg[0];
        for (int t = 1; t < T; ++t) m = lipg[t] > m ? lipg[t] : m;
        maxg = m + 1e-4;
        double lip = 0.0;
        for (int t = 0; t < T; ++t) lip += fabs(lipp[t] / maxg - lipg[t] / maxg);
        partial[2 * b + 1] = lip;
    }
}

__global__ void pose_metrics_finish_kernel(const double* __restrict__ partial, int B, int T, double* __restrict__ out) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double l2 = 0.0, lip = 0.0;
    for (int b = 0; b < B; ++b) {
        l2 += partial[2 * b];
        lip += partial[2 * b + 1];
    }
    out[0] = l2 / ((double)B * T * K121);
    out[1] = lip / ((double)B * T);
}

}  // namespace

extern "C" int sdt_pose_preprocess(const float* raw, int T, const float* mean, const float* std, int hierarchical, float* out,
                                   void* stream) {
    SDT_REQUIRE(raw && mean && std && out && T > 0, "sdt_pose_preprocess: bad arguments");
    sdt::launch(pose_preprocess_kernel, dim3(sdt::ceil_div(T * 2 * K121, 256)), dim3(256), 0, sdt::as_stream(stream), raw, T, mean, std, hierarchical, out);
    SDT_LAUNCH_OK("pose_preprocess_kernel");
    return SDT_OK;
}

extern "C" int sdt_pose_final_results(const float* poses, int B, int T, const double* mean, const double* std,
                                      const double* scale, int hierarchical, double* out, void* stream) {
    SDT_REQUIRE(poses && mean && std && scale && out && B > 0 && T > 0, "sdt_pose_final_results: bad arguments");
    const long long total = (long long)B * T * 2 * K121;
    sdt::launch(pose_final_kernel, dim3(sdt::ceil_div(total, 256)), dim3(256), 0, sdt::as_stream(stream), poses, B, T, mean, std, scale, hierarchical, out);
    SDT_LAUNCH_OK("pose_final_kernel");
    return SDT_OK;
}

extern "C" int sdt_pose_metrics(const double* pred, const double* gt, int B, int T, double* partial, double* out, void* stream) {
    SDT_REQUIRE(pred && gt && partial && out && B > 0 && T > 0, "sdt_pose_metrics: bad arguments");
    SDT_REQUIRE(T <= 1024, "sdt_pose_metrics: T=%d > 1024 unsupported", T);
    sdt::launch(pose_metrics_partial_kernel, dim3(B), dim3(256), 0, sdt::as_stream(stream), pred, gt, T, partial);
    SDT_LAUNCH_OK("pose_metrics_partial_kernel");
    sdt::launch(pose_metrics_finish_kernel, dim3(1), dim3(32), 0, sdt::as_stream(stream), partial, B, T, out);
    SDT_LAUNCH_OK("pose_metrics_finish_kernel");
    return SDT_OK;
}

extern "C" int sdt_pose_parted2global(const float* poses, int64_t n_rows, const float* mean_parted, const float* std_parted,
                                      const float* mean_global, const float* std_global, float* out, void* stream) {
    SDT_REQUIRE(poses && mean_parted && std_parted && mean_global && std_global && out && n_rows > 0, "sdt_pose_parted2global: bad arguments");
    const long long total = (long long)n_rows * 2 * K121;
    sdt::launch(pose_parted2global_kernel, dim3(sdt::ceil_div(total, 256)), dim3(256), 0, sdt::as_stream(stream), poses, total, mean_parted, std_parted,
                                                                                            mean_global, std_global, out);
    SDT_LAUNCH_OK("pose_parted2global_kernel");
    return SDT_OK;
}
