// Fused mel front end: reflect-pad + hann window + 512-pt FFT + |X|^2 + banded mel projection.
// Replaces torchaudio.transforms.MelSpectrogram as built at core/pipelines/voice2pose.py:27-30 and called at :125.
//
// One warp transforms TWO frames with one 512-point complex radix-2 FFT in shared memory (frame A in the real
// part, frame B in the imaginary part; the two real spectra are separated with the conjugate-symmetry identity),
// then projects the 257 power bins onto the 80 mel bands using the band (start,count,weights) form of the
// filterbank: 468 non-zeros instead of a 257x80 GEMM.  The spectrum never leaves shared memory;
// HBM traffic is the audio read once (L2 absorbs the 2.5x frame overlap) plus the (B,80,T) write:
// 409,704 algorithmic bytes per 64-frame clip.
#include "common.cuh"

namespace {

constexpr int kNfft = 512;
constexpr int kWin = 400;
constexpr int kHop = 160;
constexpr int kMel = 80;
constexpr int kLpad = (kNfft - kWin) / 2;  // 56 zeros each side of the window
constexpr int kCenter = kNfft / 2;         // reflect padding
constexpr int kWarps = 8;
constexpr int kFramesPerCta = kWarps * 4;  // 2 FFTs x 2 frames per warp
constexpr int kPStride = 260;

// 52 KB per CTA: four CTAs (32 warps) per SM, so the 14 x B CTAs of a 64-frame batch run as ONE wave on 148 SMs (with a
// separate 10.5 KB output staging tile it was three per SM and 448 CTAs on 444 slots: a whole second wave for 4 CTAs).
// ncu (profiles/r1_ncu_mel.txt): the kernel is bound by the shared-memory pipe -- 10 M bank-conflict cycles, short-scoreboard +
// MIO-throttle stalls -- not by HBM (1 %) or issue (21 %).  Two layout fixes:
//   * twiddles as one contiguous table PER STAGE (tw[half + pos], 511 entries): the single 256-entry table was read with stride
//     2^(9-stage), i.e. up to 16 different addresses in one bank per load;
//   * the FFT work area padded by one element every 32 (zi()): the bit-reversed scatter of the windowed samples put all 32 lanes
//     of a store into one bank.
constexpr int kZLen = kNfft + kNfft / 32;
__device__ __forceinline__ int zi(int i) { return i + (i >> 5); }
struct __align__(16) MelSmem {
    float2 tw[kNfft];                 // tw[half + pos] = exp(-2 pi i pos / (2 half)), half = 1, 2, ..., 256
    float win[kWin];
    float2 z[kWarps][kZLen];          // FFT work area; after a warp's last transform its first 320 floats stage its 4 x 80 results
    float p[kWarps][2][kPStride];
};

__device__ __forceinline__ int reflect_index(int i, int L) {
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    return i;
}

__global__ void __launch_bounds__(kWarps * 32, 4) mel_kernel(const float* __restrict__ audio, int L, int T,
                                                          const float* __restrict__ window,
                                                          const int32_t* __restrict__ fb_start,
                                                          const int32_t* __restrict__ fb_count,
                                                          const float* __restrict__ fb_weight, int fb_stride,
                                                          float* __restrict__ mel) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MelSmem& s = *reinterpret_cast<MelSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int b = blockIdx.y, t0 = blockIdx.x * kFramesPerCta;
    const float* x = audio + (size_t)b * L;

    for (int q = tid; q < kNfft; q += blockDim.x) {
        if (q == 0) continue;
        const int half = 1 << (31 - __clz(q)), pos = q - half;
        float sn, cs;
        // exp(-2 pi i pos / (2 half)); same argument values as the single table had (pos << (9 - stage)) / 512, exactly
        sincospif(-2.0f * (float)(pos * (kNfft / 2 / half)) / (float)kNfft, &sn, &cs);
        s.tw[q] = make_float2(cs, sn);
    }
    for (int i = tid; i < kWin; i += blockDim.x) s.win[i] = window[i];
    __syncthreads();

    float2* z = s.z[w];
    float keepA[3], keepB[3];             // results of the first transform (bands lane, lane+32, lane+64), kept across the second
    for (int f = 0; f < 2; ++f) {
        const int tA = t0 + w * 4 + 2 * f, tB = tA + 1;
        // windowed frames, written in bit-reversed order
#pragma unroll 4
        for (int i = 0; i < kNfft / 32; ++i) {
            const int n = lane + 32 * i;
            float xa = 0.f, xb = 0.f;
            if (n >= kLpad && n < kLpad + kWin) {
                const float wv = s.win[n - kLpad];
                if (tA < T) xa = x[reflect_index(kHop * tA - kCenter + n, L)] * wv;
                if (tB < T) xb = x[reflect_index(kHop * tB - kCenter + n, L)] * wv;
            }
            z[zi((int)(__brev((unsigned)n) >> 23))] = make_float2(xa, xb);
        }
        __syncwarp();
#pragma unroll 1
        for (int st = 1; st <= 9; ++st) {
            const int half = 1 << (st - 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = lane + 32 * i;
                const int pos = j & (half - 1);
                const int i0 = ((j >> (st - 1)) << st) + pos;
                const int i1 = i0 + half;
                const float2 tw = s.tw[half + pos];
                const float2 u = z[zi(i0)], v = z[zi(i1)];
                const float vr = v.x * tw.x - v.y * tw.y;
                const float vi = v.x * tw.y + v.y * tw.x;
                z[zi(i0)] = make_float2(u.x + vr, u.y + vi);
                z[zi(i1)] = make_float2(u.x - vr, u.y - vi);
            }
            __syncwarp();
        }
        // separate the two real spectra, power
        for (int k = lane; k <= kNfft / 2; k += 32) {
            const float2 zk = z[zi(k)], zn = z[zi((kNfft - k) & (kNfft - 1))];
            const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
            const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
            s.p[w][0][k] = ar * ar + ai * ai;
            s.p[w][1][k] = br * br + bi * bi;
        }
        __syncwarp();
        float* stage = reinterpret_cast<float*>(z);           // [kMel][4] once this warp's transforms are done
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int m = lane + 32 * i;
            float accA = 0.f, accB = 0.f;
            if (m < kMel) {
                const int k0 = fb_start[m], cnt = fb_count[m];
                const float* wt = fb_weight + (size_t)m * fb_stride;
                for (int q = 0; q < cnt; ++q) {
                    const float wq = wt[q];
                    accA = fmaf(wq, s.p[w][0][k0 + q], accA);
                    accB = fmaf(wq, s.p[w][1][k0 + q], accB);
                }
            }
            if (f == 0) {
                keepA[i] = accA;
                keepB[i] = accB;
            } else if (m < kMel) {                             // z is free now: stage all four frames of this warp
                stage[m * 4 + 0] = keepA[i];
                stage[m * 4 + 1] = keepB[i];
                stage[m * 4 + 2] = accA;
                stage[m * 4 + 3] = accB;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // coalesced write-out: 32 consecutive frames of one band per warp-wide store (frame j belongs to warp j / 4)
    for (int e = tid; e < kMel * kFramesPerCta; e += blockDim.x) {
        const int m = e / kFramesPerCta, j = e % kFramesPerCta, t = t0 + j;
        if (t < T) mel[((size_t)b * kMel + m) * T + t] = reinterpret_cast<const float*>(s.z[j >> 2])[m * 4 + (j & 3)];
    }
}

}  // namespace

extern "C" int sdt_mel_fwd(const float* audio, int B, int L, const float* window, const int32_t* fb_start,
                           const int32_t* fb_count, const float* fb_weight, int fb_stride, float* mel, void* stream) {
    SDT_REQUIRE(audio && window && fb_start && fb_count && fb_weight && mel, "sdt_mel_fwd: null pointer");
    SDT_REQUIRE(B > 0 && L > kCenter, "sdt_mel_fwd: need B > 0 and L > %d (reflect padding), got B=%d L=%d", kCenter, B, L);
    SDT_REQUIRE(fb_stride > 0, "sdt_mel_fwd: fb_stride must be positive");
    const int T = 1 + L / kHop;
    static bool attr_set = false;  // idempotent; benign race
    if (!attr_set) {
        SDT_CUDA_OK(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem)));
        SDT_CUDA_OK(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // four CTAs per SM
        attr_set = true;
    }
    dim3 grid(sdt::ceil_div(T, kFramesPerCta), B);
    sdt::launch(mel_kernel, dim3(grid), dim3(kWarps * 32), sizeof(MelSmem), sdt::as_stream(stream), audio, L, T, window, fb_start, fb_count,
                                                                                 fb_weight, fb_stride, mel);
    SDT_LAUNCH_OK("mel_kernel");
    return SDT_OK;
}
