// Shared helpers for libsdt_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "sdt_b200.h"

namespace sdt {

void set_error(const char* fmt, ...);

#define SDT_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            sdt::set_error(__VA_ARGS__);       \
            return SDT_ERR_ARG;                \
        }                                      \
    } while (0)

#define SDT_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            sdt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return SDT_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// after a kernel launch (legal during stream capture)
#define SDT_LAUNCH_OK(name)                                                                    \
    do {                                                                                       \
        cudaError_t e_ = cudaPeekAtLastError();                                                \
        if (e_ != cudaSuccess) {                                                               \
            cudaGetLastError();                                                                \
            sdt::set_error("launch of %s failed: %s", name, cudaGetErrorString(e_));           \
            return SDT_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch -----------------------------------------------------------------------------
// A train step is ~250 mostly small, mostly dependent launches.  Every kernel of this library starts with pdl_wait()
// (griddepcontrol.wait: returns once the preceding kernel of the stream has completed and its writes are visible; a no-op
// for a kernel launched without the attribute) followed by pdl_launch_dependents(), and every launch goes through
// sdt::launch(), which sets cudaLaunchAttributeProgrammaticStreamSerialization: the NEXT kernel's CTAs are scheduled, its
// parameters loaded and (tcgen05 kernels) its barriers / TMEM allocation done while this kernel's tail is still running.
// SDT_PDL=0 in the environment switches the attribute off (plain stream order).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);      // errors surface through SDT_LAUNCH_OK (cudaPeekAtLastError)
}

// the same with a thread-block cluster of (cx, cy, cz) CTAs
template <typename... KArgs, typename... Args>
inline void launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cx, int cy, int cz,
                           Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cx;
    attr[0].val.clusterDim.y = cy;
    attr[0].val.clusterDim.z = cz;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }
// derivative of LeakyReLU/ReLU from the sign of the (pre- or post-)activation value; 0 -> slope
__device__ __forceinline__ float leaky_grad(float v, float slope) { return v > 0.f ? 1.f : slope; }

// round-to-nearest (ties away) to TF32's 10 mantissa bits.  tcgen05 kind::tf32 ignores the low 13 bits of its fp32 operands,
// i.e. TRUNCATES: every product comes out low by ~7e-4 on average, a bias that normalisation layers on batch / instance
// statistics cancel but BatchNorm in eval mode does not (1.6e-2 over the 25 layers of the s2g generator, profiles/
// r2_tf32_truncation_bias.txt).  Producers of tensor-core operands therefore store RN-rounded values when asked (out_tf32):
// the MMA then sees exactly representable operands and its error is unbiased.
__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ float out_round(float v, int out_tf32) { return out_tf32 ? tf32_rna(v) : v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sdt
