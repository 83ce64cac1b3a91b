// All-reduce(sum) of the flat fp32 gradient buffer over the GPUs of one NVSwitch node, through PEER MEMORY instead of NCCL
// (reference: the DDP gradient all-reduce behind `loss.backward()`, core/pipelines/voice2pose.py:222-223,298-309, and the scalar
// reduce of trainer.py:323-327).
//
// Every rank holds the buffer at the same offset of a symmetric allocation (torch.distributed._symmetric_memory supplies the
// allocation, the peer pointers / the NVLS multicast address and the cross-GPU barrier; this file is the data path).  Rank r owns
// shard r of the element range: it sums shard r of all W buffers in rank order 0..W-1 and writes the sum back into shard r of all W
// buffers.  Ranks touch disjoint shards, so the exchange is in place, needs no staging copy, and every rank ends up with
// bit-identical values (each element is reduced exactly once, in a fixed order).
//   * multicast_ptr != 0 (NVLS): one `multimem.ld_reduce.add.v4.f32` pulls the sum of the W copies through the switch, one
//     `multimem.st.v4.f32` broadcasts it: 2 x 16 B of NVLink traffic per 4 elements instead of 2 x (W-1) x 16 B.
//   * else: W vector loads over NVLink (peer pointers), W vector stores.
// The caller brackets the launch with two cross-GPU barriers (gradients of all ranks final before; all shards written after).
// The loss / metric scalars ride along: block 0 also sums `scal_n` doubles of every rank's scalar block into a LOCAL output.
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_f4(float4* p, float4 v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f4(float4* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int MAX_WORLD = 16;
struct PeerPtrs {
    float4* p[MAX_WORLD];
    const double* s[MAX_WORLD];
};

template <bool MC>
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(const PeerPtrs pp, float4* mc, int rank, int W, long long n4, double* scal_dst,
                                                            int scal_n) {
    sdt::pdl_wait();
    sdt::pdl_launch_dependents();
    const long long per = (n4 + W - 1) / W;
    const long long lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = 4;                       // independent requests in flight per thread (remote latency is 2-3 us)
    for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
        float4 acc[U];
        if (MC) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u * stride < hi) acc[u] = mc_ld_reduce_f4(mc + i0 + u * stride);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u * stride < hi) mc_st_f4(mc + i0 + u * stride, acc[u]);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u * stride < hi) acc[u] = ld_f4(pp.p[0] + i0 + u * stride);
            for (int r = 1; r < W; ++r) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i0 + u * stride < hi) v[u] = ld_f4(pp.p[r] + i0 + u * stride);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
                }
            }
            for (int r = 0; r < W; ++r) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i0 + u * stride < hi) st_f4(pp.p[r] + i0 + u * stride, acc[u]);
            }
        }
    }
    if (blockIdx.x == 0 && scal_dst != nullptr && (int)threadIdx.x < scal_n) {
        double s = 0.0;
        for (int r = 0; r < W; ++r) {
            double v;
            asm volatile("ld.global.relaxed.sys.f64 %0, [%1];" : "=d"(v) : "l"(pp.s[r] + threadIdx.x) : "memory");
            s += v;
        }
        scal_dst[threadIdx.x] = s;
    }
}

}  // namespace

extern "C" int sdt_p2p_allreduce(const uint64_t* peer_ptrs, uint64_t multicast_ptr, int rank, int world, long long n,
                                 const uint64_t* scal_peer_ptrs, double* scal_dst, int scal_n, void* stream) {
    SDT_REQUIRE(peer_ptrs != nullptr && world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "sdt_p2p_allreduce: bad rank / world (%d / %d)", rank, world);
    SDT_REQUIRE(n > 0 && n % 4 == 0, "sdt_p2p_allreduce: the element count must be a positive multiple of 4 (n=%lld)", n);
    SDT_REQUIRE(scal_n >= 0 && scal_n <= 512 && (scal_n == 0 || (scal_peer_ptrs && scal_dst)), "sdt_p2p_allreduce: bad scalar block");
    PeerPtrs pp{};
    for (int r = 0; r < world; ++r) {
        SDT_REQUIRE(peer_ptrs[r] != 0 && (peer_ptrs[r] & 15) == 0, "sdt_p2p_allreduce: peer pointer %d is null or not 16-byte aligned", r);
        pp.p[r] = reinterpret_cast<float4*>(peer_ptrs[r]);
        pp.s[r] = scal_n ? reinterpret_cast<const double*>(scal_peer_ptrs[r]) : nullptr;
    }
    const long long n4 = n / 4, per = (n4 + world - 1) / world;
    int grid = sdt::ceil_div(per, 512 * 4);           // ~4 float4 per thread
    if (grid > 592) grid = 592;
    if (grid < 1) grid = 1;
    double* sd = scal_n ? scal_dst : nullptr;
    if (multicast_ptr != 0) {
        SDT_REQUIRE((multicast_ptr & 15) == 0, "sdt_p2p_allreduce: multicast address not 16-byte aligned");
        sdt::launch(p2p_allreduce_kernel<true>, dim3(grid), dim3(512), 0, sdt::as_stream(stream), pp, reinterpret_cast<float4*>(multicast_ptr), rank,
                    world, n4, sd, scal_n);
    } else {
        sdt::launch(p2p_allreduce_kernel<false>, dim3(grid), dim3(512), 0, sdt::as_stream(stream), pp, static_cast<float4*>(nullptr), rank, world, n4,
                    sd, scal_n);
    }
    SDT_LAUNCH_OK("p2p_allreduce_kernel");
    return SDT_OK;
}
