// Micro-benchmark: tcgen05.mma issue rate (clocks per MMA) for K-major vs MN-major shared-memory operands, M=128,
// N in {32, 96, 128, 256}, K=16, bf16.  One CTA per SM, one issuing thread, operands are fixed (uninitialised) tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I semantic_segmentation_of_stylegan2_artifacts_b200/csrc \
//        -o tools/build/mma_rate_bench tools/mma_rate_bench.cu      (analysis aid, not part of the library)
#include <stdio.h>

#include "tc_common.cuh"
using namespace msu;

__global__ void __launch_bounds__(128, 1) k(int N, int a_mn, int b_mn, int iters, int kstep_bytes_a, int kstep_bytes_b, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t ad = a_mn ? make_desc_mnmajor_sw128(a0 + ks * kstep_bytes_a, 8192) : make_desc_kmajor_sw128(a0 + ks * 32);
                const uint64_t bd = b_mn ? make_desc_mnmajor_sw128(b0 + ks * kstep_bytes_b, 8192) : make_desc_kmajor_sw128(b0 + ks * 32);
                tc_mma_bf16(tmem, ad, bd, idesc, 1);
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// A-operand collector reuse: groups of 3 MMAs share one A tile (different B tiles and accumulators), as the conv weight-gradient
// kernel issues them (A = dZ slice, B = X slab shifted per tap).  coll=1 tags them fill / use / lastuse.
#define MMA_COLL(QUAL)                                                                                                       \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                         \
                 "tcgen05.mma.cta_group::1.kind::f16" QUAL " [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), \
                 "r"(1u) : "memory")
template <int coll, int same_d>
__global__ void __launch_bounds__(128, 1) k3(int N, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
        const uint64_t ad0 = make_desc_mnmajor_sw128(smem_u32(smem), 8192), bd0 = make_desc_mnmajor_sw128(smem_u32(smem + 48 * 1024), 8192);
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t ad = ad0 + (uint64_t)(ks * 128);
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const uint64_t bd = bd0 + (uint64_t)(ks * 128 + t * 8);
                    const uint32_t d = same_d ? tmem : tmem + t * N;
                    if (!coll) MMA_COLL("");
                    else if (t == 0) MMA_COLL(".collector::a::fill");
                    else if (t == 1) MMA_COLL(".collector::a::use");
                    else MMA_COLL(".collector::a::lastuse");
                }
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// The conv kernel's K block (mode 4): operands as 32-channel chunks = 64 B rows, SWIZZLE_64B; 12 MMAs = 2 image rows x 3 pixel
// shifts (+64 B on the A start address) x 2 K halves (+32 B), accumulators r * 128.  variant 0: that pattern; 1: the same MMAs
// with an unshifted A start (dx = 0 always); 2: SWIZZLE_128B descriptors over the same bytes (rate reference).
__device__ __forceinline__ uint64_t desc_kmajor_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
template <int VAR>
__global__ void __launch_bounds__(128, 1) kconv(int N, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
        const uint64_t ad = VAR == 2 ? make_desc_kmajor_sw128(a0) : desc_kmajor_sw64(a0);
        const uint64_t bd = VAR == 2 ? make_desc_kmajor_sw128(b0) : desc_kmajor_sw64(b0);
        const uint32_t bstep = (uint32_t)(N * 4), rstep = 8704 >> 4;
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint64_t ar = ad + (uint64_t)(r * rstep);
                const uint32_t d = tmem + r * 128;
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const uint64_t a_ = VAR == 0 ? ar + 4 * dx : ar;
                    tc_mma_bf16(d, a_, bd + dx * bstep, idesc, 1);
                    tc_mma_bf16(d, a_ + 2, bd + dx * bstep + 2, idesc, 1);
                }
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

int main(int argc, char** argv) {
    if (argc > 1) {          // conv K-block pattern only
        long long* d;
        cudaMalloc(&d, 8);
        const int iters = 2000;
        for (int N : {48, 96, 128}) {
            for (int v = 0; v < 3; v++) {
                auto kern = v == 0 ? kconv<0> : v == 1 ? kconv<1> : kconv<2>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
                kern<<<148, 128, 100 * 1024>>>(N, iters, d);
                long long h = 0;
                cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                printf("conv K block, N=%3d, %s : %7.1f clk/MMA  (MMA ideal %5.1f)  %s\n", N,
                       v == 0 ? "64B swizzle, shifted A starts" : v == 1 ? "64B swizzle, unshifted       " : "128B swizzle descriptors     ",
                       (double)h / (iters * 12), 128.0 * N * 16 * 2 / 8192, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
        return 0;
    }
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2000;
    for (int N : {32, 64, 96, 128, 192, 256}) {
        for (int mode = 0; mode < 4; mode++) {
            const int a_mn = mode & 1, b_mn = mode >> 1;
            k<<<148, 128, 100 * 1024>>>(N, a_mn, b_mn, iters, 2048, 2048, d);
            long long h = 0;
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            const double clk = (double)h / (iters * 4);
            printf("N=%3d A %s B %s : %7.1f clk/MMA  (ideal %5.1f)  %s\n", N, a_mn ? "MN-major" : "K-major ", b_mn ? "MN-major" : "K-major ", clk,
                   128.0 * N * 16 * 2 / 8192, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    for (int N : {32, 96, 128, 160}) {
        for (int v = 0; v < 4; v++) {
            auto kern = v == 0 ? k3<0, 0> : v == 1 ? k3<1, 0> : v == 2 ? k3<0, 1> : k3<1, 1>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            kern<<<148, 128, 100 * 1024>>>(N, iters, d);
            long long h = 0;
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("3 MMAs per A tile, N=%3d, %s, collector %s : %7.1f clk/MMA  (MMA ideal %5.1f, smem ideal %5.1f / %5.1f with reuse)  %s\n", N,
                   (v & 2) ? "one accumulator   " : "three accumulators", (v & 1) ? "fill/use/lastuse" : "default         ",
                   (double)h / (iters * 12), 128.0 * N * 16 * 2 / 8192, (4096.0 + N * 32) / 128, (4096.0 / 3 + N * 32) / 128,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
