// Micro-benchmark: how fast can 148 CTAs write a [M, N] bf16 matrix tile by tile (128 x 192 tiles, the fc1 stage-0 GEMM
// epilogue pattern) with (a) plain coalesced st.global, (b) TMA box stores of different box shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tma_store_bench tools/tma_store_bench.cu
// (analysis aid for the GEMM epilogue design; not part of the library)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    return (EncodeTiledFn)f;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int WARPS = 12;

// each warp owns boxes of [BR rows x BC cols]; tile = 128 x 192; persistent over tiles
template <int BR, int BC>
__global__ void __launch_bounds__(WARPS * 32) tma_store_kernel(const __grid_constant__ CUtensorMap tm, int num_m_tiles, int num_n_tiles, int nbuf) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BOX = BR * BC * 2;
    uint8_t* my = smem + (size_t)warp * 2 * BOX;
    for (int i = lane; i < 2 * BOX / 16; i += 32) reinterpret_cast<uint4*>(my)[i] = make_uint4(warp, lane, i, 0x3f803f80);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    constexpr int BPR = 192 / BC, NB = (128 / BR) * BPR;   // boxes per tile
    const int tiles = num_m_tiles * num_n_tiles;
    int buf = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int mt = tile / num_n_tiles, nt = tile % num_n_tiles;
        for (int b = warp; b < NB; b += WARPS) {
            const int r0 = mt * 128 + (b / BPR) * BR, c0 = nt * 192 + (b % BPR) * BC;
            if (lane == 0) {
                if (nbuf == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            __syncwarp();
            // emulate the epilogue writing the slab
            for (int i = lane; i < BOX / 16; i += 32) reinterpret_cast<uint4*>(my + buf * BOX)[i] = make_uint4(tile, b, i, 0x3f803f80);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(smem_u32(my + buf * BOX)),
                             "r"(c0), "r"(r0) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (nbuf == 2) buf ^= 1;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// plain stores: the 12 warps write the tile as 8 rows x 64 B per instruction (what the mapped epilogue does), or full rows
__global__ void __launch_bounds__(WARPS * 32) st_kernel(uint4* out, int N, int num_m_tiles, int num_n_tiles, int mode) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = num_m_tiles * num_n_tiles;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int mt = tile / num_n_tiles, nt = tile % num_n_tiles;
        if (mode == 0) {
            // chunk = 32 rows x 32 cols; instruction = 8 rows x 64 B
            for (int ch = warp; ch < 24; ch += WARPS) {
                const int r0 = mt * 128 + (ch / 6) * 32, c0 = nt * 192 + (ch % 6) * 32;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int r = r0 + (lane >> 2) + 8 * j, c = c0 + (lane & 3) * 8;
                    out[((size_t)r * N + c) / 8] = make_uint4(tile, ch, j, lane);
                }
            }
        } else {
            // row-contiguous: a warp instruction covers 1 row x 192 cols (384 B = 24 lanes x 16 B) -> use 32 lanes over 4 rows x 128 B
            for (int rr = warp; rr < 128; rr += WARPS) {
                const int r = mt * 128 + rr;
                if (lane < 24) out[((size_t)r * N + nt * 192) / 8 + lane] = make_uint4(tile, rr, 0, lane);
            }
        }
    }
}

__global__ void fill_kernel(uint4* out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = make_uint4(i, 1, 2, 3);
}

template <int BR, int BC>
static void run_tma(void* out, int M, int N, CUtensorMapSwizzle sw, int nbuf, const char* name) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {BC, BR};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
    const int smem = WARPS * 2 * BR * BC * 2;
    cudaFuncSetAttribute(tma_store_kernel<BR, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(e0);
        tma_store_kernel<BR, BC><<<148, WARPS * 32, smem>>>(tm, M / 128, N / 192, nbuf);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-28s nbuf=%d %8.1f us  %7.1f GB/s  %s\n", name, nbuf, best * 1e3, (double)M * N * 2 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const int M = 262144, N = 384;
    void* out;
    cudaMalloc(&out, (size_t)M * N * 2);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = -1; mode < 2; mode++) {
        float best = 1e9;
        for (int it = 0; it < 5; it++) {
            cudaEventRecord(e0);
            if (mode < 0) fill_kernel<<<148 * 8, 256>>>((uint4*)out, (size_t)M * N / 8);
            else st_kernel<<<148, WARPS * 32>>>((uint4*)out, N, M / 128, N / 192, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("%-28s        %8.1f us  %7.1f GB/s\n", mode < 0 ? "fill (grid-stride 16B)" : (mode == 0 ? "st 8 rows x 64B / instr" : "st 1 row x 384B / instr"), best * 1e3,
               (double)M * N * 2 / best / 1e6);
    }
    for (int nbuf = 1; nbuf <= 2; nbuf++) {
        run_tma<32, 32>(out, M, N, CU_TENSOR_MAP_SWIZZLE_64B, nbuf, "tma 32x32 (64B rows) sw64");
        run_tma<32, 64>(out, M, N, CU_TENSOR_MAP_SWIZZLE_128B, nbuf, "tma 32x64 (128B rows) sw128");
        run_tma<16, 64>(out, M, N, CU_TENSOR_MAP_SWIZZLE_128B, nbuf, "tma 16x64 (128B rows) sw128");
        run_tma<64, 64>(out, M, N, CU_TENSOR_MAP_SWIZZLE_128B, nbuf, "tma 64x64 (128B rows) sw128");
        run_tma<32, 32>(out, M, N, CU_TENSOR_MAP_SWIZZLE_NONE, nbuf, "tma 32x32 (64B rows) none");
        run_tma<32, 64>(out, M, N, CU_TENSOR_MAP_SWIZZLE_NONE, nbuf, "tma 32x64 (128B rows) none");
    }
    return 0;
}
