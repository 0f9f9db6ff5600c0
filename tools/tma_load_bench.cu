// Micro-benchmark: TMA load throughput per SM / chip-wide for [rows x 64 bf16] boxes (128 B rows, 128B swizzle) of a row-major
// [T, C] bf16 tensor, as the weight-gradient kernels issue them.  One producer thread per CTA keeps `depth` boxes in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I semantic_segmentation_of_stylegan2_artifacts_b200/csrc \
//        -o tools/build/tma_load_bench tools/tma_load_bench.cu        (analysis aid, not part of the library)
#include <stdio.h>

#include "tc_common.cuh"
using namespace msu;

__global__ void __launch_bounds__(32, 1) k(const __grid_constant__ CUtensorMap tm, int box_bytes, int box_rows, int cboxes, int64_t rows_total, int iters,
                                           int depth, int shared_rows, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; i++) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int slabs = (int)(rows_total / box_rows);
        int slab = shared_rows ? 0 : (int)blockIdx.x * (slabs / (int)gridDim.x);
        {   // issue cost: `depth` loads back to back, then wait for all of them
            long long a0 = clock64();
            for (int s = 0; s < depth; s++) {
                mbar_arrive_expect_tx(&bar[s], box_bytes);
                tma_load_2d(smem + (size_t)s * box_bytes, &tm, &bar[s], 0, (slab + s < slabs ? slab + s : slab + s - slabs) * box_rows);
            }
            long long a1 = clock64();
            for (int s = 0; s < depth; s++) mbar_wait(&bar[s], 0);
            long long a2 = clock64();
            if (blockIdx.x == 0) { out[200] = a1 - a0; out[201] = a2 - a0; }
        }
        long long t0 = clock64();
        uint32_t phase[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
        for (int it = 0; it < iters + depth; it++) {
            const int s = it & (depth - 1);
            const int cbx = 0; (void)cbx;
            if (it >= depth) { mbar_wait(&bar[s], phase[s]); phase[s] ^= 1; }
            if (it < iters) {
                mbar_arrive_expect_tx(&bar[s], box_bytes);
                const int cb = cboxes == 1 ? 0 : (cboxes == 2 ? (it & 1) : (it % cboxes));
                tma_load_2d(smem + (size_t)s * box_bytes, &tm, &bar[s], cb * 64, slab * box_rows);
                if (cb == cboxes - 1) { slab++; if (slab >= slabs) slab = 0; }
            }
        }
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
}

static CUtensorMap make(const void* ptr, int64_t rows, int64_t cols, int box_rows) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    tc_get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return tm;
}

int main() {
    void* buf;
    const size_t bytes = (size_t)1 << 31;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    long long* d;
    cudaMalloc(&d, 256 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int C, box_rows, grid, depth, shared; int64_t rows; const char* name; };
    Cfg cfgs[] = {
        {96, 64, 148, 8, 0, 4194304, "C=96  box 64 rows, HBM stream (conv dZ/X)"},
        {96, 64, 148, 8, 0, 131072, "C=96  box 64 rows, 25 MB (L2 resident)"},
        {96, 64, 148, 8, 1, 131072, "C=96  box 64 rows, all CTAs same rows"},
        {384, 64, 148, 8, 0, 262144, "C=384 box 64 rows, 200 MB stream"},
        {384, 64, 148, 8, 0, 16384, "C=384 box 64 rows, 12.6 MB (L2 resident)"},
        {1536, 64, 148, 8, 0, 16384, "C=1536 box 64 rows, 50 MB (L2 resident)"},
        {1536, 64, 148, 8, 1, 16384, "C=1536 box 64 rows, all CTAs same rows"},
        {1536, 64, 32, 8, 0, 16384, "C=1536 box 64 rows, L2 resident, 32 CTAs"},
        {1536, 64, 8, 8, 0, 16384, "C=1536 box 64 rows, L2 resident, 8 CTAs"},
        {1536, 128, 148, 8, 0, 16384, "C=1536 box 128 rows, L2 resident"},
        {1536, 64, 148, 16, 0, 16384, "C=1536 box 64 rows, L2 resident, depth 16"},
        {1536, 64, 148, 4, 0, 16384, "C=1536 box 64 rows, L2 resident, depth 4"},
        {1536, 256, 148, 4, 0, 16384, "C=1536 box 256 rows, L2 resident, depth 4"},
        {1536, 128, 148, 4, 0, 16384, "C=1536 box 128 rows, L2 resident, depth 4"},
        {1536, 32, 148, 8, 0, 16384, "C=1536 box 32 rows, L2 resident"},
        {1536, 16, 148, 8, 0, 16384, "C=1536 box 16 rows, L2 resident"},
        {96, 128, 148, 8, 0, 131072, "C=96  box 128 rows, 25 MB (L2 resident)"},
        {96, 256, 148, 4, 0, 131072, "C=96  box 256 rows, 25 MB (L2 resident), depth 4"},
        {96, 32, 148, 8, 0, 131072, "C=96  box 32 rows, 25 MB (L2 resident)"},
    };
    for (const Cfg& c : cfgs) {
        CUtensorMap tm = make(buf, c.rows, c.C, c.box_rows);
        const int box_bytes = c.box_rows * 128;
        const int iters = 4000;
        const int cboxes = (c.C + 63) / 64;
        for (int rep = 0; rep < 2; rep++) k<<<c.grid, 32, c.depth * box_bytes + 1024>>>(tm, box_bytes, c.box_rows, cboxes, c.rows, iters, c.depth, c.shared, d);
        long long h[256];
        cudaMemcpy(h, d, 256 * 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        long long mx = 0;
        for (int i = 0; i < c.grid; i++) mx = h[i] > mx ? h[i] : mx;
        const double bpc = (double)iters * box_bytes / mx;
        printf("%-48s grid %3d depth %2d: %6.1f B/clk/SM  %7.0f B/clk chip  (%.0f clk/box; issue %lld clk for %d loads, all landed after %lld)  %s\n", c.name, c.grid, c.depth, bpc, bpc * c.grid, (double)mx / iters, h[200], c.depth, h[201],
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
