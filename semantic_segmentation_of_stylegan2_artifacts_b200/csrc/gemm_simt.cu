// Generic fp32-accumulate SIMT GEMM with mapped operands and fused epilogues.
//
//   C[m,n] = epilogue( sum_k A(m,k) * B(n,k) )
//
// This is the fp32 "parity mode" engine (TF32/bf16-free, so the <=1e-3 gate of BASELINE.json's north_star
// can be met through 28-52 Swin blocks) and the fallback for operand patterns the tcgen05 path
// (gemm_tc.cu) does not cover yet.  Operand/row maps fold torch.cat, window partition/reverse, einops
// depth-to-space and the 3x3 im2col into index math (see include/msunet_b200.h).
#include "common.cuh"

namespace msu {

__device__ __forceinline__ float fetch_operand(const MsuOperand& op, int64_t i, int64_t kk, int64_t I, int64_t Kd) {
    if (i >= I || kk >= Kd) return 0.f;
    int64_t r = op.orient ? kk : i;
    int c = (int)(op.orient ? i : kk);
    const void* base = op.ptr;
    int64_t ld = op.ld;
    if (op.ptr2 != nullptr && c >= op.k_split) {
        base = op.ptr2;
        ld = op.ld2;
        c -= op.k_split;
    }
    RowCol rc = map_rc(op.map, op.geo, r, c);
    if (rc.row < 0) return 0.f;
    float v = ld_as_f(base, rc.row * ld + rc.col, op.dtype);
    if (op.rowscale != nullptr) v *= op.rowscale[rc.row / op.rows_per_sample];
    return v;
}

__device__ __forceinline__ void epilogue_store(const MsuEpilogue& E, int64_t m, int n, float acc) {
    float v = acc;
    if (E.bias != nullptr) v += E.bias[n];
    RowCol rc = map_rc(E.map, E.geo, m, n);
    if (rc.row < 0) return;  // padding token: output dropped (TV:models/swin_transformer.py:227)
    const int64_t o = rc.row * E.ldc + rc.col;
    if (E.Cpre != nullptr) st_from_f(E.Cpre, o, E.dtype, E.act == 2 ? gelu_grad_f(v) : v);   // act 2: Cpre = GELU'(pre)
    if (E.act == 1 || E.act == 2) v = gelu_f(v);
    if (E.H != nullptr) v *= (E.act == 3 ? ld_as_f(E.H, m * E.ldh + n, E.dtype) : gelu_grad_f(ld_as_f(E.H, m * E.ldh + n, E.dtype)));
    if (E.rowscale != nullptr) v *= E.rowscale[rc.row / E.rows_per_sample];
    if (E.R != nullptr) v += ld_as_f(E.R, rc.row * E.ldr + rc.col, E.dtype);
    if (E.out_f32) {
        float* C = reinterpret_cast<float*>(E.C);
        C[o] = E.accumulate ? C[o] + v : v;
    } else {
        if (E.accumulate) v += ld_as_f(E.C, o, E.dtype);
        st_from_f(E.C, o, E.dtype, v);
    }
}

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int TM = 8, TN = 4;

__global__ void __launch_bounds__(NT) gemm_simt_kernel(MsuOperand A, MsuOperand B, MsuEpilogue E, int64_t M, int64_t N,
                                                      int64_t K, int64_t k_per_split, float* splitk_ws) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int64_t n0 = (int64_t)blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = imin(K, kbeg + k_per_split);
    const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads -> (TN*16=64) x (TM*16=128)

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    float ra[BM * BK / NT], rb[BN * BK / NT];
    auto gload = [&](int64_t k0) {
#pragma unroll
        for (int j = 0; j < BM * BK / NT; j++) {
            const int idx = tid + NT * j;
            int i, k;
            if (A.orient == 0) { k = idx % BK; i = idx / BK; } else { i = idx % BM; k = idx / BM; }
            ra[j] = (k0 + k < kend) ? fetch_operand(A, m0 + i, k0 + k, M, K) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < BN * BK / NT; j++) {
            const int idx = tid + NT * j;
            int i, k;
            if (B.orient == 0) { k = idx % BK; i = idx / BK; } else { i = idx % BN; k = idx / BN; }
            rb[j] = (k0 + k < kend) ? fetch_operand(B, n0 + i, k0 + k, N, K) : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < BM * BK / NT; j++) {
            const int idx = tid + NT * j;
            int i, k;
            if (A.orient == 0) { k = idx % BK; i = idx / BK; } else { i = idx % BM; k = idx / BM; }
            As[buf][k][i] = ra[j];
        }
#pragma unroll
        for (int j = 0; j < BN * BK / NT; j++) {
            const int idx = tid + NT * j;
            int i, k;
            if (B.orient == 0) { k = idx % BK; i = idx / BK; } else { i = idx % BN; k = idx / BN; }
            Bs[buf][k][i] = rb[j];
        }
    };

    int buf = 0;
    if (kbeg < kend) {
        gload(kbeg);
        sstore(0);
    }
    __syncthreads();
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) gload(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int64_t n = n0 + tx * 4 + j;
            if (n >= N) continue;
            if (splitk_ws != nullptr) splitk_ws[((int64_t)blockIdx.z * M + m) * N + n] = acc[i][j];
            else epilogue_store(E, m, (int)n, acc[i][j]);
        }
    }
}

// `sscale` (optional): per-sample scale of the contraction rows (stochastic depth on the gradient rows); every split
// lies inside one sample (splits_per_sample of them each), so the scale is applied to whole partials.
__global__ void splitk_reduce_kernel(MsuEpilogue E, int64_t M, int64_t N, int splits, const float* ws, const float* bws, int brows,
                                     const float* sscale, int splits_per_sample) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) {
        // bias-gradient partials [split][M] of the fused column sums (fixed order: deterministic)
        const int64_t m = idx - M * N;
        if (bws != nullptr && m < M) {
            float b = 0.f;
            for (int z = 0; z < brows; z++) {
                const float v = bws[(int64_t)z * M + m];
                b += sscale != nullptr ? v * sscale[z / splits_per_sample] : v;
            }
            E.colsum[m] = b;
        }
        return;
    }
    float s = 0.f;
    if (sscale != nullptr) {
        for (int z = 0; z < splits; z++) s = fmaf(ws[(int64_t)z * M * N + idx], sscale[z / splits_per_sample], s);
    } else {
        for (int z = 0; z < splits; z++) s += ws[(int64_t)z * M * N + idx];  // fixed order: deterministic
    }
    epilogue_store(E, idx / N, (int)(idx % N), s);
}

void launch_splitk_reduce(const MsuEpilogue& E, int64_t M, int64_t N, int splits, const float* ws, cudaStream_t st,
                          const float* bws, int brows, const float* sscale, int splits_per_sample) {
    const int64_t tot = M * N + (bws != nullptr ? M : 0);
    splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(E, M, N, splits, ws, bws, brows, sscale, splits_per_sample);
    count_launch();
}

int gemm_simt(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t M, int64_t N, int64_t K,
              float* splitk_ws, int64_t splitk_ws_elems, cudaStream_t st) {
    const int64_t gm = (M + BM - 1) / BM, gn = (N + BN - 1) / BN;
    int splits = 1;
    if (splitk_ws != nullptr && gm * gn < 2 * num_sms() && K >= 4096) {
        splits = (int)imin((4 * num_sms() + gm * gn - 1) / (gm * gn), (K + 1023) / 1024);
        while (splits > 1 && (int64_t)splits * M * N > splitk_ws_elems) splits--;
    }
    int64_t kps = (K + splits - 1) / splits;
    kps = (kps + BK - 1) / BK * BK;
    splits = (int)((K + kps - 1) / kps);
    MSU_REQUIRE(gn <= 65535 && splits <= 65535, "gemm_simt: grid too large (N=%lld)", (long long)N);
    dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
    gemm_simt_kernel<<<grid, NT, 0, st>>>(*A, *B, *E, M, N, K, kps, splits > 1 ? splitk_ws : nullptr);
    count_launch();
    if (splits > 1) {
        const int64_t tot = M * N;
        splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(*E, M, N, splits, splitk_ws, nullptr, 0, nullptr, 1);
        count_launch();
    }
    return check_launch("gemm_simt");
}

// ---- column sums (bias gradients) ---------------------------------------------------------------
__global__ void colsum_partial_kernel(MsuOperand X, int64_t M, int64_t N, int64_t rows_per_block, float* ws) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = imin(M, r0 + rows_per_block);
    float s = 0.f;
    for (int64_t r = r0; r < r1; r++) s += fetch_operand(X, r, n, M, N);
    ws[(int64_t)blockIdx.y * N + n] = s;
}
// fast path: dense [M, N] rows, 64/128-bit loads, each block sums a row slab (HBM-bound: M*N*e bytes)
template <typename T>
__global__ void __launch_bounds__(256) colsum_dense_kernel(const T* __restrict__ X, int64_t M, int N, int64_t ld,
                                                          int64_t rows_per_block, float* __restrict__ ws) {
    __shared__ float4 sm[8][32];
    const int vc = blockIdx.x * 32 + threadIdx.x;          // vector column (4 elements)
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = imin(M, r0 + rows_per_block);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (vc * 4 < N) {
        int64_t r = r0 + threadIdx.y;
        for (; r + 8 < r1; r += 16) {                       // two independent loads in flight
            const float4 u = Vec4<T>::ld(X + r * ld + vc * 4);
            const float4 v = Vec4<T>::ld(X + (r + 8) * ld + vc * 4);
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
        }
        for (; r < r1; r += 8) {
            const float4 u = Vec4<T>::ld(X + r * ld + vc * 4);
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
        }
    }
    sm[threadIdx.y][threadIdx.x] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    __syncthreads();
    if (threadIdx.y == 0 && vc * 4 < N) {
        float4 s = sm[0][threadIdx.x];
#pragma unroll
        for (int i = 1; i < 8; i++) { const float4 t = sm[i][threadIdx.x]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        *reinterpret_cast<float4*>(ws + (int64_t)blockIdx.y * N + vc * 4) = s;
    }
}

// out[n] = sum_p ws[p][n]: 32 columns x 8 partial-row lanes per block, fixed-order combine (deterministic)
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* ws, int parts, int64_t N, float* out, int accumulate) {
    __shared__ float sm[8][33];
    const int n = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (n < N)
        for (int p = threadIdx.y; p < parts; p += 8) s += ws[(int64_t)p * N + n];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) t += sm[i][threadIdx.x];
        out[n] = accumulate ? out[n] + t : t;
    }
}

}  // namespace msu

using namespace msu;

extern "C" int msu_colsum(const MsuOperand* X, int64_t M, int64_t N, float* out, int accumulate, float* ws,
                          int64_t ws_elems, void* stream) {
    MSU_REQUIRE(X && out && ws, "msu_colsum: null pointer");
    MSU_REQUIRE(X->orient == 0, "msu_colsum: operand must be [row, col]");
    if (X->map == MSU_MAP_NONE && X->rowscale == nullptr && X->ptr2 == nullptr && N % 4 == 0 && X->ld % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(X->ptr) & 15) == 0) {
        int parts = (int)imin(4 * num_sms(), imax(1, M / 128));
        while (parts > 1 && (int64_t)parts * N > ws_elems) parts--;
        const int64_t rpb = (M + parts - 1) / parts;
        parts = (int)((M + rpb - 1) / rpb);
        cudaStream_t st = (cudaStream_t)stream;
        dim3 grid((unsigned)((N / 4 + 31) / 32), (unsigned)parts), block(32, 8);
        if (X->dtype == MSU_F32) colsum_dense_kernel<float><<<grid, block, 0, st>>>((const float*)X->ptr, M, (int)N, X->ld, rpb, ws);
        else colsum_dense_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)X->ptr, M, (int)N, X->ld, rpb, ws);
        colsum_final_kernel<<<(unsigned)((N + 31) / 32), dim3(32, 8), 0, st>>>(ws, parts, N, out, accumulate);
        count_launch(2);
        return check_launch("msu_colsum");
    }
    int parts = (int)imin(256, imax(1, M / 64));
    while (parts > 1 && (int64_t)parts * N > ws_elems) parts--;
    const int64_t rpb = (M + parts - 1) / parts;
    parts = (int)((M + rpb - 1) / rpb);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)parts);
    colsum_partial_kernel<<<grid, 128, 0, st>>>(*X, M, N, rpb, ws);
    colsum_final_kernel<<<(unsigned)((N + 31) / 32), dim3(32, 8), 0, st>>>(ws, parts, N, out, accumulate);
    count_launch(2);
    return check_launch("msu_colsum");
}
