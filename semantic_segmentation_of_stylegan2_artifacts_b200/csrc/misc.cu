// Layout / dtype preparation kernels: weight shadows (fp32 master -> compute dtype, rearranged for the
// GEMM operand conventions), the 4x4 patch im2col of PatchEmbed, casts and adds.  All memory-bound.
#include "common.cuh"

namespace msu {

template <typename T>
__global__ void prep_weight_kernel(int mode, const float* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t C) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    switch (mode) {
        case 0: {  // cast [R,C]
            if (idx < R * C) dst[idx] = from_f<T>(src[idx]);
            break;
        }
        case 1: {  // dst[C,R] = src[R,C]^T   (idx walks dst)
            if (idx < R * C) {
                const int64_t c = idx / R, r = idx % R;
                dst[idx] = from_f<T>(src[r * C + c]);
            }
            break;
        }
        case 2: {  // conv [co=R, ci=C, 3,3] -> [co, (tap ci)]
            if (idx < R * C * 9) {
                const int64_t co = idx / (9 * C);
                const int rem = (int)(idx % (9 * C)), tap = rem / (int)C, ci = rem % (int)C;
                dst[idx] = from_f<T>(src[(co * C + ci) * 9 + tap]);
            }
            break;
        }
        case 3: {  // conv [co=R, ci=C, 3,3] -> [ci, (tap' co)], tap' = 8 - tap (dgrad = correlation with flipped kernel)
            if (idx < R * C * 9) {
                const int64_t ci = idx / (9 * R);
                const int rem = (int)(idx % (9 * R)), tapf = rem / (int)R, co = rem % (int)R;
                dst[idx] = from_f<T>(src[((int64_t)co * C + ci) * 9 + (8 - tapf)]);
            }
            break;
        }
        case 5: {  // patch-embed conv [E=R, 48] -> [E, 64], zero padded K
            if (idx < R * 64) {
                const int64_t e = idx / 64;
                const int k = (int)(idx % 64);
                dst[idx] = from_f<T>(k < 48 ? src[e * 48 + k] : 0.f);
            }
            break;
        }
    }
}
// fp32 -> fp32 gradient re-layouts
__global__ void prep_grad_kernel(int mode, const float* __restrict__ src, float* __restrict__ dst, int64_t R, int64_t C) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (mode == 4) {  // [co, (tap ci)] -> [co, ci, 3, 3]   (idx walks dst)
        if (idx < R * C * 9) {
            const int64_t co = idx / (9 * C);
            const int rem = (int)(idx % (9 * C)), ci = rem / 9, tap = rem % 9;
            dst[idx] = src[(co * 9 + tap) * C + ci];
        }
    } else if (mode == 6) {  // [E, 64] -> [E, 48]
        if (idx < R * 48) dst[idx] = src[(idx / 48) * 64 + idx % 48];
    }
}

// ---- every weight shadow of a model in ONE launch ------------------------------------------------------------------------------
// block -> (job, tile) through device maps (like the AdamW step).  Pair jobs (mode 0) read a 32 x 128 tile of the fp32 master once
// and write both the plain cast [R, C] and the transposed cast [C, R] (through shared memory, both coalesced); the few re-laid-out
// tensors (conv / patch-embed weights, modes 2 / 3 / 5 of msu_prep_weight) go through the same index formulas, SH_CHUNK outputs
// per block.
constexpr int SH_TR = 32, SH_TC = 128, SH_CHUNK = 4096, SH_THREADS = 256;

__device__ __forceinline__ void shadow_store(void* dst, int dtype, int64_t i, float v) {
    if (dtype == MSU_BF16) reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16(v);
    else reinterpret_cast<float*>(dst)[i] = v;
}

__global__ void __launch_bounds__(SH_THREADS) refresh_shadows_kernel(const MsuShadowJob* __restrict__ jobs,
                                                                    const int32_t* __restrict__ blk_job,
                                                                    const int32_t* __restrict__ blk_tile) {
    __shared__ float tile[SH_TR][SH_TC + 1];
    const MsuShadowJob j = jobs[blk_job[blockIdx.x]];
    const int t = blk_tile[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t R = j.R, C = j.C;
    if (j.mode == 0) {
        const int tiles_c = (int)((C + SH_TC - 1) / SH_TC);
        const int64_t r0 = (int64_t)(t / tiles_c) * SH_TR, c0 = (int64_t)(t % tiles_c) * SH_TC;
#pragma unroll
        for (int i = 0; i < SH_TR / 8; i++) {
            const int rr = warp + 8 * i;
            const int64_t r = r0 + rr;
#pragma unroll
            for (int q = 0; q < SH_TC / 32; q++) {
                const int cc = lane + 32 * q;
                const int64_t c = c0 + cc;
                const bool in = r < R && c < C;
                const float v = in ? __ldg(j.src + r * C + c) : 0.f;
                tile[rr][cc] = v;
                if (in && j.dst != nullptr) shadow_store(j.dst, j.dtype, r * C + c, v);
            }
        }
        __syncthreads();
        if (j.dst_t != nullptr) {
            const int64_t r = r0 + lane;
#pragma unroll 4
            for (int i = 0; i < SH_TC / 8; i++) {
                const int cc = warp + 8 * i;
                const int64_t c = c0 + cc;
                if (c < C && r < R) shadow_store(j.dst_t, j.dtype, c * R + r, tile[lane][cc]);
            }
        }
        return;
    }
    const int64_t n = j.mode == 5 ? R * 64 : R * C * 9;
    const int64_t base = (int64_t)t * SH_CHUNK;
    for (int k = 0; k < SH_CHUNK / SH_THREADS; k++) {
        const int64_t idx = base + k * SH_THREADS + threadIdx.x;
        if (idx >= n) break;
        float v;
        if (j.mode == 2) {          // conv [co=R, ci=C, 3,3] -> [co, (tap ci)]
            const int64_t co = idx / (9 * C);
            const int rem = (int)(idx % (9 * C)), tap = rem / (int)C, ci = rem % (int)C;
            v = j.src[(co * C + ci) * 9 + tap];
        } else if (j.mode == 3) {   // conv -> [ci, (tap' co)], tap' = 8 - tap
            const int64_t ci = idx / (9 * R);
            const int rem = (int)(idx % (9 * R)), tapf = rem / (int)R, co = rem % (int)R;
            v = j.src[((int64_t)co * C + ci) * 9 + (8 - tapf)];
        } else {                    // patch-embed [E=R, 48] -> [E, 64], zero padded K
            const int64_t e = idx / 64;
            const int kk = (int)(idx % 64);
            v = kk < 48 ? j.src[e * 48 + kk] : 0.f;
        }
        shadow_store(j.dst, j.dtype, idx, v);
    }
}

// image [B,3,S,S] NCHW fp32 -> rows [(b, py, px), 64]; col = c*16 + ky*4 + kx (Conv2d weight order), cols 48..63 zero
template <typename T>
__global__ void patchify4_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int S) {
    const int P = S / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (token, c, ky): 4 contiguous pixels
    const int64_t total = (int64_t)B * P * P * 16;
    if (idx >= total) return;
    const int64_t tok = idx / 16;
    const int sub = (int)(idx % 16);
    T* o = out + tok * 64 + sub * 4;
    if (sub >= 12) {
        Vec4<T>::st(o, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    const int c = sub / 4, ky = sub % 4;
    const int64_t b = tok / (P * P);
    const int t = (int)(tok % (P * P)), py = t / P, px = t % P;
    const float4 v = *reinterpret_cast<const float4*>(img + ((b * 3 + c) * S + (py * 4 + ky)) * (int64_t)S + px * 4);
    Vec4<T>::st(o, v);
}

// dst[m, n] = src[map(m, n)] * rowscale[sample]  (zero where the map has no source): materialises a
// window-ordered / inverse-depth-to-space view so that the following GEMMs read dense, TMA-able rows.
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t M, int N, int64_t ld_src, int map,
                                   MsuOperand geo_holder) {
    const int nv = N / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * nv) return;
    const int64_t m = idx / nv;
    const int n = (int)(idx - m * nv) * 4;
    const RowCol rc = map_rc(map, geo_holder.geo, m, n);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rc.row >= 0) {
        v = Vec4<T>::ld(src + rc.row * ld_src + rc.col);
        if (geo_holder.rowscale != nullptr) {
            const float s = geo_holder.rowscale[rc.row / geo_holder.rows_per_sample];
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        }
    }
    Vec4<T>::st(dst + m * N + n, v);
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = from_f<TD>(to_f<TS>(s[i]));
}
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        const float4 x = Vec4<T>::ld(a + i * 4), z = Vec4<T>::ld(b + i * 4);
        Vec4<T>::st(y + i * 4, make_float4(x.x + z.x, x.y + z.y, x.z + z.z, x.w + z.w));
    }
}

}  // namespace msu

using namespace msu;

extern "C" int msu_prep_weight(int mode, int dst_dtype, const float* src, void* dst, int64_t R, int64_t C, void* stream) {
    MSU_REQUIRE(src && dst && R > 0 && C > 0, "msu_prep_weight: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n;
    if (mode == 0 || mode == 1) n = R * C;
    else if (mode == 2 || mode == 3 || mode == 4) n = R * C * 9;
    else if (mode == 5) n = R * 64;
    else if (mode == 6) n = R * 48;
    else MSU_REQUIRE(false, "msu_prep_weight: bad mode %d", mode);
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (mode == 4 || mode == 6) prep_grad_kernel<<<grid, 256, 0, st>>>(mode, src, (float*)dst, R, C);
    else if (dst_dtype == MSU_F32) prep_weight_kernel<float><<<grid, 256, 0, st>>>(mode, src, (float*)dst, R, C);
    else if (dst_dtype == MSU_BF16) prep_weight_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(mode, src, (__nv_bfloat16*)dst, R, C);
    else MSU_REQUIRE(false, "msu_prep_weight: bad dtype %d", dst_dtype);
    count_launch();
    return check_launch("msu_prep_weight");
}

extern "C" int msu_shadow_blocks(int mode, int64_t R, int64_t C) {
    if (R <= 0 || C <= 0) return -1;
    int64_t nb;
    if (mode == 0) nb = ((R + SH_TR - 1) / SH_TR) * ((C + SH_TC - 1) / SH_TC);
    else if (mode == 2 || mode == 3) nb = (R * C * 9 + SH_CHUNK - 1) / SH_CHUNK;
    else if (mode == 5) nb = (R * 64 + SH_CHUNK - 1) / SH_CHUNK;
    else return -1;
    return nb < (1ll << 30) ? (int)nb : -1;
}

extern "C" int msu_refresh_shadows(const MsuShadowJob* jobs, const int32_t* blk_job, const int32_t* blk_tile, int32_t n_blocks,
                                   void* stream) {
    MSU_REQUIRE(jobs && blk_job && blk_tile && n_blocks >= 0, "msu_refresh_shadows: bad arguments");
    if (n_blocks == 0) return 0;
    refresh_shadows_kernel<<<n_blocks, SH_THREADS, 0, (cudaStream_t)stream>>>(jobs, blk_job, blk_tile);
    count_launch();
    return check_launch("msu_refresh_shadows");
}

extern "C" int msu_patchify4(int dst_dtype, const float* img, void* out, int32_t B, int32_t S, void* stream) {
    MSU_REQUIRE(img && out && B > 0 && S > 0 && S % 4 == 0, "msu_patchify4: bad arguments");
    const int64_t total = (int64_t)B * (S / 4) * (S / 4) * 16;
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == MSU_F32) patchify4_kernel<float><<<grid, 256, 0, st>>>(img, (float*)out, B, S);
    else if (dst_dtype == MSU_BF16) patchify4_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(img, (__nv_bfloat16*)out, B, S);
    else MSU_REQUIRE(false, "msu_patchify4: bad dtype %d", dst_dtype);
    count_launch();
    return check_launch("msu_patchify4");
}

extern "C" int msu_gather_rows(const MsuOperand* src, void* dst, int64_t M, int64_t N, void* stream) {
    MSU_REQUIRE(src && src->ptr && dst, "msu_gather_rows: null pointer");
    MSU_REQUIRE(src->orient == 0 && src->ptr2 == nullptr, "msu_gather_rows: operand must be a single [row, col] source");
    MSU_REQUIRE(N % 4 == 0 && src->ld % 4 == 0, "msu_gather_rows: N and ld must be multiples of 4");
    if (src->map == MSU_MAP_SHUFFLE || src->map == MSU_MAP_CONV3 || src->map == MSU_MAP_MERGE)
        MSU_REQUIRE((src->map == MSU_MAP_SHUFFLE ? src->geo[3] : src->geo[2]) % 4 == 0, "msu_gather_rows: chunk width must be a multiple of 4");
    if (M == 0) return 0;
    const int64_t total = M * (N / 4);
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (src->dtype == MSU_F32) gather_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)src->ptr, (float*)dst, M, (int)N, src->ld, src->map, *src);
    else if (src->dtype == MSU_BF16) gather_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src->ptr, (__nv_bfloat16*)dst, M, (int)N, src->ld, src->map, *src);
    else MSU_REQUIRE(false, "msu_gather_rows: bad dtype %d", src->dtype);
    count_launch();
    return check_launch("msu_gather_rows");
}

extern "C" int msu_cast(int sd, int dd, const void* src, void* dst, int64_t n, void* stream) {
    MSU_REQUIRE(src && dst, "msu_cast: null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (sd == MSU_F32 && dd == MSU_BF16) cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
    else if (sd == MSU_BF16 && dd == MSU_F32) cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
    else if (sd == MSU_F16 && dd == MSU_F32) cast_kernel<__half, float><<<grid, 256, 0, st>>>((const __half*)src, (float*)dst, n);
    else if (sd == MSU_F32 && dd == MSU_F16) cast_kernel<float, __half><<<grid, 256, 0, st>>>((const float*)src, (__half*)dst, n);
    else if (sd == MSU_F32 && dd == MSU_F32) cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, n);
    else if (sd == MSU_BF16 && dd == MSU_BF16) cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
    else MSU_REQUIRE(false, "msu_cast: unsupported %d -> %d", sd, dd);
    count_launch();
    return check_launch("msu_cast");
}

extern "C" int msu_add(int dtype, const void* a, const void* b, void* y, int64_t n, void* stream) {
    MSU_REQUIRE(a && b && y && n % 4 == 0, "msu_add: bad arguments");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n / 4 + 255) / 256);
    if (dtype == MSU_F32) add_kernel<float><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, (float*)y, n / 4);
    else if (dtype == MSU_BF16) add_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, n / 4);
    else MSU_REQUIRE(false, "msu_add: bad dtype %d", dtype);
    count_launch();
    return check_launch("msu_add");
}
