// LayerNorm forward/backward, fp32 statistics, 64/128-bit vector accesses.
// A row is owned by `lpr` lanes (4..32, power of two) so that narrow rows (C = 48..192) put several rows of
// independent loads in flight per warp: HBM-bound, algorithmic bytes 2*rows*C*e forward, (3-4)*rows*C*e backward.
// The gather/scatter variants fold window partition (+zero padding, roll), PatchMerging's 2x2 concat and
// the head's LN + 1x1 conv into the same pass (include/msunet_b200.h for the maps and citations).
#include "common.cuh"

namespace msu {

constexpr float LN_EPS = 1e-5f;
constexpr int LN_WARPS = 8;

__device__ __forceinline__ float group_sum(float v, int lpr) {
    for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct MergeGeo { int H, W; };
// source offset (in elements) of logical column c of LayerNorm row lr when the row is a 2x2 neighbourhood concat
__device__ __forceinline__ int64_t merge_off(const MergeGeo& mg, int64_t lr, int c, int Cin) {
    const int q = c / Cin, ci = c - q * Cin;
    const int h2w2 = (mg.H / 2) * (mg.W / 2);
    const int64_t b = lr / h2w2;
    const int t = (int)(lr - b * h2w2);
    const int y = 2 * (t / (mg.W / 2)) + (q & 1), x = 2 * (t % (mg.W / 2)) + (q >> 1);
    return (b * (int64_t)(mg.H * mg.W) + y * mg.W + x) * Cin + ci;
}

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const T* __restrict__ X, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, T* __restrict__ Y,
                                                              float* __restrict__ mean, float* __restrict__ rstd,
                                                              int64_t rows, int C, int lpr, int in_map, int out_map,
                                                              WinGeo wg, MergeGeo mg, const float* __restrict__ dotw) {
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr, sub = lane / lpr, l = lane - sub * lpr;
    const int64_t r = ((int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * rpw + sub;
    const bool in_range = r < rows;
    int64_t lr = r;  // LayerNorm row (statistics index)
    bool pad = false;
    if (in_range && out_map == MSU_MAP_WINDOW) {
        lr = win_to_pix(wg, r);
        pad = lr < 0;  // zero padding token (not masked: TV:models/swin_transformer.py:152-156)
    }
    const bool live = in_range && !pad;
    const int Cin = C / 4;
    float4 x[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (l + lpr * j) * 4;
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && c < C) {
            const T* p = (in_map == MSU_MAP_MERGE) ? X + merge_off(mg, lr, c, Cin) : X + lr * C + c;
            x[j] = Vec4<T>::ld(p);
            s += x[j].x + x[j].y + x[j].z + x[j].w;
        }
    }
    const float mu = group_sum(s, lpr) / C;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (l + lpr * j) * 4;
        if (c < C) {
            const float a = x[j].x - mu, b = x[j].y - mu, cc = x[j].z - mu, d = x[j].w - mu;
            v += a * a + b * b + cc * cc + d * d;
        }
    }
    const float rs = 1.0f / sqrtf(group_sum(v, lpr) / C + LN_EPS);
    if (live && l == 0) {
        mean[lr] = mu;
        rstd[lr] = rs;
    }
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (l + lpr * j) * 4;
        if (in_range && c < C) {
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!pad) {
                const float4 g = *reinterpret_cast<const float4*>(gamma + c);
                const float4 b = *reinterpret_cast<const float4*>(beta + c);
                y.x = (x[j].x - mu) * rs * g.x + b.x;
                y.y = (x[j].y - mu) * rs * g.y + b.y;
                y.z = (x[j].z - mu) * rs * g.z + b.z;
                y.w = (x[j].w - mu) * rs * g.w + b.w;
            }
            if (dotw != nullptr) {
                const float4 w = *reinterpret_cast<const float4*>(dotw + c);
                dot += y.x * w.x + y.y * w.y + y.z * w.z + y.w * w.w;
            } else {
                Vec4<T>::st(Y + r * C + c, y);
            }
        }
    }
    if (dotw != nullptr) {
        dot = group_sum(dot, lpr);
        if (in_range && l == 0) Y[r] = from_f<T>(dot);
    }
}

// Backward: warp w walks row groups  w, w + P, ...  (P = number of warps) and keeps the per-column parameter-gradient
// partial sums in registers; partial[w][3][C] is reduced afterwards in a fixed order (deterministic).
template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const T* __restrict__ dY, const T* __restrict__ X,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const T* __restrict__ dRes, T* __restrict__ dX, int64_t rows,
                                                              int C, int lpr, int dy_map, int dx_map, WinGeo wg, MergeGeo mg,
                                                              const float* __restrict__ dotw, float* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr, sub = lane / lpr, l = lane - sub * lpr;
    const int64_t wid = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const int64_t P = (int64_t)gridDim.x * LN_WARPS;
    const int Cin = C / 4;
    float4 ag[NV], ab[NV], aw[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) ag[j] = ab[j] = aw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 gm[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (l + lpr * j) * 4;
        gm[j] = (c < C) ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (int64_t g0 = wid * rpw; g0 < rows; g0 += P * rpw) {
        const int64_t lr = g0 + sub;
        const bool live = lr < rows;
        const float mu = live ? mean[lr] : 0.f, rs = live ? rstd[lr] : 0.f;
        const int64_t dyr = (live && dy_map == MSU_MAP_WINDOW) ? pix_to_win(wg, lr) : lr;
        const float dl = (live && dotw != nullptr) ? to_f<T>(dY[lr]) : 0.f;
        float4 x[NV], dy[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (l + lpr * j) * 4;
            x[j] = dy[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && c < C) {
                const int64_t xo = (dx_map == MSU_MAP_MERGE) ? merge_off(mg, lr, c, Cin) : lr * C + c;
                x[j] = Vec4<T>::ld(X + xo);
                if (dotw != nullptr) {
                    const float4 w = *reinterpret_cast<const float4*>(dotw + c);
                    dy[j] = make_float4(dl * w.x, dl * w.y, dl * w.z, dl * w.w);
                } else {
                    dy[j] = Vec4<T>::ld(dY + dyr * C + c);
                }
                // x <- normalised x-hat
                x[j].x = (x[j].x - mu) * rs; x[j].y = (x[j].y - mu) * rs; x[j].z = (x[j].z - mu) * rs; x[j].w = (x[j].w - mu) * rs;
                const float g0_ = dy[j].x * gm[j].x, g1_ = dy[j].y * gm[j].y, g2_ = dy[j].z * gm[j].z, g3_ = dy[j].w * gm[j].w;
                s1 += g0_ + g1_ + g2_ + g3_;
                s2 += g0_ * x[j].x + g1_ * x[j].y + g2_ * x[j].z + g3_ * x[j].w;
            }
        }
        s1 = group_sum(s1, lpr) / C;
        s2 = group_sum(s2, lpr) / C;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (l + lpr * j) * 4;
            if (live && c < C) {
                const int64_t xo = (dx_map == MSU_MAP_MERGE) ? merge_off(mg, lr, c, Cin) : lr * C + c;
                float4 dx;
                dx.x = rs * (dy[j].x * gm[j].x - s1 - x[j].x * s2);
                dx.y = rs * (dy[j].y * gm[j].y - s1 - x[j].y * s2);
                dx.z = rs * (dy[j].z * gm[j].z - s1 - x[j].z * s2);
                dx.w = rs * (dy[j].w * gm[j].w - s1 - x[j].w * s2);
                if (dRes != nullptr) {
                    const float4 d = Vec4<T>::ld(dRes + xo);
                    dx.x += d.x; dx.y += d.y; dx.z += d.z; dx.w += d.w;
                }
                Vec4<T>::st(dX + xo, dx);
                ag[j].x += dy[j].x * x[j].x; ag[j].y += dy[j].y * x[j].y; ag[j].z += dy[j].z * x[j].z; ag[j].w += dy[j].w * x[j].w;
                ab[j].x += dy[j].x; ab[j].y += dy[j].y; ab[j].z += dy[j].z; ab[j].w += dy[j].w;
                if (dotw != nullptr) {
                    const float4 b = *reinterpret_cast<const float4*>(beta + c);
                    aw[j].x += dl * (x[j].x * gm[j].x + b.x); aw[j].y += dl * (x[j].y * gm[j].y + b.y);
                    aw[j].z += dl * (x[j].z * gm[j].z + b.z); aw[j].w += dl * (x[j].w * gm[j].w + b.w);
                }
            }
        }
    }
    // fold the row groups of this warp (lanes l, l+lpr, ...) in a fixed order, then one partial row per warp
    float* pg = partial + wid * 3 * (int64_t)C;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        for (int o = lpr; o < 32; o <<= 1) {
            ag[j].x += __shfl_xor_sync(0xffffffffu, ag[j].x, o); ag[j].y += __shfl_xor_sync(0xffffffffu, ag[j].y, o);
            ag[j].z += __shfl_xor_sync(0xffffffffu, ag[j].z, o); ag[j].w += __shfl_xor_sync(0xffffffffu, ag[j].w, o);
            ab[j].x += __shfl_xor_sync(0xffffffffu, ab[j].x, o); ab[j].y += __shfl_xor_sync(0xffffffffu, ab[j].y, o);
            ab[j].z += __shfl_xor_sync(0xffffffffu, ab[j].z, o); ab[j].w += __shfl_xor_sync(0xffffffffu, ab[j].w, o);
            if (dotw != nullptr) {
                aw[j].x += __shfl_xor_sync(0xffffffffu, aw[j].x, o); aw[j].y += __shfl_xor_sync(0xffffffffu, aw[j].y, o);
                aw[j].z += __shfl_xor_sync(0xffffffffu, aw[j].z, o); aw[j].w += __shfl_xor_sync(0xffffffffu, aw[j].w, o);
            }
        }
        const int c = (l + lpr * j) * 4;
        if (sub == 0 && c < C) {
            *reinterpret_cast<float4*>(pg + c) = ag[j];
            *reinterpret_cast<float4*>(pg + C + c) = ab[j];
            *reinterpret_cast<float4*>(pg + 2 * C + c) = aw[j];
        }
    }
}

// out[a][c] = sum_p partial[p][a][c]; block = 32 columns x 32 row-lanes, fixed-order tree => deterministic.
__global__ void __launch_bounds__(1024) ln_param_reduce_kernel(const float* __restrict__ partial, int P, int C,
                                                               float* dgamma, float* dbeta, float* ddotw, int accumulate) {
    __shared__ float sm[32][33];
    const int a = blockIdx.y;
    float* out = a == 0 ? dgamma : (a == 1 ? dbeta : ddotw);
    if (out == nullptr) return;
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < C)
        for (int p = threadIdx.y; p < P; p += 32) s += partial[((int64_t)p * 3 + a) * C + c];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i++) t += sm[i][threadIdx.x];
        out[c] = accumulate ? out[c] + t : t;
    }
}

// lanes per row: power of two in [4, 32] giving about 3-4 vectors (of 4 elements) per lane
static int pick_lpr(int C) {
    const int nvec = C / 4;
    int lpr = 4;
    while (lpr < 32 && lpr * 4 < nvec) lpr <<= 1;
    return lpr;
}

template <typename T, int NV>
static int launch_fwd(const void* X, const float* gamma, const float* beta, void* Y, float* mean, float* rstd,
                      int64_t rows, int C, int lpr, int in_map, int out_map, const int32_t* geo, const float* dotw,
                      cudaStream_t st) {
    WinGeo wg{0, 0, 0, 0, 0, 0};
    MergeGeo mg{0, 0};
    if (out_map == MSU_MAP_WINDOW) wg = make_wingeo(geo);
    if (in_map == MSU_MAP_MERGE) { mg.H = geo[0]; mg.W = geo[1]; }
    const int rpb = LN_WARPS * (32 / lpr);
    const unsigned grid = (unsigned)((rows + rpb - 1) / rpb);
    ln_fwd_kernel<T, NV><<<grid, LN_WARPS * 32, 0, st>>>((const T*)X, gamma, beta, (T*)Y, mean, rstd, rows, C, lpr, in_map,
                                                        out_map, wg, mg, dotw);
    count_launch();
    return check_launch("msu_ln_fwd");
}

template <typename T, int NV>
static int launch_bwd(const void* dY, const void* X, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, const void* dRes, void* dX, int64_t rows, int C, int lpr, int dy_map, int dx_map,
                      const int32_t* geo, const float* dotw, float* partial, int grid, cudaStream_t st) {
    WinGeo wg{0, 0, 0, 0, 0, 0};
    MergeGeo mg{0, 0};
    if (dy_map == MSU_MAP_WINDOW) wg = make_wingeo(geo);
    if (dx_map == MSU_MAP_MERGE) { mg.H = geo[0]; mg.W = geo[1]; }
    ln_bwd_kernel<T, NV><<<grid, LN_WARPS * 32, 0, st>>>((const T*)dY, (const T*)X, gamma, beta, mean, rstd,
                                                        (const T*)dRes, (T*)dX, rows, C, lpr, dy_map, dx_map, wg, mg,
                                                        dotw, partial);
    count_launch();
    return check_launch("msu_ln_bwd");
}

#define LN_DISPATCH_NV(FN, T, ...)                                         \
    do {                                                                   \
        const int nv = (C / 4 + pick_lpr(C) - 1) / pick_lpr(C);            \
        if (nv <= 1) return FN<T, 1>(__VA_ARGS__);                         \
        if (nv <= 2) return FN<T, 2>(__VA_ARGS__);                         \
        if (nv <= 3) return FN<T, 3>(__VA_ARGS__);                         \
        if (nv <= 4) return FN<T, 4>(__VA_ARGS__);                         \
        if (nv <= 6) return FN<T, 6>(__VA_ARGS__);                         \
        if (nv <= 8) return FN<T, 8>(__VA_ARGS__);                         \
        if (nv <= 12) return FN<T, 12>(__VA_ARGS__);                       \
        if (nv <= 16) return FN<T, 16>(__VA_ARGS__);                       \
        msu::set_error("layernorm: C=%d too large (max 2048)", C);         \
        return -1;                                                         \
    } while (0)

}  // namespace msu

using namespace msu;

extern "C" int msu_ln_fwd(int dtype, const void* X, const float* gamma, const float* beta, void* Y, float* mean,
                          float* rstd, int64_t rows, int32_t C, int32_t in_map, int32_t out_map, const int32_t* geo,
                          const float* dotw, void* stream) {
    MSU_REQUIRE(X && gamma && beta && Y && mean && rstd, "msu_ln_fwd: null pointer");
    MSU_REQUIRE(C % 4 == 0 && C > 0, "msu_ln_fwd: C=%d must be a positive multiple of 4", C);
    MSU_REQUIRE(in_map == MSU_MAP_NONE || in_map == MSU_MAP_MERGE, "msu_ln_fwd: bad in_map %d", in_map);
    MSU_REQUIRE(out_map == MSU_MAP_NONE || out_map == MSU_MAP_WINDOW, "msu_ln_fwd: bad out_map %d", out_map);
    MSU_REQUIRE((in_map == 0 && out_map == 0) || geo != nullptr, "msu_ln_fwd: geo required for mapped rows");
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_fwd, float, X, gamma, beta, Y, mean, rstd, rows, C, pick_lpr(C), in_map, out_map, geo, dotw, st);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_fwd, __nv_bfloat16, X, gamma, beta, Y, mean, rstd, rows, C, pick_lpr(C), in_map, out_map, geo, dotw, st);
    MSU_REQUIRE(false, "msu_ln_fwd: unsupported dtype %d", dtype);
}

// number of partial rows P = grid*8 warps: enough warps to cover the SMs, bounded by a 32 MiB workspace
extern "C" int msu_ln_bwd_partial_rows(int64_t rows, int32_t C) {
    const int rpw = 32 / pick_lpr(C);
    int64_t P = imin((rows + rpw - 1) / rpw, (int64_t)num_sms() * 32);
    P = imin(P, (8ll << 20) / (3ll * C));
    const int grid = (int)imax(1, (P + LN_WARPS - 1) / LN_WARPS);
    return grid * LN_WARPS;
}

extern "C" int msu_ln_bwd(int dtype, const void* dY, const void* X, const float* gamma, const float* beta,
                          const float* mean, const float* rstd, const void* dRes, void* dX, int64_t rows, int32_t C,
                          int32_t dy_map, int32_t dx_map, const int32_t* geo, const float* dotw, float* partial,
                          void* stream) {
    MSU_REQUIRE(dY && X && gamma && beta && mean && rstd && dX && partial, "msu_ln_bwd: null pointer");
    MSU_REQUIRE(C % 4 == 0 && C > 0, "msu_ln_bwd: C=%d must be a positive multiple of 4", C);
    MSU_REQUIRE(dy_map == MSU_MAP_NONE || dy_map == MSU_MAP_WINDOW, "msu_ln_bwd: bad dy_map %d", dy_map);
    MSU_REQUIRE(dx_map == MSU_MAP_NONE || dx_map == MSU_MAP_MERGE, "msu_ln_bwd: bad dx_map %d", dx_map);
    const int grid = msu_ln_bwd_partial_rows(rows, C) / LN_WARPS;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_bwd, float, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, pick_lpr(C), dy_map, dx_map, geo, dotw, partial, grid, st);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_bwd, __nv_bfloat16, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, pick_lpr(C), dy_map, dx_map, geo, dotw, partial, grid, st);
    MSU_REQUIRE(false, "msu_ln_bwd: unsupported dtype %d", dtype);
}

extern "C" int msu_ln_param_reduce(const float* partial, int32_t P, int32_t C, float* dgamma, float* dbeta,
                                   float* ddotw, int accumulate, void* stream) {
    MSU_REQUIRE(partial && P > 0 && C > 0, "msu_ln_param_reduce: bad arguments");
    dim3 grid((unsigned)((C + 31) / 32), 3), block(32, 32);
    ln_param_reduce_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(partial, P, C, dgamma, dbeta, ddotw, accumulate);
    count_launch();
    return check_launch("msu_ln_param_reduce");
}
