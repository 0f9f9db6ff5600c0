// LayerNorm forward/backward, one warp per row, 64/128-bit vectorised, fp32 statistics.
// Memory-bound: algorithmic bytes are 2*rows*C*e forward and (3 or 4)*rows*C*e backward.
// The gather/scatter variants fold window partition (+zero padding, roll), PatchMerging's 2x2 concat and
// the head's LN + 1x1 conv into the same pass (include/msunet_b200.h for the maps and citations).
#include "common.cuh"

namespace msu {

constexpr float LN_EPS = 1e-5f;
constexpr int LN_WARPS = 8;

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const T* __restrict__ X, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, T* __restrict__ Y,
                                                              float* __restrict__ mean, float* __restrict__ rstd,
                                                              int64_t rows, int C, int in_map, int out_map,
                                                              WinGeo wg, int mH, int mW, const float* __restrict__ dotw) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    if (r >= rows) return;
    int64_t lr = r;  // LayerNorm row (statistics index)
    if (out_map == MSU_MAP_WINDOW) {
        lr = win_to_pix(wg, r);
        if (lr < 0) {  // zero padding token (not masked: TV:models/swin_transformer.py:152-156)
#pragma unroll
            for (int j = 0; j < NV; j++) {
                const int c = (lane + 32 * j) * 4;
                if (c < C) Vec4<T>::st(Y + r * C + c, make_float4(0.f, 0.f, 0.f, 0.f));
            }
            return;
        }
    }
    const int Cin = C / 4;
    float4 x[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (lane + 32 * j) * 4;
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C) {
            const T* p;
            if (in_map == MSU_MAP_MERGE) {
                const int q = c / Cin, ci = c - q * Cin;
                const int h2w2 = (mH / 2) * (mW / 2);
                const int64_t b = lr / h2w2;
                const int t = (int)(lr - b * h2w2);
                const int y = 2 * (t / (mW / 2)) + (q & 1), xx = 2 * (t % (mW / 2)) + (q >> 1);
                p = X + (b * (int64_t)(mH * mW) + y * mW + xx) * Cin + ci;
            } else {
                p = X + lr * C + c;
            }
            x[j] = Vec4<T>::ld(p);
            s += x[j].x + x[j].y + x[j].z + x[j].w;
        }
    }
    const float mu = warp_sum(s) / C;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (lane + 32 * j) * 4;
        if (c < C) {
            const float a = x[j].x - mu, b = x[j].y - mu, cc = x[j].z - mu, d = x[j].w - mu;
            v += a * a + b * b + cc * cc + d * d;
        }
    }
    const float rs = 1.0f / sqrtf(warp_sum(v) / C + LN_EPS);
    if (lane == 0) {
        mean[lr] = mu;
        rstd[lr] = rs;
    }
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (lane + 32 * j) * 4;
        if (c < C) {
            const float4 g = *reinterpret_cast<const float4*>(gamma + c);
            const float4 b = *reinterpret_cast<const float4*>(beta + c);
            float4 y;
            y.x = (x[j].x - mu) * rs * g.x + b.x;
            y.y = (x[j].y - mu) * rs * g.y + b.y;
            y.z = (x[j].z - mu) * rs * g.z + b.z;
            y.w = (x[j].w - mu) * rs * g.w + b.w;
            if (dotw != nullptr) {
                const float4 w = *reinterpret_cast<const float4*>(dotw + c);
                dot += y.x * w.x + y.y * w.y + y.z * w.z + y.w * w.w;
            } else {
                Vec4<T>::st(Y + r * C + c, y);
            }
        }
    }
    if (dotw != nullptr) {
        dot = warp_sum(dot);
        if (lane == 0) Y[r] = from_f<T>(dot);
    }
}

// Backward: each warp walks LayerNorm rows  warp_id, warp_id + P, ...  and keeps the per-column
// parameter-gradient partial sums in registers; partial[warp_id][3][C] is reduced afterwards in a fixed order.
template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const T* __restrict__ dY, const T* __restrict__ X,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const T* __restrict__ dRes, T* __restrict__ dX, int64_t rows,
                                                              int C, int dy_map, int dx_map, WinGeo wg, int mH, int mW,
                                                              const float* __restrict__ dotw, float* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const int64_t P = (int64_t)gridDim.x * LN_WARPS;
    const int Cin = C / 4;
    float4 ag[NV], ab[NV], aw[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) ag[j] = ab[j] = aw[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int64_t lr = wid; lr < rows; lr += P) {
        const float mu = mean[lr], rs = rstd[lr];
        const int64_t dyr = (dy_map == MSU_MAP_WINDOW) ? pix_to_win(wg, lr) : lr;
        const float dl = (dotw != nullptr) ? to_f<T>(dY[lr]) : 0.f;
        // merge geometry for this row
        int64_t mb = 0; int my = 0, mx = 0;
        if (dx_map == MSU_MAP_MERGE) {
            const int h2w2 = (mH / 2) * (mW / 2);
            mb = lr / h2w2;
            const int t = (int)(lr - mb * h2w2);
            my = 2 * (t / (mW / 2));
            mx = 2 * (t % (mW / 2));
        }
        auto xoff = [&](int c) -> int64_t {
            if (dx_map == MSU_MAP_MERGE) {
                const int q = c / Cin, ci = c - q * Cin;
                return (mb * (int64_t)(mH * mW) + (my + (q & 1)) * mW + (mx + (q >> 1))) * Cin + ci;
            }
            return lr * C + c;
        };
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (lane + 32 * j) * 4;
            if (c < C) {
                const float4 x = Vec4<T>::ld(X + xoff(c));
                const float4 g = *reinterpret_cast<const float4*>(gamma + c);
                float4 dy;
                if (dotw != nullptr) {
                    const float4 w = *reinterpret_cast<const float4*>(dotw + c);
                    dy = make_float4(dl * w.x, dl * w.y, dl * w.z, dl * w.w);
                } else {
                    dy = Vec4<T>::ld(dY + dyr * C + c);
                }
                const float xh0 = (x.x - mu) * rs, xh1 = (x.y - mu) * rs, xh2 = (x.z - mu) * rs, xh3 = (x.w - mu) * rs;
                const float g0 = dy.x * g.x, g1 = dy.y * g.y, g2 = dy.z * g.z, g3 = dy.w * g.w;
                s1 += g0 + g1 + g2 + g3;
                s2 += g0 * xh0 + g1 * xh1 + g2 * xh2 + g3 * xh3;
            }
        }
        s1 = warp_sum(s1) / C;
        s2 = warp_sum(s2) / C;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (lane + 32 * j) * 4;
            if (c < C) {
                const int64_t xo = xoff(c);
                const float4 x = Vec4<T>::ld(X + xo);  // L1-resident re-read
                const float4 g = *reinterpret_cast<const float4*>(gamma + c);
                float4 dy;
                if (dotw != nullptr) {
                    const float4 w = *reinterpret_cast<const float4*>(dotw + c);
                    const float4 b = *reinterpret_cast<const float4*>(beta + c);
                    dy = make_float4(dl * w.x, dl * w.y, dl * w.z, dl * w.w);
                    aw[j].x += dl * ((x.x - mu) * rs * g.x + b.x);
                    aw[j].y += dl * ((x.y - mu) * rs * g.y + b.y);
                    aw[j].z += dl * ((x.z - mu) * rs * g.z + b.z);
                    aw[j].w += dl * ((x.w - mu) * rs * g.w + b.w);
                } else {
                    dy = Vec4<T>::ld(dY + dyr * C + c);
                }
                const float xh0 = (x.x - mu) * rs, xh1 = (x.y - mu) * rs, xh2 = (x.z - mu) * rs, xh3 = (x.w - mu) * rs;
                float4 dx;
                dx.x = rs * (dy.x * g.x - s1 - xh0 * s2);
                dx.y = rs * (dy.y * g.y - s1 - xh1 * s2);
                dx.z = rs * (dy.z * g.z - s1 - xh2 * s2);
                dx.w = rs * (dy.w * g.w - s1 - xh3 * s2);
                if (dRes != nullptr) {
                    const float4 d = Vec4<T>::ld(dRes + xo);
                    dx.x += d.x; dx.y += d.y; dx.z += d.z; dx.w += d.w;
                }
                Vec4<T>::st(dX + xo, dx);
                ag[j].x += dy.x * xh0; ag[j].y += dy.y * xh1; ag[j].z += dy.z * xh2; ag[j].w += dy.w * xh3;
                ab[j].x += dy.x; ab[j].y += dy.y; ab[j].z += dy.z; ab[j].w += dy.w;
            }
        }
    }
    float* pg = partial + wid * 3 * (int64_t)C;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (lane + 32 * j) * 4;
        if (c < C) {
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

template <typename T, int NV>
static int launch_fwd(const void* X, const float* gamma, const float* beta, void* Y, float* mean, float* rstd,
                      int64_t rows, int C, int in_map, int out_map, const int32_t* geo, const float* dotw,
                      cudaStream_t st) {
    WinGeo wg{0, 0, 0, 0, 0, 0};
    int mH = 0, mW = 0;
    if (out_map == MSU_MAP_WINDOW) wg = make_wingeo(geo);
    if (in_map == MSU_MAP_MERGE) { mH = geo[0]; mW = geo[1]; }
    const unsigned grid = (unsigned)((rows + LN_WARPS - 1) / LN_WARPS);
    ln_fwd_kernel<T, NV><<<grid, LN_WARPS * 32, 0, st>>>((const T*)X, gamma, beta, (T*)Y, mean, rstd, rows, C, in_map,
                                                        out_map, wg, mH, mW, dotw);
    count_launch();
    return check_launch("msu_ln_fwd");
}

template <typename T, int NV>
static int launch_bwd(const void* dY, const void* X, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, const void* dRes, void* dX, int64_t rows, int C, int dy_map, int dx_map,
                      const int32_t* geo, const float* dotw, float* partial, int grid, cudaStream_t st) {
    WinGeo wg{0, 0, 0, 0, 0, 0};
    int mH = 0, mW = 0;
    if (dy_map == MSU_MAP_WINDOW) wg = make_wingeo(geo);
    if (dx_map == MSU_MAP_MERGE) { mH = geo[0]; mW = geo[1]; }
    ln_bwd_kernel<T, NV><<<grid, LN_WARPS * 32, 0, st>>>((const T*)dY, (const T*)X, gamma, beta, mean, rstd,
                                                        (const T*)dRes, (T*)dX, rows, C, dy_map, dx_map, wg, mH, mW,
                                                        dotw, partial);
    count_launch();
    return check_launch("msu_ln_bwd");
}

#define LN_DISPATCH_NV(FN, T, ...)                                         \
    do {                                                                   \
        const int nv = (C / 4 + 31) / 32;                                  \
        if (nv <= 1) return FN<T, 1>(__VA_ARGS__);                         \
        if (nv <= 2) return FN<T, 2>(__VA_ARGS__);                         \
        if (nv <= 3) return FN<T, 3>(__VA_ARGS__);                         \
        if (nv <= 4) return FN<T, 4>(__VA_ARGS__);                         \
        if (nv <= 6) return FN<T, 6>(__VA_ARGS__);                         \
        if (nv <= 8) return FN<T, 8>(__VA_ARGS__);                         \
        if (nv <= 12) return FN<T, 12>(__VA_ARGS__);                       \
        if (nv <= 16) return FN<T, 16>(__VA_ARGS__);                       \
        if (nv <= 24) return FN<T, 24>(__VA_ARGS__);                       \
        if (nv <= 32) return FN<T, 32>(__VA_ARGS__);                       \
        msu::set_error("layernorm: C=%d too large (max 4096)", C);         \
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
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_fwd, float, X, gamma, beta, Y, mean, rstd, rows, C, in_map, out_map, geo, dotw, st);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_fwd, __nv_bfloat16, X, gamma, beta, Y, mean, rstd, rows, C, in_map, out_map, geo, dotw, st);
    MSU_REQUIRE(false, "msu_ln_fwd: unsupported dtype %d", dtype);
}

// number of partial rows P = grid*8: enough warps to cover the SMs, bounded by a 32 MiB workspace
extern "C" int msu_ln_bwd_partial_rows(int64_t rows, int32_t C) {
    int64_t P = imin(rows, (int64_t)num_sms() * 32);
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
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_bwd, float, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, dy_map, dx_map, geo, dotw, partial, grid, st);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_bwd, __nv_bfloat16, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, dy_map, dx_map, geo, dotw, partial, grid, st);
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
