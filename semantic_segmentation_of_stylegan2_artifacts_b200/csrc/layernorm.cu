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

// 16-byte vectors: 4 floats or 8 bf16
template <typename T> struct VecW;
template <> struct VecW<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void ld(const float* p, float* o) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    static __device__ __forceinline__ void st(float* p, const float* o) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    }
};
template <> struct VecW<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
        const uint4 r = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 f = __bfloat1622float2(h[i]);
            o[2 * i] = f.x; o[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* o) {
        uint4 r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};
// raw 16 B vector (kept packed in registers) and its conversion
template <typename T> __device__ __forceinline__ void cvt_raw(const uint4& r, float* o);
template <> __device__ __forceinline__ void cvt_raw<float>(const uint4& r, float* o) {
    o[0] = __uint_as_float(r.x); o[1] = __uint_as_float(r.y); o[2] = __uint_as_float(r.z); o[3] = __uint_as_float(r.w);
}
template <> __device__ __forceinline__ void cvt_raw<__nv_bfloat16>(const uint4& r, float* o) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 f = __bfloat1622float2(h[i]);
        o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void ldf(const float* p, float* o, int n) {
    for (int i = 0; i < n; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + i);
        o[i] = v.x; o[i + 1] = v.y; o[i + 2] = v.z; o[i + 3] = v.w;
    }
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

// Both kernels are templated on the row map so that each instantiation is straight-line code (ncu showed the generic
// version issue-bound at ~23 instructions per element: runtime map branches, gamma / beta reloaded per row, IEEE
// divisions); gamma / beta / the dot weight live in registers for narrow rows, statistics use rsqrtf and a host-side 1/C.
constexpr int LNM_PLAIN = 0, LNM_WINDOW = 1, LNM_MERGE = 2, LNM_DOT = 3, LNM_UNSHUFFLE = 4;
constexpr int LNM_DUAL = 5;   // backward only: dX in pixel order AND a per-sample-scaled copy in window order (wg)

__device__ __forceinline__ float group_sum_u(float v, int lpr) {   // branch-free: always five shuffles, selects for lpr < 32
    float t;
    t = __shfl_xor_sync(0xffffffffu, v, 16); v += lpr > 16 ? t : 0.f;
    t = __shfl_xor_sync(0xffffffffu, v, 8);  v += lpr > 8 ? t : 0.f;
    t = __shfl_xor_sync(0xffffffffu, v, 4);  v += lpr > 4 ? t : 0.f;
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Forward: every warp owns LN_RG row groups (rpw rows each) whose 16 B loads are all issued before the first
// reduction, so ~2x the bytes are in flight per warp (the one-group version topped out at 3.4 TB/s).
constexpr int LN_RG = 2;
template <typename T, int NV, int MODE>
__global__ void __launch_bounds__(LN_WARPS * 32, NV <= 3 ? 3 : 1) ln_fwd_kernel(const T* __restrict__ X, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, T* __restrict__ Y,
                                                              float* __restrict__ mean, float* __restrict__ rstd,
                                                              int64_t rows, int C, float invC, int lpr,
                                                              WinGeo wg, MergeGeo mg, const float* __restrict__ dotw) {
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr, sub = lane / lpr, l = lane - sub * lpr;
    constexpr int VW = VecW<T>::N;
    const int Cin = C / 4;
    const int64_t wbase = ((int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * LN_RG;
    int64_t r[LN_RG], lr[LN_RG];
    bool live[LN_RG];
    uint4 xr[LN_RG][NV];
    bool vld[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) vld[j] = (l + lpr * j) * VW < C;
#pragma unroll
    for (int g = 0; g < LN_RG; g++) {
        r[g] = (wbase + g) * rpw + sub;
        lr[g] = r[g];       // LayerNorm row (statistics index)
        live[g] = r[g] < rows;
        if (MODE == LNM_WINDOW && live[g]) {
            lr[g] = win_to_pix(wg, r[g]);
            live[g] = lr[g] >= 0;  // zero padding token (not masked: TV:models/swin_transformer.py:152-156)
        }
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (l + lpr * j) * VW;
            xr[g][j] = make_uint4(0, 0, 0, 0);
            if (live[g] && vld[j]) {
                const T* p = (MODE == LNM_MERGE) ? X + merge_off(mg, lr[g], c, Cin) : X + lr[g] * C + c;
                xr[g][j] = *reinterpret_cast<const uint4*>(p);
            }
        }
    }
#pragma unroll
    for (int g = 0; g < LN_RG; g++) {
        float x[NV][VW];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            cvt_raw<T>(xr[g][j], x[j]);
#pragma unroll
            for (int e = 0; e < VW; e++) s += x[j][e];
        }
        const float mu = group_sum_u(s, lpr) * invC;
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            if (vld[j]) {
#pragma unroll
                for (int e = 0; e < VW; e++) { const float a = x[j][e] - mu; v = fmaf(a, a, v); }
            }
        }
        const float rs = rsqrtf(group_sum_u(v, lpr) * invC + LN_EPS);
        const float nmr = -mu * rs;
        if (live[g] && l == 0) {
            mean[lr[g]] = mu;
            rstd[lr[g]] = rs;
        }
        const bool in_range = r[g] < rows;
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (l + lpr * j) * VW;
            if (in_range && vld[j]) {
                float y[VW];
                if (live[g]) {
                    float gl[VW], bl[VW];
                    ldf(gamma + c, gl, VW);
                    ldf(beta + c, bl, VW);
#pragma unroll
                    for (int e = 0; e < VW; e++) y[e] = fmaf(fmaf(x[j][e], rs, nmr), gl[e], bl[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < VW; e++) y[e] = 0.f;
                }
                if (MODE == LNM_DOT) {
                    float wl[VW];
                    ldf(dotw + c, wl, VW);
#pragma unroll
                    for (int e = 0; e < VW; e++) dot = fmaf(y[e], wl[e], dot);
                } else {
                    VecW<T>::st(Y + r[g] * C + c, y);
                }
            }
        }
        if (MODE == LNM_DOT) {
            dot = group_sum_u(dot, lpr);
            if (in_range && l == 0) Y[r[g]] = from_f<T>(dot);
        }
    }
}

// Backward: warp w walks row groups  w, w + P, ...  (P = number of warps) and keeps the per-column parameter-gradient
// partial sums in registers; partial[w][3][C] is reduced afterwards in a fixed order (deterministic).
// Operands stay packed (16 B) in registers between the two passes over a row and the residual gradient is requested
// together with x and dy, so a row group costs ONE memory round trip and the kernel fits 2 blocks (16 warps) per SM:
// ~75 KB of loads in flight per SM instead of 24 KB (the first version ran at 8 warps/SM and ~1.4 TB/s).
template <typename T, int NV, int MODE, bool RES>
__global__ void __launch_bounds__(LN_WARPS * 32, NV <= 3 ? 2 : 1) ln_bwd_kernel(const T* __restrict__ dY, const T* __restrict__ X,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const T* __restrict__ dRes, T* __restrict__ dX, int64_t rows,
                                                              int C, float invC, int lpr, WinGeo wg, MergeGeo mg,
                                                              const float* __restrict__ dotw, float* __restrict__ partial,
                                                              MsuOperand ug, T* __restrict__ dXw, const float* __restrict__ wscale,
                                                              int rows_per_sample) {
    constexpr bool DOT = MODE == LNM_DOT;
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr, sub = lane / lpr, l = lane - sub * lpr;
    const int64_t wid = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const int64_t P = (int64_t)gridDim.x * LN_WARPS;
    constexpr int VW = VecW<T>::N;
    const int Cin = C / 4;
    bool vld[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) vld[j] = (l + lpr * j) * VW < C;
    // ag = sum dy * x-hat (DOT: sum dl * x-hat), ab = sum dy (DOT: the scalar sum dl lives in ab[0][0])
    float ag[NV][VW], ab[DOT ? 1 : NV][DOT ? 1 : VW];
#pragma unroll
    for (int j = 0; j < NV; j++) {
#pragma unroll
        for (int e = 0; e < VW; e++) {
            ag[j][e] = 0.f;
            if (!DOT) ab[DOT ? 0 : j][DOT ? 0 : e] = 0.f;
        }
    }
    if (DOT) ab[0][0] = 0.f;

    for (int64_t g0 = wid * rpw; g0 < rows; g0 += P * rpw) {
        const int64_t lr = g0 + sub;
        const bool live = lr < rows;
        const float mu = live ? mean[lr] : 0.f, rs = live ? rstd[lr] : 0.f;
        const float nmr = -mu * rs;
        const int64_t dyr = (MODE == LNM_WINDOW && live) ? pix_to_win(wg, lr) : lr;
        // DUAL: the same gradient row, scaled per sample, also goes to its window-order row (the padding rows of that
        // buffer are zeroed once by the caller and never written)
        const int64_t wrow = (MODE == LNM_DUAL && live) ? pix_to_win(wg, lr) : 0;
        const float wsc = (MODE == LNM_DUAL && live && wscale != nullptr) ? wscale[lr / rows_per_sample] : 1.0f;
        const float dl = (DOT && live) ? to_f<T>(dY[lr]) : 0.f;
        const int64_t rowoff = lr * C;
        uint4 xr[NV], yr[DOT ? 1 : NV], rr[RES ? NV : 1];
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (l + lpr * j) * VW;
            if (live && vld[j]) {
                const int64_t xo = (MODE == LNM_MERGE) ? merge_off(mg, lr, c, Cin) : rowoff + c;
                xr[j] = *reinterpret_cast<const uint4*>(X + xo);
                if (!DOT) yr[DOT ? 0 : j] = *reinterpret_cast<const uint4*>(dY + dyr * C + c);
                if (RES) rr[RES ? j : 0] = *reinterpret_cast<const uint4*>(dRes + xo);
            }
        }
        // pass 1: row sums of g = dy * gamma and g * x-hat
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            if (live && vld[j]) {
                const int c = (l + lpr * j) * VW;
                float x[VW], dy[VW], gl[VW];
                cvt_raw<T>(xr[j], x);
                ldf(gamma + c, gl, VW);
                if (DOT) {
                    ldf(dotw + c, dy, VW);
#pragma unroll
                    for (int e = 0; e < VW; e++) dy[e] *= dl;
                } else {
                    cvt_raw<T>(yr[DOT ? 0 : j], dy);
                }
#pragma unroll
                for (int e = 0; e < VW; e++) {
                    const float g = dy[e] * gl[e];
                    s1 += g;
                    s2 = fmaf(g, fmaf(x[e], rs, nmr), s2);
                }
            }
        }
        s1 = group_sum_u(s1, lpr) * invC;
        s2 = group_sum_u(s2, lpr) * invC;
        const float k1 = -s1 * rs, k2 = -s2 * rs;
        if (DOT) ab[0][0] += dl;
        // pass 2: dx = rs * (g - s1 - xh * s2) + dres, parameter gradients d gamma += dy * x-hat, d beta += dy
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (l + lpr * j) * VW;
            if (live && vld[j]) {
                float x[VW], dy[VW], gl[VW], dx[VW];
                cvt_raw<T>(xr[j], x);
                ldf(gamma + c, gl, VW);
                if (DOT) {
                    ldf(dotw + c, dy, VW);
#pragma unroll
                    for (int e = 0; e < VW; e++) dy[e] *= dl;
                } else {
                    cvt_raw<T>(yr[DOT ? 0 : j], dy);
                }
                if (RES) cvt_raw<T>(rr[RES ? j : 0], dx);
                else {
#pragma unroll
                    for (int e = 0; e < VW; e++) dx[e] = 0.f;
                }
#pragma unroll
                for (int e = 0; e < VW; e++) {
                    const float xh = fmaf(x[e], rs, nmr);
                    dx[e] += fmaf(dy[e] * gl[e], rs, fmaf(xh, k2, k1));
                    ag[j][e] = fmaf(DOT ? dl : dy[e], xh, ag[j][e]);
                    if (!DOT) ab[DOT ? 0 : j][DOT ? 0 : e] += dy[e];
                }
                int64_t wo = (MODE == LNM_MERGE) ? merge_off(mg, lr, c, Cin) : rowoff + c;
                if (MODE == LNM_UNSHUFFLE) {   // gradient written straight in the inverse depth-to-space layout
                    const RowCol rc = map_rc(MSU_MAP_UNSHUFFLE, ug.geo, lr, c);
                    wo = rc.row * (int64_t)(ug.geo[2] * ug.geo[2] * ug.geo[3]) + rc.col;
                }
                VecW<T>::st(dX + wo, dx);
                if (MODE == LNM_DUAL) {
#pragma unroll
                    for (int e = 0; e < VW; e++) dx[e] *= wsc;
                    VecW<T>::st(dXw + wrow * C + c, dx);
                }
            }
        }
    }
    // fold the row groups of this warp (lanes l, l+lpr, ...) in a fixed order, then one partial row per warp
    float* pg = partial + wid * 3 * (int64_t)C;
    float s0 = 0.f;
    if (DOT) {
        s0 = ab[0][0];
        for (int o = lpr; o < 32; o <<= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    }
#pragma unroll
    for (int j = 0; j < NV; j++) {
        for (int o = lpr; o < 32; o <<= 1) {
#pragma unroll
            for (int e = 0; e < VW; e++) {
                ag[j][e] += __shfl_xor_sync(0xffffffffu, ag[j][e], o);
                if (!DOT) ab[DOT ? 0 : j][DOT ? 0 : e] += __shfl_xor_sync(0xffffffffu, ab[DOT ? 0 : j][DOT ? 0 : e], o);
            }
        }
        const int c = (l + lpr * j) * VW;
        if (sub == 0 && c < C) {
            if (DOT) {   // dy = dl * w: dgamma = w S1, dbeta = w S0, d(dot weight) = gamma S1 + beta S0
                float w[VW], gl[VW], bt[VW];
                ldf(dotw + c, w, VW);
                ldf(gamma + c, gl, VW);
                ldf(beta + c, bt, VW);
#pragma unroll
                for (int e = 0; e < VW; e += 4) {
                    *reinterpret_cast<float4*>(pg + c + e) = make_float4(w[e] * ag[j][e], w[e + 1] * ag[j][e + 1], w[e + 2] * ag[j][e + 2], w[e + 3] * ag[j][e + 3]);
                    *reinterpret_cast<float4*>(pg + C + c + e) = make_float4(w[e] * s0, w[e + 1] * s0, w[e + 2] * s0, w[e + 3] * s0);
                    *reinterpret_cast<float4*>(pg + 2 * C + c + e) = make_float4(fmaf(gl[e], ag[j][e], bt[e] * s0), fmaf(gl[e + 1], ag[j][e + 1], bt[e + 1] * s0),
                                                                                 fmaf(gl[e + 2], ag[j][e + 2], bt[e + 2] * s0), fmaf(gl[e + 3], ag[j][e + 3], bt[e + 3] * s0));
                }
            } else {
#pragma unroll
                for (int e = 0; e < VW; e += 4) {
                    *reinterpret_cast<float4*>(pg + c + e) = make_float4(ag[j][e], ag[j][e + 1], ag[j][e + 2], ag[j][e + 3]);
                    *reinterpret_cast<float4*>(pg + C + c + e) = make_float4(ab[DOT ? 0 : j][DOT ? 0 : e], ab[DOT ? 0 : j][DOT ? 0 : e + 1], ab[DOT ? 0 : j][DOT ? 0 : e + 2], ab[DOT ? 0 : j][DOT ? 0 : e + 3]);
                }
            }
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
static int pick_lpr(int C, int vw) {
    const int nvec = (C + vw - 1) / vw;
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
    const int rpb = LN_WARPS * (32 / lpr) * LN_RG;
    const unsigned grid = (unsigned)((rows + rpb - 1) / rpb);
    const float invC = 1.0f / (float)C;
#define LN_FWD_LAUNCH(MODE)                                                                                               \
    ln_fwd_kernel<T, NV, MODE><<<grid, LN_WARPS * 32, 0, st>>>((const T*)X, gamma, beta, (T*)Y, mean, rstd, rows, C, invC, lpr, \
                                                              wg, mg, dotw)
    if (dotw != nullptr) {
        if (in_map != MSU_MAP_NONE || out_map != MSU_MAP_NONE) { set_error("msu_ln_fwd: dotw with a row map is not supported"); return -1; }
        LN_FWD_LAUNCH(LNM_DOT);
    } else if (out_map == MSU_MAP_WINDOW) {
        if (in_map != MSU_MAP_NONE) { set_error("msu_ln_fwd: in_map and out_map together are not supported"); return -1; }
        LN_FWD_LAUNCH(LNM_WINDOW);
    } else if (in_map == MSU_MAP_MERGE) {
        LN_FWD_LAUNCH(LNM_MERGE);
    } else {
        LN_FWD_LAUNCH(LNM_PLAIN);
    }
#undef LN_FWD_LAUNCH
    count_launch();
    return check_launch("msu_ln_fwd");
}

template <typename T, int NV>
static int launch_bwd(const void* dY, const void* X, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, const void* dRes, void* dX, int64_t rows, int C, int lpr, int dy_map, int dx_map,
                      const int32_t* geo, const float* dotw, float* partial, int grid, cudaStream_t st, void* dXw = nullptr,
                      const float* wscale = nullptr, int rows_per_sample = 1) {
    WinGeo wg{0, 0, 0, 0, 0, 0};
    MergeGeo mg{0, 0};
    if (dy_map == MSU_MAP_WINDOW || dXw != nullptr) wg = make_wingeo(geo);
    if (dx_map == MSU_MAP_MERGE) { mg.H = geo[0]; mg.W = geo[1]; }
    MsuOperand ug{};
    if (dx_map == MSU_MAP_UNSHUFFLE) for (int k = 0; k < 4; k++) ug.geo[k] = geo[k];
    const float invC = 1.0f / (float)C;
#define LN_BWD_LAUNCH(MODE, RES)                                                                                           \
    ln_bwd_kernel<T, NV, MODE, RES><<<grid, LN_WARPS * 32, 0, st>>>((const T*)dY, (const T*)X, gamma, beta, mean, rstd,     \
                                                                   (const T*)dRes, (T*)dX, rows, C, invC, lpr, wg, mg, dotw, partial, ug, \
                                                                   (T*)dXw, wscale, rows_per_sample)
    const int nmaps = (dy_map != MSU_MAP_NONE) + (dx_map != MSU_MAP_NONE) + (dotw != nullptr);
    if (nmaps > 1) { set_error("msu_ln_bwd: at most one of dy_map / dx_map / dotw"); return -1; }
    if (dXw != nullptr) {
        if (nmaps != 0) { set_error("msu_ln_bwd_dual: no other row map allowed"); return -1; }
        if (dRes) LN_BWD_LAUNCH(LNM_DUAL, true); else LN_BWD_LAUNCH(LNM_DUAL, false);
    } else if (dotw != nullptr) {
        if (dRes) LN_BWD_LAUNCH(LNM_DOT, true); else LN_BWD_LAUNCH(LNM_DOT, false);
    } else if (dy_map == MSU_MAP_WINDOW) {
        if (dRes) LN_BWD_LAUNCH(LNM_WINDOW, true); else LN_BWD_LAUNCH(LNM_WINDOW, false);
    } else if (dx_map == MSU_MAP_MERGE) {
        if (dRes) LN_BWD_LAUNCH(LNM_MERGE, true); else LN_BWD_LAUNCH(LNM_MERGE, false);
    } else if (dx_map == MSU_MAP_UNSHUFFLE) {
        if (dRes) LN_BWD_LAUNCH(LNM_UNSHUFFLE, true); else LN_BWD_LAUNCH(LNM_UNSHUFFLE, false);
    } else {
        if (dRes) LN_BWD_LAUNCH(LNM_PLAIN, true); else LN_BWD_LAUNCH(LNM_PLAIN, false);
    }
#undef LN_BWD_LAUNCH
    count_launch();
    return check_launch("msu_ln_bwd");
}

#define LN_DISPATCH_NV(FN, T, ...)                                         \
    do {                                                                   \
        const int vw_ = VecW<T>::N;                                        \
        const int nv = ((C + vw_ - 1) / vw_ + pick_lpr(C, vw_) - 1) / pick_lpr(C, vw_); \
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
    MSU_REQUIRE(C > 0 && C % (dtype == MSU_F32 ? 4 : 8) == 0, "msu_ln_fwd: C=%d must be a positive multiple of the 16-byte vector width", C);
    MSU_REQUIRE(in_map == MSU_MAP_NONE || in_map == MSU_MAP_MERGE, "msu_ln_fwd: bad in_map %d", in_map);
    MSU_REQUIRE(out_map == MSU_MAP_NONE || out_map == MSU_MAP_WINDOW, "msu_ln_fwd: bad out_map %d", out_map);
    MSU_REQUIRE((in_map == 0 && out_map == 0) || geo != nullptr, "msu_ln_fwd: geo required for mapped rows");
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_fwd, float, X, gamma, beta, Y, mean, rstd, rows, C, pick_lpr(C, dtype == MSU_F32 ? 4 : 8), in_map, out_map, geo, dotw, st);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_fwd, __nv_bfloat16, X, gamma, beta, Y, mean, rstd, rows, C, pick_lpr(C, dtype == MSU_F32 ? 4 : 8), in_map, out_map, geo, dotw, st);
    MSU_REQUIRE(false, "msu_ln_fwd: unsupported dtype %d", dtype);
}

// number of partial rows P = grid*8 warps: 2 resident blocks per SM, bounded so that the fp32 partial rows stay small
extern "C" int msu_ln_bwd_partial_rows(int dtype, int64_t rows, int32_t C) {
    const int rpw = 32 / pick_lpr(C, dtype == MSU_F32 ? 4 : 8);
    int64_t P = imin((rows + rpw - 1) / rpw, (int64_t)num_sms() * 2 * LN_WARPS);
    P = imin(P, (8ll << 20) / (3ll * C));
    P = imin(P, imax(LN_WARPS, rows / 8));    // keep the fp32 partial rows (3*C floats each) small next to the row data
    const int grid = (int)imax(1, (P + LN_WARPS - 1) / LN_WARPS);
    return grid * LN_WARPS;
}

extern "C" int msu_ln_bwd(int dtype, const void* dY, const void* X, const float* gamma, const float* beta,
                          const float* mean, const float* rstd, const void* dRes, void* dX, int64_t rows, int32_t C,
                          int32_t dy_map, int32_t dx_map, const int32_t* geo, const float* dotw, float* partial,
                          void* stream) {
    MSU_REQUIRE(dY && X && gamma && beta && mean && rstd && dX && partial, "msu_ln_bwd: null pointer");
    MSU_REQUIRE(C > 0 && C % (dtype == MSU_F32 ? 4 : 8) == 0, "msu_ln_bwd: C=%d must be a positive multiple of the 16-byte vector width", C);
    MSU_REQUIRE(dy_map == MSU_MAP_NONE || dy_map == MSU_MAP_WINDOW, "msu_ln_bwd: bad dy_map %d", dy_map);
    MSU_REQUIRE(dx_map == MSU_MAP_NONE || dx_map == MSU_MAP_MERGE || dx_map == MSU_MAP_UNSHUFFLE, "msu_ln_bwd: bad dx_map %d", dx_map);
    MSU_REQUIRE(dx_map != MSU_MAP_UNSHUFFLE || (geo != nullptr && geo[3] % (dtype == MSU_F32 ? 4 : 8) == 0), "msu_ln_bwd: UNSHUFFLE needs geo with a vector-aligned chunk width");
    const int grid = msu_ln_bwd_partial_rows(dtype, rows, C) / LN_WARPS;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_bwd, float, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, pick_lpr(C, dtype == MSU_F32 ? 4 : 8), dy_map, dx_map, geo, dotw, partial, grid, st);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_bwd, __nv_bfloat16, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, pick_lpr(C, dtype == MSU_F32 ? 4 : 8), dy_map, dx_map, geo, dotw, partial, grid, st);
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

// LayerNorm backward that also writes scale[sample] * dX in window order (rows pix -> (b, window, i) of wgeo): the operand of the
// attention projection's dgrad / wgrad, without a separate gather pass.  Padding rows of dXw are not written (keep them zero).
extern "C" int msu_ln_bwd_dual(int dtype, const void* dY, const void* X, const float* gamma, const float* beta,
                               const float* mean, const float* rstd, const void* dRes, void* dX, void* dXw, int64_t rows,
                               int32_t C, const int32_t* wgeo, const float* rowscale, int32_t rows_per_sample,
                               float* partial, void* stream) {
    MSU_REQUIRE(dY && X && gamma && beta && mean && rstd && dX && dXw && wgeo && partial, "msu_ln_bwd_dual: null pointer");
    MSU_REQUIRE(C > 0 && C % (dtype == MSU_F32 ? 4 : 8) == 0, "msu_ln_bwd_dual: C=%d must be a positive multiple of the 16-byte vector width", C);
    MSU_REQUIRE(rowscale == nullptr || rows_per_sample > 0, "msu_ln_bwd_dual: rows_per_sample required with rowscale");
    const int grid = msu_ln_bwd_partial_rows(dtype, rows, C) / LN_WARPS;
    cudaStream_t st = (cudaStream_t)stream;
    const int rps = rows_per_sample > 0 ? rows_per_sample : 1;
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_bwd, float, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, pick_lpr(C, 4), 0, 0, wgeo, nullptr, partial, grid, st, dXw, rowscale, rps);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_bwd, __nv_bfloat16, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, pick_lpr(C, 8), 0, 0, wgeo, nullptr, partial, grid, st, dXw, rowscale, rps);
    MSU_REQUIRE(false, "msu_ln_bwd_dual: unsupported dtype %d", dtype);
}
