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

// Head forward (LayerNorm + 1x1 conv as a dot product, network/model_parts.py:842-846; the default bf16 path has it in the second
// conv's epilogue instead, MsuEpilogue.lnd_*).  Persistent warps: warp w walks super-groups w, w + P, ... of LN_RG row groups (rpw
// rows each); the 16 B loads of the NEXT super-group are issued before the current one is reduced (register double buffer).
// y is never formed:
//   logit = sum_c ((x_c - mu) rs gamma_c + beta_c) w_c = rs (sum_c x_c gw_c - mu G) + Bw,   gw = gamma w, G = sum gw, Bw = sum beta w
// i.e. 3 FMAs per element on top of the statistics, gw in registers (NV <= 3 vectors per lane).
constexpr int LN_RG = 2;
template <typename T, int NV>
struct LnGroup {
    int64_t r[LN_RG];
    uint4 xr[LN_RG][NV];
};
// window row -> pixel row with 32-bit arithmetic (row counts are far below 2^31; the 64-bit divisions of win_to_pix cost
// ~100 instructions per row)
__device__ __forceinline__ int64_t win_to_pix32(const WinGeo& g, uint32_t wr) {
    const uint32_t per_img = (uint32_t)(g.nwin() * WT);
    const uint32_t b = wr / per_img, r = wr - b * per_img;
    const uint32_t w = r / WT, i = r - w * WT;
    const uint32_t nwx = (uint32_t)g.nwx();
    const uint32_t wy = w / nwx, wx = w - wy * nwx, iy = i / WS, ix = i - iy * WS;
    int py = (int)(wy * WS + iy) + g.sh; if (py >= g.Ph) py -= g.Ph;
    int px = (int)(wx * WS + ix) + g.sw; if (px >= g.Pw) px -= g.Pw;
    if (py >= g.H || px >= g.W) return -1;
    return (int64_t)b * (g.H * g.W) + py * g.W + px;
}
__device__ __forceinline__ int64_t pix_to_win32(const WinGeo& g, uint32_t pr) {
    const uint32_t hw = (uint32_t)(g.H * g.W);
    const uint32_t b = pr / hw, r = pr - b * hw;
    const uint32_t y = r / (uint32_t)g.W, x = r - y * (uint32_t)g.W;
    int ry = (int)y - g.sh; if (ry < 0) ry += g.Ph;
    int rx = (int)x - g.sw; if (rx < 0) rx += g.Pw;
    const int wy = ry / WS, wx = rx / WS;
    return (int64_t)b * (g.nwin() * WT) + (wy * g.nwx() + wx) * WT + (ry - wy * WS) * WS + (rx - wx * WS);
}

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, 2) ln_fwd_dot_kernel(const T* __restrict__ X, const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta, T* __restrict__ Y,
                                                                      float* __restrict__ mean, float* __restrict__ rstd,
                                                                      int64_t rows, int C, float invC, int lpr,
                                                                      const float* __restrict__ dotw, float* __restrict__ dot_m2) {
    static_assert(NV <= 3, "gamma * w lives in registers");
    constexpr int VW = VecW<T>::N;
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr, sub = lane / lpr, l = lane - sub * lpr;
    const int64_t P = (int64_t)gridDim.x * LN_WARPS;
    const int64_t wid = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const int64_t nsuper = (rows + (int64_t)rpw * LN_RG - 1) / ((int64_t)rpw * LN_RG);
    bool vld[NV];
    float pg[NV][VW];
    float Gsum = 0.f, Bsum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (l + lpr * j) * VW;
        vld[j] = c < C;
#pragma unroll
        for (int e = 0; e < VW; e++) pg[j][e] = 0.f;
        if (vld[j]) {
            float wl[VW], bt[VW];
            ldf(gamma + c, pg[j], VW);
            ldf(dotw + c, wl, VW);
            ldf(beta + c, bt, VW);
#pragma unroll
            for (int e = 0; e < VW; e++) {
                pg[j][e] *= wl[e];
                Gsum += pg[j][e];
                Bsum = fmaf(bt[e], wl[e], Bsum);
            }
        }
    }
    Gsum = group_sum_u(Gsum, lpr);
    Bsum = group_sum_u(Bsum, lpr);

    auto load = [&](int64_t sg, LnGroup<T, NV>& G) {
#pragma unroll
        for (int g = 0; g < LN_RG; g++) {
            G.r[g] = (sg * LN_RG + g) * rpw + sub;
#pragma unroll
            for (int j = 0; j < NV; j++) {
                G.xr[g][j] = make_uint4(0, 0, 0, 0);
                if (G.r[g] < rows && vld[j]) G.xr[g][j] = *reinterpret_cast<const uint4*>(X + G.r[g] * C + (l + lpr * j) * VW);
            }
        }
    };

    LnGroup<T, NV> cur;
    if (wid < nsuper) load(wid, cur);
    for (int64_t sg = wid; sg < nsuper; sg += P) {
        LnGroup<T, NV> nxt;
        if (sg + P < nsuper) load(sg + P, nxt);
#pragma unroll
        for (int g = 0; g < LN_RG; g++) {
            float x[NV][VW];
            float s = 0.f, d = 0.f;
#pragma unroll
            for (int j = 0; j < NV; j++) {
                cvt_raw<T>(cur.xr[g][j], x[j]);
#pragma unroll
                for (int e = 0; e < VW; e++) {
                    s += x[j][e];
                    d = fmaf(x[j][e], pg[j][e], d);
                }
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
            d = group_sum_u(d, lpr);
            if (cur.r[g] < rows && l == 0) {
                const float dc = fmaf(-mu, Gsum, d) * rs;           // sum_c gw_c x-hat_c
                mean[cur.r[g]] = mu;
                rstd[cur.r[g]] = rs;
                Y[cur.r[g]] = from_f<T>(dc + Bsum);
                dot_m2[cur.r[g]] = dc * invC;                         // the backward's row mean of g * x-hat, per unit d(logit)
            }
        }
        cur = nxt;
    }
}

// Forward: every warp owns LN_RG row groups (rpw rows each) whose 16 B loads are all issued before the first
// reduction, so ~2x the bytes are in flight per warp (the one-group version topped out at 3.4 TB/s).
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
            lr[g] = win_to_pix32(wg, (uint32_t)r[g]);
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
// The operands of a row group (x, dy, residual gradient: 16 B vectors) travel global -> shared memory with cp.async, each lane
// into its own slots of a two-stage per-warp ring, so the NEXT group's loads are in flight while the current one is
// reduced and no register is spent on data in flight (the accumulators already take 2 * NV * VW of them).  A lane only ever
// reads back what it copied itself: cp.async.wait_group is all the synchronisation there is.  2 blocks (16 warps) per SM
// keep >= 74 KB of loads in flight continuously (the register-staged version had them in flight a third of the time).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 16 : 0;           // 0: nothing is read, the slot is zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NV> __host__ __device__ constexpr int ln_bwd_stages() { return NV <= 6 ? 2 : 1; }

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
    constexpr int NOPS = (DOT ? 1 : 2) + (RES ? 1 : 0);      // x, dy, dres
    constexpr int OP_DY = 1, OP_RES = DOT ? 1 : 2;
    constexpr int STAGES = ln_bwd_stages<NV>();
    extern __shared__ __align__(16) uint4 ln_ring[];          // [warp][stage][op][j][lane]
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr, sub = lane / lpr, l = lane - sub * lpr;
    const int64_t wid = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const int64_t P = (int64_t)gridDim.x * LN_WARPS;
    constexpr int VW = VecW<T>::N;
    const int Cin = C / 4;
    uint4* ring = ln_ring + (size_t)(threadIdx.x >> 5) * (STAGES * NOPS * NV * 32) + lane;
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

    struct RowInfo { int64_t lr, wrow; bool live; float mu, rs, dl, wsc; };
    auto issue = [&](int64_t g0, int stage, RowInfo& ri) {
        ri.lr = g0 + sub;
        ri.live = ri.lr < rows;
        ri.mu = ri.live ? mean[ri.lr] : 0.f;
        ri.rs = ri.live ? rstd[ri.lr] : 0.f;
        const int64_t dyr = (MODE == LNM_WINDOW && ri.live) ? pix_to_win32(wg, (uint32_t)ri.lr) : ri.lr;
        // DUAL: the same gradient row, scaled per sample, also goes to its window-order row (the padding rows of that
        // buffer are zeroed once by the caller and never written)
        ri.wrow = (MODE == LNM_DUAL && ri.live) ? pix_to_win32(wg, (uint32_t)ri.lr) : 0;
        ri.wsc = (MODE == LNM_DUAL && ri.live && wscale != nullptr) ? wscale[ri.lr / rows_per_sample] : 1.0f;
        ri.dl = (DOT && ri.live) ? to_f<T>(dY[ri.lr]) : 0.f;
        uint4* st = ring + (size_t)stage * (NOPS * NV * 32);
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int c = (l + lpr * j) * VW;
            const bool on = ri.live && vld[j];
            const int64_t xo = !on ? 0 : ((MODE == LNM_MERGE) ? merge_off(mg, ri.lr, c, Cin) : ri.lr * C + c);
            cp_async16(st + (0 * NV + j) * 32, X + xo, on);
            if (!DOT) cp_async16(st + (OP_DY * NV + j) * 32, dY + (on ? dyr * C + c : 0), on);
            if (RES) cp_async16(st + (OP_RES * NV + j) * 32, dRes + xo, on);
        }
        cp_async_commit();
    };

    RowInfo cur, nxt;
    int stage = 0;
    if (wid * rpw < rows) issue(wid * rpw, 0, cur);
    for (int64_t g0 = wid * rpw; g0 < rows; g0 += P * rpw) {
        const bool more = g0 + P * rpw < rows;
        if (STAGES == 2) {
            if (more) issue(g0 + P * rpw, stage ^ 1, nxt);
            if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        } else {
            cp_async_wait<0>();
        }
        const uint4* st = ring + (size_t)stage * (NOPS * NV * 32);
        const int64_t lr = cur.lr;
        const bool live = cur.live;
        const float rs = cur.rs, nmr = -cur.mu * cur.rs, dl = cur.dl;
        const int64_t rowoff = lr * C;
        // pass 1: row sums of g = dy * gamma and g * x-hat
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            if (live && vld[j]) {
                const int c = (l + lpr * j) * VW;
                float x[VW], dy[VW], gl[VW];
                cvt_raw<T>(st[(0 * NV + j) * 32], x);
                ldf(gamma + c, gl, VW);
                if (DOT) {
                    ldf(dotw + c, dy, VW);
#pragma unroll
                    for (int e = 0; e < VW; e++) dy[e] *= dl;
                } else {
                    cvt_raw<T>(st[(OP_DY * NV + j) * 32], dy);
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
                cvt_raw<T>(st[(0 * NV + j) * 32], x);
                ldf(gamma + c, gl, VW);
                if (DOT) {
                    ldf(dotw + c, dy, VW);
#pragma unroll
                    for (int e = 0; e < VW; e++) dy[e] *= dl;
                } else {
                    cvt_raw<T>(st[(OP_DY * NV + j) * 32], dy);
                }
                if (RES) cvt_raw<T>(st[(OP_RES * NV + j) * 32], dx);
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
                    for (int e = 0; e < VW; e++) dx[e] *= cur.wsc;
                    VecW<T>::st(dXw + cur.wrow * C + c, dx);
                }
            }
        }
        if (STAGES == 2) { cur = nxt; stage ^= 1; }
        else if (more) issue(g0 + P * rpw, 0, cur);
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

// Head backward (LayerNorm + 1x1 conv as a dot product, network/model_parts.py:842-846): with dy_c = dl w_c the two row
// reductions of the generic backward are known in closed form from what the forward saved,
//   mean_c(dy_c gamma_c) = dl G / C,   mean_c(dy_c gamma_c x-hat_c) = dl m2      (gw = gamma w, G = sum gw, m2 from the forward)
//   dx_c = rs dl (gw_c - G/C) - rs^2 dl m2 (x_c - mu)
// so a row needs no shuffle at all: 2 FMAs + 1 FMUL per element for dx and one FMA for the parameter sums
// S1_c = sum_rows dl x-hat_c, S0 = sum_rows dl  (d gamma = w S1, d beta = w S0, d w = gamma S1 + beta S0).
// The generic kernel spent ~14 instructions per element here and was issue-bound at 3.8 TB/s.
template <typename T, int NV, bool RES>
__global__ void __launch_bounds__(LN_WARPS * 32, 2) ln_bwd_dot_kernel(const T* __restrict__ dL, const T* __restrict__ X,
                                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                      const float* __restrict__ m2, const T* __restrict__ dRes,
                                                                      T* __restrict__ dX, int64_t rows, int C, float invC, int lpr,
                                                                      const float* __restrict__ dotw, float* __restrict__ partial) {
    constexpr int VW = VecW<T>::N;
    constexpr int RG = 2;
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr, sub = lane / lpr, l = lane - sub * lpr;
    const int64_t wid = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const int64_t P = (int64_t)gridDim.x * LN_WARPS;
    bool vld[NV];
    float gwc[NV][VW], ag[NV][VW];
    float G = 0.f, s0 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int c = (l + lpr * j) * VW;
        vld[j] = c < C;
#pragma unroll
        for (int e = 0; e < VW; e++) { gwc[j][e] = 0.f; ag[j][e] = 0.f; }
        if (vld[j]) {
            float wl[VW];
            ldf(gamma + c, gwc[j], VW);
            ldf(dotw + c, wl, VW);
#pragma unroll
            for (int e = 0; e < VW; e++) { gwc[j][e] *= wl[e]; G += gwc[j][e]; }
        }
    }
    G = group_sum_u(G, lpr) * invC;
#pragma unroll
    for (int j = 0; j < NV; j++) {
#pragma unroll
        for (int e = 0; e < VW; e++) gwc[j][e] -= G;
    }
    // x (and the residual gradient) of RG row groups per stage go through a two-stage per-warp cp.async ring (see ln_bwd_kernel)
    constexpr int NOPS = RES ? 2 : 1;
    extern __shared__ __align__(16) uint4 ln_ring[];          // [warp][stage][group][op][j][lane]
    uint4* ring = ln_ring + (size_t)(threadIdx.x >> 5) * (2 * RG * NOPS * NV * 32) + lane;
    struct Grp { int lr[RG]; float mu[RG], a[RG], b[RG]; };   // rows < 2^31
    auto issue = [&](int64_t g0, int stage, Grp& Gp) {
        uint4* st = ring + (size_t)stage * (RG * NOPS * NV * 32);
#pragma unroll
        for (int g = 0; g < RG; g++) {
            const int64_t lr = g0 + (int64_t)g * rpw + sub;
            const bool live = lr < rows;
            Gp.lr[g] = live ? (int)lr : -1;
            float mu = 0.f, rs = 0.f, mm = 0.f, dl = 0.f;
            if (live) { mu = mean[lr]; rs = rstd[lr]; mm = m2[lr]; dl = to_f<T>(dL[lr]); }
            s0 += dl;
            Gp.mu[g] = mu;
            Gp.a[g] = rs * dl;
            Gp.b[g] = -rs * dl * mm * rs;
#pragma unroll
            for (int j = 0; j < NV; j++) {
                const int c = (l + lpr * j) * VW;
                const bool on = live && vld[j];
                const int64_t xo = on ? lr * C + c : 0;
                cp_async16(st + ((g * NOPS + 0) * NV + j) * 32, X + xo, on);
                if (RES) cp_async16(st + ((g * NOPS + 1) * NV + j) * 32, dRes + xo, on);
            }
        }
        cp_async_commit();
    };
    Grp cur, nxt;
    int stage = 0;
    const int64_t step = P * rpw * RG;
    if (wid * rpw * RG < rows) issue(wid * rpw * RG, 0, cur);
    for (int64_t g0 = wid * rpw * RG; g0 < rows; g0 += step) {
        const bool more = g0 + step < rows;
        if (more) issue(g0 + step, stage ^ 1, nxt);
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        const uint4* st = ring + (size_t)stage * (RG * NOPS * NV * 32);
#pragma unroll
        for (int g = 0; g < RG; g++) {
            const float mu = cur.mu[g], a = cur.a[g], b = cur.b[g];
            const bool live = cur.lr[g] >= 0;
#pragma unroll
            for (int j = 0; j < NV; j++) {
                if (live && vld[j]) {
                    float x[VW], dx[VW];
                    cvt_raw<T>(st[((g * NOPS + 0) * NV + j) * 32], x);
                    if (RES) cvt_raw<T>(st[((g * NOPS + (RES ? 1 : 0)) * NV + j) * 32], dx);
                    else {
#pragma unroll
                        for (int e = 0; e < VW; e++) dx[e] = 0.f;
                    }
#pragma unroll
                    for (int e = 0; e < VW; e++) {
                        const float xm = x[e] - mu;
                        dx[e] += fmaf(b, xm, a * gwc[j][e]);
                        ag[j][e] = fmaf(a, xm, ag[j][e]);
                    }
                    VecW<T>::st(dX + (int64_t)cur.lr[g] * C + (l + lpr * j) * VW, dx);
                }
            }
        }
        cur = nxt;
        stage ^= 1;
    }
    // fold the rows of this warp (lanes l, l + lpr, ...) in a fixed order, then one partial row per warp
    float* pg = partial + wid * 3 * (int64_t)C;
    for (int o = lpr; o < 32; o <<= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
#pragma unroll
    for (int j = 0; j < NV; j++) {
        for (int o = lpr; o < 32; o <<= 1) {
#pragma unroll
            for (int e = 0; e < VW; e++) ag[j][e] += __shfl_xor_sync(0xffffffffu, ag[j][e], o);
        }
        const int c = (l + lpr * j) * VW;
        if (sub == 0 && c < C) {
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
        }
    }
}

// out[a][c] = sum_p partial[p][a][c]; block = 8 columns (one 32 B sector per partial row) x 128 row-lanes, fixed-order tree =>
// deterministic.  (32 columns x 32 row-lanes gave 9 CTAs at C = 96, each thread walking 74 rows: 8.6 us per call inside the step.)
constexpr int LPR_COLS = 8, LPR_ROWS = 128;
__global__ void __launch_bounds__(LPR_COLS * LPR_ROWS) ln_param_reduce_kernel(const float* __restrict__ partial, int P, int C,
                                                                              float* dgamma, float* dbeta, float* ddotw, int accumulate) {
    __shared__ float sm[LPR_ROWS][LPR_COLS + 1];
    __shared__ float sm2[4][LPR_COLS];
    const int a = blockIdx.y;
    float* out = a == 0 ? dgamma : (a == 1 ? dbeta : ddotw);
    if (out == nullptr) return;
    const int c = blockIdx.x * LPR_COLS + threadIdx.x;
    float s = 0.f;
    if (c < C)
        for (int p = threadIdx.y; p < P; p += LPR_ROWS) s += partial[((int64_t)p * 3 + a) * C + c];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y < 4) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < LPR_ROWS / 4; i++) t += sm[threadIdx.y * (LPR_ROWS / 4) + i][threadIdx.x];
        sm2[threadIdx.y][threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        const float t = (sm2[0][threadIdx.x] + sm2[1][threadIdx.x]) + (sm2[2][threadIdx.x] + sm2[3][threadIdx.x]);
        out[c] = accumulate ? out[c] + t : t;
    }
}

// lanes per row: power of two in [4, 32] giving at most 3 vectors (16 B) per lane while the row fits 32 lanes x 3
static int pick_lpr(int C, int vw) {
    const int nvec = (C + vw - 1) / vw;
    int lpr = 4;
    while (lpr < 32 && lpr * 3 < nvec) lpr <<= 1;
    return lpr;
}

template <typename T, int NV>
static int launch_fwd(const void* X, const float* gamma, const float* beta, void* Y, float* mean, float* rstd,
                      int64_t rows, int C, int lpr, int in_map, int out_map, const int32_t* geo, const float* dotw,
                      float* dot_m2, cudaStream_t st) {
    WinGeo wg{0, 0, 0, 0, 0, 0};
    MergeGeo mg{0, 0};
    if (out_map == MSU_MAP_WINDOW) wg = make_wingeo(geo);
    if (in_map == MSU_MAP_MERGE) { mg.H = geo[0]; mg.W = geo[1]; }
    const int rpb = LN_WARPS * (32 / lpr) * LN_RG;
    const int64_t want = (rows + rpb - 1) / rpb;
    const unsigned grid = (unsigned)want;
    const float invC = 1.0f / (float)C;
#define LN_FWD_LAUNCH(MODE)                                                                                               \
    ln_fwd_kernel<T, NV, MODE><<<grid, LN_WARPS * 32, 0, st>>>((const T*)X, gamma, beta, (T*)Y, mean, rstd, rows, C, invC, lpr, \
                                                              wg, mg, dotw)
    if (dotw != nullptr) {
        if (in_map != MSU_MAP_NONE || out_map != MSU_MAP_NONE) { set_error("msu_ln_fwd: dotw with a row map is not supported"); return -1; }
        if (NV > 3 || dot_m2 == nullptr) { set_error("msu_ln_fwd: dotw needs C <= %d and a dot_m2 output", 96 * VecW<T>::N); return -1; }
        // persistent, register-prefetched, y never formed: the resident block count (2 per SM)
        if constexpr (NV <= 3) {
            const unsigned pgrid = (unsigned)imax(1, imin(want, (int64_t)num_sms() * 2));
            ln_fwd_dot_kernel<T, NV><<<pgrid, LN_WARPS * 32, 0, st>>>((const T*)X, gamma, beta, (T*)Y, mean, rstd, rows, C, invC, lpr, dotw, dot_m2);
        }
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
                      const float* wscale = nullptr, int rows_per_sample = 1, const float* dot_m2 = nullptr) {
    WinGeo wg{0, 0, 0, 0, 0, 0};
    MergeGeo mg{0, 0};
    if (dy_map == MSU_MAP_WINDOW || dXw != nullptr) wg = make_wingeo(geo);
    if (dx_map == MSU_MAP_MERGE) { mg.H = geo[0]; mg.W = geo[1]; }
    MsuOperand ug{};
    if (dx_map == MSU_MAP_UNSHUFFLE) for (int k = 0; k < 4; k++) ug.geo[k] = geo[k];
    const float invC = 1.0f / (float)C;
#define LN_BWD_LAUNCH(MODE, RES)                                                                                           \
    do {                                                                                                                   \
        constexpr int nops_ = ((MODE) == LNM_DOT ? 1 : 2) + ((RES) ? 1 : 0);                                               \
        constexpr int smem_ = LN_WARPS * ln_bwd_stages<NV>() * nops_ * NV * 32 * 16;                                        \
        static PerDeviceOnce attr_;                                                                                        \
        if (smem_ > 48 * 1024 && attr_.need()) {                                                                           \
            cudaFuncSetAttribute(ln_bwd_kernel<T, NV, MODE, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_);     \
            attr_.set();                                                                                                   \
        }                                                                                                                  \
        ln_bwd_kernel<T, NV, MODE, RES><<<grid, LN_WARPS * 32, smem_, st>>>((const T*)dY, (const T*)X, gamma, beta, mean, rstd, \
                                                                   (const T*)dRes, (T*)dX, rows, C, invC, lpr, wg, mg, dotw, partial, ug, \
                                                                   (T*)dXw, wscale, rows_per_sample);                      \
    } while (0)
    const int nmaps = (dy_map != MSU_MAP_NONE) + (dx_map != MSU_MAP_NONE) + (dotw != nullptr);
    if (nmaps > 1) { set_error("msu_ln_bwd: at most one of dy_map / dx_map / dotw"); return -1; }
    if (dXw != nullptr) {
        if (nmaps != 0) { set_error("msu_ln_bwd_dual: no other row map allowed"); return -1; }
        if (dRes) LN_BWD_LAUNCH(LNM_DUAL, true); else LN_BWD_LAUNCH(LNM_DUAL, false);
    } else if (dotw != nullptr) {
        if constexpr (NV <= 3) {
            if (dot_m2 == nullptr) { set_error("msu_ln_bwd: dotw needs the forward's dot_m2"); return -1; }
            constexpr int sm1 = LN_WARPS * 2 * 2 * NV * 32 * 16;          // [warps][stages][RG][op][NV][lane] x 16 B
            static PerDeviceOnce attr_;
            if (attr_.need()) {
                cudaFuncSetAttribute(ln_bwd_dot_kernel<T, NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * sm1);
                cudaFuncSetAttribute(ln_bwd_dot_kernel<T, NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm1);
                attr_.set();
            }
            if (dRes) ln_bwd_dot_kernel<T, NV, true><<<grid, LN_WARPS * 32, 2 * sm1, st>>>((const T*)dY, (const T*)X, gamma, beta, mean, rstd, dot_m2,
                                                                                         (const T*)dRes, (T*)dX, rows, C, invC, lpr, dotw, partial);
            else ln_bwd_dot_kernel<T, NV, false><<<grid, LN_WARPS * 32, sm1, st>>>((const T*)dY, (const T*)X, gamma, beta, mean, rstd, dot_m2,
                                                                                   (const T*)dRes, (T*)dX, rows, C, invC, lpr, dotw, partial);
        } else {
            set_error("msu_ln_bwd: dotw needs C <= %d", 96 * VecW<T>::N);
            return -1;
        }
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
                          const float* dotw, float* dot_m2, void* stream) {
    MSU_REQUIRE(X && gamma && beta && Y && mean && rstd, "msu_ln_fwd: null pointer");
    MSU_REQUIRE(C > 0 && C % (dtype == MSU_F32 ? 4 : 8) == 0, "msu_ln_fwd: C=%d must be a positive multiple of the 16-byte vector width", C);
    MSU_REQUIRE(in_map == MSU_MAP_NONE || in_map == MSU_MAP_MERGE, "msu_ln_fwd: bad in_map %d", in_map);
    MSU_REQUIRE(out_map == MSU_MAP_NONE || out_map == MSU_MAP_WINDOW, "msu_ln_fwd: bad out_map %d", out_map);
    MSU_REQUIRE((in_map == 0 && out_map == 0) || geo != nullptr, "msu_ln_fwd: geo required for mapped rows");
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_fwd, float, X, gamma, beta, Y, mean, rstd, rows, C, pick_lpr(C, dtype == MSU_F32 ? 4 : 8), in_map, out_map, geo, dotw, dot_m2, st);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_fwd, __nv_bfloat16, X, gamma, beta, Y, mean, rstd, rows, C, pick_lpr(C, dtype == MSU_F32 ? 4 : 8), in_map, out_map, geo, dotw, dot_m2, st);
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
                          int32_t dy_map, int32_t dx_map, const int32_t* geo, const float* dotw, const float* dot_m2,
                          float* partial, void* stream) {
    MSU_REQUIRE(dY && X && gamma && beta && mean && rstd && dX && partial, "msu_ln_bwd: null pointer");
    MSU_REQUIRE(C > 0 && C % (dtype == MSU_F32 ? 4 : 8) == 0, "msu_ln_bwd: C=%d must be a positive multiple of the 16-byte vector width", C);
    MSU_REQUIRE(dy_map == MSU_MAP_NONE || dy_map == MSU_MAP_WINDOW, "msu_ln_bwd: bad dy_map %d", dy_map);
    MSU_REQUIRE(dx_map == MSU_MAP_NONE || dx_map == MSU_MAP_MERGE || dx_map == MSU_MAP_UNSHUFFLE, "msu_ln_bwd: bad dx_map %d", dx_map);
    MSU_REQUIRE(dx_map != MSU_MAP_UNSHUFFLE || (geo != nullptr && geo[3] % (dtype == MSU_F32 ? 4 : 8) == 0), "msu_ln_bwd: UNSHUFFLE needs geo with a vector-aligned chunk width");
    const int grid = msu_ln_bwd_partial_rows(dtype, rows, C) / LN_WARPS;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSU_F32) LN_DISPATCH_NV(launch_bwd, float, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, pick_lpr(C, dtype == MSU_F32 ? 4 : 8), dy_map, dx_map, geo, dotw, partial, grid, st, nullptr, nullptr, 1, dot_m2);
    if (dtype == MSU_BF16) LN_DISPATCH_NV(launch_bwd, __nv_bfloat16, dY, X, gamma, beta, mean, rstd, dRes, dX, rows, C, pick_lpr(C, dtype == MSU_F32 ? 4 : 8), dy_map, dx_map, geo, dotw, partial, grid, st, nullptr, nullptr, 1, dot_m2);
    MSU_REQUIRE(false, "msu_ln_bwd: unsupported dtype %d", dtype);
}

extern "C" int msu_ln_param_reduce(const float* partial, int32_t P, int32_t C, float* dgamma, float* dbeta,
                                   float* ddotw, int accumulate, void* stream) {
    MSU_REQUIRE(partial && P > 0 && C > 0, "msu_ln_param_reduce: bad arguments");
    dim3 grid((unsigned)((C + LPR_COLS - 1) / LPR_COLS), 3), block(LPR_COLS, LPR_ROWS);
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
