// Shifted-window attention core (49 tokens x head-dim 32 per (window, head)), forward and backward.
//
// Operates on window-ordered qkv rows (produced by the LN1 window gather + QKV GEMM), so torch.roll,
// window_partition and the zero padding never materialise; the relative-position bias comes from the
// expanded [nH,49,49] table and the -100 shift mask is derived from the window index on the fly
// (TV:models/swin_transformer.py:181-214).  Softmax statistics are fp32.  Backward recomputes P
// (nothing of size [windows, heads, 49, 49] is ever written to HBM) and reduces d(bias) deterministically.
//
// v1 maps one warp to one (window, head) with fp32 FMA; lanes own query rows (phase 1) or key rows
// (phase 2) and the other operand is broadcast from shared memory.
#include "common.cuh"

namespace msu {

constexpr int AW = 4;                   // warps per CTA
constexpr float QK_SCALE = 0.17677669529663687f;  // 32^-1/2 (TV:...:188; qk_scale never reaches torchvision)

struct MaskInfo {
    bool any;      // this window carries a mask
    int ty, tx;    // local row/col threshold: region bit = (a >= ty), (b >= tx); 7 => never
};
__device__ __forceinline__ MaskInfo mask_info(const WinGeo& g, int w) {
    MaskInfo m;
    const int wy = w / g.nwx(), wx = w % g.nwx();
    m.ty = (g.sh > 0 && wy == g.Ph / WS - 1) ? WS - g.sh : WS;
    m.tx = (g.sw > 0 && wx == g.Pw / WS - 1) ? WS - g.sw : WS;
    m.any = (m.ty < WS) || (m.tx < WS);
    return m;
}
__device__ __forceinline__ int region_of(const MaskInfo& m, int t) {
    const int a = t / WS, b = t - a * WS;
    return (a >= m.ty ? 2 : 0) + (b >= m.tx ? 1 : 0);
}

// cooperative (one warp) load of a [49,32] head slice into smem as fp32, optional scale
template <typename T>
__device__ __forceinline__ void load_tile(float* dst, const T* src, int64_t ld, int lane, float scale) {
    for (int idx = lane; idx < WT * (HD / 4); idx += 32) {
        const int r = idx >> 3, v = idx & 7;
        float4 x = Vec4<T>::ld(src + r * ld + v * 4);
        x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
        *reinterpret_cast<float4*>(dst + r * HD + v * 4) = x;
    }
}
template <typename T>
__device__ __forceinline__ void load_row(float* reg, const T* src, float scale) {
#pragma unroll
    for (int v = 0; v < HD / 4; v++) {
        const float4 x = Vec4<T>::ld(src + v * 4);
        reg[4 * v] = x.x * scale; reg[4 * v + 1] = x.y * scale; reg[4 * v + 2] = x.z * scale; reg[4 * v + 3] = x.w * scale;
    }
}
__device__ __forceinline__ float dot_row(const float* reg, const float* srow) {
    float a = 0.f;
#pragma unroll
    for (int v = 0; v < HD / 4; v++) {
        const float4 k = *reinterpret_cast<const float4*>(srow + v * 4);
        a = fmaf(reg[4 * v], k.x, a); a = fmaf(reg[4 * v + 1], k.y, a);
        a = fmaf(reg[4 * v + 2], k.z, a); a = fmaf(reg[4 * v + 3], k.w, a);
    }
    return a;
}
__device__ __forceinline__ void axpy_row(float* reg, float s, const float* srow) {
#pragma unroll
    for (int v = 0; v < HD / 4; v++) {
        const float4 k = *reinterpret_cast<const float4*>(srow + v * 4);
        reg[4 * v] = fmaf(s, k.x, reg[4 * v]); reg[4 * v + 1] = fmaf(s, k.y, reg[4 * v + 1]);
        reg[4 * v + 2] = fmaf(s, k.z, reg[4 * v + 2]); reg[4 * v + 3] = fmaf(s, k.w, reg[4 * v + 3]);
    }
}
template <typename T>
__device__ __forceinline__ void store_row(T* dst, const float* reg, float scale) {
#pragma unroll
    for (int v = 0; v < HD / 4; v++)
        Vec4<T>::st(dst + v * 4, make_float4(reg[4 * v] * scale, reg[4 * v + 1] * scale, reg[4 * v + 2] * scale,
                                              reg[4 * v + 3] * scale));
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(AW * 32) winattn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ bias,
                                                             T* __restrict__ O, int64_t n_windows, int nH, WinGeo g, AttnDrop ad,
                                                             float* __restrict__ lse) {
    extern __shared__ __align__(16) float smem[];
    float* sbias = smem;                                   // [49*49]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sK = smem + 2404 + warp * (2 * WT * HD + WT * 32);
    float* sV = sK + WT * HD;
    float* sc = sV + WT * HD;                              // [49][32] scores, lane-private columns
    const int h = blockIdx.y;
    const int C = nH * HD;
    for (int i = threadIdx.x; i < WT * WT; i += blockDim.x) sbias[i] = bias[h * WT * WT + i];
    __syncthreads();
    const int nwin_img = g.nwin();
    const uint32_t ds0 = ad.thr ? ad.seed[0] : 0u, ds1 = ad.thr ? ad.seed[1] : 0u;
    for (int64_t win = (int64_t)blockIdx.x * AW + warp; win < n_windows; win += (int64_t)gridDim.x * AW) {
        const T* base = qkv + win * WT * 3 * (int64_t)C + h * HD;
        load_tile<T>(sK, base + C, 3 * C, lane, 1.f);
        load_tile<T>(sV, base + 2 * C, 3 * C, lane, 1.f);
        __syncwarp();
        const MaskInfo mi = mask_info(g, (int)(win % nwin_img));
        for (int rnd = 0; rnd < 2; rnd++) {
            const int i = lane + 32 * rnd;
            const bool act = i < WT;
            const int ic = act ? i : WT - 1;
            float q[HD];
            load_row<T>(q, base + (int64_t)ic * 3 * C, QK_SCALE);
            const int ri = region_of(mi, ic);
            float mx = -INFINITY;
            for (int j = 0; j < WT; j++) {
                float s = dot_row(q, sK + j * HD) + sbias[ic * WT + j];
                if (mi.any && region_of(mi, j) != ri) s += -100.0f;
                sc[j * 32 + lane] = s;
                mx = fmaxf(mx, s);
            }
            float o[HD];
#pragma unroll
            for (int d = 0; d < HD; d++) o[d] = 0.f;
            float sum = 0.f;
            const uint32_t rowkey = (uint32_t)((win * nH + h) * WT + ic);
            for (int j = 0; j < WT; j++) {
                const float p = expf(sc[j * 32 + lane] - mx);
                sum += p;
                if (ad.thr == 0 || attn_drop_keep(rowkey, j, ds0, ds1, ad.thr)) axpy_row(o, p, sV + j * HD);
            }
            if (act) store_row<T>(O + (win * WT + i) * (int64_t)C + h * HD, o, ad.inv_keep / sum);
            if (act && lse != nullptr) lse[(win * WT + i) * nH + h] = (mx + logf(sum)) * 1.4426950408889634f;   // log2 domain
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(AW * 32, 1) winattn_bwd_kernel(const T* __restrict__ qkv, const float* __restrict__ bias,
                                                                const T* __restrict__ O, const T* __restrict__ dO,
                                                                T* __restrict__ dqkv, float* __restrict__ dbias_partial,
                                                                int64_t n_windows, int nH, WinGeo g, AttnDrop ad) {
    extern __shared__ __align__(16) float smem[];
    float* sbias = smem;  // [2401] (+3 pad)
    const uint32_t ds0 = ad.thr ? ad.seed[0] : 0u, ds1 = ad.thr ? ad.seed[1] : 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PER_WARP = 4 * WT * HD + WT * 32 + 3 * 64 + 2404;
    float* sQ = smem + 2404 + warp * PER_WARP;  // scaled q
    float* sK = sQ + WT * HD;
    float* sV = sK + WT * HD;
    float* sD = sV + WT * HD;       // dO
    float* sc = sD + WT * HD;       // [49][32]
    float* sM = sc + WT * 32;       // row max   [64]
    float* sL = sM + 64;            // 1/rowsum  [64]
    float* sDl = sL + 64;           // delta     [64]
    float* sdb = sDl + 64;          // per-warp d(bias) accumulator [2401]
    const int h = blockIdx.y;
    const int C = nH * HD;
    for (int i = threadIdx.x; i < WT * WT; i += blockDim.x) sbias[i] = bias[h * WT * WT + i];
    for (int i = lane; i < WT * WT; i += 32) sdb[i] = 0.f;
    __syncthreads();
    const int nwin_img = g.nwin();
    for (int64_t win = (int64_t)blockIdx.x * AW + warp; win < n_windows; win += (int64_t)gridDim.x * AW) {
        const int64_t row0 = win * WT;
        const T* base = qkv + row0 * 3 * (int64_t)C + h * HD;
        const T* dob = dO + row0 * (int64_t)C + h * HD;
        const T* ob = O + row0 * (int64_t)C + h * HD;
        T* dbase = dqkv + row0 * 3 * (int64_t)C + h * HD;
        load_tile<T>(sQ, base, 3 * C, lane, QK_SCALE);
        load_tile<T>(sK, base + C, 3 * C, lane, 1.f);
        load_tile<T>(sV, base + 2 * C, 3 * C, lane, 1.f);
        load_tile<T>(sD, dob, C, lane, 1.f);
        __syncwarp();
        const MaskInfo mi = mask_info(g, (int)(win % nwin_img));
        // ---- phase 1: lanes own query rows: softmax stats, delta, dQ
        for (int rnd = 0; rnd < 2; rnd++) {
            const int i = lane + 32 * rnd;
            const bool act = i < WT;
            const int ic = act ? i : WT - 1;
            float q[HD], dq[HD];
            float dot_o = 0.f;
            {
                float orow[HD], drow[HD];
                load_row<T>(orow, ob + (int64_t)ic * C, 1.f);
                load_row<T>(drow, dob + (int64_t)ic * C, 1.f);
#pragma unroll
                for (int d = 0; d < HD; d++) dot_o = fmaf(orow[d], drow[d], dot_o);  // delta_i = dO_i . O_i
            }
#pragma unroll
            for (int d = 0; d < HD; d++) { q[d] = sQ[ic * HD + d]; dq[d] = 0.f; }
            const int ri = region_of(mi, ic);
            float mx = -INFINITY;
            for (int j = 0; j < WT; j++) {
                float s = dot_row(q, sK + j * HD) + sbias[ic * WT + j];
                if (mi.any && region_of(mi, j) != ri) s += -100.0f;
                sc[j * 32 + lane] = s;
                mx = fmaxf(mx, s);
            }
            float sum = 0.f;
            for (int j = 0; j < WT; j++) {
                const float p = expf(sc[j * 32 + lane] - mx);
                sc[j * 32 + lane] = p;
                sum += p;
            }
            const float inv = 1.0f / sum;
            // dO_i in registers (reuse q[] for it after S is done)
#pragma unroll
            for (int d = 0; d < HD; d++) q[d] = sD[ic * HD + d];
            const uint32_t rowkey = (uint32_t)((win * nH + h) * WT + ic);
            for (int j = 0; j < WT; j++) {
                const float p = sc[j * 32 + lane] * inv;
                // dP = m * dP~ with the forward's dropout mask m in {0, 1/keep}; delta_i = dO_i . O_i still equals sum_j P dP
                const float m = (ad.thr == 0 || attn_drop_keep(rowkey, j, ds0, ds1, ad.thr)) ? ad.inv_keep : 0.f;
                const float dp = dot_row(q, sV + j * HD) * m;
                axpy_row(dq, p * (dp - dot_o), sK + j * HD);
            }
            if (act) {
                store_row<T>(dbase + (int64_t)i * 3 * C, dq, QK_SCALE);
                sM[i] = mx; sL[i] = inv; sDl[i] = dot_o;
            }
        }
        __syncwarp();
        // ---- phase 2: lanes own key rows: dK, dV, d(bias)
        for (int rnd = 0; rnd < 2; rnd++) {
            const int j = lane + 32 * rnd;
            const bool act = j < WT;
            const int jc = act ? j : WT - 1;
            float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
            for (int d = 0; d < HD; d++) { k[d] = sK[jc * HD + d]; v[d] = sV[jc * HD + d]; dk[d] = 0.f; dv[d] = 0.f; }
            const int rj = region_of(mi, jc);
            for (int i = 0; i < WT; i++) {
                float s = dot_row(k, sQ + i * HD) + sbias[i * WT + jc];
                if (mi.any && region_of(mi, i) != rj) s += -100.0f;
                const float p = expf(s - sM[i]) * sL[i];
                const float m = (ad.thr == 0 || attn_drop_keep((uint32_t)((win * nH + h) * WT + i), jc, ds0, ds1, ad.thr)) ? ad.inv_keep : 0.f;
                const float dp = dot_row(v, sD + i * HD) * m;
                const float ds = p * (dp - sDl[i]);
                axpy_row(dv, p * m, sD + i * HD);
                axpy_row(dk, ds, sQ + i * HD);
                if (act) sdb[i * WT + j] += ds;
            }
            if (act) {
                store_row<T>(dbase + (int64_t)j * 3 * C + C, dk, 1.f);
                store_row<T>(dbase + (int64_t)j * 3 * C + 2 * C, dv, 1.f);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // fixed-order reduction over the CTA's warps -> partial[blockIdx.x][h][2401]
    float* out = dbias_partial + ((int64_t)blockIdx.x * nH + h) * (WT * WT);
    for (int i = threadIdx.x; i < WT * WT; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < AW; w++) s += smem[2404 + w * PER_WARP + (PER_WARP - 2404) + i];
        out[i] = s;
    }
}

// bias[h,i,j] = table[rel_index(i,j), h]; rel_index as TV:models/swin_transformer.py:272-284
__global__ void relbias_expand_kernel(const float* __restrict__ table, float* __restrict__ bias, int nH) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nH * WT * WT) return;
    const int h = idx / (WT * WT), r = idx % (WT * WT), i = r / WT, j = r % WT;
    const int dy = i / WS - j / WS + WS - 1, dx = i % WS - j % WS + WS - 1;
    bias[idx] = table[(dy * (2 * WS - 1) + dx) * nH + h];
}
// part[0][h][e] = sum_c part[c][h][e]  (in place, fixed order; each thread touches only its own column)
__global__ void relbias_sum_kernel(float* __restrict__ part, int grid, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float s = 0.f;
    for (int c = 0; c < grid; c++) s += part[(int64_t)c * n + idx];
    part[idx] = s;
}
// dtable[t,h] = sum over (i,j) with rel_index(i,j)=t of the CTA-summed d(bias)[h,i,j]; fixed order.
__global__ void relbias_reduce_kernel(const float* __restrict__ part, int grid, int nH, float* __restrict__ dtable,
                                      int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int NT_ = (2 * WS - 1) * (2 * WS - 1);
    if (idx >= NT_ * nH) return;
    const int t = idx / nH, h = idx % nH;
    const int dy = t / (2 * WS - 1) - (WS - 1), dx = t % (2 * WS - 1) - (WS - 1);
    float s = 0.f;
    for (int yi = 0; yi < WS; yi++) {
        const int yj = yi - dy;
        if (yj < 0 || yj >= WS) continue;
        for (int xi = 0; xi < WS; xi++) {
            const int xj = xi - dx;
            if (xj < 0 || xj >= WS) continue;
            const int e = (yi * WS + xi) * WT + (yj * WS + xj);
            s += part[(int64_t)h * (WT * WT) + e];   // slot 0 holds the sum over CTAs (relbias_sum_kernel)
        }
    }
    dtable[idx] = accumulate ? dtable[idx] + s : s;
}

constexpr int FWD_SMEM = (2404 + AW * (2 * WT * HD + WT * 32)) * 4;
constexpr int BWD_SMEM = (2404 + AW * (4 * WT * HD + WT * 32 + 3 * 64 + 2404)) * 4;

static int attn_grid(int64_t n_windows, int nH) {
    // enough CTAs for ~2 waves, but few enough that each warp amortises its bias/d(bias) tiles
    int64_t gx = (n_windows + AW - 1) / AW;
    const int64_t target = imax(1, (int64_t)num_sms() * 2 / nH);
    return (int)imax(1, imin(gx, target));
}

}  // namespace msu

namespace msu {
int winattn_fwd_tc(const void* qkv, const float* bias, void* O, int64_t n_windows, int nH, const WinGeo& g, const AttnDrop& ad,
                   float* lse, cudaStream_t st);
int winattn_bwd_tc_grid(int64_t n_windows, int nH);
int winattn_bwd_tc(const void* qkv, const float* bias, const void* dO, void* dqkv, float* dbias_partial, float* dtable, int64_t n_windows,
                   int nH, const WinGeo& g, const AttnDrop& ad, const float* lse, cudaStream_t st);
static int g_attn_backend = 0;  // 0 auto (tcgen05 for bf16), 1 force the SIMT kernels
}

using namespace msu;

extern "C" int msu_set_attn_backend(int backend) { g_attn_backend = backend; return 0; }

extern "C" int msu_winattn_fwd(int dtype, const void* qkv, const float* bias, void* O, int64_t n_windows, int32_t nH,
                               const int32_t* geo, float p_drop, const uint32_t* seed, float* lse, void* stream) {
    MSU_REQUIRE(qkv && bias && O && geo, "msu_winattn_fwd: null pointer");
    MSU_REQUIRE(nH > 0 && nH <= 65535, "msu_winattn_fwd: bad head count %d", nH);
    MSU_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "msu_winattn_fwd: bad dropout probability %f", (double)p_drop);
    if (n_windows == 0) return 0;
    WinGeo g = make_wingeo(geo);
    const AttnDrop ad = make_attn_drop(p_drop, seed);
    dim3 grid(attn_grid(n_windows, nH), nH);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSU_BF16 && g_attn_backend == 0) {
        const int rc = winattn_fwd_tc(qkv, bias, O, n_windows, nH, g, ad, lse, st);
        if (rc != 1) return rc;
    }
    if (dtype == MSU_F32) {
        static PerDeviceOnce attr;
        if (attr.need()) { cudaFuncSetAttribute(winattn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM); attr.set(); }
        winattn_fwd_kernel<float><<<grid, AW * 32, FWD_SMEM, st>>>((const float*)qkv, bias, (float*)O, n_windows, nH, g, ad, lse);
    } else if (dtype == MSU_BF16) {
        static PerDeviceOnce attr;
        if (attr.need()) { cudaFuncSetAttribute(winattn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM); attr.set(); }
        winattn_fwd_kernel<__nv_bfloat16><<<grid, AW * 32, FWD_SMEM, st>>>((const __nv_bfloat16*)qkv, bias, (__nv_bfloat16*)O, n_windows, nH, g, ad, lse);
    } else {
        MSU_REQUIRE(false, "msu_winattn_fwd: unsupported dtype %d", dtype);
    }
    count_launch();
    return check_launch("msu_winattn_fwd");
}

extern "C" int msu_winattn_bwd_grid(int dtype, int64_t n_windows, int32_t nH) {
    if (dtype == MSU_BF16 && g_attn_backend == 0) return 2 * winattn_bwd_tc_grid(n_windows, nH);
    return attn_grid(n_windows, nH);
}

// O (the forward output) supplies delta_i = dO_i . O_i without a second P.V product.
extern "C" int msu_winattn_bwd_direct(int dtype) {
    return (dtype == MSU_BF16 && g_attn_backend == 0 && !deterministic_mode()) ? 1 : 0;
}

extern "C" int msu_winattn_bwd(int dtype, const void* qkv, const float* bias, const void* O, const void* dO, void* dqkv,
                               float* dbias_partial, float* dtable, int64_t n_windows, int32_t nH, const int32_t* geo, float p_drop,
                               const uint32_t* seed, const float* lse, void* stream) {
    MSU_REQUIRE(qkv && bias && O && dO && dqkv && geo, "msu_winattn_bwd: null pointer");
    const bool direct = dtable != nullptr;
    MSU_REQUIRE(!direct || msu_winattn_bwd_direct(dtype), "msu_winattn_bwd: dtable given, but this dtype / backend / deterministic mode "
                "writes partials (ask msu_winattn_bwd_direct first)");
    MSU_REQUIRE(direct || dbias_partial, "msu_winattn_bwd: neither dbias_partial nor dtable");
    MSU_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "msu_winattn_bwd: bad dropout probability %f", (double)p_drop);
    WinGeo g = make_wingeo(geo);
    const AttnDrop ad = make_attn_drop(p_drop, seed);
    const int gx = attn_grid(n_windows, nH);
    dim3 grid(gx, nH);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSU_BF16 && g_attn_backend == 0) {
        const int rc = winattn_bwd_tc(qkv, bias, dO, dqkv, dbias_partial, dtable, n_windows, nH, g, ad, lse, st);
        if (rc != 1) return rc;
        MSU_REQUIRE(false, "msu_winattn_bwd: tcgen05 path unavailable for these pointers (workspace was sized for it)");
    }
    if (dtype == MSU_F32) {
        static PerDeviceOnce attr;
        if (attr.need()) { cudaFuncSetAttribute(winattn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM); attr.set(); }
        winattn_bwd_kernel<float><<<grid, AW * 32, BWD_SMEM, st>>>((const float*)qkv, bias, (const float*)O, (const float*)dO,
                                                                  (float*)dqkv, dbias_partial, n_windows, nH, g, ad);
    } else if (dtype == MSU_BF16) {
        static PerDeviceOnce attr;
        if (attr.need()) { cudaFuncSetAttribute(winattn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM); attr.set(); }
        winattn_bwd_kernel<__nv_bfloat16><<<grid, AW * 32, BWD_SMEM, st>>>((const __nv_bfloat16*)qkv, bias, (const __nv_bfloat16*)O,
                                                                          (const __nv_bfloat16*)dO, (__nv_bfloat16*)dqkv,
                                                                          dbias_partial, n_windows, nH, g, ad);
    } else {
        MSU_REQUIRE(false, "msu_winattn_bwd: unsupported dtype %d", dtype);
    }
    count_launch();
    return check_launch("msu_winattn_bwd");
}

extern "C" int msu_relbias_expand(const float* table, float* bias, int32_t nH, void* stream) {
    MSU_REQUIRE(table && bias && nH > 0, "msu_relbias_expand: bad arguments");
    const int n = nH * WT * WT;
    relbias_expand_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(table, bias, nH);
    count_launch();
    return check_launch("msu_relbias_expand");
}

extern "C" int msu_relbias_reduce(const float* dbias_partial, int32_t grid, int32_t nH, float* dtable, int accumulate,
                                  void* stream) {
    MSU_REQUIRE(dbias_partial && dtable && grid > 0 && nH > 0, "msu_relbias_reduce: bad arguments");
    const int n = 169 * nH;
    const int ne = nH * WT * WT;
    relbias_sum_kernel<<<(ne + 255) / 256, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(dbias_partial), grid, ne);
    relbias_reduce_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(dbias_partial, grid, nH, dtable, accumulate);
    count_launch(2);
    return check_launch("msu_relbias_reduce");
}
