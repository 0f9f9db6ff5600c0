// Shared device/host helpers for libmsunet_sm100.so (sm_100a only).
#pragma once
#include <utility>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/msunet_b200.h"

namespace msu {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int deterministic_mode();        // msu_set_deterministic / MSU_DETERMINISTIC: 1 = fixed-order reductions only (no atomics)
int check_launch(const char* what);  // cudaGetLastError -> code, records message

#define MSU_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            msu::set_error(__VA_ARGS__); \
            return -1;                  \
        }                               \
    } while (0)

__host__ __device__ inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }

constexpr int WS = 7;
constexpr int WT = 49;  // tokens per window
constexpr int HD = 32;  // head dim (C / num_heads is 32 for every config of the reference)

// ---- dtype helpers --------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float ld_as_f(const void* p, int64_t idx, int dtype) {
    return dtype == MSU_F32 ? reinterpret_cast<const float*>(p)[idx]
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
}
__device__ __forceinline__ void st_from_f(void* p, int64_t idx, int dtype, float v) {
    if (dtype == MSU_F32) reinterpret_cast<float*>(p)[idx] = v;
    else reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
}

// 4-element vector load/store converting to/from fp32 (16 B for float, 8 B for bf16)
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
    static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
        uint2 r = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
        __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<uint32_t*>(&a);
        r.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = r;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact (erf) GELU and its derivative — nn.GELU() default, TV:ops/misc.py:264-305
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// ---- window geometry (index math that replaces pad/roll/partition/reverse/crop) -------------
// TV:models/swin_transformer.py:152-172 (forward gather) and :219-227 (reverse scatter).
struct WinGeo {
    int H, W, Ph, Pw, sh, sw;
    __host__ __device__ int nwx() const { return Pw / WS; }
    __host__ __device__ int nwin() const { return (Ph / WS) * (Pw / WS); }
};
__host__ __device__ inline WinGeo make_wingeo(const int32_t* g) { return WinGeo{g[0], g[1], g[2], g[3], g[4], g[5]}; }

// window-order row (b, w, i) -> source pixel row (b, y, x) or -1 for a padding token
__device__ __forceinline__ int64_t win_to_pix(const WinGeo& g, int64_t wr) {
    const int per_img = g.nwin() * WT;
    const int64_t b = wr / per_img;
    const int r = (int)(wr - b * per_img);
    const int w = r / WT, i = r - w * WT;
    const int ry = (w / g.nwx()) * WS + i / WS, rx = (w % g.nwx()) * WS + i % WS;
    int py = ry + g.sh; if (py >= g.Ph) py -= g.Ph;
    int px = rx + g.sw; if (px >= g.Pw) px -= g.Pw;
    if (py >= g.H || px >= g.W) return -1;
    return b * (int64_t)(g.H * g.W) + py * g.W + px;
}
// source pixel row -> window-order row (every real pixel appears exactly once)
__device__ __forceinline__ int64_t pix_to_win(const WinGeo& g, int64_t pr) {
    const int hw = g.H * g.W;
    const int64_t b = pr / hw;
    const int r = (int)(pr - b * hw);
    const int y = r / g.W, x = r - y * g.W;
    int ry = y - g.sh; if (ry < 0) ry += g.Ph;
    int rx = x - g.sw; if (rx < 0) rx += g.Pw;
    const int w = (ry / WS) * g.nwx() + rx / WS;
    const int i = (ry % WS) * WS + rx % WS;
    return b * (int64_t)(g.nwin() * WT) + w * WT + i;
}

// generic row/col mapping shared by GEMM operands and outputs.  Returns memory row (or -1) and col.
struct RowCol { int64_t row; int col; };
__device__ __forceinline__ RowCol map_rc(int map, const int32_t* geo, int64_t r, int c) {
    switch (map) {
        case MSU_MAP_WINDOW: {
            WinGeo g = make_wingeo(geo);
            return RowCol{win_to_pix(g, r), c};
        }
        case MSU_MAP_SHUFFLE: {  // geo = {H, W, p, cc}
            const int H = geo[0], W = geo[1], p = geo[2], cc = geo[3];
            const int q = c / cc, p1 = q / p, p2 = q - p1 * p;
            const int hw = H * W;
            const int64_t b = r / hw;
            const int t = (int)(r - b * hw);
            const int h = t / W, w = t - h * W;
            return RowCol{(b * (H * p) + (h * p + p1)) * (int64_t)(W * p) + (w * p + p2), c - q * cc};
        }
        case MSU_MAP_UNSHUFFLE: {  // geo = {H, W, p, cc}: r walks the shuffled map [(b, h*p+p1, w*p+p2)]
            const int H = geo[0], W = geo[1], p = geo[2], cc = geo[3];
            const int Wp = W * p, HWp = H * p * Wp;
            const int64_t b = r / HWp;
            const int t = (int)(r - b * HWp);
            const int y = t / Wp, x = t - y * Wp;
            const int h = y / p, p1 = y - h * p, w = x / p, p2 = x - w * p;
            return RowCol{(b * H + h) * (int64_t)W + w, (p1 * p + p2) * cc + c};
        }
        case MSU_MAP_CONV3: {  // geo = {H, W, C}
            const int H = geo[0], W = geo[1], C = geo[2];
            const int tap = c / C, ci = c - tap * C;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            const int hw = H * W;
            const int64_t b = r / hw;
            const int t = (int)(r - b * hw);
            const int y = t / W + dy, x = t % W + dx;
            if (y < 0 || y >= H || x < 0 || x >= W) return RowCol{-1, ci};
            return RowCol{b * hw + y * W + x, ci};
        }
        case MSU_MAP_MERGE: {  // geo = {H, W, C}: r indexes (b, h/2, w/2)
            const int H = geo[0], W = geo[1], C = geo[2];
            const int q = c / C, ci = c - q * C;
            const int h2w2 = (H / 2) * (W / 2);
            const int64_t b = r / h2w2;
            const int t = (int)(r - b * h2w2);
            const int y = 2 * (t / (W / 2)) + (q & 1), x = 2 * (t % (W / 2)) + (q >> 1);
            return RowCol{b * (int64_t)(H * W) + y * W + x, ci};
        }
        default:
            return RowCol{r, c};
    }
}

// ---- attention dropout (TV:models/swin_transformer.py:205 F.dropout on the softmax output) -----------------
// Counter-based mask shared by forward and backward: one lowbias32 hash per (window, head, query row, key pair),
// its two 16-bit halves decide the two keys; keep iff half >= thr, thr = round(p * 65536).  oracle/msunet_oracle.py
// restates it (attn_drop_keep) so the parity tests apply the identical mask to the PyTorch reference.
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t attn_drop_hash(uint32_t rowkey, int jpair, uint32_t s0, uint32_t s1) {
    return lowbias32((((rowkey << 5) + (uint32_t)jpair) + s0) * 0x9E3779B1u ^ s1);
}
__host__ __device__ __forceinline__ bool attn_drop_keep(uint32_t rowkey, int j, uint32_t s0, uint32_t s1, uint32_t thr) {
    const uint32_t h = attn_drop_hash(rowkey, j >> 1, s0, s1);
    return ((j & 1) ? (h >> 16) : (h & 0xffffu)) >= thr;
}
struct AttnDrop {           // thr == 0: no dropout
    uint32_t thr;
    float inv_keep;
    const uint32_t* seed;   // device pointer to two 32-bit words (written by the caller's RNG, CUDA-graph friendly)
};
inline AttnDrop make_attn_drop(float p, const uint32_t* seed) {
    AttnDrop d{0u, 1.0f, nullptr};
    if (p > 0.f && seed != nullptr) {
        d.thr = (uint32_t)(p * 65536.0f + 0.5f);
        d.inv_keep = 1.0f / (1.0f - p);
        d.seed = seed;
    }
    return d;
}

// Function attributes (max dynamic shared memory) live per device: one flag per (call site, device) so that several devices
// driven from one process (nn.DataParallel threads, trainer.py:96-97) each get theirs.
struct PerDeviceOnce {
    bool done[64] = {};
    bool need() {
        int dev = 0;
        cudaGetDevice(&dev);
        slot = dev & 63;
        return !done[slot];
    }
    void set() { done[slot] = true; }
    static thread_local int slot;
};
inline thread_local int PerDeviceOnce::slot = 0;

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while the previous kernel of its stream is
// still running (as soon as every CTA of that kernel has called pdl_trigger() or exited, and an SM has room); it runs its prologue
// (barrier init, TMEM allocation) and then blocks in pdl_wait() until the previous kernel has completed and its writes are visible.
// A kernel that is NOT launched this way is unaffected by its predecessor's trigger.
// OFF by default: measured on the training step (three A/B/C runs per box) the step was 0.2-0.3 ms SLOWER with the GEMM kernel
// launched this way (22.3-22.5 -> 22.6-22.7 ms) and slower again with the weight-gradient / attention kernels included — the early
// resident 200 KB CTAs take SMs the side-stream kernels would otherwise fill — and the pinned-host e2e loop showed outliers.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline int pdl_level() {          // MSU_PDL: 0 off, 1 the forward / dgrad GEMM kernel only, 2 also the weight-gradient and attention kernels
    static const int lv = getenv("MSU_PDL") ? atoi(getenv("MSU_PDL")) : 0;
    return lv;
}
template <int LEVEL = 1, typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
        n++;
    }
    if (pdl_level() >= LEVEL) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    cfg.attrs = at;
    cfg.numAttrs = (unsigned)n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace msu
