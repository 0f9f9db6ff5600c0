// Fused DynamicLoss (BCE-with-logits + Tversky, loss/DynamicLoss.py:73-111) forward/backward and the
// Dice/IoU counting of scripts/validation_functions.py:106-108, 214-309 — single-pass, memory-bound,
// warp-shuffle reductions, no host synchronisation (the reference does 1+3B syncs per loss call).
#include "common.cuh"

namespace msu {

constexpr int LM_BLOCKS = 64;   // blocks per sample
constexpr int LM_THREADS = 256;
constexpr int LS = 16;          // floats per sample in the partial/stat records
// record layout: 0 a=sum[max(x,0)+log1p(exp(-|x|))], 1 sum p, 2 max t, 3 -, then per interpretation k in {raw, >127.5}:
// 4+6k: sum x t, 5+6k: TP, 6+6k: FP, 7+6k: FN, 8+6k: sum t, 9+6k: -
constexpr float TV_SMOOTH = 1e-6f;

template <int NV>
__device__ __forceinline__ void block_reduce(float* v, float* smem, bool is_max_at2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = (is_max_at2 && k == 2) ? warp_max(v[k]) : warp_sum(v[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) smem[warp * NV + k] = v[k];
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            float x = lane < nw ? smem[lane * NV + k] : ((is_max_at2 && k == 2) ? -INFINITY : 0.f);
            v[k] = (is_max_at2 && k == 2) ? warp_max(x) : warp_sum(x);
        }
    }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T>
__global__ void __launch_bounds__(LM_THREADS) loss_partial_kernel(const T* __restrict__ logits, const float* __restrict__ target,
                                                                 int64_t N, float* __restrict__ ws) {
    __shared__ float sm[(LM_THREADS / 32) * LS];
    const int b = blockIdx.y;
    const T* x = logits + (int64_t)b * N;
    const float* t = target + (int64_t)b * N;
    float v[LS];
#pragma unroll
    for (int k = 0; k < LS; k++) v[k] = 0.f;
    v[2] = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * LM_THREADS + threadIdx.x; i < N; i += (int64_t)gridDim.x * LM_THREADS) {
        const float xv = to_f<T>(x[i]);
        const float tv = t[i];
        const float tb = tv > 127.5f ? 1.f : 0.f;
        const float p = sigmoid_f(xv);
        v[0] += fmaxf(xv, 0.f) + log1pf(expf(-fabsf(xv)));
        v[1] += p;
        v[2] = fmaxf(v[2], tv);
        v[4] += xv * tv;  v[5] += p * tv;  v[6] += p * (1.f - tv);  v[7] += (1.f - p) * tv;  v[8] += tv;
        v[10] += xv * tb; v[11] += p * tb; v[12] += p * (1.f - tb); v[13] += (1.f - p) * tb; v[14] += tb;
    }
    block_reduce<LS>(v, sm, true);
    if (threadIdx.x == 0) {
        float* o = ws + ((int64_t)b * gridDim.x + blockIdx.x) * LS;
#pragma unroll
        for (int k = 0; k < LS; k++) o[k] = v[k];
    }
}

// Vector path (N % 8 == 0, 16 B aligned rows): 8 pixels per thread per step (one 16 B load of bf16/fp16 logits or two of
// fp32, two 16 B loads of the target), one exp and one log per pixel (sigmoid and softplus share exp(-|x|)), and only the
// sums that cannot be derived: FP = sum p - TP and FN = sum t - TP are formed once per block.
template <typename T> struct Ld8;
template <> struct Ld8<float> {
    static __device__ __forceinline__ void ld(const float* p, float* o) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    }
    static __device__ __forceinline__ void st(float* p, const float* o) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
};
template <> struct Ld8<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
        const uint4 r = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* o) {
        uint4 r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};
template <> struct Ld8<__half> {
    static __device__ __forceinline__ void ld(const __half* p, float* o) {
        const uint4 r = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) { const float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void st(__half* p, const float* o) {
        uint4 r;
        __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};
// p = sigmoid(x) and softplus(x) = max(x, 0) + log(1 + exp(-|x|)) from one exponential
__device__ __forceinline__ void sigmoid_softplus(float x, float& p, float& sp) {
    const float e = __expf(-fabsf(x));
    const float r = __fdividef(1.0f, 1.0f + e);
    p = x >= 0.f ? r : e * r;
    sp = fmaxf(x, 0.f) + __logf(1.0f + e);
}

template <typename T>
__global__ void __launch_bounds__(LM_THREADS) loss_partial_vec_kernel(const T* __restrict__ logits, const float* __restrict__ target,
                                                                     int64_t N, float* __restrict__ ws) {
    __shared__ float sm[(LM_THREADS / 32) * LS];
    const int b = blockIdx.y;
    const T* x = logits + (int64_t)b * N;
    const float* t = target + (int64_t)b * N;
    float a = 0.f, sp_ = 0.f, tmax = -INFINITY, xt = 0.f, tp = 0.f, stt = 0.f, xtb = 0.f, tpb = 0.f, stb = 0.f;
    for (int64_t i = ((int64_t)blockIdx.x * LM_THREADS + threadIdx.x) * 8; i < N; i += (int64_t)gridDim.x * LM_THREADS * 8) {
        float xv[8], tv[8];
        Ld8<T>::ld(x + i, xv);
        Ld8<float>::ld(t + i, tv);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            float p, sp;
            sigmoid_softplus(xv[e], p, sp);
            const float tb = tv[e] > 127.5f ? 1.f : 0.f;
            a += sp; sp_ += p; tmax = fmaxf(tmax, tv[e]);
            xt = fmaf(xv[e], tv[e], xt); tp = fmaf(p, tv[e], tp); stt += tv[e];
            xtb = fmaf(xv[e], tb, xtb); tpb = fmaf(p, tb, tpb); stb += tb;
        }
    }
    float v[LS];
#pragma unroll
    for (int k = 0; k < LS; k++) v[k] = 0.f;
    v[0] = a; v[1] = sp_; v[2] = tmax;
    v[4] = xt; v[5] = tp; v[6] = sp_ - tp; v[7] = stt - tp; v[8] = stt;
    v[10] = xtb; v[11] = tpb; v[12] = sp_ - tpb; v[13] = stb - tpb; v[14] = stb;
    block_reduce<LS>(v, sm, true);
    if (threadIdx.x == 0) {
        float* o = ws + ((int64_t)b * gridDim.x + blockIdx.x) * LS;
#pragma unroll
        for (int k = 0; k < LS; k++) o[k] = v[k];
    }
}

template <typename T>
__global__ void __launch_bounds__(LM_THREADS) loss_bwd_vec_kernel(const T* __restrict__ logits, const float* __restrict__ target,
                                                                 int64_t N, float alpha, float beta, const float* __restrict__ stats,
                                                                 const int32_t* __restrict__ flag, const float* __restrict__ gscale,
                                                                 T* __restrict__ dlogits) {
    const int b = blockIdx.y;
    const float* s = stats + (int64_t)b * 8;
    const float g = gscale ? *gscale : 1.f;
    const float wb = s[0] * g, mt = s[1] * g, Nn = s[2], D = s[3];
    const float invD2 = 1.0f / (D * D);
    const int k = *flag;
    // dtv = -p' (t D - Nn (t + alpha (1 - t) - beta t)) / D^2 = p' (c0 + c1 t):  c0 = Nn alpha / D^2, c1 = (Nn (1 - alpha - beta) - D) / D^2
    const float c0 = mt * Nn * alpha * invD2, c1 = mt * (Nn * (1.f - alpha - beta) - D) * invD2;
    const T* x = logits + (int64_t)b * N;
    const float* t = target + (int64_t)b * N;
    T* dx = dlogits + (int64_t)b * N;
    for (int64_t i = ((int64_t)blockIdx.x * LM_THREADS + threadIdx.x) * 8; i < N; i += (int64_t)gridDim.x * LM_THREADS * 8) {
        float xv[8], tv[8], o[8];
        Ld8<T>::ld(x + i, xv);
        Ld8<float>::ld(t + i, tv);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float tt = k ? (tv[e] > 127.5f ? 1.f : 0.f) : tv[e];
            const float ex = __expf(-fabsf(xv[e]));
            const float r = __fdividef(1.0f, 1.0f + ex);
            const float p = xv[e] >= 0.f ? r : ex * r;
            const float dp = p * (1.f - p);
            o[e] = fmaf(wb, p - tt, dp * fmaf(c1, tt, c0));
        }
        Ld8<T>::st(dx + i, o);
    }
}

// one block: combines partials (fixed order, double), decides the {0,255} interpretation, writes per-sample
// backward coefficients stats[b] = {w_bce/(N*B), m_b/B, Nn, D, loss_b} and the batch-mean loss.
// per_sample != 0: every image decides its own {0,255} interpretation and is a batch of one (the reference validates one
// image per DynamicLoss call, validation_functions.py:89-104), flag / loss may be null.
__global__ void loss_final_kernel(const float* __restrict__ ws, int B, int nblk, int64_t N, float alpha, float beta,
                                  float mix, float* __restrict__ stats, int32_t* __restrict__ flag, float* __restrict__ loss,
                                  int per_sample) {
    __shared__ float smax[256];
    __shared__ double sloss[256];
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < B * nblk; i += blockDim.x) mx = fmaxf(mx, ws[(int64_t)i * LS + 2]);
    smax[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + s]);
        __syncthreads();
    }
    const int kg = smax[0] > 1.0f ? 1 : 0;  // loss/DynamicLoss.py:87-88
    if (threadIdx.x == 0 && flag) *flag = kg;
    // one warp per sample: lanes add the block records (double, fixed lane order), then a shuffle tree
    double acc = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int b = warp; b < B; b += nwarp) {
        double r6[6] = {0, 0, 0, 0, 0, 0};
        int k = kg;
        if (per_sample) {
            float m1 = -INFINITY;
            for (int j = lane; j < nblk; j += 32) m1 = fmaxf(m1, ws[((int64_t)b * nblk + j) * LS + 2]);
            for (int o = 16; o > 0; o >>= 1) m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
            k = m1 > 1.0f ? 1 : 0;
        }
        for (int j = lane; j < nblk; j += 32) {
            const float* r = ws + ((int64_t)b * nblk + j) * LS;
            r6[0] += r[0]; r6[1] += r[4 + 6 * k]; r6[2] += r[5 + 6 * k]; r6[3] += r[6 + 6 * k]; r6[4] += r[7 + 6 * k]; r6[5] += r[8 + 6 * k];
        }
#pragma unroll
        for (int q = 0; q < 6; q++)
            for (int o = 16; o > 0; o >>= 1) r6[q] += __shfl_xor_sync(0xffffffffu, r6[q], o);
        if (lane != 0) continue;
        const double a = r6[0], xt = r6[1], tp = r6[2], fp = r6[3], fn = r6[4], st = r6[5];
        const float bce = (float)((a - xt) / (double)N);
        const float Nn = (float)tp + TV_SMOOTH;
        const float D = (float)tp + alpha * (float)fp + beta * (float)fn + TV_SMOOTH;
        const float tv = 1.0f - Nn / D;
        const bool pos = st != 0.0;
        const float m = pos ? mix : 0.f;
        const float lb = (1.f - m) * bce + m * tv;
        float* s = stats + (int64_t)b * 8;
        s[0] = (1.f - m) / ((float)N * (float)B);
        s[1] = m / (float)B;
        s[2] = Nn;
        s[3] = D;
        s[4] = lb;
        s[5] = bce;
        s[6] = tv;
        s[7] = pos ? 1.f : 0.f;
        acc += lb;
    }
    sloss[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sloss[threadIdx.x] += sloss[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && loss) *loss = (float)(sloss[0] / (double)B);
}

template <typename T>
__global__ void __launch_bounds__(LM_THREADS) loss_bwd_kernel(const T* __restrict__ logits, const float* __restrict__ target,
                                                             int64_t N, float alpha, float beta, const float* __restrict__ stats,
                                                             const int32_t* __restrict__ flag, const float* __restrict__ gscale,
                                                             T* __restrict__ dlogits) {
    const int b = blockIdx.y;
    const float* s = stats + (int64_t)b * 8;
    const float g = gscale ? *gscale : 1.f;
    const float wb = s[0] * g, mt = s[1] * g, Nn = s[2], D = s[3];
    const float invD2 = 1.0f / (D * D);
    const int k = *flag;
    const T* x = logits + (int64_t)b * N;
    const float* t = target + (int64_t)b * N;
    T* dx = dlogits + (int64_t)b * N;
    for (int64_t i = (int64_t)blockIdx.x * LM_THREADS + threadIdx.x; i < N; i += (int64_t)gridDim.x * LM_THREADS) {
        const float xv = to_f<T>(x[i]);
        float tv = t[i];
        if (k) tv = tv > 127.5f ? 1.f : 0.f;
        const float p = sigmoid_f(xv);
        const float dp = p * (1.f - p);
        const float ct = tv + alpha * (1.f - tv) - beta * tv;  // d(TP + a FP + b FN)/dp
        const float dtv = -dp * (tv * D - Nn * ct) * invD2;
        dx[i] = from_f<T>(wb * (p - tv) + mt * dtv);
    }
}

// ---------------------------------------------------------------------------------------------
// metrics: record = {tp, fp, fn, tn} (int64) + 8 doubles
template <typename T>
__global__ void __launch_bounds__(LM_THREADS) metrics_partial_kernel(int from_logits, const T* __restrict__ in,
                                                                    const void* __restrict__ label_or_gt,
                                                                    const uint8_t* __restrict__ pred_bin, int64_t N, float thr,
                                                                    long long* __restrict__ wc, double* __restrict__ wsft,
                                                                    T* __restrict__ pred_out) {
    __shared__ long long sc[(LM_THREADS / 32) * 4];
    __shared__ double sd[(LM_THREADS / 32) * 8];
    const int b = blockIdx.y;
    const T* x = in + (int64_t)b * N;
    long long c[4] = {0, 0, 0, 0};
    double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * LM_THREADS + threadIdx.x; i < N; i += (int64_t)gridDim.x * LM_THREADS) {
        float p;
        bool pb, g;
        if (from_logits == 2) {
            // probabilities kept in fp32 whatever the logits dtype (the reference thresholds an fp16 / fp32 sigmoid,
            // validation_functions.py:78,106-107: rounding it to bf16 would move the 0.5 boundary by 2^-9); pred_out is float
            p = sigmoid_f(to_f<T>(x[i]));
            pb = p > thr;
            g = reinterpret_cast<const float*>(label_or_gt)[(int64_t)b * N + i] > 0.f;
            if (pred_out) reinterpret_cast<float*>(pred_out)[(int64_t)b * N + i] = p;
        } else if (from_logits) {
            // sigmoid in fp32 rounded to the logits dtype, THEN compared
            p = to_f<T>(from_f<T>(sigmoid_f(to_f<T>(x[i]))));
            pb = p > thr;
            g = reinterpret_cast<const float*>(label_or_gt)[(int64_t)b * N + i] > 0.f;
            if (pred_out) pred_out[(int64_t)b * N + i] = from_f<T>(p);
        } else {
            p = to_f<T>(x[i]);
            pb = pred_bin[(int64_t)b * N + i] != 0;
            g = reinterpret_cast<const uint8_t*>(label_or_gt)[(int64_t)b * N + i] != 0;
        }
        c[0] += (pb && g); c[1] += (pb && !g); c[2] += (!pb && g); c[3] += (!pb && !g);
        const float gf = g ? 1.f : 0.f;
        d[0] += (double)(p * gf); d[1] += (double)((1.f - gf) * p); d[2] += (double)(gf * (1.f - p));
        d[3] += (double)((1.f - p) * (1.f - gf)); d[4] += (double)(p * p); d[5] += (double)gf; d[6] += (double)p; d[7] += (double)gf;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
#pragma unroll
    for (int k = 0; k < 8; k++)
        for (int o = 16; o > 0; o >>= 1) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
    if (lane == 0) {
        for (int k = 0; k < 4; k++) sc[warp * 4 + k] = c[k];
        for (int k = 0; k < 8; k++) sd[warp * 8 + k] = d[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        long long s = 0;
        for (int w = 0; w < LM_THREADS / 32; w++) s += sc[w * 4 + threadIdx.x];
        wc[((int64_t)b * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = s;
    } else if (threadIdx.x >= 32 && threadIdx.x < 40) {
        const int k = threadIdx.x - 32;
        double s = 0;
        for (int w = 0; w < LM_THREADS / 32; w++) s += sd[w * 8 + k];
        wsft[((int64_t)b * gridDim.x + blockIdx.x) * 8 + k] = s;
    }
}
__global__ void metrics_final_kernel(const long long* __restrict__ wc, const double* __restrict__ wsft, int B, int nblk,
                                     long long* __restrict__ counts, double* __restrict__ soft) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * 12) return;
    const int b = idx / 12, k = idx % 12;
    if (k < 4) {
        long long s = 0;
        for (int j = 0; j < nblk; j++) s += wc[((int64_t)b * nblk + j) * 4 + k];
        counts[b * 4 + k] = s;
    } else {
        double s = 0;
        for (int j = 0; j < nblk; j++) s += wsft[((int64_t)b * nblk + j) * 8 + (k - 4)];
        soft[b * 8 + (k - 4)] = s;
    }
}

static int lm_blocks(int64_t N) { return (int)imax(1, imin(LM_BLOCKS, (N + LM_THREADS * 4 - 1) / (LM_THREADS * 4))); }

}  // namespace msu

using namespace msu;

/* ws: fp32, at least B*64*16 floats */
static int loss_fwd_impl(int dtype, const void* logits, const float* target, int32_t B, int64_t N, float alpha,
                         float beta, float mix, float* ws, float* stats, int32_t* flag, float* loss, int per_sample, void* stream) {
    MSU_REQUIRE(logits && target && ws && stats && (per_sample || (flag && loss)), "msu_loss_fwd: null pointer");
    MSU_REQUIRE(B > 0 && N > 0 && B <= 65535, "msu_loss_fwd: bad shape B=%d N=%lld", B, (long long)N);
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = lm_blocks(N);
    dim3 grid(nblk, B);
    const bool vec = (N % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(target)) & 15) == 0;
    if (vec && dtype == MSU_F32) loss_partial_vec_kernel<float><<<grid, LM_THREADS, 0, st>>>((const float*)logits, target, N, ws);
    else if (vec && dtype == MSU_BF16) loss_partial_vec_kernel<__nv_bfloat16><<<grid, LM_THREADS, 0, st>>>((const __nv_bfloat16*)logits, target, N, ws);
    else if (vec && dtype == MSU_F16) loss_partial_vec_kernel<__half><<<grid, LM_THREADS, 0, st>>>((const __half*)logits, target, N, ws);
    else if (dtype == MSU_F32) loss_partial_kernel<float><<<grid, LM_THREADS, 0, st>>>((const float*)logits, target, N, ws);
    else if (dtype == MSU_BF16) loss_partial_kernel<__nv_bfloat16><<<grid, LM_THREADS, 0, st>>>((const __nv_bfloat16*)logits, target, N, ws);
    else if (dtype == MSU_F16) loss_partial_kernel<__half><<<grid, LM_THREADS, 0, st>>>((const __half*)logits, target, N, ws);
    else MSU_REQUIRE(false, "msu_loss_fwd: unsupported dtype %d", dtype);
    loss_final_kernel<<<1, 256, 0, st>>>(ws, B, nblk, N, alpha, beta, mix, stats, flag, loss, per_sample);
    count_launch(2);
    return check_launch("msu_loss_fwd");
}

extern "C" int msu_loss_fwd(int dtype, const void* logits, const float* target, int32_t B, int64_t N, float alpha,
                            float beta, float mix, float* ws, float* stats, int32_t* flag, float* loss, void* stream) {
    return loss_fwd_impl(dtype, logits, target, B, N, alpha, beta, mix, ws, stats, flag, loss, 0, stream);
}

extern "C" int msu_loss_per_sample(int dtype, const void* logits, const float* target, int32_t B, int64_t N, float alpha,
                                   float beta, float mix, float* ws, float* stats, void* stream) {
    return loss_fwd_impl(dtype, logits, target, B, N, alpha, beta, mix, ws, stats, nullptr, nullptr, 1, stream);
}

extern "C" int msu_loss_bwd(int dtype, const void* logits, const float* target, int32_t B, int64_t N, float alpha,
                            float beta, float mix, const float* stats, const int32_t* flag, const float* gscale,
                            void* dlogits, void* stream) {
    MSU_REQUIRE(logits && target && stats && flag && dlogits, "msu_loss_bwd: null pointer");
    (void)mix;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)imax(1, imin(4096, (N + LM_THREADS * 4 - 1) / (LM_THREADS * 4))), B);
    const bool vec = (N % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0;
    if (vec) {
        dim3 gv((unsigned)imax(1, imin(2048, (N + LM_THREADS * 8 - 1) / (LM_THREADS * 8))), B);
        if (dtype == MSU_F32) loss_bwd_vec_kernel<float><<<gv, LM_THREADS, 0, st>>>((const float*)logits, target, N, alpha, beta, stats, flag, gscale, (float*)dlogits);
        else if (dtype == MSU_BF16) loss_bwd_vec_kernel<__nv_bfloat16><<<gv, LM_THREADS, 0, st>>>((const __nv_bfloat16*)logits, target, N, alpha, beta, stats, flag, gscale, (__nv_bfloat16*)dlogits);
        else if (dtype == MSU_F16) loss_bwd_vec_kernel<__half><<<gv, LM_THREADS, 0, st>>>((const __half*)logits, target, N, alpha, beta, stats, flag, gscale, (__half*)dlogits);
        else MSU_REQUIRE(false, "msu_loss_bwd: unsupported dtype %d", dtype);
        count_launch();
        return check_launch("msu_loss_bwd");
    }
    if (dtype == MSU_F32) loss_bwd_kernel<float><<<grid, LM_THREADS, 0, st>>>((const float*)logits, target, N, alpha, beta, stats, flag, gscale, (float*)dlogits);
    else if (dtype == MSU_BF16) loss_bwd_kernel<__nv_bfloat16><<<grid, LM_THREADS, 0, st>>>((const __nv_bfloat16*)logits, target, N, alpha, beta, stats, flag, gscale, (__nv_bfloat16*)dlogits);
    else if (dtype == MSU_F16) loss_bwd_kernel<__half><<<grid, LM_THREADS, 0, st>>>((const __half*)logits, target, N, alpha, beta, stats, flag, gscale, (__half*)dlogits);
    else MSU_REQUIRE(false, "msu_loss_bwd: unsupported dtype %d", dtype);
    count_launch();
    return check_launch("msu_loss_bwd");
}

/* wc: int64 [B*64*4], wsft: double [B*64*8] workspaces */
extern "C" int msu_metrics(int dtype, int from_logits, const void* in, const void* label_or_gt, const uint8_t* pred_bin,
                           int32_t B, int64_t N, float thr, long long* wc, double* wsft, long long* counts, double* soft,
                           void* pred_out, void* stream) {
    MSU_REQUIRE(in && label_or_gt && wc && wsft && counts && soft, "msu_metrics: null pointer");
    MSU_REQUIRE(from_logits || pred_bin, "msu_metrics: pred_bin required when from_logits=0");
    MSU_REQUIRE(B > 0 && N > 0 && B <= 65535, "msu_metrics: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = lm_blocks(N);
    dim3 grid(nblk, B);
    if (dtype == MSU_F32) metrics_partial_kernel<float><<<grid, LM_THREADS, 0, st>>>(from_logits, (const float*)in, label_or_gt, pred_bin, N, thr, wc, wsft, (float*)pred_out);
    else if (dtype == MSU_BF16) metrics_partial_kernel<__nv_bfloat16><<<grid, LM_THREADS, 0, st>>>(from_logits, (const __nv_bfloat16*)in, label_or_gt, pred_bin, N, thr, wc, wsft, (__nv_bfloat16*)pred_out);
    else if (dtype == MSU_F16) metrics_partial_kernel<__half><<<grid, LM_THREADS, 0, st>>>(from_logits, (const __half*)in, label_or_gt, pred_bin, N, thr, wc, wsft, (__half*)pred_out);
    else MSU_REQUIRE(false, "msu_metrics: unsupported dtype %d", dtype);
    metrics_final_kernel<<<(B * 12 + 127) / 128, 128, 0, st>>>(wc, wsft, B, nblk, counts, soft);
    count_launch(2);
    return check_launch("msu_metrics");
}
