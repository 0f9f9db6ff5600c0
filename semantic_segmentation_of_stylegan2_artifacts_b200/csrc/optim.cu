// Fused multi-tensor AdamW step (the step right after the hot path, trainer.py:143-152, 315).
// One launch updates every parameter: block -> (tensor, 8192-element chunk) through a device table; per element the
// arithmetic follows torch.optim.AdamW (decoupled decay, lerp first moment, bias-corrected denominator) in fp32.
// HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"

namespace msu {

constexpr int AD_CHUNK = 8192;
constexpr int AD_THREADS = 256;

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const MsuAdamTensor& t) {
    p *= t.decay;                                   // param.mul_(1 - lr * weight_decay)
    m = m + (g - m) * (1.0f - t.beta1);             // exp_avg.lerp_(grad, 1 - beta1)
    v = v * t.beta2 + (1.0f - t.beta2) * g * g;     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) * t.inv_bias2_sqrt + t.eps;
    p -= t.step_size * (m / denom);                 // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(AD_THREADS) adamw_kernel(const MsuAdamTensor* __restrict__ tab, const int32_t* __restrict__ blk_tensor,
                                                          const int32_t* __restrict__ blk_chunk, const float* __restrict__ inv_scale,
                                                          const float* __restrict__ found_inf) {
    if (found_inf != nullptr && *found_inf != 0.f) return;     // GradScaler: skip the whole step on overflow
    const MsuAdamTensor t = tab[blk_tensor[blockIdx.x]];
    const int64_t base = (int64_t)blk_chunk[blockIdx.x] * AD_CHUNK;
    const int64_t end = base + AD_CHUNK < t.n ? base + AD_CHUNK : t.n;
    const float gs = inv_scale != nullptr ? *inv_scale : 1.0f;
    float* p = reinterpret_cast<float*>(t.p);
    const float* g = reinterpret_cast<const float*>(t.g);
    float* m = reinterpret_cast<float*>(t.m);
    float* v = reinterpret_cast<float*>(t.v);
    const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
    if (vec) {
        const int64_t nv = (end - base) / 4;
        for (int64_t i = threadIdx.x; i < nv; i += AD_THREADS) {
            const int64_t o = base + i * 4;
            float4 pp = *reinterpret_cast<float4*>(p + o), mm = *reinterpret_cast<float4*>(m + o), vv = *reinterpret_cast<float4*>(v + o);
            const float4 gg = *reinterpret_cast<const float4*>(g + o);
            adamw_elem(pp.x, gg.x * gs, mm.x, vv.x, t);
            adamw_elem(pp.y, gg.y * gs, mm.y, vv.y, t);
            adamw_elem(pp.z, gg.z * gs, mm.z, vv.z, t);
            adamw_elem(pp.w, gg.w * gs, mm.w, vv.w, t);
            *reinterpret_cast<float4*>(p + o) = pp;
            *reinterpret_cast<float4*>(m + o) = mm;
            *reinterpret_cast<float4*>(v + o) = vv;
        }
        for (int64_t o = base + nv * 4 + threadIdx.x; o < end; o += AD_THREADS) adamw_elem(p[o], g[o] * gs, m[o], v[o], t);
    } else {
        for (int64_t o = base + threadIdx.x; o < end; o += AD_THREADS) adamw_elem(p[o], g[o] * gs, m[o], v[o], t);
    }
}

}  // namespace msu

using namespace msu;

extern "C" int msu_adamw_chunk(void) { return AD_CHUNK; }

extern "C" int msu_adamw_step(const MsuAdamTensor* table, const int32_t* blk_tensor, const int32_t* blk_chunk, int32_t n_blocks,
                              const float* inv_scale, const float* found_inf, void* stream) {
    MSU_REQUIRE(table && blk_tensor && blk_chunk && n_blocks >= 0, "msu_adamw_step: bad arguments");
    if (n_blocks == 0) return 0;
    adamw_kernel<<<n_blocks, AD_THREADS, 0, (cudaStream_t)stream>>>(table, blk_tensor, blk_chunk, inv_scale, found_inf);
    count_launch();
    return check_launch("msu_adamw_step");
}
