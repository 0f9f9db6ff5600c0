// Device-side input staging: the tail of the reference's per-sample transform, dataset/dataset.py:13-16, 49-63 —
// horizontal flip, uint8 HWC image -> float32 / 255 in CHW, label -> (label > 127) as float — done for a whole batch on the GPU so
// that the host hands over 4 bytes per pixel instead of 16.  HBM-bound: 4 B read + 16 B written per pixel.
#include "common.cuh"

namespace msu {

constexpr int SG_THREADS = 256;

// four consecutive pixels of one image row per thread (W % 4 == 0): 12 image bytes + 4 label bytes in, four float4 out.
// `tpr` (a power of two) threads share one image row, so the only division is one 32-bit row / H per row.
__global__ void __launch_bounds__(SG_THREADS) stage_u8_vec_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ lab,
                                                                 const uint8_t* __restrict__ flip, float* __restrict__ out,
                                                                 float* __restrict__ lab_out, int B, int H, int W, int tpr_log2) {
    const int W4 = W >> 2;
    const int rows = B * H, rpb = SG_THREADS >> tpr_log2;
    const int64_t plane = (int64_t)H * W;
    const int lane = threadIdx.x & ((1 << tpr_log2) - 1);
    for (int row = blockIdx.x * rpb + (threadIdx.x >> tpr_log2); row < rows; row += gridDim.x * rpb) {
        const int b = row / H;
        const bool f = flip != nullptr && flip[b] != 0;
        const int64_t o_row = (int64_t)b * 3 * plane + (int64_t)(row - b * H) * W;
        for (int xq = lane; xq < W4; xq += 1 << tpr_log2) {
            const int xs = f ? W - 4 - 4 * xq : 4 * xq;   // source group; its pixels are read in reverse when flipped
            const uint32_t* ip = reinterpret_cast<const uint32_t*>(img + ((int64_t)row * W + xs) * 3);
            const uint32_t w0 = __ldg(ip), w1 = __ldg(ip + 1), w2 = __ldg(ip + 2);
            uint32_t px[12];                              // byte j of the 12-byte group = channel j % 3 of pixel j / 3
#pragma unroll
            for (int k = 0; k < 4; k++) {
                px[k] = (w0 >> (8 * k)) & 0xFFu;
                px[4 + k] = (w1 >> (8 * k)) & 0xFFu;
                px[8 + k] = (w2 >> (8 * k)) & 0xFFu;
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) v[k] = __fdiv_rn((float)(f ? px[3 * (3 - k) + c] : px[3 * k + c]), 255.0f);
                __stcs(reinterpret_cast<float4*>(out + o_row + c * plane + 4 * xq), make_float4(v[0], v[1], v[2], v[3]));
            }
            if (lab != nullptr) {
                const uint32_t lw = __ldg(reinterpret_cast<const uint32_t*>(lab + (int64_t)row * W + xs));
                float l[4];
#pragma unroll
                for (int k = 0; k < 4; k++) l[k] = ((f ? lw >> (8 * (3 - k)) : lw >> (8 * k)) & 0xFFu) > 127u ? 1.0f : 0.0f;
                *reinterpret_cast<float4*>(lab_out + (int64_t)row * W + 4 * xq) = make_float4(l[0], l[1], l[2], l[3]);
            }
        }
    }
}

// any width / alignment: one block per image row, one pixel per thread
__global__ void __launch_bounds__(SG_THREADS) stage_u8_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ lab,
                                                             const uint8_t* __restrict__ flip, float* __restrict__ out,
                                                             float* __restrict__ lab_out, int B, int H, int W) {
    const int64_t plane = (int64_t)H * W;
    for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
        const int b = row / H;
        const bool f = flip != nullptr && flip[b] != 0;
        const int64_t o_row = (int64_t)b * 3 * plane + (int64_t)(row - b * H) * W;
        for (int x = threadIdx.x; x < W; x += SG_THREADS) {
            const int xs = f ? W - 1 - x : x;
            const uint8_t* ip = img + ((int64_t)row * W + xs) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) out[o_row + c * plane + x] = __fdiv_rn((float)ip[c], 255.0f);
            if (lab != nullptr) lab_out[(int64_t)row * W + x] = lab[(int64_t)row * W + xs] > 127 ? 1.0f : 0.0f;
        }
    }
}

}  // namespace msu

using namespace msu;

extern "C" int msu_stage_u8(const uint8_t* img_hwc, const uint8_t* label, const uint8_t* flip, float* image_out, float* label_out,
                            int32_t B, int32_t H, int32_t W, void* stream) {
    MSU_REQUIRE(img_hwc && image_out && B >= 0 && H > 0 && W > 0, "msu_stage_u8: bad arguments");
    MSU_REQUIRE((label == nullptr) == (label_out == nullptr), "msu_stage_u8: label and label_out go together");
    if (B == 0) return 0;
    const bool vec = (W % 4 == 0) && (((uintptr_t)img_hwc | (uintptr_t)label) % 4 == 0) &&
                     (((uintptr_t)image_out | (uintptr_t)label_out) % 16 == 0);
    MSU_REQUIRE((int64_t)B * H < (1ll << 31), "msu_stage_u8: B * H too large");
    int tpr_log2 = 5;                                     // threads per image row: next power of two >= W / 4, in [32, 256]
    while ((1 << tpr_log2) < W / 4 && tpr_log2 < 8) tpr_log2++;
    const int64_t want = vec ? (((int64_t)B * H << tpr_log2) + SG_THREADS - 1) / SG_THREADS : (int64_t)B * H;
    const int64_t cap = (int64_t)num_sms() * 8 * 4;
    const int grid = (int)(want < cap ? want : cap);
    if (vec)
        stage_u8_vec_kernel<<<grid, SG_THREADS, 0, (cudaStream_t)stream>>>(img_hwc, label, flip, image_out, label_out, B, H, W,
                                                                          tpr_log2);
    else
        stage_u8_kernel<<<grid, SG_THREADS, 0, (cudaStream_t)stream>>>(img_hwc, label, flip, image_out, label_out, B, H, W);
    count_launch();
    return check_launch("msu_stage_u8");
}
