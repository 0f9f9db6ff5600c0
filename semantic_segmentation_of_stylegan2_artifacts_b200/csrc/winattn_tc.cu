// Shifted-window attention core on tcgen05 (bf16): S = Q K^T, O = P V and the five backward products run on the 5th-gen tensor
// cores with TMEM accumulators; the softmax (scale, relative-position bias, -100 shift mask, fp32 statistics) runs on the TMEM
// lanes (forward: one thread per query row; backward: two).  TV:models/swin_transformer.py:181-214.
//
// Tiny-tile strategy (49 tokens x head-dim 32): two windows are packed into one M=128 MMA; the Q / K / V (/ dO) tiles of a unit
// (window pair x head) are [64 rows x 32] bf16 blocks (64B swizzle) that arrive as ONE 4-D TMA box straight from the window-ordered
// qkv rows (rows 49..63 of a block belong to the next window: finite values that meet exact zeros in P), the unit's outputs leave as
// ONE TMA box store.  Forward: 128-thread CTAs, three per SM, operand tiles double buffered.  Backward: one persistent
// warp-specialised CTA per SM (softmax group / epilogue group / MMA-issuing warp, two TMEM buffers, three operand stages).
// Kernel-specific notes precede each kernel.
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace msu {

constexpr float ATT_SCALE = 0.17677669529663687f;   // 32^-1/2
constexpr float LOG2E = 1.4426950408889634f;

// K-major, 64B-swizzled descriptor (rows of 32 bf16 = 64 B, 8-row groups 512 B apart)
__device__ __forceinline__ uint64_t make_desc_kmajor_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
    return d;
}
// MN-major, 64B-swizzled descriptor for a [k rows x 32] tile: one 64 B chunk along MN, 8-row k groups 512 B apart;
// `lbo_bytes` = distance to the next 32-column chunk (N = 64: the same rows of the pair's second window, 4 KB further)
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw64(uint32_t saddr, uint32_t lbo_bytes = 16) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

struct MaskInfoTc { int any, ty, tx; };
__device__ __forceinline__ MaskInfoTc mask_info_tc(const WinGeo& g, int w) {
    MaskInfoTc m;
    const uint32_t nwx = (uint32_t)g.nwx();
    const uint32_t wy = (uint32_t)w / nwx, wx = (uint32_t)w - wy * nwx;
    m.ty = (g.sh > 0 && (int)wy == g.Ph / WS - 1) ? WS - g.sh : WS;
    m.tx = (g.sw > 0 && (int)wx == g.Pw / WS - 1) ? WS - g.sw : WS;
    m.any = (m.ty < WS) || (m.tx < WS);
    return m;
}
// Shift mask of query token t as a bitmask over the 49 keys (bit j set = same region = attends normally; clear = -100,
// TV:models/swin_transformer.py:186-209).  Only the windows of the last window row / column of a shifted block have one, so the
// per-key work sits behind a warp-uniform branch (a warp's 32 rows belong to one window): the branch-free per-element
// region compares were 45 % of the kernel's instructions, paid by every window.
__device__ __forceinline__ uint64_t mask_allowed_tc(const MaskInfoTc& m, int t) {
    const uint32_t a = (uint32_t)t / WS, b = (uint32_t)t - a * WS;
    uint32_t cols = (1u << m.tx) - 1u;                 // keys left of the column split
    if ((int)b >= m.tx) cols = ~cols & 0x7fu;
    uint32_t rows = (1u << m.ty) - 1u;                 // key rows above the row split
    if ((int)a >= m.ty) rows = ~rows & 0x7fu;
    uint64_t allowed = 0;
#pragma unroll
    for (int r = 0; r < WS; r++)
        if ((rows >> r) & 1u) allowed |= (uint64_t)cols << (WS * r);
    return allowed;
}

// this head's [49, 49] relative-position bias, pre-multiplied by log2(e), rows padded to 50 floats (8 B aligned: the softmax
// reads it with 25 LDS.64 per row, conflict free for 16 consecutive rows).  All loads are issued before the first store: the
// rolled loop was 19 dependent global round trips, 12 % of the forward kernel's time.
constexpr int AT_BROW = 50;
__device__ __forceinline__ void stage_bias_tc(float* sBias, const float* __restrict__ bias, int h, int tid) {
    constexpr int N = WT * WT, PER = (N + 127) / 128;
    float v[PER];
    const float* src = bias + (int64_t)h * N;
#pragma unroll
    for (int m = 0; m < PER; m++) {
        const int k = tid + 128 * m;
        v[m] = k < N ? src[k] : 0.f;
    }
#pragma unroll
    for (int m = 0; m < PER; m++) {
        const int k = tid + 128 * m;
        if (k < N) {
            const int r = k / WT;
            sBias[r * AT_BROW + (k - r * WT)] = v[m] * LOG2E;
        }
    }
    if (tid < WT) sBias[tid * AT_BROW + WT] = 0.f;
}

__device__ __forceinline__ uint32_t pk2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int AT_P_BYTES = 16 * 1024;   // compact P tile: row = query (128), 64 own-window keys (128 B rows, 128B swizzle)
constexpr int AT_OPS_BYTES = 24 * 1024;  // one stage of Q, K, V tiles (8 KB each: two [64 rows x 32] boxes)
constexpr int AT_BIAS_BYTES = 49 * 50 * 4 + 8;  // this head's [49,49] bias (rows padded to 50), pre-multiplied by log2(e)
constexpr int AT_SMEM = AT_P_BYTES + 2 * AT_OPS_BYTES + AT_BIAS_BYTES + 64 + 1024;

// Forward.  Unit = two windows x one head.  The operand tiles are double buffered: the TMA loads of unit n+1 are issued
// before unit n is touched, so a CTA never waits for HBM; 75 KB of shared memory and 128 TMEM columns per CTA let three
// CTAs per SM overlap each other's MMA / softmax / store phases.
//   MMA 1:  S[128 x 128] = [Q_a; Q_b] [K_a; K_b]^T      (K = 32; the two off-diagonal 64x64 blocks are unused)
//   softmax per row on its own 64-column block -> P row (bf16, keys >= 49 zero) in the compact K-major tile
//   MMA 2:  O_w[128 x 32] = P[:, own keys] V_w for w = a, b into TMEM columns [32 w, 32 w + 32): rows of the other
//           window hold unused values; V is read MN-major straight from its TMA tile (no transpose)
__global__ void __launch_bounds__(128, 3)
winattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm4,
                      const __grid_constant__ CUtensorMap tmO, const float* __restrict__ bias,
                      int64_t n_windows, int nH, WinGeo g, AttnDrop ad, float* __restrict__ lse) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sP = smem;
    uint8_t* sOps = smem + AT_P_BYTES;        // [2][Q 8K | K 8K | V 8K]
    float* sBias = reinterpret_cast<float*>(sOps + 2 * AT_OPS_BYTES);
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + AT_BIAS_BYTES);   // [2]
    uint64_t* bar_mma = bar_load + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int C = nH * HD;
    const int h = blockIdx.y;

    if (tid == 0) {
        prefetch_tmap(&tm);
        prefetch_tmap(&tm4);
        prefetch_tmap(&tmO);
        mbar_init(&bar_load[0], 1);
        mbar_init(&bar_load[1], 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    pdl_wait();           // PDL: the bias table below comes from the previous kernel (msu_relbias_expand)
    stage_bias_tc(sBias, bias, h, tid);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t n_pairs = (n_windows + 1) / 2;
    const int nwin_img = g.nwin();
    const int half = tid >> 6;             // which window of the pair this row belongs to
    const int i = tid & 63;                // token index inside the window (valid if < 49)
    const uint32_t idesc1 = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc2 = make_idesc_bf16(128, 64, 0, 1);
    uint32_t ph_mma = 0;
    const uint32_t ds0 = ad.thr ? ad.seed[0] : 0u, ds1 = ad.thr ? ad.seed[1] : 0u;

    auto issue_loads = [&](int64_t pair, int buf) {
        uint8_t* q = sOps + buf * AT_OPS_BYTES;
        mbar_arrive_expect_tx(&bar_load[buf], 6 * 4096);
        if (pair + 1 < n_pairs) {
            // ONE box for the unit: [32 channels] x [64 rows] x [2 windows, 49 rows apart] x [q, k, v, C channels apart] lands as
            // Q_a Q_b K_a K_b V_a V_b (4 KB each).  An SM's TMA unit serves a box in a fixed ~222 clk up to ~12 KB (then ~57 B/clk):
            // six 4 KB boxes cost 1330 clk per unit, one 24 KB box ~430.  Rows 49..63 of a window's tile are the next window's
            // first rows (finite values that meet exact zeros in P); only the LAST pair would read past the tensor, so it keeps
            // the bounds-checked 2-D boxes below.
            tma_load_4d(q, &tm4, &bar_load[buf], h * HD, 0, (int)(pair * 2), 0);
            return;
        }
        const int r0 = (int)(pair * 2 * WT), r1 = r0 + WT;
        tma_load_2d(q, &tm, &bar_load[buf], h * HD, r0);
        tma_load_2d(q + 4096, &tm, &bar_load[buf], h * HD, r1);
        tma_load_2d(q + 8192, &tm, &bar_load[buf], C + h * HD, r0);
        tma_load_2d(q + 12288, &tm, &bar_load[buf], C + h * HD, r1);
        tma_load_2d(q + 16384, &tm, &bar_load[buf], 2 * C + h * HD, r0);
        tma_load_2d(q + 20480, &tm, &bar_load[buf], 2 * C + h * HD, r1);
    };
    if (tid == 0 && (int64_t)blockIdx.x < n_pairs) issue_loads(blockIdx.x, 0);

    int it = 0;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, it++) {
        const int buf = it & 1;
        const int64_t win = pair * 2 + half;
        uint8_t* sQ = sOps + buf * AT_OPS_BYTES;
        uint8_t* sK = sQ + 8192;
        uint8_t* sV = sQ + 16384;
        if (warp == 0 && elect_one()) {   // == thread 0 (the warp is converged here); elect: see tc_common.cuh
            mbar_wait(&bar_load[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint64_t qd = make_desc_kmajor_sw64(smem_u32(sQ)), kd = make_desc_kmajor_sw64(smem_u32(sK));
            tc_mma_bf16(tmem, qd, kd, idesc1, 0);
            tc_mma_bf16(tmem, qd + 2, kd + 2, idesc1, 1);   // +32 B: second K=16 slice of the 64 B rows
            // the previous unit's output box (staged in the P tile) must have left shared memory before any thread rewrites the
            // tile: waited for while the MMAs run, and before the commit that releases the softmax threads
            tma_store_wait_read0();
            tc_commit(bar_mma);
            // the other stage was last read by the MMAs of unit it-1, which completed before that unit's epilogue
            if (pair + gridDim.x < n_pairs) issue_loads(pair + gridDim.x, buf ^ 1);
        }
        // ---- softmax on this thread's row
        const bool row_ok = (i < WT) && (win < n_windows);
        const MaskInfoTc mi = mask_info_tc(g, (int)((uint32_t)win % (uint32_t)nwin_img));
        const float2* brow = reinterpret_cast<const float2*>(sBias + i * AT_BROW);
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        float s[64];
        {
            uint32_t* sr = reinterpret_cast<uint32_t*>(s);
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + half * 64;
            tc_ld32_nowait(trow, sr);
            tc_ld32_nowait(trow + 32, sr + 32);
            tc_ld_wait();
        }
        float inv = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int j = 24; j < 32; j++) pk[j] = 0u;
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 25; j++) {                       // log2-domain logits (column 49 of the padded bias row is 0)
                const float2 b2 = brow[j];
                s[2 * j] = fmaf(s[2 * j], ATT_SCALE * LOG2E, b2.x);
                s[2 * j + 1] = fmaf(s[2 * j + 1], ATT_SCALE * LOG2E, b2.y);
            }
            if (mi.any) {                                        // warp-uniform: windows on the shifted block's last row / column
                const uint64_t allowed = mask_allowed_tc(mi, i);
#pragma unroll
                for (int j = 0; j < WT; j++)
                    if (!((allowed >> j) & 1ull)) s[j] += -100.0f * LOG2E;
            }
            // four independent chains per reduction: with one warp per scheduler and CTA, a 49-long dependent chain of
            // FMNMX / FADD costs its full latency (~4 clk per link)
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < WT; j++) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            // normalise by what the tensor core will actually sum (the bf16-rounded probabilities): the rounding errors of a row
            // then cancel in its common component (max error of the 1024^2 logits 3.1e-2 -> 2.4e-2 of their range)
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < WT; j += 2) {
                const float p0 = ex2_fast(s[j] - mx);
                const float p1 = (j + 1 < WT) ? ex2_fast(s[j + 1] - mx) : 0.f;
                const uint32_t u = pk2(p0, p1);
                pk[j >> 1] = u;
                s4[(j >> 1) & 3] += __uint_as_float(u << 16) + __uint_as_float(u & 0xffff0000u);
            }
            const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            inv = __fdividef(ad.inv_keep, sum);
            if (lse != nullptr) {      // log2-domain log-sum-exp of the row: the backward forms P = 2^(l - lse) from it
                float lg;
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(sum));
                lse[(win * WT + i) * nH + h] = mx + lg;
            }
            if (ad.thr) {   // attention dropout: dropped probabilities leave the P tile (the row sum above is the undropped one)
                const uint32_t rowkey = (uint32_t)((win * nH + h) * WT + i);
#pragma unroll
                for (int jp = 0; jp < 25; jp++) {
                    const uint32_t hs = attn_drop_hash(rowkey, jp, ds0, ds1);
                    if ((hs & 0xffffu) < ad.thr) pk[jp] &= 0xffff0000u;
                    if ((hs >> 16) < ad.thr) pk[jp] &= 0x0000ffffu;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 25; j++) pk[j] = 0u;
        }
        // P row -> compact tile: row tid, 8 x 16 B chunks, 128B swizzle (chunk ^ (row & 7)); padding rows are zero
        {
            uint8_t* prow = sP + tid * 128;
#pragma unroll
            for (int c = 0; c < 8; c++)
                *reinterpret_cast<uint4*>(prow + ((c ^ (tid & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();   // all S rows are read (MMA 2 overwrites columns 0..63) and all P rows are written
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
#pragma unroll
            for (int k = 0; k < 4; k++) {   // 64 own keys = 4 K-steps (+32 B inside the 128 B rows), V advances 16 rows = 1 KB;
                // N = 64: columns 0..31 multiply V_a, 32..63 V_b (second 32-column chunk 4 KB further): rows of window w
                // hold O_w in columns [32 w, 32 w + 32), the other half of each row is unused
                const uint64_t ad = make_desc_kmajor_sw128(pa + k * 32);
                const uint64_t bd = make_desc_mnmajor_sw64(va + k * 1024, 4096);
                tc_mma_bf16(tmem, ad, bd, idesc2, k != 0);
            }
            tc_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        float o[32];
        {
            uint32_t* orr = reinterpret_cast<uint32_t*>(o);
            tc_ld32_nowait(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, orr);
            tc_ld_wait();
        }
        // O row -> staging box of its window ([49 x 32] bf16, 64B swizzle) in the P tile (MMA 2 has consumed it); the
        // two boxes leave as TMA stores: per-thread 64 B row stores were 32 separate lines per instruction
        if (i < WT) {
            // staging box [2 windows][49 rows][32] (64 B rows, 64B swizzle on the box-row index): the pair leaves as ONE TMA store
            const int R = half * WT + i;
            uint8_t* orow = sP + R * 64;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint4 v;
                v.x = pk2(o[8 * c] * inv, o[8 * c + 1] * inv);
                v.y = pk2(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
                v.z = pk2(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
                v.w = pk2(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
                *reinterpret_cast<uint4*>(orow + ((c ^ ((R >> 1) & 3)) << 4)) = v;
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();   // TMEM columns are free for the next unit; the staged rows are complete
        if (warp == 0 && elect_one()) {
            tma_store_3d(sP, &tmO, h * HD, 0, (int)(pair * 2));      // a window index past the end (odd count) is clipped by TMA
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
    }
}

// Window-pair tensor map over a window-ordered [n_windows * 49, parts * C] bf16 matrix:
//   dims  (channel, row in window, window, part)   sizes (C, box_rows, n_windows, parts)
//   strides        (row pitch, 49 rows, C channels) box   (32, box_rows, 2, parts)
// box_rows = 64: operand loads (the 15 extra rows of a window's tile are the next window's first rows — the caller must not
// use this map for the LAST pair, which would run past the allocation; out-of-range window indices are zero-filled);
// box_rows = 49: stores of exactly the windows' rows (out-of-range windows are clipped).  parts = 1 drops the last dimension.
static bool make_pair_map(CUtensorMap* tm, const void* ptr, int64_t n_windows, int C, int parts, int box_rows) {
    const cuuint64_t pitch = (cuuint64_t)parts * C * 2;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)box_rows, (cuuint64_t)n_windows, (cuuint64_t)parts};
    cuuint64_t gstr[3] = {pitch, (cuuint64_t)WT * pitch, (cuuint64_t)C * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)box_rows, 2, (cuuint32_t)parts};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return tc_get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, parts > 1 ? 4 : 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns 0 launched, 1 unsupported
int winattn_fwd_tc(const void* qkv, const float* bias, void* O, int64_t n_windows, int nH, const WinGeo& g, const AttnDrop& ad,
                   float* lse, cudaStream_t st) {
    if (tc_get_encode() == nullptr) return 1;
    const int C = nH * HD;
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(O) & 15)) return 1;
    const int64_t rows = n_windows * WT;
    CUtensorMap tm, tm4, tmO;
    if (!make_pair_map(&tmO, O, n_windows, C, 1, WT) || !make_pair_map(&tm4, qkv, n_windows, C, 3, 64)) return 1;
    cuuint64_t gdim[2] = {(cuuint64_t)(3 * C), (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)(3 * C) * 2};
    cuuint32_t box[2] = {32, 64};
    cuuint32_t estr[2] = {1, 1};
    if (tc_get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    static PerDeviceOnce attr;
    if (attr.need()) {
        cudaError_t e = cudaFuncSetAttribute(winattn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
        if (e != cudaSuccess) { set_error("winattn_fwd_tc: %s", cudaGetErrorString(e)); return (int)e; }
        attr.set();
    }
    const int64_t pairs = (n_windows + 1) / 2;
    // whole waves: 3 CTAs per SM are resident; a few CTAs more would start late and double the tail
    const int gx = (int)imax(1, imin(pairs, ((int64_t)num_sms() * 3) / nH));
    dim3 grid(gx, nH);
    {
        const cudaError_t e = launch_pdl<2>(winattn_fwd_tc_kernel, grid, dim3(128), (size_t)AT_SMEM, st, 1, tm, tm4, tmO, bias, n_windows, nH, g, ad, lse);
        if (e != cudaSuccess) { set_error("winattn_fwd_tc: launch: %s", cudaGetErrorString(e)); return (int)e; }
    }
    count_launch();
    return check_launch("winattn_fwd_tc");
}

// ================================================================================================
// Backward on tcgen05.  Per unit (two windows x one head), P is recomputed (nothing of size 49x49 is saved):
//   MMA a: S  = [Q_a;Q_b] [K_a;K_b]^T            -> TMEM cols [0,128)
//   MMA b: dP = [dO_a;dO_b] [V_a;V_b]^T          -> TMEM cols [128,256)
//   rows : P = softmax(S*s + bias + mask), delta = sum_j P dP, dS = P (dP - delta), d(bias) += dS (registers)
//          P and dS rows (bf16) go to two compact [128 rows x 64 own-window keys] 128B-swizzled tiles
//   MMA c/d: dV_w = P_w^T dO_w, dK_w = dS_w^T Q_w (A read MN-major from the same tiles, contraction over the 64
//            rows of window w; key-indexed results land in TMEM lanes 0..63 for both windows)
//   MMA e:   dQ_w = dS K_w (A = all 128 rows K-major; lanes of the other window hold unused values)
// dQ and dK are scaled by 32^-1/2 in the epilogue (S = s Q K^T).  d(bias) partial sums stay in registers over
// all units of the CTA and are written once: partial[2*blockIdx.x + half][h][49*49] (fixed order => deterministic).
// ================================================================================================
constexpr int AB_TILE = 8 * 1024;          // one [128 x 32] bf16 operand tile (two 64-row boxes)
constexpr int AB_X = 16 * 1024;            // one [128 x 64] bf16 P / dS tile
constexpr int AB_OPS = 4 * AB_TILE;        // one stage of Q, K, V, dO
constexpr int AB_XCH = 3 * 2 * 128 * 4;    // row-pair exchange slots: [max | sum | dot][part][row]
constexpr int AB_STAGES = 3;               // operand stages
constexpr int AB_NTAB = (2 * WS - 1) * (2 * WS - 1);   // 169 entries of one head's relative-position-bias table
constexpr int AB_SMEM = 2 * 2 * AB_X + AB_STAGES * AB_OPS + AT_BIAS_BYTES + AB_XCH + 256 + 1024 + 1024;
constexpr int AB_THREADS = 13 * 32;
constexpr int AB_KSPLIT = 24;              // part 0 owns keys [0, 24), part 1 keys [24, 49)

__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void pair_sync(int wq) {   // the two warps that share TMEM lane quadrant wq
    switch (wq) {           // immediate ids: a register id makes ptxas reserve all 16 named barriers
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}

// Backward kernel.  256 threads: TWO threads per query row (the tensor memory caps a CTA at 256 columns = 2 CTAs per SM, and
// with one thread per row that was 2 warps per scheduler, each a ~1100-instruction serial stream per unit: issue slots 35 %
// used).  Warps w and w + 4 read the same TMEM lane quadrant; part 0 (warps 0-3) owns keys [0, 24) of the row, part 1 keys
// [24, 49): three 16 B chunks of the P / dS tile rows each (+ the zero tail), 24 / 25 d(bias) accumulators, and half of the
// dq / dk / dv epilogue columns.  Row maximum, row sum and sum_j e_j dP_j are exchanged through shared memory with a 64-thread
// named barrier per warp pair (two exchanges per unit).
// Operand tiles are double buffered (the loads of unit n+1 are issued before unit n is touched).  dV / dK of window w
// use ONE MN-major A descriptor over the compact tile with the two 64-key chunks 8 KB apart (= the two windows' row
// blocks): the products of window w's rows land in TMEM lanes 64 w .. 64 w + 63 (the other half holds unused values),
// so every row's dq, dk and dv come from its own lane — a balanced epilogue.
template <int PART>
__device__ __forceinline__ void winattn_bwd_rows(uint32_t trow, const float* __restrict__ brow_base, bool row_ok, const MaskInfoTc& mi,
                                                 int i, int row, int wq, float* xch, const AttnDrop& ad, uint32_t rowkey,
                                                 uint32_t ds0, uint32_t ds1, float (&acc)[25], uint8_t* prow, uint8_t* srow,
                                                 uint64_t* tile_free, uint32_t tile_parity, bool tile_wait, bool have_lse, float lse_row) {
    constexpr int KB = PART == 0 ? 0 : AB_KSPLIT;           // first key of this part
    constexpr int NK = PART == 0 ? AB_KSPLIT : WT - AB_KSPLIT;   // 24 / 25 keys
    float s[32], dp[32];
    {
        uint32_t* sr = reinterpret_cast<uint32_t*>(s);
        uint32_t* dr = reinterpret_cast<uint32_t*>(dp);
        if (PART == 0) {
            tc_ld16_nowait(trow, sr);
            tc_ld8_nowait(trow + 16, sr + 16);
            tc_ld16_nowait(trow + 128, dr);
            tc_ld8_nowait(trow + 128 + 16, dr + 16);
        } else {
            tc_ld8_nowait(trow + 24, sr);
            tc_ld16_nowait(trow + 32, sr + 8);
            tc_ld8_nowait(trow + 48, sr + 24);
            tc_ld8_nowait(trow + 128 + 24, dr);
            tc_ld16_nowait(trow + 128 + 32, dr + 8);
            tc_ld8_nowait(trow + 128 + 48, dr + 24);
        }
        tc_ld_wait();
    }
    float* x_max = xch, *x_sum = xch + 256, *x_dot = xch + 512;
    float mpart = -INFINITY;
    if (row_ok) {
        const float2* brow = reinterpret_cast<const float2*>(brow_base + KB);
#pragma unroll
        for (int j = 0; j < (NK + 1) / 2; j++) {            // log2-domain logits (column 49 of the padded bias row is 0)
            const float2 b2 = brow[j];
            s[2 * j] = fmaf(s[2 * j], ATT_SCALE * LOG2E, b2.x);
            s[2 * j + 1] = fmaf(s[2 * j + 1], ATT_SCALE * LOG2E, b2.y);
        }
        if (mi.any) {                                        // warp-uniform (see the forward kernel)
            const uint64_t allowed = mask_allowed_tc(mi, i) >> KB;
#pragma unroll
            for (int j = 0; j < NK; j++)
                if (!((allowed >> j) & 1ull)) s[j] += -100.0f * LOG2E;
        }
        if (!have_lse) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < NK; j++) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
            mpart = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
    }
    // with the forward's log-sum-exp the probabilities come out normalised (P = 2^(l - lse)): no row maximum, no row sum, one
    // exchange (sum_j P dP) instead of two
    float mx = lse_row;
    if (!have_lse) {
        x_max[PART * 128 + row] = mpart;
        pair_sync(wq);
        mx = fmaxf(mpart, x_max[(PART ^ 1) * 128 + row]);
    }
    float spart = 0.f, dpart = 0.f;
    if (row_ok) {
        if (ad.thr) {   // attention dropout: dP = m * dP~ with the forward's mask m in {0, 1/keep}
#pragma unroll
            for (int jp = 0; jp < (NK + 1) / 2; jp++) {
                const uint32_t hs = attn_drop_hash(rowkey, KB / 2 + jp, ds0, ds1);
                dp[2 * jp] *= ((hs & 0xffffu) >= ad.thr) ? ad.inv_keep : 0.f;
                if (2 * jp + 1 < NK) dp[2 * jp + 1] *= ((hs >> 16) >= ad.thr) ? ad.inv_keep : 0.f;
            }
        }
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < NK; j++) {
            s[j] = ex2_fast(s[j] - mx);
            s4[j & 3] += s[j];
            d4[j & 3] = fmaf(s[j], dp[j], d4[j & 3]);
        }
        spart = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        dpart = (d4[0] + d4[1]) + (d4[2] + d4[3]);
    }
    if (!have_lse) x_sum[PART * 128 + row] = spart;
    x_dot[PART * 128 + row] = dpart;
    pair_sync(wq);
    uint32_t pk[16], dk_[16];
#pragma unroll
    for (int j = 0; j < 16; j++) { pk[j] = 0u; dk_[j] = 0u; }
    if (row_ok) {
        // both threads of the row add the two partial sums in the same order: identical inv / delta
        float delta = x_dot[row] + x_dot[128 + row];
        if (!have_lse) {
            const float inv = __fdividef(1.0f, x_sum[row] + x_sum[128 + row]);
            delta *= inv;
#pragma unroll
            for (int j = 0; j < NK; j++) s[j] *= inv;                  // P
        }
        if (!ad.thr) {
#pragma unroll
            for (int j = 0; j < NK; j += 2) pk[j >> 1] = pk2(s[j], (j + 1 < NK) ? s[j + 1] : 0.f);
        } else {        // with dropout the dV operand is P~ = m * P (the kept entries scaled by 1/keep)
#pragma unroll
            for (int jp = 0; jp < (NK + 1) / 2; jp++) {
                const uint32_t hs = attn_drop_hash(rowkey, KB / 2 + jp, ds0, ds1);
                const float m0 = ((hs & 0xffffu) >= ad.thr) ? ad.inv_keep : 0.f;
                const float m1 = ((hs >> 16) >= ad.thr) ? ad.inv_keep : 0.f;
                pk[jp] = pk2(s[2 * jp] * m0, (2 * jp + 1 < NK) ? s[2 * jp + 1] * m1 : 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < NK; j++) {
            dp[j] = s[j] * (dp[j] - delta);   // dS
            acc[j] += dp[j];
        }
#pragma unroll
        for (int j = 0; j < NK; j += 2) dk_[j >> 1] = pk2(dp[j], (j + 1 < NK) ? dp[j + 1] : 0.f);
    }
    // own 16 B chunks of row `row` of both compact tiles (128 B rows, 128B swizzle); zeros for padding rows and for the
    // key tail 49..63 (they are contracted over in dV / dK).  The tile buffer was the staging box of the unit two back:
    // its TMA store must have drained.
    if (tile_wait) mbar_wait(tile_free, tile_parity);
    constexpr int C0 = PART == 0 ? 0 : 3, NC = PART == 0 ? 3 : 5;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const int sw = ((C0 + c) ^ (row & 7)) << 4;
        const uint4 pv = c < 4 ? make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]) : make_uint4(0, 0, 0, 0);
        const uint4 sv = c < 4 ? make_uint4(dk_[4 * c], dk_[4 * c + 1], dk_[4 * c + 2], dk_[4 * c + 3]) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(prow + sw) = pv;
        *reinterpret_cast<uint4*>(srow + sw) = sv;
    }
}

// Persistent, warp-specialised: ONE CTA per SM with all 512 TMEM columns = two S / dP buffers, so that two units are in
// flight per CTA and the softmax group never waits for its own unit's MMAs:
//   warps 0-7   softmax group (two threads per row, see winattn_bwd_rows): unit n in TMEM buffer n & 1, P / dS tiles n & 1
//   warps 8-11  epilogue group (one thread per row): dq / dk / dv of unit n from buffer n & 1 -> staging box -> TMA store;
//               warp 8 lane 0 also issues the TMA loads (unit n + 3 into the operand stage unit n has just released)
//   warp  12    MMA issuer (one thread): dV / dK / dQ of unit n when its tiles are written, S / dP of unit n + 2 as soon as
//               the epilogue group has drained buffer n & 1.  It has a warp of its own: the twelve dV / dK / dQ MMAs execute in
//               ~1300 clk (MN-major operands: shared-memory-read bound) and the issuing thread blocks behind them.
// (Register files are allocated to warps in groups of four: 13 warps count as 16, 128 registers per thread.)
// Per TMEM buffer the cycle is S / dP MMAs -> softmax -> dV / dK / dQ MMAs -> drain; with two buffers a unit costs
// max(softmax, half that cycle).  The one-unit-at-a-time version (2 CTAs per SM) spent a third of every warp's time waiting
// for its own unit's MMAs (~850 clk for S / dP, ~1150 clk for the twelve shared-memory-read-bound dV / dK / dQ MMAs).
__global__ void __launch_bounds__(AB_THREADS, 1)
winattn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmQKV4, const __grid_constant__ CUtensorMap tmDO3,
                      const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ bias, float* __restrict__ dbias_partial,
                      float* __restrict__ dtable,
                      int64_t n_windows, int nH, WinGeo g, AttnDrop ad, const float* __restrict__ lse, long long* trace) {
#ifdef MSU_ATT_TRACE_BUILD   // phase timeline (debug builds only), first 8 units of CTAs with blockIdx.y == 0: [cta][unit][16 events]
#define AB_TRACE(ev) do { if (trace != nullptr && lane == 0 && blockIdx.y == 0 && n < 8) trace[((size_t)blockIdx.x * 8 + n) * 16 + (ev)] = clock64(); } while (0)
#else
#define AB_TRACE(ev) do { } while (0)
#endif
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sTiles = smem;                             // [2][P 16 KB | dS 16 KB]
    uint8_t* sOps = sTiles + 2 * 2 * AB_X;              // [AB_STAGES][Q | K | V | dO]
    float* sBias = reinterpret_cast<float*>(sOps + AB_STAGES * AB_OPS);
    float* sXch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sBias) + AT_BIAS_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sXch) + AB_XCH);
    uint64_t* full = bars;                              // [AB_STAGES] operands landed
    uint64_t* s_ready = full + AB_STAGES;               // [2] S / dP in TMEM
    uint64_t* tiles_ready = s_ready + 2;                // [2] P / dS tiles written (8 softmax warps)
    uint64_t* out_ready = tiles_ready + 2;              // [2] dV / dK / dQ in TMEM (their MMAs retired: tiles and operand stage free)
    uint64_t* buf_free = out_ready + 2;                 // [2] the epilogue group has drained the TMEM buffer (4 warps)
    uint64_t* tile_free = buf_free + 2;                 // [2] the staging box has left the tile buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_free + 2);
    float* sTab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);     // [169] this CTA's bias-table gradient (dtable mode)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = nH * HD;
    const int h = blockIdx.y;
    if (tid >= 128 && tid < 128 + AB_NTAB) sTab[tid - 128] = 0.f;
    if (tid == 0) {
        prefetch_tmap(&tmQKV);
        prefetch_tmap(&tmDO);
        prefetch_tmap(&tmQKV4);
        prefetch_tmap(&tmDO3);
        prefetch_tmap(&tmOut);
        for (int k = 0; k < AB_STAGES; k++) mbar_init(&full[k], 1);
        for (int k = 0; k < 2; k++) {
            mbar_init(&s_ready[k], 1);
            mbar_init(&tiles_ready[k], 8);
            mbar_init(&out_ready[k], 1);
            mbar_init(&buf_free[k], 4);
            mbar_init(&tile_free[k], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    pdl_wait();           // PDL: everything below reads what earlier kernels wrote
    if (tid < 128) stage_bias_tc(sBias, bias, h, tid);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t n_pairs = (n_windows + 1) / 2;
    const int n_units = (int)((n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x);    // pairs blockIdx.x, + gridDim.x, ...
    const int nwin_img = g.nwin();

    if (warp < 8) {
        // ===================== softmax group =====================
        const int wq = warp & 3, part = warp >> 2;
        const int row = wq * 32 + lane;             // query row of the unit's 128 (TMEM lane)
        const int half = row >> 6, i = row & 63;
        const uint32_t ds0 = ad.thr ? ad.seed[0] : 0u, ds1 = ad.thr ? ad.seed[1] : 0u;
        float acc[25];
#pragma unroll
        for (int j = 0; j < 25; j++) acc[j] = 0.f;
        const float* brow = sBias + i * AT_BROW;
        for (int n = 0; n < n_units; n++) {
            const int b = n & 1;
            const uint32_t par = (uint32_t)(n >> 1) & 1u;
            const int64_t pair = blockIdx.x + (int64_t)n * gridDim.x;
            const int64_t win = pair * 2 + half;
            const bool row_ok = (i < WT) && (win < n_windows);
            const MaskInfoTc mi = mask_info_tc(g, (int)((uint32_t)win % (uint32_t)nwin_img));
            const uint32_t rowkey = (uint32_t)((win * nH + h) * WT + i);
            uint8_t* sXP = sTiles + b * 2 * AB_X;
            uint8_t* sXS = sXP + AB_X;
            const bool have_lse = lse != nullptr;
            const float lse_row = (have_lse && row_ok) ? lse[(win * WT + i) * nH + h] : 0.f;
            if (warp == 0) AB_TRACE(0);
            mbar_wait(&s_ready[b], par);
            tc_fence_after();
            if (warp == 0) AB_TRACE(1);
            const uint32_t trow = tmem + b * 256 + ((uint32_t)(wq * 32) << 16) + half * 64;
            if (part == 0) winattn_bwd_rows<0>(trow, brow, row_ok, mi, i, row, wq, sXch, ad, rowkey, ds0, ds1, acc, sXP + row * 128, sXS + row * 128,
                                               &tile_free[b], par ^ 1u, n >= 2, have_lse, lse_row);
            else winattn_bwd_rows<1>(trow, brow, row_ok, mi, i, row, wq, sXch, ad, rowkey, ds0, ds1, acc, sXP + row * 128, sXS + row * 128,
                                     &tile_free[b], par ^ 1u, n >= 2, have_lse, lse_row);
            fence_proxy_async_smem();               // tile rows (generic proxy) -> visible to the MMAs (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tiles_ready[b]);
            if (warp == 0) AB_TRACE(2);
        }
        // d(bias) partial of this CTA: slab (2*blockIdx.x + half), head h, row i, this part's keys
        if (dtable != nullptr) {
            // straight into the bias-table gradient [169, nH] (zeroed or accumulating, the caller's choice): the entry of (query i, key j)
            // is (yi - yj + 6) * 13 + (xi - xj + 6).  The CTA folds its 2 x 49 x 49 values into 169 in shared memory first, then adds
            // those with red.global.add: no partial slabs, no reduce kernels (two launches and 24 us per block before).  Adding the
            // 4802 values straight to global memory serialised on the 169 addresses (stage 2: 47 -> 106 us).
            if (i < WT) {
                const int nk = part ? WT - AB_KSPLIT : AB_KSPLIT;
                const int base_t = (i / WS + WS - 1) * (2 * WS - 1) + (i % WS) + WS - 1;
#pragma unroll
                for (int j = 0; j < 25; j++) {
                    if (j < nk) {
                        const int jj = j + (part ? AB_KSPLIT : 0);
                        atomicAdd(&sTab[base_t - ((jj / WS) * (2 * WS - 1) + (jj % WS))], acc[j]);
                    }
                }
            }
            asm volatile("bar.sync 6, 256;" ::: "memory");           // the eight softmax warps
            if (tid < AB_NTAB) atomicAdd(&dtable[tid * nH + h], sTab[tid]);
        } else if (i < WT) {
            float* out = dbias_partial + (((int64_t)blockIdx.x * 2 + half) * nH + h) * (WT * WT) + i * WT + (part ? AB_KSPLIT : 0);
            const int nk = part ? WT - AB_KSPLIT : AB_KSPLIT;
#pragma unroll
            for (int j = 0; j < 25; j++)
                if (j < nk) out[j] = acc[j];
        }
    } else if (warp < 12) {
        // ===================== epilogue group (warp 8 lane 0 also issues the TMA loads and the TMA store) =====================
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int half = row >> 6, i = row & 63;
        const bool is_tma = (warp == 8 && lane == 0);
        auto issue_loads = [&](int n) {                  // unit n -> stage n % AB_STAGES (the caller knows the stage is free)
            const int st = n % AB_STAGES;
            const int64_t pair = blockIdx.x + (int64_t)n * gridDim.x;
            uint8_t* q = sOps + st * AB_OPS;
            mbar_arrive_expect_tx(&full[st], 8 * 4096);
            if (pair + 1 < n_pairs) {   // two boxes per unit instead of eight (see the forward kernel); the last pair stays bounds-checked
                tma_load_4d(q, &tmQKV4, &full[st], h * HD, 0, (int)(pair * 2), 0);
                tma_load_3d(q + 3 * AB_TILE, &tmDO3, &full[st], h * HD, 0, (int)(pair * 2));
                return;
            }
            const int r0 = (int)(pair * 2 * WT), r1 = r0 + WT;
            tma_load_2d(q, &tmQKV, &full[st], h * HD, r0);
            tma_load_2d(q + 4096, &tmQKV, &full[st], h * HD, r1);
            tma_load_2d(q + AB_TILE, &tmQKV, &full[st], C + h * HD, r0);
            tma_load_2d(q + AB_TILE + 4096, &tmQKV, &full[st], C + h * HD, r1);
            tma_load_2d(q + 2 * AB_TILE, &tmQKV, &full[st], 2 * C + h * HD, r0);
            tma_load_2d(q + 2 * AB_TILE + 4096, &tmQKV, &full[st], 2 * C + h * HD, r1);
            tma_load_2d(q + 3 * AB_TILE, &tmDO, &full[st], h * HD, r0);
            tma_load_2d(q + 3 * AB_TILE + 4096, &tmDO, &full[st], h * HD, r1);
        };
        if (is_tma)
            for (int n = 0; n < AB_STAGES && n < n_units; n++) issue_loads(n);
        for (int n = 0; n < n_units; n++) {
            const int b = n & 1;
            const uint32_t par = (uint32_t)(n >> 1) & 1u;
            const int64_t pair = blockIdx.x + (int64_t)n * gridDim.x;
            uint8_t* sStage = sTiles + b * 2 * AB_X;     // the unit's own tile buffer, once its MMAs have retired
            mbar_wait(&out_ready[b], par);
            tc_fence_after();
            if (warp == 8) AB_TRACE(8);
            if (is_tma && n + AB_STAGES < n_units) issue_loads(n + AB_STAGES);   // the unit's operand stage is free
            // own query row of dQ and own key row of dK, dV (lanes 64..127 hold window b)
            float dq[32], dkk[32], dvv[32];
            const uint32_t lane_base = tmem + b * 256 + ((uint32_t)(wq * 32) << 16) + half * 32;
            tc_ld32_nowait(lane_base + 128, reinterpret_cast<uint32_t*>(dq));
            tc_ld32_nowait(lane_base + 64, reinterpret_cast<uint32_t*>(dkk));
            tc_ld32_nowait(lane_base, reinterpret_cast<uint32_t*>(dvv));
            tc_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&buf_free[b]);    // S / dP of unit n + 2 may overwrite the buffer
            // rows -> ONE staging box [dq | dk | dv][2 windows][49 rows][32] (64 B rows, 64B swizzle on the box-row index):
            // the unit's 18.4 KB of gradients leave as one TMA store
            if (i < WT) {
                const int R0 = half * WT + i, R1 = R0 + 2 * WT, R2 = R0 + 4 * WT;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    *reinterpret_cast<uint4*>(sStage + R0 * 64 + ((c ^ ((R0 >> 1) & 3)) << 4)) = make_uint4(pk2(dq[8 * c] * ATT_SCALE, dq[8 * c + 1] * ATT_SCALE),
                        pk2(dq[8 * c + 2] * ATT_SCALE, dq[8 * c + 3] * ATT_SCALE), pk2(dq[8 * c + 4] * ATT_SCALE, dq[8 * c + 5] * ATT_SCALE),
                        pk2(dq[8 * c + 6] * ATT_SCALE, dq[8 * c + 7] * ATT_SCALE));
                    *reinterpret_cast<uint4*>(sStage + R1 * 64 + ((c ^ ((R1 >> 1) & 3)) << 4)) = make_uint4(pk2(dkk[8 * c] * ATT_SCALE, dkk[8 * c + 1] * ATT_SCALE),
                        pk2(dkk[8 * c + 2] * ATT_SCALE, dkk[8 * c + 3] * ATT_SCALE), pk2(dkk[8 * c + 4] * ATT_SCALE, dkk[8 * c + 5] * ATT_SCALE),
                        pk2(dkk[8 * c + 6] * ATT_SCALE, dkk[8 * c + 7] * ATT_SCALE));
                    *reinterpret_cast<uint4*>(sStage + R2 * 64 + ((c ^ ((R2 >> 1) & 3)) << 4)) = make_uint4(pk2(dvv[8 * c], dvv[8 * c + 1]), pk2(dvv[8 * c + 2], dvv[8 * c + 3]),
                        pk2(dvv[8 * c + 4], dvv[8 * c + 5]), pk2(dvv[8 * c + 6], dvv[8 * c + 7]));
                }
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync 5, 128;" ::: "memory");          // the four epilogue warps
            if (is_tma) {
                tma_store_4d(sStage, &tmOut, h * HD, 0, (int)(pair * 2), 0);     // a window index past the end (odd count) is clipped by TMA
                tma_store_commit();
                tma_store_wait_read0();                  // the box has left shared memory: the softmax group may reuse the tiles
                mbar_arrive(&tile_free[b]);
                AB_TRACE(9);
            }
        }
        if (is_tma) tma_store_wait_all();
    } else if (elect_one()) {
        // ===================== MMA issuer (one thread) =====================
        const uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);     // S, dP
        const uint32_t id_kv = make_idesc_bf16(128, 64, 1, 1);     // dV, dK: A MN-major, B MN-major (both windows' columns)
        const uint32_t id_q = make_idesc_bf16(128, 64, 0, 1);      // dQ: A K-major, B MN-major
        auto mma1 = [&](int n) {                         // S and dP of unit n into TMEM buffer n & 1 (the caller knows it is free)
            const int st = n % AB_STAGES;
            mbar_wait(&full[st], (uint32_t)(n / AB_STAGES) & 1u);
            tc_fence_after();
            const uint8_t* q = sOps + st * AB_OPS;
            const uint64_t qd = make_desc_kmajor_sw64(smem_u32(q)), kd = make_desc_kmajor_sw64(smem_u32(q + AB_TILE));
            const uint64_t vd = make_desc_kmajor_sw64(smem_u32(q + 2 * AB_TILE)), dd = make_desc_kmajor_sw64(smem_u32(q + 3 * AB_TILE));
            const uint32_t t = tmem + (n & 1) * 256;
            tc_mma_bf16(t, qd, kd, id_s, 0);
            tc_mma_bf16(t, qd + 2, kd + 2, id_s, 1);
            tc_mma_bf16(t + 128, dd, vd, id_s, 0);
            tc_mma_bf16(t + 128, dd + 2, vd + 2, id_s, 1);
            tc_commit(&s_ready[n & 1]);
        };
        for (int n = 0; n < 2 && n < n_units; n++) mma1(n);
        for (int n = 0; n < n_units; n++) {
            const int b = n & 1, st = n % AB_STAGES;
            const uint32_t par = (uint32_t)(n >> 1) & 1u;
            AB_TRACE(4);
            mbar_wait(&tiles_ready[b], par);
            tc_fence_after();
            AB_TRACE(5);
            const uint8_t* q = sOps + st * AB_OPS;
            const uint32_t xp = smem_u32(sTiles + b * 2 * AB_X), xs = xp + AB_X;
            const uint32_t qa = smem_u32(q), ka = qa + AB_TILE, da = qa + 3 * AB_TILE;
            const uint32_t t = tmem + b * 256;
#pragma unroll
            for (int k = 0; k < 4; k++) {   // contraction over 16 rows of both windows per step (2 KB of X, 1 KB of B per window)
                // A: M chunk 0 = keys of window a (rows [16k, 16k+16) of block a), chunk 1 = 8 KB further = block b.
                // B: N = 64, columns 0..31 from window a's dO / Q rows, 32..63 from window b's (4 KB further): lanes
                // [64 w, 64 w + 64) hold window w's result in columns [32 w, 32 w + 32), the rest is unused
                const uint64_t ap = make_desc_mnmajor_sw128(xp + k * 2048, 8192);
                const uint64_t as = make_desc_mnmajor_sw128(xs + k * 2048, 8192);
                tc_mma_bf16(t, ap, make_desc_mnmajor_sw64(da + k * 1024, 4096), id_kv, k != 0);        // dV
                tc_mma_bf16(t + 64, as, make_desc_mnmajor_sw64(qa + k * 1024, 4096), id_kv, k != 0);   // dK
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {   // contraction over the 64 own keys: +32 B per step inside the 128 B rows
                const uint64_t as = make_desc_kmajor_sw128(xs + k * 32);
                tc_mma_bf16(t + 128, as, make_desc_mnmajor_sw64(ka + k * 1024, 4096), id_q, k != 0);   // dQ
            }
            tc_commit(&out_ready[b]);
            AB_TRACE(6);
            if (n + 2 < n_units) {                       // S / dP of unit n + 2 as soon as the epilogue group has drained the buffer
                mbar_wait(&buf_free[b], par);
                mma1(n + 2);
                AB_TRACE(7);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

static bool make_attn_map(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int box_rows = 64) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return tc_get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int winattn_bwd_tc_grid(int64_t n_windows, int nH) {
    const int64_t pairs = (n_windows + 1) / 2;
    return (int)imax(1, imin(pairs, (int64_t)num_sms() / nH));         // persistent: one CTA per SM
}

// returns 0 launched (dbias_partial holds 2*winattn_bwd_tc_grid slabs), 1 unsupported
int winattn_bwd_tc(const void* qkv, const float* bias, const void* dO, void* dqkv, float* dbias_partial, float* dtable, int64_t n_windows,
                   int nH, const WinGeo& g, const AttnDrop& ad, const float* lse, cudaStream_t st) {
    if (tc_get_encode() == nullptr) return 1;
    const int C = nH * HD;
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(dO) & 15) || (reinterpret_cast<uintptr_t>(dqkv) & 15)) return 1;
    CUtensorMap tmQKV, tmDO, tmQKV4, tmDO3, tmOut;
    if (!make_attn_map(&tmQKV, qkv, n_windows * WT, 3 * C) || !make_attn_map(&tmDO, dO, n_windows * WT, C) ||
        !make_pair_map(&tmQKV4, qkv, n_windows, C, 3, 64) || !make_pair_map(&tmDO3, dO, n_windows, C, 1, 64) ||
        !make_pair_map(&tmOut, dqkv, n_windows, C, 3, WT))
        return 1;
    static PerDeviceOnce attr;
    if (attr.need()) {
        cudaError_t e = cudaFuncSetAttribute(winattn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
        if (e != cudaSuccess) { set_error("winattn_bwd_tc: %s", cudaGetErrorString(e)); return (int)e; }
        attr.set();
    }
    dim3 grid(winattn_bwd_tc_grid(n_windows, nH), nH);
    long long* trace_buf = nullptr;
#ifdef MSU_ATT_TRACE_BUILD
    static const int trace_on = getenv("MSU_ATT_TRACE") ? atoi(getenv("MSU_ATT_TRACE")) : 0;
    const size_t trace_n = (size_t)grid.x * 8 * 16;
    if (trace_on) {
        cudaMalloc(&trace_buf, trace_n * sizeof(long long));
        cudaMemsetAsync(trace_buf, 0, trace_n * sizeof(long long), st);
    }
#endif
    {
        const cudaError_t e = launch_pdl<2>(winattn_bwd_tc_kernel, grid, dim3(AB_THREADS), (size_t)AB_SMEM, st, 1, tmQKV, tmDO, tmQKV4, tmDO3, tmOut,
                                         bias, dbias_partial, dtable, n_windows, nH, g, ad, lse, trace_buf);
        if (e != cudaSuccess) { set_error("winattn_bwd_tc: launch: %s", cudaGetErrorString(e)); return (int)e; }
    }
#ifdef MSU_ATT_TRACE_BUILD
    if (trace_on) {   // synchronous dump of the per-unit role timeline of two CTAs
        long long* host = (long long*)malloc(trace_n * sizeof(long long));
        cudaStreamSynchronize(st);
        cudaMemcpy(host, trace_buf, trace_n * sizeof(long long), cudaMemcpyDeviceToHost);
        static const char* names[10] = {"sm_wait", "sm_S_ready", "sm_tiles_done", "-", "mma_wait_tiles", "mma_tiles_ready", "mma2_issued",
                                        "mma1_next_issued", "epi_out_ready", "epi_stored"};
        for (int cta : {0, (int)grid.x / 2}) {
            const long long t0 = host[(size_t)cta * 128];
            fprintf(stderr, "[att trace] nwin=%lld nH=%d grid=%d cta %d\n", (long long)n_windows, nH, (int)grid.x, cta);
            for (int it = 0; it < 8; it++) {
                fprintf(stderr, "  unit %d:", it);
                for (int ev = 0; ev < 10; ev++)
                    if (ev != 3) fprintf(stderr, " %s=%lld", names[ev], host[((size_t)cta * 8 + it) * 16 + ev] ? host[((size_t)cta * 8 + it) * 16 + ev] - t0 : -1);
                fprintf(stderr, "\n");
            }
        }
        free(host);
        cudaFree(trace_buf);
    }
#endif
    count_launch();
    return check_launch("winattn_bwd_tc");
}

}  // namespace msu
