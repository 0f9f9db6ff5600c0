// tcgen05 / TMEM / TMA GEMM for sm_100a: C[m,n] = epilogue(sum_k A(m,k) B(n,k)), bf16 operands, fp32 accumulate.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer (+TMEM
// alloc), warps 2..2+EPW = self-contained epilogue warps: each takes 32-row x 32-column blocks of the
// accumulator (its TMEM lane quadrant, column chunks round-robin among the warps of the quadrant):
// tcgen05.ld -> +bias -> bf16 -> a private 2.5 KB shared-memory transpose -> fused epilogue -> coalesced
// 16 B global stores (8 rows x 64 B per instruction).  No barrier between epilogue warps; the residual /
// GELU' operands are requested before the TMEM load is waited for.  Operands are staged in 128B-swizzled
// shared memory by cp.async.bulk.tensor; accumulators live in TMEM (2 x 256 columns so the epilogue of
// tile i overlaps the MMAs of tile i+1).  UMMA shape M=128, N=BN (16..256), K=16.  The epilogue scratch is
// only EPW x 2.5 KB, which leaves room for 4-5 operand stages: measured TMA round trips are ~1700 cycles
// under load, so mainloop throughput = operand bytes in flight / latency.
//
// A-operand modes (all K-major, i.e. contraction index contiguous in memory):
//   plain   : [M, K] rows                              (2-D tensor map)
//   dual    : cat([X1, X2], -1) without the concat      (two 2-D maps; network/model_parts.py:792,804,823)
//   conv3x3 : implicit im2col of an NHWC image          (4-D map, zero fill outside the image does the
//             padding; network/model_parts.py:468-471)
// B is always a K-major [N, K] bf16 weight shadow.  K tails are zero-filled by TMA (out-of-bounds).
// Every output map / fused epilogue of MsuEpilogue is honoured (each epilogue thread owns one output row).
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

#include "common.cuh"

namespace msu {

// ---- CTA pair (cta_group::2) helpers: two CTAs of a cluster (the two SMs of a TPC) run ONE M = 256 UMMA; each holds its own
// 128 rows of A and HALF of B's rows, so the B bytes an SM stages and its tensor core reads are halved.  In the
// shared::cluster window the CTA rank of a pair is bit 24 of a shared-memory address: clearing it addresses the same offset
// in the even (leader) CTA.
constexpr uint32_t TC_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads whose completion bytes are counted by the LEADER CTA's mbarrier (issued by both CTAs, each into its own shared memory)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)), "l"(tm), "r"(smem_u32(bar) & TC_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)), "l"(tm), "r"(smem_u32(bar) & TC_PEER_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)), "l"(tm), "r"(smem_u32(bar) & TC_PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
// arrives (once the MMAs issued so far have retired) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
        smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the leader CTA's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & TC_PEER_MASK) : "memory");
}

// ------------------------------------------------------------------------------------------------
struct TcParams {
    int64_t M;
    int N, K, BN, num_m_tiles, num_n_tiles;
    int mode;                 // 0 plain, 1 dual, 2 conv3x3 (shifted boxes), 3 conv3x3 halo rows, 4 conv3x3 halo rows x 2 image rows (BK = 32)
    int nacc;                 // accumulators (128-row segments) per tile: 2 in mode 4
    int kb1, kb2, k_split;    // k-blocks (of 64) from source 1 / source 2; column where source 2 starts in B
    int C, H, W, bmw, bmh, cblocks, tiles_x, tiles_y;  // conv geometry
    int stages;               // operand pipeline depth
    int pair;                 // mode 4: CTA pairs (cta_group::2, M = 256): each CTA stages half of every weight tile
    int box2;                 // mode 4, MSU_CONV_2BOX=1: the two halo rows arrive as one TMA box and the three weight tiles as one 3-D
                              // box (2 boxes per K block instead of 5).  Parity-tested, but measured the same 687 us: with the box
                              // rate out of the way the kernel sits on its shared-memory bandwidth bound, so it stays off
    int epi_tma;              // 1: per-warp swizzled slabs + TMA stores (Cpre out / R or H in through TMA too); depth-to-space maps
                              //    ride on a 5-D tensor map of the output (the TMA unit does the scatter)
    int epi_bytes;            // shared memory per epilogue warp
    int a_stage, b_stage;     // bytes per pipeline stage of the A / B operand rings
    long long* trace;         // debug (MSU_TC_TRACE=1): clock64 stamps [cta][tile < 8][16 events]
    int dbg_skip;             // debug (MSU_CONV_SKIP bit mask, pair conv only): 1 no weight loads, 2 no halo loads, 4 no output stores after a CTA's first tile
    MsuEpilogue E;
};
#define TC_TRACE(ev) do { if (p.trace != nullptr && lane == 0 && it < 8) p.trace[((size_t)blockIdx.x * 8 + it) * 16 + (ev)] = clock64(); } while (0)

constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
// epilogue warps (multiple of 4: warp % 4 selects the TMEM lane quadrant).  The kernel is instantiated per epilogue path so that
// each gets its own register budget: the TMA-slab path runs 16 warps (its ALU-heavy GELU / GELU' epilogues want issue slots:
// fc1 stage 0 115 -> 103 us, head expand 440 -> 401 us), the generic scatter path 12 (more registers per thread).
constexpr int TC_EPW_TMA = 16, TC_EPW_GEN = 12;
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_CW = 32;                          // epilogue chunk width (columns)
constexpr int TC_CPITCH_B = (TC_CW + 8) * 2;       // 80 B scratch row pitch: 16 B accesses of 32 lanes are conflict free
constexpr int TC_SCRATCH_BYTES = 32 * TC_CPITCH_B; // per epilogue warp

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 t = __bfloat1622float2(p[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// GELU(x) = x * Phi(x) for the bf16 epilogues.  Phi(x) = 0.5 (1 + erf(x / sqrt 2)) is evaluated as
// 0.5 + 0.5 tanh(x (c0 + c1 x^2 + c2 x^4)) with minimax-fitted c (|Phi error| <= 5.1e-5 before the MUFU.TANH error of
// 2^-11 relative, i.e. <= 3e-4 absolute on Phi: below the bf16 rounding of an O(1) output).  8 issue slots per element
// with ONE transcendental: ncu showed the XU pipe (MUFU + F2F conversions, 16 lanes/clk/SM) as the top pipe of the
// short-K GEMMs at 3 XU ops per element (ex2, rcp, F2F.BF16), so the epilogues use tanh and never F2F.
// The fp32 parity mode keeps erff (common.cuh).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float phi_cdf_fast(float x, float& x2c) {
    x2c = fminf(x * x, 50.0f);                     // beyond |x| = 7.07 Phi is 0 / 1 to fp32 precision
    float p = fmaf(-3.710398891e-04f, x2c, 3.708714633e-02f);
    p = fmaf(p, x2c, 7.975576149e-01f);
    return fmaf(0.5f, tanh_approx(x * p), 0.5f);
}
__device__ __forceinline__ float gelu_fast(float x) {
    float x2c;
    return x * phi_cdf_fast(x, x2c);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
    float x2c;
    const float cdf = phi_cdf_fast(x, x2c);
    const float e = ex2_approx(fmaf(x2c, -0.72134752044448170368f, -1.32574806473616222f));   // phi(x) = exp(-x^2 / 2) / sqrt(2 pi)
    return fmaf(x, e, cdf);
}
// GELU(x) and GELU'(x) together (they share Phi): the MLP forward stores the derivative instead of the pre-activation
// (MsuEpilogue.act = 2), so the backward epilogue is a plain multiply (act = 3) instead of 15 instructions per element
__device__ __forceinline__ float gelu_and_grad_fast(float x, float& grad) {
    float x2c;
    const float cdf = phi_cdf_fast(x, x2c);
    const float e = ex2_approx(fmaf(x2c, -0.72134752044448170368f, -1.32574806473616222f));   // phi(x) = exp(-x^2 / 2) / sqrt(2 pi)
    grad = fmaf(x, e, cdf);
    return x * cdf;
}
// the two bf16 halves of a packed word as fp32 (ALU shifts / masks; no XU conversion)
__device__ __forceinline__ float bf16lo_f(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi_f(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// K-major, 64B-swizzled descriptor (rows of 32 bf16 = 64 B, 8-row groups 512 B apart)
__device__ __forceinline__ uint64_t make_desc_kmajor_sw64_g(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
    return d;
}
constexpr int TC_HALO32_BYTES = 8704;       // 130 pixels x 64 B (8320) rounded up to the 512 B swizzle period

// first logical row of accumulator `r` of m-tile `mt` (conv modes: tiles are image-row segments)
__device__ __forceinline__ int64_t tile_row0(const TcParams& p, int mt, int r) {
    if (p.mode < 2) return (int64_t)mt * TC_BM;
    const int per_img = p.tiles_x * p.tiles_y;
    const int cb = mt / per_img, rr = mt % per_img;
    return ((int64_t)cb * p.H + (rr / p.tiles_x) * p.bmh + r) * p.W + (rr % p.tiles_x) * p.bmw;
}

// 16 B chunk g of row `row` inside a [32 rows x 64 B] SWIZZLE_64B slab (1 KB aligned): chunk ^= (row / 2) % 4
__device__ __forceinline__ uint32_t slab_off(int row, int g) { return (uint32_t)(row * 64 + ((g ^ ((row >> 1) & 3)) << 4)); }

// tile -> (m tile, n tile).  CTA pairs of the plain / dual-source modes: the two CTAs of a cluster (consecutive tile numbers) take the
// two 128-row halves of one 256-row M tile and the SAME n tile (they share its B rows); n tiles fastest among the pair tiles.
template <bool PAIR>
__device__ __forceinline__ void tile_mn(const TcParams& p, int tile, int& mt, int& nt) {
    if (PAIR && p.mode < 2) {
        const int q = tile >> 1;
        nt = q % p.num_n_tiles;
        mt = (q / p.num_n_tiles) * 2 + (tile & 1);
    } else {
        mt = tile / p.num_n_tiles;
        nt = tile % p.num_n_tiles;
    }
}

template <bool EPI_TMA, int TC_EPI_WARPS, bool PAIR = false, bool LND = false>
__global__ void __launch_bounds__(32 * (2 + TC_EPI_WARPS), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmAux, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int B_BYTES = p.BN * TC_BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)p.stages * p.a_stage;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * p.b_stage);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;   // [2]
    uint64_t* tempty = tfull + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint64_t* auxbar = reinterpret_cast<uint64_t*>(tmem_slot + 4);   // [TC_EPI_WARPS][2] aux-in slab landed (TMA path)
    uint8_t* sScratch = smem + (size_t)p.stages * (p.a_stage + p.b_stage) + 1024;   // [TC_EPI_WARPS][epi_bytes], 1 KB aligned
    float* sGw = reinterpret_cast<float*>(sScratch + (size_t)TC_EPI_WARPS * p.epi_bytes);   // LND: gamma * w [N] (+ 1 KB of shared memory)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int KB = p.kb1 + p.kb2;
    // PAIR (mode 4 only): the two CTAs of a cluster take adjacent tiles (same trip count: tile and grid counts are even); the even
    // CTA issues the M = 256 MMAs for both, every barrier the MMA thread waits on lives in its shared memory
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (p.kb2 > 0) prefetch_tmap(&tmA2);
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], PAIR ? 2 * TC_EPI_WARPS : TC_EPI_WARPS); }
        for (int a = 0; a < 2 * TC_EPI_WARPS; a++) mbar_init(&auxbar[a], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (EPI_TMA && warp == 2 && lane == 0) {
        if (p.E.C != nullptr) prefetch_tmap(&tmC);
        if (p.E.Cpre != nullptr || p.E.R != nullptr || p.E.H != nullptr) prefetch_tmap(&tmAux);
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    // PDL: everything above overlapped the previous kernel's tail; nothing below may run before that kernel has completed
    pdl_trigger();
    pdl_wait();
    if (LND) {
        for (int n = threadIdx.x; n < p.N; n += blockDim.x) sGw[n] = p.E.lnd_gamma[n] * p.E.lnd_w[n];
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
                int mt, nt;
            tile_mn<PAIR>(p, tile, mt, nt);
                int cb = 0, cy = 0, cx = 0;
                TC_TRACE(0);
                if (p.mode >= 2) {
                    const int per_img = p.tiles_x * p.tiles_y;
                    cb = mt / per_img;
                    const int r = mt % per_img;
                    cy = (r / p.tiles_x) * p.bmh;
                    cx = (r % p.tiles_x) * p.bmw;
                }
                for (int kb = 0; kb < KB; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* a_dst = sA + (size_t)stage * p.a_stage;
                    uint8_t* b_dst = sB + (size_t)stage * p.b_stage;
                    if (p.mode == 4) {
                        // two image rows per tile share the three weight tiles of (dy, 32-channel block): halo rows of
                        // 130 pixels x 32 channels (64B swizzle), K = 9C walked in exact 32-channel steps (no zero padding)
                        const int dyi = kb / p.cblocks, c0 = (kb % p.cblocks) * 32;
                        if (PAIR) {
                            // own two halo rows + own HALF of the three weight tiles (rows [BN/2 rank, +BN/2) of each); the bytes
                            // of both CTAs are counted by the leader's barrier
                            const int hb = (p.BN / 2) * 64;
                            if (p.box2 && p.dbg_skip != 0 && it > 0) {   // timing experiments only (MSU_CONV_SKIP): results are wrong
                                const bool la = !(p.dbg_skip & 2), lb = !(p.dbg_skip & 1);
                                if (leader && (la || lb)) mbar_arrive_expect_tx(&full[stage], 2 * ((la ? 2 * 130 * 64 : 0) + (lb ? 3 * hb : 0)));
                                if (la) tma_load_4d_pair(a_dst, &tmA, &full[stage], c0, cx - 1, cy + dyi - 1, cb);
                                if (lb) tma_load_3d_pair(b_dst, &tmB, &full[stage], dyi * 3 * p.C + c0, nt * p.BN + (int)cta_rank * (p.BN / 2), 0);
                                if (!la && !lb && leader) mbar_arrive(&full[stage]);
                                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                                continue;
                            }
                            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (2 * 130 * 64 + 3 * hb));
                            if (p.box2) {   // two boxes per K block: both halo rows, all three taps' half tiles
                                tma_load_4d_pair(a_dst, &tmA, &full[stage], c0, cx - 1, cy + dyi - 1, cb);
                                tma_load_3d_pair(b_dst, &tmB, &full[stage], dyi * 3 * p.C + c0, nt * p.BN + (int)cta_rank * (p.BN / 2), 0);
                                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                                continue;
                            }
                            tma_load_4d_pair(a_dst, &tmA, &full[stage], c0, cx - 1, cy + dyi - 1, cb);
                            tma_load_4d_pair(a_dst + TC_HALO32_BYTES, &tmA, &full[stage], c0, cx - 1, cy + dyi, cb);
                            for (int dx = 0; dx < 3; dx++)
                                tma_load_2d_pair(b_dst + dx * hb, &tmB, &full[stage], (dyi * 3 + dx) * p.C + c0, nt * p.BN + (int)cta_rank * (p.BN / 2));
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        const int bb = p.BN * 64;
                        mbar_arrive_expect_tx(&full[stage], 2 * 130 * 64 + 3 * bb);
                        if (p.box2) {
                            tma_load_4d(a_dst, &tmA, &full[stage], c0, cx - 1, cy + dyi - 1, cb);            // rows y+dy-1, y+dy
                            tma_load_3d(b_dst, &tmB, &full[stage], dyi * 3 * p.C + c0, nt * p.BN, 0);        // taps dx = 0, 1, 2
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        tma_load_4d(a_dst, &tmA, &full[stage], c0, cx - 1, cy + dyi - 1, cb);
                        tma_load_4d(a_dst + TC_HALO32_BYTES, &tmA, &full[stage], c0, cx - 1, cy + dyi, cb);
                        for (int dx = 0; dx < 3; dx++)
                            tma_load_2d(b_dst + dx * bb, &tmB, &full[stage], (dyi * 3 + dx) * p.C + c0, nt * p.BN);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (p.mode == 3) {
                        // halo row: 130 pixels x 64 channels once per (dy, channel block); the three dx taps are
                        // row-shifted views of it.  Three weight tiles ride in the same stage.
                        const int dyi = kb / p.cblocks, c0 = (kb % p.cblocks) * TC_BK;
                        mbar_arrive_expect_tx(&full[stage], 130 * 128 + 3 * B_BYTES);
                        tma_load_4d(a_dst, &tmA, &full[stage], c0, cx - 1, cy + dyi - 1, cb);
                        for (int dx = 0; dx < 3; dx++)
                            tma_load_2d(b_dst + dx * B_BYTES, &tmB, &full[stage], (dyi * 3 + dx) * p.C + c0, nt * p.BN);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (PAIR) {     // plain / dual-source: own 128 rows of A, own half of the B tile's rows; bytes counted by the leader
                        const int hb = (p.BN / 2) * TC_BK * 2;
                        if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (TC_A_BYTES + hb));
                        int bk;
                        if (kb < p.kb1) {
                            tma_load_2d_pair(a_dst, &tmA, &full[stage], kb * TC_BK, mt * TC_BM);
                            bk = kb * TC_BK;
                        } else {
                            tma_load_2d_pair(a_dst, &tmA2, &full[stage], (kb - p.kb1) * TC_BK, mt * TC_BM);
                            bk = p.k_split + (kb - p.kb1) * TC_BK;
                        }
                        tma_load_2d_pair(b_dst, &tmB, &full[stage], bk, nt * p.BN + (int)cta_rank * (p.BN / 2));
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full[stage], TC_A_BYTES + B_BYTES);
                    int bk;
                    if (p.mode == 2) {
                        const int tap = kb / p.cblocks, c0 = (kb % p.cblocks) * TC_BK;
                        tma_load_4d(a_dst, &tmA, &full[stage], c0, cx + tap % 3 - 1, cy + tap / 3 - 1, cb);
                        bk = tap * p.C + c0;
                    } else if (kb < p.kb1) {
                        tma_load_2d(a_dst, &tmA, &full[stage], kb * TC_BK, mt * TC_BM);
                        bk = kb * TC_BK;
                    } else {
                        tma_load_2d(a_dst, &tmA2, &full[stage], (kb - p.kb1) * TC_BK, mt * TC_BM);
                        bk = p.k_split + (kb - p.kb1) * TC_BK;
                    }
                    tma_load_2d(b_dst, &tmB, &full[stage], bk, nt * p.BN);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                TC_TRACE(1);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread; the leader CTA's in a pair) =====================
        if (leader && elect_one()) {
            const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * TC_BM : TC_BM, p.BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                TC_TRACE(2);
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < KB; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (kb == 0) TC_TRACE(3);
                    const uint32_t a_addr = smem_u32(sA + (size_t)stage * p.a_stage);
                    const uint32_t b_addr = smem_u32(sB + (size_t)stage * p.b_stage);
                    if (p.mode == 4) {
                        // one descriptor pair per stage, advanced through the 14-bit address field (the issuing thread is
                        // close to the critical path at 12 MMAs per K block): pixel shift dx = +64 B = +4 (the swizzle
                        // follows the address bits), second image row = +TC_HALO32_BYTES, weight tile of tap dx = +BN*64 B
                        const uint64_t ad = make_desc_kmajor_sw64_g(a_addr), bd = make_desc_kmajor_sw64_g(b_addr);
                        const uint32_t bstep = (uint32_t)((PAIR ? p.BN / 2 : p.BN) * 4);         // a CTA of a pair holds half of a weight tile's rows
                        const uint32_t rstep = p.box2 ? (130 * 64) >> 4 : TC_HALO32_BYTES >> 4;   // one box: the rows are contiguous
                        const uint32_t acc0 = kb != 0;
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            const uint64_t ar = ad + (uint64_t)(r * rstep);
                            const uint32_t d = d_tmem + r * 128;
                            if (PAIR) {
                                tc_mma_bf16_pair(d, ar, bd, idesc, acc0);
                                tc_mma_bf16_pair(d, ar + 2, bd + 2, idesc, 1);
                                tc_mma_bf16_pair(d, ar + 4, bd + bstep, idesc, 1);
                                tc_mma_bf16_pair(d, ar + 6, bd + bstep + 2, idesc, 1);
                                tc_mma_bf16_pair(d, ar + 8, bd + 2 * bstep, idesc, 1);
                                tc_mma_bf16_pair(d, ar + 10, bd + 2 * bstep + 2, idesc, 1);
                            } else {
                                tc_mma_bf16(d, ar, bd, idesc, acc0);
                                tc_mma_bf16(d, ar + 2, bd + 2, idesc, 1);
                                tc_mma_bf16(d, ar + 4, bd + bstep, idesc, 1);
                                tc_mma_bf16(d, ar + 6, bd + bstep + 2, idesc, 1);
                                tc_mma_bf16(d, ar + 8, bd + 2 * bstep, idesc, 1);
                                tc_mma_bf16(d, ar + 10, bd + 2 * bstep + 2, idesc, 1);
                            }
                        }
                    } else if (p.mode == 3) {
                        for (int dx = 0; dx < 3; dx++) {
                            // A rows shifted by dx pixels = start address + dx*128 B.  The 128B swizzle is a function of
                            // the shared-memory address bits (measured: base-offset field must stay 0), so a row-shifted
                            // view of the TMA-written halo tile is read back consistently.
                            const uint64_t adesc = make_desc_kmajor_sw128(a_addr + dx * 128);
                            const uint64_t bdesc = make_desc_kmajor_sw128(b_addr + dx * B_BYTES);
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; k++)
                                tc_mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | dx | k) != 0);
                        }
                    } else {
                        const uint64_t adesc = make_desc_kmajor_sw128(a_addr);
                        const uint64_t bdesc = make_desc_kmajor_sw128(b_addr);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; k++) {   // +32 B per K=16 step inside the 128 B swizzle row
                            if (PAIR) tc_mma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                            else tc_mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                    }
                    if (PAIR) tc_commit_pair(&empty[stage]); else tc_commit(&empty[stage]);   // frees the smem slot when these MMAs retire
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (PAIR) tc_commit_pair(&tfull[acc]); else tc_commit(&tfull[acc]);           // accumulator ready for the epilogue
                TC_TRACE(4);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if constexpr (EPI_TMA) {
        // ===================== epilogue warps, unmapped output: TMEM -> fused epilogue in registers -> swizzled slab -> TMA store ==========
        // Each lane owns one accumulator row (32 columns per chunk).  The residual / GELU' operand arrives as a TMA-loaded
        // [32 x 32] slab one chunk ahead; the result (and the pre-activation copy) leave as TMA box stores, so there is no
        // per-row address arithmetic and no transposition.
        const MsuEpilogue& E = p.E;
        const int quad = warp & 3;
        const int sub = (warp - 2) >> 2;
        constexpr int NSUB = TC_EPI_WARPS / 4;
        const int nchunks = p.BN / TC_CW;               // BN is a multiple of 32 on this path
        uint8_t* slab_out = sScratch + (size_t)(warp - 2) * p.epi_bytes;
        uint8_t* slab_aux = slab_out + 2048;            // Cpre staging (out) or R/H operand (in, 2 slabs)
        uint64_t* abar = auxbar + 2 * (warp - 2);
        const bool aux_in = (E.R != nullptr || E.H != nullptr);
        constexpr bool lnd = LND;            // fused head LayerNorm + dot: its own instantiation (registers of the common epilogues)
        float lnd_g = 0.f, lnd_bw = 0.f;     // G = sum_c gamma_c w_c, Bw = sum_c beta_c w_c (the same in every lane)
        if (lnd) {
            for (int n = 0; n < p.N; n++) {
                lnd_g += sGw[n];
                lnd_bw = fmaf(E.lnd_beta[n], E.lnd_w[n], lnd_bw);
            }
        }
        uint32_t aux_phase[2] = {0, 0};
        int aux_buf = 0;
        int acc = 0; uint32_t acc_phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            int mt, nt;
            tile_mn<PAIR>(p, tile, mt, nt);
            // chunk list of this warp: cc = r * nchunks + c over the tile's accumulators r (consecutive row segments).
            // Fused LayerNorm + dot (lnd): a warp owns ALL chunks of one 32-row block (its lanes carry the rows' statistics from
            // chunk to chunk); the blocks of consecutive tiles rotate over the NSUB warps of the quadrant.
            const int nch_tot = p.nacc * nchunks;
            int c_first = (sub + it) % NSUB, c_end = nch_tot, c_step = NSUB;
            if (lnd) {
                const int r_mine = (sub - (it * p.nacc) % NSUB + NSUB) % NSUB;
                c_first = r_mine < p.nacc ? r_mine * nchunks : nch_tot;
                c_end = r_mine < p.nacc ? c_first + nchunks : nch_tot;
                c_step = 1;
            }
            float st_c0 = 0.f, st_s = 0.f, st_q = 0.f, st_d = 0.f;
            auto row0_of = [&](int cc) { return (int)(tile_row0(p, mt, cc / nchunks) + quad * 32); };
            if (aux_in && c_first < nch_tot && lane == 0) {   // operand slab of the first chunk
                mbar_arrive_expect_tx(&abar[aux_buf], 2048);
                tma_load_2d(slab_aux + aux_buf * 2048, &tmAux, &abar[aux_buf], nt * p.BN + (c_first % nchunks) * TC_CW, row0_of(c_first));
            }
            if (warp == 2) TC_TRACE(5);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            if (warp == 2) TC_TRACE(6);
            const uint32_t t_base = tmem_base + acc * 256 + ((uint32_t)(quad * 32) << 16);
            bool released = false;
            bool first = true;
            for (int cc = c_first; cc < c_end; cc += c_step) {
                const int r = cc / nchunks, c = cc - r * nchunks;
                uint32_t raw[TC_CW];
                tc_ld32_nowait(t_base + r * 128 + c * TC_CW, raw);
                const int n0 = nt * p.BN + c * TC_CW;
                const int row0 = row0_of(cc);
                float rs = 1.0f;
                if (E.rowscale != nullptr) {
                    const int64_t m_own = (int64_t)row0 + lane;
                    rs = m_own < p.M ? E.rowscale[m_own / E.rows_per_sample] : 0.0f;
                }
                if (aux_in && cc + c_step < c_end && lane == 0) {   // next chunk's operand slab (its buffer was drained a chunk ago)
                    const int cn = cc + c_step;
                    mbar_arrive_expect_tx(&abar[aux_buf ^ 1], 2048);
                    tma_load_2d(slab_aux + (aux_buf ^ 1) * 2048, &tmAux, &abar[aux_buf ^ 1], nt * p.BN + (cn % nchunks) * TC_CW, row0_of(cn));
                }
                tc_ld_wait();
                if (warp == 2 && first) TC_TRACE(10);
                if (cc + c_step >= c_end) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
                    released = true;
                }
                float v[TC_CW];
#pragma unroll
                for (int i = 0; i < TC_CW; i++) v[i] = __uint_as_float(raw[i]);
                if (E.bias != nullptr) {
#pragma unroll
                    for (int g4 = 0; g4 < TC_CW / 4; g4++) {
                        if (n0 + g4 * 4 < p.N) {
                            const float4 b0 = *reinterpret_cast<const float4*>(E.bias + n0 + g4 * 4);
                            v[g4 * 4] += b0.x; v[g4 * 4 + 1] += b0.y; v[g4 * 4 + 2] += b0.z; v[g4 * 4 + 3] += b0.w;
                        }
                    }
                }
                // the previous chunk's stores must have finished reading the slabs before they are overwritten
                if (lane == 0) tma_store_wait_read0();
                __syncwarp();
                if (E.Cpre != nullptr) {
                    // the activation sees the bf16-rounded pre-activation, exactly like a separate GELU pass over Cpre
                    // (rounded by the pack itself and unpacked with shifts: F2F.BF16 would run on the XU pipe)
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            if (E.act == 2) {      // Cpre receives GELU'(pre) instead of pre: nothing stores the pre-activation, so GELU
                                float g0, g1;      // and GELU' both see it unrounded (3 instructions per pair less than the round trip)
                                v[g * 8 + 2 * i] = gelu_and_grad_fast(v[g * 8 + 2 * i], g0);
                                v[g * 8 + 2 * i + 1] = gelu_and_grad_fast(v[g * 8 + 2 * i + 1], g1);
                                pk[i] = pack_bf16x2(g0, g1);
                            } else {
                                pk[i] = pack_bf16x2(v[g * 8 + 2 * i], v[g * 8 + 2 * i + 1]);
                                v[g * 8 + 2 * i] = bf16lo_f(pk[i]);
                                v[g * 8 + 2 * i + 1] = bf16hi_f(pk[i]);
                            }
                        }
                        *reinterpret_cast<uint4*>(slab_aux + slab_off(lane, g)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                if (E.act == 1) {
#pragma unroll
                    for (int i = 0; i < TC_CW; i++) v[i] = gelu_fast(v[i]);
                }
                if (aux_in) {
                    mbar_wait(&abar[aux_buf], aux_phase[aux_buf]);
                    aux_phase[aux_buf] ^= 1;
                    const uint8_t* ab = slab_aux + aux_buf * 2048;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        float f[8];
                        unpack8(*reinterpret_cast<const uint4*>(ab + slab_off(lane, g)), f);
                        if (E.H != nullptr) {
                            if (E.act == 3) {      // H already holds GELU'(pre)
#pragma unroll
                                for (int i = 0; i < 8; i++) v[g * 8 + i] *= f[i];
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; i++) v[g * 8 + i] *= gelu_grad_fast(f[i]);
                            }
                            if (E.rowscale != nullptr) {
#pragma unroll
                                for (int i = 0; i < 8; i++) v[g * 8 + i] *= rs;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; i++) v[g * 8 + i] = fmaf(v[g * 8 + i], rs, f[i]);
                        }
                    }
                    aux_buf ^= 1;
                } else if (E.rowscale != nullptr) {
#pragma unroll
                    for (int i = 0; i < TC_CW; i++) v[i] *= rs;
                }
                if (lnd) {
                    // head LayerNorm + 1x1 conv on the rows as they are stored (bf16-rounded): shifted one-pass statistics
                    // (shift = the row's first value) and sum_c v_c gw_c with gw = gamma w from shared memory (broadcast reads)
                    if (c == 0) { st_s = st_q = st_d = 0.f; }
#pragma unroll
                    for (int g4 = 0; g4 < TC_CW / 4; g4++) {
                        const float4 gq = *reinterpret_cast<const float4*>(sGw + n0 + g4 * 4);
                        const float gw[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
                        for (int i = 0; i < 4; i += 2) {
                            const uint32_t u = pack_bf16x2(v[g4 * 4 + i], v[g4 * 4 + i + 1]);
                            const float a0 = bf16lo_f(u), a1 = bf16hi_f(u);
                            if (c == 0 && g4 == 0 && i == 0) st_c0 = a0;
                            const float d0 = a0 - st_c0, d1 = a1 - st_c0;
                            st_s += d0 + d1;
                            st_q = fmaf(d0, d0, fmaf(d1, d1, st_q));
                            st_d = fmaf(a0, gw[i], fmaf(a1, gw[i + 1], st_d));
                        }
                    }
                    if (c == nchunks - 1) {
                        const float invn = 1.0f / (float)p.N;
                        const float ms = st_s * invn;
                        const float mu = st_c0 + ms;
                        const float rs = rsqrtf(fmaxf(fmaf(-ms, ms, st_q * invn), 0.f) + 1e-5f);
                        const float dc = fmaf(-mu, lnd_g, st_d) * rs;                // sum_c gw_c x-hat_c
                        const int64_t m_own = (int64_t)row0 + lane;
                        if (m_own < p.M) {
                            reinterpret_cast<__nv_bfloat16*>(E.lnd_logits)[m_own] = __float2bfloat16(dc + lnd_bw);
                            E.lnd_mean[m_own] = mu;
                            E.lnd_rstd[m_own] = rs;
                            E.lnd_m2[m_own] = dc * invn;
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; g++)
                    *reinterpret_cast<uint4*>(slab_out + slab_off(lane, g)) =
                        make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                                   pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
                fence_proxy_async_smem();
                __syncwarp();
                if (warp == 2 && first) TC_TRACE(8);
                if (lane == 0 && E.C != nullptr && !((p.dbg_skip & 4) && it > 0)) {
                    if (E.map == MSU_MAP_NONE) {
                        tma_store_2d(slab_out, &tmC, n0, row0);
                        if (E.Cpre != nullptr) tma_store_2d(slab_aux, &tmAux, n0, row0);
                    } else if (E.map == MSU_MAP_SHUFFLE) {
                        // rows = 32 tokens (b, h, w0..w0+31), columns = 32 channels of one (p1, p2): box [32 c, 1, 32 w, 1, 1]
                        const int Wt = E.geo[1], pp = E.geo[2], cc = E.geo[3];
                        const int q = n0 / cc, c0 = n0 - q * cc, p1 = q / pp, p2 = q - p1 * pp;
                        const int bh = row0 / Wt, w0 = row0 - bh * Wt;
                        tma_store_5d(slab_out, &tmC, c0, p2, w0, p1, bh);
                        if (E.Cpre != nullptr) tma_store_5d(slab_aux, &tmAux, c0, p2, w0, p1, bh);
                    } else {
                        // UNSHUFFLE: rows = 32 pixels (b, y, x0..x0+31) of the shuffled map, box [32 c, p (p2), 32/p (w), 1, 1]
                        const int Wt = E.geo[1], pp = E.geo[2];
                        const int Wp = Wt * pp;
                        const int by = row0 / Wp, x0 = row0 - by * Wp;          // by = b * (H p) + y
                        const int Hp = E.geo[0] * pp;
                        const int b = by / Hp, y = by - b * Hp;
                        tma_store_5d(slab_out, &tmC, n0, 0, x0 / pp, y % pp, b * E.geo[0] + y / pp);
                    }
                    tma_store_commit();
                }
                if (warp == 2 && first) TC_TRACE(9);
                first = false;
            }
            if (!released) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
            }
            if (warp == 2) TC_TRACE(7);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all();               // global writes complete before the CTA retires its shared memory
    } else {
        // ===================== epilogue warps: TMEM -> (+bias) -> private transpose -> fused epilogue -> global =====================
        const MsuEpilogue& E = p.E;
        __nv_bfloat16* Cp = reinterpret_cast<__nv_bfloat16*>(E.C);
        __nv_bfloat16* Cpre = reinterpret_cast<__nv_bfloat16*>(E.Cpre);
        const __nv_bfloat16* Rp = reinterpret_cast<const __nv_bfloat16*>(E.R);
        const __nv_bfloat16* Hp = reinterpret_cast<const __nv_bfloat16*>(E.H);
        const int quad = warp & 3;                      // TMEM lane quadrant this warp may read
        const int sub = (warp - 2) >> 2;                // index among the warps of the quadrant
        constexpr int NSUB = TC_EPI_WARPS / 4;
        const int nchunks = (p.BN + TC_CW - 1) / TC_CW;
        const int rsub = lane >> 2, q = lane & 3;       // store phase: rows rsub + 8 j (j < 4), 16 B quarter q of the 64 B row
        uint8_t* scr = sScratch + (size_t)(warp - 2) * TC_SCRATCH_BYTES;
        const bool passthrough = ((E.act == 0 || E.act == 3) && Hp == nullptr && E.rowscale == nullptr && Rp == nullptr);
        int acc = 0; uint32_t acc_phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            int mt, nt;
            tile_mn<PAIR>(p, tile, mt, nt);
            bool released = false;
            bool first = true;
            for (int racc = 0; racc < p.nacc; racc++) {
            // ---- output row of accumulator lane (quad, lane), computed once per (tile, accumulator) by its owner lane
            int64_t m_own, ro_own = -1;
            int coff_own = 0;
            float rs_own = 1.0f;
            {
                const int rl = quad * 32 + lane;
                if (p.mode == 2) {
                    const int per_img = p.tiles_x * p.tiles_y;
                    const int cb = mt / per_img, rr = mt % per_img;
                    const int y = (rr / p.tiles_x) * p.bmh + rl / p.bmw, x = (rr % p.tiles_x) * p.bmw + rl % p.bmw;
                    m_own = ((int64_t)cb * p.H + y) * p.W + x;
                } else {
                    m_own = tile_row0(p, mt, racc) + rl;
                }
                if (m_own < p.M) {
                    if (E.map == MSU_MAP_WINDOW) {
                        ro_own = win_to_pix(make_wingeo(E.geo), m_own);
                    } else if (E.map == MSU_MAP_UNSHUFFLE) {
                        const RowCol rc = map_rc(MSU_MAP_UNSHUFFLE, E.geo, m_own, 0);
                        ro_own = rc.row;
                        coff_own = rc.col;
                    } else if (E.map == MSU_MAP_SHUFFLE) {   // base row; (p1, p2) offsets are added per column chunk
                        const int hw = E.geo[0] * E.geo[1], pp = E.geo[2];
                        const int64_t b = m_own / hw;
                        const int t = (int)(m_own - b * hw);
                        const int hh = t / E.geo[1], ww = t - hh * E.geo[1];
                        ro_own = (b * (E.geo[0] * pp) + hh * pp) * (int64_t)(E.geo[1] * pp) + ww * pp;
                    } else {
                        ro_own = m_own;
                    }
                }
                if (E.rowscale != nullptr && ro_own >= 0) rs_own = E.rowscale[ro_own / E.rows_per_sample];   // shuffle offsets stay inside the sample
            }
            // rows this lane stores: hand the owner lanes' maps over
            int64_t m_r[4], ro_r[4];
            int coff_r[4];
            float rs_r[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int src = rsub + 8 * j;
                m_r[j] = __shfl_sync(0xffffffffu, m_own, src);
                ro_r[j] = __shfl_sync(0xffffffffu, ro_own, src);
                coff_r[j] = __shfl_sync(0xffffffffu, coff_own, src);
                rs_r[j] = __shfl_sync(0xffffffffu, rs_own, src);
            }
            if (racc == 0) {
                if (warp == 2) TC_TRACE(5);
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                if (warp == 2) TC_TRACE(6);
            }
            const uint32_t t_base = tmem_base + acc * 256 + racc * 128 + ((uint32_t)(quad * 32) << 16);
            for (int c = (sub + it) % NSUB; c < nchunks; c += NSUB) {
                const int cols = p.BN - c * TC_CW >= TC_CW ? TC_CW : 16;   // BN is a multiple of 16
                uint32_t raw[TC_CW];
                if (cols == TC_CW) tc_ld32_nowait(t_base + c * TC_CW, raw);
                else tc_ld16_nowait(t_base + c * TC_CW, raw);
                // ---- addresses and operand prefetch of the 4 (row, quarter) vectors this lane stores
                const int n0 = nt * p.BN + c * TC_CW;
                const int n = n0 + q * 8;
                int64_t o[4];
                uint4 hraw[4], rraw[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    o[j] = -1;
                    if (ro_r[j] < 0 || n >= p.N || q * 8 >= cols) continue;
                    int64_t rw = ro_r[j];
                    int co = n + coff_r[j];
                    if (E.map == MSU_MAP_SHUFFLE) {
                        const int pp = E.geo[2], cc = E.geo[3];
                        const int qq = n / cc, p1 = qq / pp, p2 = qq - p1 * pp;
                        rw += (int64_t)p1 * (E.geo[1] * pp) + p2;
                        co = n - qq * cc;
                    }
                    o[j] = rw * E.ldc + co;
                    if (Hp != nullptr) hraw[j] = *reinterpret_cast<const uint4*>(Hp + m_r[j] * E.ldh + n);
                    if (Rp != nullptr) rraw[j] = *reinterpret_cast<const uint4*>(Rp + rw * E.ldr + co);
                }
                tc_ld_wait();
                if (warp == 2 && first) TC_TRACE(10);
                if (c + NSUB >= nchunks && racc == p.nacc - 1) {   // this warp's last read of the accumulators: hand TMEM back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
                    released = true;
                }
                // ---- own accumulator row -> bf16 -> scratch (row = lane)
#pragma unroll
                for (int g8 = 0; g8 < TC_CW / 8; g8++) {
                    if (g8 * 8 < cols) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) v[i] = __uint_as_float(raw[g8 * 8 + i]);
                        if (E.bias != nullptr && n0 + g8 * 8 < p.N) {
                            const float4 b0 = *reinterpret_cast<const float4*>(E.bias + n0 + g8 * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(E.bias + n0 + g8 * 8 + 4);
                            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                        }
                        *reinterpret_cast<uint4*>(scr + lane * TC_CPITCH_B + g8 * 16) =
                            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    }
                }
                __syncwarp();
                if (warp == 2 && first) TC_TRACE(8);
                // ---- transposed read: 8 rows x 64 B per instruction -> fused epilogue -> coalesced stores
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (o[j] < 0) continue;
                    const uint4 tv = *reinterpret_cast<const uint4*>(scr + (rsub + 8 * j) * TC_CPITCH_B + q * 16);
                    if (Cpre != nullptr && E.act != 2) *reinterpret_cast<uint4*>(Cpre + o[j]) = tv;
                    if (passthrough) {
                        *reinterpret_cast<uint4*>(Cp + o[j]) = tv;
                        continue;
                    }
                    float w[8];
                    unpack8(tv, w);
                    if (E.act == 1) {
#pragma unroll
                        for (int i = 0; i < 8; i++) w[i] = gelu_fast(w[i]);
                    } else if (E.act == 2) {       // Cpre receives GELU'(pre) instead of pre
                        float gq[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) w[i] = gelu_and_grad_fast(w[i], gq[i]);
                        if (Cpre != nullptr)
                            *reinterpret_cast<uint4*>(Cpre + o[j]) = make_uint4(pack_bf16x2(gq[0], gq[1]), pack_bf16x2(gq[2], gq[3]),
                                                                                pack_bf16x2(gq[4], gq[5]), pack_bf16x2(gq[6], gq[7]));
                    }
                    if (Hp != nullptr) {
                        float hf[8];
                        unpack8(hraw[j], hf);
                        if (E.act == 3) {          // H already holds GELU'(pre)
#pragma unroll
                            for (int i = 0; i < 8; i++) w[i] *= hf[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; i++) w[i] *= gelu_grad_fast(hf[i]);
                        }
                    }
                    if (E.rowscale != nullptr) {
#pragma unroll
                        for (int i = 0; i < 8; i++) w[i] *= rs_r[j];
                    }
                    if (Rp != nullptr) {
                        float rf[8];
                        unpack8(rraw[j], rf);
#pragma unroll
                        for (int i = 0; i < 8; i++) w[i] += rf[i];
                    }
                    *reinterpret_cast<uint4*>(Cp + o[j]) = make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]),
                                                                      pack_bf16x2(w[4], w[5]), pack_bf16x2(w[6], w[7]));
                }
                __syncwarp();                              // scratch is rewritten by the next chunk
                if (warp == 2 && first) TC_TRACE(9);
                first = false;
            }
            }   // accumulators of the tile
            if (!released) {                               // no chunk of the last accumulator fell to this warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
            }
            if (warp == 2) TC_TRACE(7);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

// 2-D bf16 map over [rows, cols] (cols contiguous, row pitch ld elements), box = [box_rows, 64], 128B swizzle.
static bool make_map_2d(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// 4-D bf16 map over NHWC [B, H, W, C]; box = [64 ch, bmw, bmh, 1]
static bool make_map_nhwc(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int bmw, int bmh) {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)bmw, (cuuint32_t)bmh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [rows, cols] bf16 output / epilogue operand, box = [32 rows, 32 cols] (one epilogue warp's chunk), 64B swizzle
static bool make_map_slab(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_CW, 32};
    cuuint32_t estr[2] = {1, 1};
    return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int pick_bn(int64_t N, int cap = 256, int step = 16) {
    // largest tile <= cap (multiple of step) that wastes the least of the last N tile
    if (N <= cap) return (int)((N + step - 1) / step * step);
    int best = cap;
    int64_t best_waste = (N + cap - 1) / cap * cap - N;
    for (int bn = cap; bn >= 96; bn -= step) {
        const int64_t waste = (N + bn - 1) / bn * bn - N;
        if (waste < best_waste) { best = bn; best_waste = waste; }
    }
    return best;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// returns 0 = launched, 1 = pattern unsupported (caller uses the SIMT engine), other = error
int gemm_tc(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t M, int64_t N, int64_t K,
            float* /*splitk_ws*/, int64_t /*splitk_ws_elems*/, cudaStream_t st) {
    // ---- eligibility
    if (A->dtype != MSU_BF16 || B->dtype != MSU_BF16 || E->dtype != MSU_BF16 || E->out_f32 || E->accumulate) return 1;
    if (A->orient != 0 || B->orient != 0 || B->map != MSU_MAP_NONE || B->ptr2 != nullptr) return 1;
    if (A->rowscale != nullptr || B->rowscale != nullptr) return 1;
    if (!(A->map == MSU_MAP_NONE || A->map == MSU_MAP_CONV3)) return 1;
    if (!(E->map == MSU_MAP_NONE || E->map == MSU_MAP_WINDOW || E->map == MSU_MAP_SHUFFLE || E->map == MSU_MAP_UNSHUFFLE)) return 1;
    if (N % 8 != 0 || (E->ldc % 8) != 0 || (E->R && E->ldr % 8 != 0) || (E->H && E->ldh % 8 != 0)) return 1;
    if ((E->map == MSU_MAP_SHUFFLE || E->map == MSU_MAP_UNSHUFFLE) && E->geo[3] % 8 != 0) return 1;
    if ((A->ld % 8) != 0 || (B->ld % 8) != 0 || !aligned16(A->ptr) || !aligned16(B->ptr) || !aligned16(E->C)) return 1;
    if ((E->Cpre && !aligned16(E->Cpre)) || (E->R && !aligned16(E->R)) || (E->H && !aligned16(E->H))) return 1;
    if (E->bias && !aligned16(E->bias)) return 1;
    if (M < 64) return 1;  // tiny problems: not worth a 128-row tile
    if (get_encode() == nullptr) return 1;
    const bool lnd = E->lnd_w != nullptr;      // fused head LayerNorm + dot: only this path implements it (the caller fails on 1)
    if (lnd && (!E->lnd_gamma || !E->lnd_beta || !E->lnd_logits || !E->lnd_mean || !E->lnd_rstd || !E->lnd_m2 || E->map != MSU_MAP_NONE ||
                E->R || E->H || E->Cpre || E->act || E->rowscale || N % 32 != 0 || N > 256 || !aligned16(E->lnd_gamma) ||
                !aligned16(E->lnd_beta) || !aligned16(E->lnd_w)))
        return 1;
    if (!lnd && E->C == nullptr) return 1;

    TcParams p{};
    p.M = M; p.N = (int)N; p.K = (int)K;
    p.nacc = 1;
    static const int env_bn = getenv("MSU_TC_BN") ? atoi(getenv("MSU_TC_BN")) : 0;
    static const int env_stages = getenv("MSU_TC_STAGES") ? atoi(getenv("MSU_TC_STAGES")) : 0;
    static const int env_epi = getenv("MSU_TC_EPI") ? atoi(getenv("MSU_TC_EPI")) : 0;   // 1: force the generic epilogue
    // unmapped outputs whose tile rows are consecutive leave through TMA stores (conv tiles: whole-row tiles only)
    bool epi_tma = env_epi != 1 && E->map == MSU_MAP_NONE && !(E->R && E->H) && !(E->Cpre && (E->R || E->H)) &&
                   E->ldr == E->ldc;
    // depth-to-space output maps: every 32-row x 32-column slab is one box of a 5-D view of the output
    if (env_epi != 1 && E->map == MSU_MAP_SHUFFLE && !E->R && !E->H && E->geo[1] % 32 == 0 && E->geo[3] % 32 == 0 &&
        E->ldc == E->geo[3] && N == (int64_t)E->geo[2] * E->geo[2] * E->geo[3] && A->map == MSU_MAP_NONE)
        epi_tma = true;
    if (env_epi != 1 && E->map == MSU_MAP_UNSHUFFLE && !E->R && !E->Cpre && 32 % E->geo[2] == 0 &&
        (E->geo[1] * E->geo[2]) % 32 == 0 && N == E->geo[3] && E->ldc == (int64_t)E->geo[2] * E->geo[2] * E->geo[3])
        epi_tma = true;
    if (A->map == MSU_MAP_CONV3 && A->geo[1] % 128 != 0) epi_tma = false;
    if (lnd && !epi_tma) return 1;
    const int naux = E->Cpre ? 1 : ((E->R || E->H) ? 2 : 0);
    p.epi_tma = epi_tma ? 1 : 0;
    p.epi_bytes = epi_tma ? 2048 * (1 + naux) : ((TC_SCRATCH_BYTES + 1023) / 1024) * 1024;
    p.BN = pick_bn(N, env_bn ? env_bn : 256, epi_tma ? 32 : 16);
    p.num_n_tiles = (int)((N + p.BN - 1) / p.BN);
    if (lnd && p.num_n_tiles != 1) return 1;
    p.E = *E;
    CUtensorMap tmA, tmA2, tmB;
    if (A->map == MSU_MAP_CONV3) {
        if (A->ptr2 != nullptr) return 1;
        const int H = A->geo[0], W = A->geo[1], C = A->geo[2];
        if (A->ld != C || C % 8 != 0 || K != 9 * (int64_t)C || M % ((int64_t)H * W) != 0) return 1;
        int bmw, bmh;
        if (W % 128 == 0) { bmw = 128; bmh = 1; }
        else if (W % 64 == 0 && H % 2 == 0) { bmw = 64; bmh = 2; }
        else if (W % 32 == 0 && H % 4 == 0) { bmw = 32; bmh = 4; }
        else if (W % 16 == 0 && H % 8 == 0) { bmw = 16; bmh = 8; }
        else return 1;
        const int Bn = (int)(M / ((int64_t)H * W));
        p.mode = 2; p.C = C; p.H = H; p.W = W; p.bmw = bmw; p.bmh = bmh;
        p.cblocks = (C + TC_BK - 1) / TC_BK;
        p.tiles_x = W / bmw; p.tiles_y = H / bmh;
        p.num_m_tiles = Bn * p.tiles_x * p.tiles_y;
        p.kb1 = 9 * p.cblocks; p.kb2 = 0; p.k_split = 0;
        static const int halo_off = getenv("MSU_CONV_HALO") ? (atoi(getenv("MSU_CONV_HALO")) == 0) : 0;
        static const int rows2_off = getenv("MSU_CONV_ROWS2") ? (atoi(getenv("MSU_CONV_ROWS2")) == 0) : 0;
        p.nacc = 1;
        if (!halo_off && !rows2_off && bmw == 128 && H % 2 == 0 && C % 32 == 0 && N <= 128) {
            // two image rows per tile share every weight tile (35 % less operand traffic per pixel), K walked in exact
            // 32-channel steps (no zero-padded channel block): BK = 32, 64B swizzle
            p.mode = 4;
            p.nacc = 2;
            p.bmh = 2;
            p.tiles_y = H / 2;
            p.num_m_tiles = Bn * p.tiles_x * p.tiles_y;
            p.cblocks = C / 32;
            p.kb1 = 3 * p.cblocks;
            cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bn};
            cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
            // CTA pairs: the kernel sits on two on-chip limits at once, shared-memory bandwidth (operand reads of the tensor core
            // + TMA writes) and the TMA unit's box rate; a pair halves the weight-tile share of the first, two boxes per K block
            // instead of five lift the second (either alone measured no gain).  Needs whole-row TMA-store tiles, one N tile, an
            // even tile count and half tiles of whole 8-row swizzle groups (MSU_CONV_PAIR=0 / MSU_CONV_2BOX=0: off).
            static const int pair_on = getenv("MSU_CONV_PAIR") ? atoi(getenv("MSU_CONV_PAIR")) : 1;
            p.pair = (pair_on && epi_tma && p.num_n_tiles == 1 && p.BN % 16 == 0 && p.num_m_tiles % 2 == 0 && num_sms() >= 2) ? 1 : 0;
            static const int box2_on = getenv("MSU_CONV_2BOX") ? atoi(getenv("MSU_CONV_2BOX")) : -1;
            p.box2 = box2_on < 0 ? p.pair : box2_on;
            cuuint32_t box[4] = {32, 130, (cuuint32_t)(p.box2 ? 2 : 1), 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            if (get_encode()(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(A->ptr), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return 1;
        } else if (!halo_off && bmw == 128) {
            // one 130-pixel halo row per (dy, channel block) instead of nine shifted 128-pixel boxes: 2.9x less A traffic
            p.mode = 3;
            p.kb1 = 3 * p.cblocks;
            if (!make_map_nhwc(&tmA, A->ptr, Bn, H, W, C, 130, 1)) return 1;
        } else if (!make_map_nhwc(&tmA, A->ptr, Bn, H, W, C, bmw, bmh)) return 1;
        tmA2 = tmA;
    } else {
        p.mode = A->ptr2 ? 1 : 0;
        p.num_m_tiles = (int)((M + TC_BM - 1) / TC_BM);
        if (A->ptr2) {
            if ((A->ld2 % 8) != 0 || !aligned16(A->ptr2) || A->k_split % 8 != 0) return 1;
            const int64_t K1 = A->k_split, K2 = K - K1;
            p.kb1 = (int)((K1 + TC_BK - 1) / TC_BK); p.kb2 = (int)((K2 + TC_BK - 1) / TC_BK); p.k_split = (int)K1;
            if (!make_map_2d(&tmA, A->ptr, M, K1, A->ld, TC_BM)) return 1;
            if (!make_map_2d(&tmA2, A->ptr2, M, K2, A->ld2, TC_BM)) return 1;
        } else {
            if (K % 8 != 0) return 1;
            p.kb1 = (int)((K + TC_BK - 1) / TC_BK); p.kb2 = 0; p.k_split = 0;
            if (!make_map_2d(&tmA, A->ptr, M, K, A->ld, TC_BM)) return 1;
            tmA2 = tmA;
        }
        // CTA pairs for the long-K shapes: a pair shares the B tile, (128 + BN) -> (128 + BN / 2) operand rows per CTA and K block.
        // Measured (tools/gemm_case.py, pairs on / off): 8192 x 4096 x 4096 183 / 193 us (1504 TFLOP/s), 4096 x 3072 x 768 27.5 / 31.2,
        // 65536 x 192 x 768 35.0 / 36.3, 16384 x 384 x 1536 27.2 / 27.9; at K = 384 the two-tile-per-CTA kernels are bound by their
        // prologue and epilogue tail and pairs cost 3-6 %, hence the K threshold.  The M tile count is rounded up to whole pairs:
        // a tile past M loads zeros and stores nothing (MSU_TC_PAIR=0: off, MSU_TC_PAIR_K: threshold).
        static const int tpair_on = getenv("MSU_TC_PAIR") ? atoi(getenv("MSU_TC_PAIR")) : 1;
        static const int tpair_k = getenv("MSU_TC_PAIR_K") ? atoi(getenv("MSU_TC_PAIR_K")) : 768;
        // K >= 384 with the stored-derivative multiply (act 3, the MLP backward's dh): that epilogue waits on its operand slabs, not on
        // issue slots, and the pair's smaller L2 traffic shows (16384 x 1536 x 384: 37.9 -> 33.7 us; the GELU epilogues of the same shape
        // lose 12 % because the slower CTA's epilogue gates both)
        const bool pair_k384 = K >= 384 && E->H != nullptr && E->act == 3;
        if (tpair_on && (K >= tpair_k || pair_k384) && p.BN % 16 == 0 && p.num_m_tiles >= 2 && num_sms() >= 2) {
            p.pair = 1;
            p.num_m_tiles = (p.num_m_tiles + 1) & ~1;
        }
    }
    if (p.mode == 4 && p.box2) {   // the three dx tiles of (dy, channel block) as one box: third dimension = tap, C columns apart
        cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)N, 3};
        cuuint64_t gstr[2] = {(cuuint64_t)B->ld * 2, (cuuint64_t)p.C * 2};
        cuuint32_t box[3] = {32, (cuuint32_t)(p.pair ? p.BN / 2 : p.BN), 3};
        cuuint32_t estr[3] = {1, 1, 1};
        if (get_encode()(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(B->ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 1;
    } else if (p.mode == 4) {   // weight tiles [BN rows, 32 k] with the 64B swizzle of the halo rows (a CTA of a pair: half the rows)
        cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
        cuuint64_t gstr[1] = {(cuuint64_t)B->ld * 2};
        cuuint32_t box[2] = {32, (cuuint32_t)(p.pair ? p.BN / 2 : p.BN)};
        cuuint32_t estr[2] = {1, 1};
        if (get_encode()(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(B->ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 1;
    } else if (!make_map_2d(&tmB, B->ptr, N, K, B->ld, p.pair ? p.BN / 2 : p.BN)) return 1;

    p.a_stage = p.mode == 4 ? 2 * TC_HALO32_BYTES : (p.mode == 3 ? 17 * 1024 : TC_A_BYTES);
    p.b_stage = p.mode == 4 ? 3 * (p.pair ? p.BN / 2 : p.BN) * 64 : (p.mode == 3 ? 3 : 1) * (p.pair ? p.BN / 2 : p.BN) * TC_BK * 2;
    const int stage_bytes = p.a_stage + p.b_stage;
    // [operand stages][1 KB: barriers, TMEM slot][epilogue slabs] + 1 KB alignment slack
    const int TC_EPI_WARPS = epi_tma ? TC_EPW_TMA : TC_EPW_GEN;
    auto stages_for = [&](int sb) { int s_ = (227 * 1024 - 2048 - TC_EPI_WARPS * p.epi_bytes - (lnd ? 1024 : 0)) / sb; return s_ > 8 ? 8 : s_; };
    p.stages = stages_for(stage_bytes);
    if (p.mode < 3 && p.BN > 192 && p.stages < 4 && !env_bn && !p.pair) {
        // operand bytes in flight bound the mainloop: a narrower N tile that buys the 4th stage wins
        const int bn2 = pick_bn(N, 192, epi_tma ? 32 : 16);
        const int sb2 = p.a_stage + bn2 * TC_BK * 2;
        if (stages_for(sb2) >= 4) {
            p.BN = bn2;
            p.num_n_tiles = (int)((N + p.BN - 1) / p.BN);
            p.b_stage = p.BN * TC_BK * 2;
            if (!make_map_2d(&tmB, B->ptr, N, K, B->ld, p.BN)) return 1;
            p.stages = stages_for(sb2);
        }
    }
    const int stage_bytes_final = p.a_stage + p.b_stage;
    const int fixed_bytes = 2048 + TC_EPI_WARPS * p.epi_bytes + (lnd ? 1024 : 0);
    if (env_stages && p.stages > env_stages) p.stages = env_stages;
    if (p.stages < 2) return 1;
    const int smem = p.stages * stage_bytes_final + fixed_bytes;
    CUtensorMap tmC = tmB, tmAux = tmB;
    if (epi_tma) {
        const void* aux = E->Cpre ? E->Cpre : (E->R ? E->R : E->H);
        const int64_t ldaux = E->Cpre ? E->ldc : (E->R ? E->ldr : E->ldh);
        if (E->map == MSU_MAP_NONE) {
            if (E->C != nullptr && !make_map_slab(&tmC, E->C, M, N, E->ldc)) return 1;
            if (aux != nullptr && !make_map_slab(&tmAux, aux, M, N, ldaux)) return 1;
        } else {
            // 5-D view (c, p2, w, p1, b*H + h) of the depth-to-space tensor: SHUFFLE memory is [(b, h p + p1, w p + p2), c],
            // UNSHUFFLE memory is [(b, h, w), (p1 p2 c)]
            const int Ht = E->geo[0], Wt = E->geo[1], pp = E->geo[2], cc = E->geo[3];
            const int64_t BH = (E->map == MSU_MAP_SHUFFLE ? M : M / ((int64_t)pp * pp)) / Wt;
            cuuint64_t gdim[5] = {(cuuint64_t)cc, (cuuint64_t)pp, (cuuint64_t)Wt, (cuuint64_t)pp, (cuuint64_t)BH};
            cuuint64_t gstr[4];
            cuuint32_t box[5] = {32, 1, 32, 1, 1};
            if (E->map == MSU_MAP_SHUFFLE) {
                gstr[0] = (cuuint64_t)cc * 2; gstr[1] = (cuuint64_t)pp * cc * 2; gstr[2] = (cuuint64_t)Wt * pp * cc * 2;
                gstr[3] = (cuuint64_t)pp * Wt * pp * cc * 2;
            } else {
                gstr[0] = (cuuint64_t)cc * 2; gstr[1] = (cuuint64_t)pp * pp * cc * 2; gstr[2] = (cuuint64_t)pp * cc * 2;
                gstr[3] = (cuuint64_t)Wt * pp * pp * cc * 2;
                box[1] = (cuuint32_t)pp; box[2] = (cuuint32_t)(32 / pp);
            }
            (void)Ht;
            cuuint32_t estr[5] = {1, 1, 1, 1, 1};
            auto enc5 = [&](CUtensorMap* tm, const void* ptr) {
                return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            };
            if (!enc5(&tmC, E->C)) return 1;
            if (E->Cpre != nullptr && !enc5(&tmAux, E->Cpre)) return 1;
            if (E->H != nullptr && !make_map_slab(&tmAux, E->H, M, N, E->ldh)) return 1;   // gelu' operand: unmapped [M, N]
        }
    }
    static PerDeviceOnce smem_set;
    if (smem_set.need()) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<true, TC_EPW_TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_tc_kernel<false, TC_EPW_GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_tc_kernel<true, TC_EPW_TMA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_tc_kernel<false, TC_EPW_GEN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_tc_kernel<true, TC_EPW_TMA, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_tc_kernel<true, TC_EPW_TMA, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        smem_set.set();
    }
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    static const int dbg_skip = getenv("MSU_CONV_SKIP") ? atoi(getenv("MSU_CONV_SKIP")) : 0;
    p.dbg_skip = dbg_skip;
    static const int trace_on = getenv("MSU_TC_TRACE") ? atoi(getenv("MSU_TC_TRACE")) : 0;
    static long long* trace_buf = nullptr;
    if (trace_on) {
        if (trace_buf == nullptr) cudaMalloc(&trace_buf, 148 * 8 * 16 * sizeof(long long));
        cudaMemsetAsync(trace_buf, 0, 148 * 8 * 16 * sizeof(long long), st);
        p.trace = trace_buf;
    }
    {
        const dim3 blk(32 * (2 + (epi_tma ? TC_EPW_TMA : TC_EPW_GEN)), 1, 1);
        const dim3 grd((unsigned)(p.pair ? (grid & ~1) : grid), 1, 1);   // CTA pairs: whole clusters of two
        const int cl = p.pair ? 2 : 1;
        cudaError_t e;
        if (p.pair) {
            e = lnd ? launch_pdl(gemm_tc_kernel<true, TC_EPW_TMA, true, true>, grd, blk, (size_t)smem, st, cl, tmA, tmA2, tmB, tmC, tmAux, p)
                    : (epi_tma ? launch_pdl(gemm_tc_kernel<true, TC_EPW_TMA, true>, grd, blk, (size_t)smem, st, cl, tmA, tmA2, tmB, tmC, tmAux, p)
                               : launch_pdl(gemm_tc_kernel<false, TC_EPW_GEN, true>, grd, blk, (size_t)smem, st, cl, tmA, tmA2, tmB, tmC, tmAux, p));
        } else {
            e = lnd ? launch_pdl(gemm_tc_kernel<true, TC_EPW_TMA, false, true>, grd, blk, (size_t)smem, st, cl, tmA, tmA2, tmB, tmC, tmAux, p)
                    : (epi_tma ? launch_pdl(gemm_tc_kernel<true, TC_EPW_TMA>, grd, blk, (size_t)smem, st, cl, tmA, tmA2, tmB, tmC, tmAux, p)
                               : launch_pdl(gemm_tc_kernel<false, TC_EPW_GEN>, grd, blk, (size_t)smem, st, cl, tmA, tmA2, tmB, tmC, tmAux, p));
        }
        if (e != cudaSuccess) { set_error("gemm_tc: launch: %s", cudaGetErrorString(e)); return (int)e; }
    }
    count_launch();
    if (trace_on) {   // debug only: synchronous dump of the per-tile role timeline of two CTAs
        static long long host[148 * 8 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost);
        static const char* names[11] = {"prod_start", "prod_end", "mma_tempty", "mma_full0", "mma_commit", "epi_mapped",
                                        "epi_tfull", "epi_done", "epi_c0_scratch", "epi_c0_stored", "epi_c0_ldtm"};
        for (int cta : {0, 100}) {
            if (cta >= grid) continue;
            const long long t0 = host[(size_t)cta * 8 * 16];
            fprintf(stderr, "[tc trace] M=%lld N=%d K=%d BN=%d stages=%d epw=%d cta %d\n", (long long)M, (int)N, (int)K, p.BN, p.stages, TC_EPI_WARPS, cta);
            for (int it = 0; it < 8; it++) {
                fprintf(stderr, "  tile %d:", it);
                for (int ev = 0; ev < 11; ev++) {
                    const long long v = host[((size_t)cta * 8 + it) * 16 + ev];
                    fprintf(stderr, " %s=%lld", names[ev], v ? v - t0 : -1);
                }
                fprintf(stderr, "\n");
            }
        }
    }
    return check_launch("gemm_tc");
}

// ================================================================================================
// Weight-gradient GEMM on tcgen05:  C[i, j] = sum_t P[t, i] * Q[t, j]   (contraction over token rows)
//
// Both operands are read exactly as autograd leaves them — row-major [tokens, channels] — i.e. MN-major for
// the tensor core (no transposes are materialised).  One CTA owns a token slab (deterministic split-K) and
// up to MT 128-row accumulators in TMEM, so every activation element is fetched once per N tile.
// Partials go to an fp32 workspace [split][I][J]; splitk_reduce_kernel adds them in a fixed order.
// ================================================================================================
struct WgParams {
    int64_t T;          // tokens (contraction length)
    int I, J;           // logical output [I, J]
    int Pn, Qn;         // extent of the M-side / N-side operands (== I,J or J,I when swapped)
    int swap;           // 1: M side is the logical j index
    int MT, BN, n_super, n_q_tiles, splits;
    int64_t tok_per_split;
    int stages, qboxes;
    float* ws;
    float* bws;         // bias-gradient partials [split][Pn] (column sums of the M-side operand), or nullptr
    // red = 1: no workspace, no reduce kernel.  The output [I, J] fp32 (zeroed by the host unless it accumulates) takes every
    // split's tile as a TMA reduce-add (L2 atomics; the order of the fp32 additions is then not fixed) and the bias partials
    // as red.global.add; the per-sample scale of the token rows is applied to the split's tile in registers.
    int red, sps;
    const float* sscale;
    float* colsum;
};

constexpr int WG_BK = 64;            // tokens per stage
constexpr int WG_BOX_BYTES = 64 * 64 * 2;
constexpr int WG_THREADS = 192;


__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ,
                const __grid_constant__ CUtensorMap tmW, const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int A_BYTES = p.MT * 2 * WG_BOX_BYTES;
    const int B_BYTES = p.qboxes * WG_BOX_BYTES;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)p.stages * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * B_BYTES);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    uint8_t* sSlab = smem + (size_t)p.stages * (A_BYTES + B_BYTES) + 1024;   // [4 warps][2][4 KB] fp32 output slabs
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work decomposition: blockIdx.x = (split * n_super + sup) * n_q_tiles + qt
    const int qt = blockIdx.x % p.n_q_tiles;
    const int sup = (blockIdx.x / p.n_q_tiles) % p.n_super;
    const int split = blockIdx.x / (p.n_q_tiles * p.n_super);
    const int64_t t0 = (int64_t)split * p.tok_per_split;
    int64_t t1 = t0 + p.tok_per_split;
    if (t1 > p.T) t1 = p.T;
    const int KB = (int)((t1 - t0 + WG_BK - 1) / WG_BK);
    const int p0 = sup * p.MT * 128;   // first M-side channel of this CTA
    const int q0 = qt * p.BN;
    // bias gradient = column sums of the call's A operand (the M side, or the N side when swapped): the four epilogue
    // warps, idle during the main loop, add the operand tiles up from shared memory as the stages land (an extra
    // MMA against a tile of ones cost +20 % kernel time: MN-major MMAs are shared-memory-read bound).  Only the CTAs of the
    // first tile along the other axis do it (sharing the stages among all tiles measured slower: every CTA then pays
    // the later slot release).
    const bool do_bias = (p.bws != nullptr || (p.red && p.colsum != nullptr)) && (p.swap ? sup == 0 : qt == 0);
    constexpr int nq = 1, myq = 0;
    const float red_scale = (p.red && p.sscale != nullptr) ? p.sscale[split / p.sps] : 1.f;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmP);
        prefetch_tmap(&tmQ);
        prefetch_tmap(&tmW);
        // a stage is free when its MMAs have retired (+ when the four column-sum warps have read it)
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], do_bias ? 5 : 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    pdl_wait();           // PDL: barrier init / TMEM allocation above overlapped the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
                const int tok = (int)(t0 + (int64_t)kb * WG_BK);
                uint8_t* a_dst = sA + (size_t)stage * A_BYTES;
                uint8_t* b_dst = sB + (size_t)stage * B_BYTES;
                for (int bx = 0; bx < p.MT * 2; bx++) tma_load_2d(a_dst + bx * WG_BOX_BYTES, &tmP, &full[stage], p0 + bx * 64, tok);
                for (int bx = 0; bx < p.qboxes; bx++) tma_load_2d(b_dst + bx * WG_BOX_BYTES, &tmQ, &full[stage], q0 + bx * 64, tok);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, p.BN, 1, 1);
            int stage = 0; uint32_t phase = 0;
            // The issuing thread is the critical resource of this loop (ncu / SASS: ~21 instructions per MMA when every
            // descriptor was rebuilt): descriptors are built once and advanced by adding to the 14-bit address field
            // (shared memory < 256 KB, so the field cannot carry): +128 per 16-token K step, +1024 per 128-row M tile.
            const uint64_t ad0 = make_desc_mnmajor_sw128(smem_u32(sA), WG_BOX_BYTES);
            const uint64_t bd0 = make_desc_mnmajor_sw128(smem_u32(sB), WG_BOX_BYTES);
            const uint32_t a_step = (uint32_t)(A_BYTES >> 4), b_step = (uint32_t)(B_BYTES >> 4);
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t ad = ad0 + (uint64_t)(stage * a_step), bd = bd0 + (uint64_t)(stage * b_step);
                const uint32_t acc0 = kb != 0;
                for (int mt = 0; mt < p.MT; mt++) {
                    const uint64_t am = ad + (uint64_t)(mt * (2 * WG_BOX_BYTES >> 4));
                    const uint32_t d = tmem_base + mt * p.BN;
                    tc_mma_bf16(d, am, bd, idesc, acc0);                  // 16 tokens = 2048 B per K step
                    tc_mma_bf16(d, am + 128, bd + 128, idesc, 1);
                    tc_mma_bf16(d, am + 256, bd + 256, idesc, 1);
                    tc_mma_bf16(d, am + 384, bd + 384, idesc, 1);
                }
                tc_commit(&empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            tc_commit(tfull);
        }
    } else {
        // epilogue: TMEM -> fp32 [32 x 32] slab (128B swizzle) -> TMA store into this split's partial [I, J]
        // (coalesced 128 B rows, out-of-range rows / columns clipped by the tensor map)
        const int quad = warp & 3;
        uint8_t* slab = sSlab + (size_t)(warp - 2) * 2 * 4096;
        if (do_bias) {
            // warp w adds tokens [16 w, 16 w + 16) of every stage; lane = channel pair of a 64-channel box (one 128 B row
            // of the swizzled tile per load instruction: conflict free).  Fixed order => deterministic.
            const int ew = warp - 2;
            const int nbx = p.swap ? p.qboxes : p.MT * 2;          // <= 8 boxes of [64 tokens x 64 channels]
            float2 acc[8];
#pragma unroll
            for (int b = 0; b < 8; b++) acc[b] = make_float2(0.f, 0.f);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&full[stage], phase);
                const uint8_t* src = p.swap ? sB + (size_t)stage * B_BYTES : sA + (size_t)stage * A_BYTES;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    if (b < nbx && kb % nq == myq) {
#pragma unroll
                        for (int tt = 0; tt < 16; tt++) {
                            const int t = ew * 16 + tt;
                            const uint32_t u = *reinterpret_cast<const uint32_t*>(src + b * WG_BOX_BYTES + t * 128 + (((lane >> 2) ^ (t & 7)) << 4) + (lane & 3) * 4);
                            acc[b].x += bf16lo_f(u);
                            acc[b].y += bf16hi_f(u);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            // combine the four token quarters through the (still unused) output slabs, then one partial row per split
            float2* comb = reinterpret_cast<float2*>(sSlab);       // [4 warps][8 boxes][32 pairs]
#pragma unroll
            for (int b = 0; b < 8; b++) comb[(ew * 8 + b) * 32 + lane] = acc[b];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int org = p.swap ? q0 : p0;
            const int ext = p.swap ? (p.Qn < q0 + p.BN ? p.Qn : q0 + p.BN) : p.Pn;   // channels this CTA owns
            const int ldb = p.swap ? p.Qn : p.Pn;
            for (int idx = ew * 32 + lane; idx < nbx * 32; idx += 128) {
                const int b = idx >> 5, l = idx & 31;
                float2 t = comb[(0 * 8 + b) * 32 + l];
                for (int w = 1; w < 4; w++) { const float2 u = comb[(w * 8 + b) * 32 + l]; t.x += u.x; t.y += u.y; }
                const int c = org + b * 64 + l * 2;
                if (p.red) {
                    if (c < ext) atomicAdd(&p.colsum[c], t.x * red_scale);
                    if (c + 1 < ext) atomicAdd(&p.colsum[c + 1], t.y * red_scale);
                    continue;
                }
                if (c < ext) p.bws[((int64_t)split * nq + myq) * ldb + c] = t.x;
                if (c + 1 < ext) p.bws[((int64_t)split * nq + myq) * ldb + c + 1] = t.y;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");         // the slabs are reused by the stores below
        }
        mbar_wait(tfull, 0);
        tc_fence_after();
        int nb = 0;
        for (int mt = 0; mt < p.MT; mt++) {
            const int prow0 = p0 + mt * 128 + quad * 32;       // first M-side channel of this warp's slab
            for (int c = 0; c < (p.BN + 31) / 32; c++) {
                const int cols = p.BN - c * 32 >= 32 ? 32 : 16;
                uint32_t raw[32];
                if (cols == 32) tc_ld32_nowait(tmem_base + mt * p.BN + c * 32 + ((uint32_t)(quad * 32) << 16), raw);
                else tc_ld16_nowait(tmem_base + mt * p.BN + c * 32 + ((uint32_t)(quad * 32) << 16), raw);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // slab nb was stored two chunks ago
                __syncwarp();
                tc_ld_wait();
                if (prow0 >= p.Pn) continue;                   // warp-uniform: the whole slab is out of range
                uint8_t* sl = slab + nb * 4096;
                if (KB == 0) {
#pragma unroll
                    for (int i = 0; i < 32; i++) raw[i] = 0u;
                }
                if (p.red && p.sscale != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; i++) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * red_scale);
                }
                if (!p.swap) {          // slab[row = lane (i)][col = j]: 8 x 16 B per lane
#pragma unroll
                    for (int g = 0; g < 8; g++)
                        if (g * 4 < cols)
                            *reinterpret_cast<uint4*>(sl + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                                make_uint4(raw[g * 4], raw[g * 4 + 1], raw[g * 4 + 2], raw[g * 4 + 3]);
                } else {                // transposed: slab[row = q (i)][col = lane (j)]
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        if (i < cols)
                            *reinterpret_cast<uint32_t*>(sl + i * 128 + (((lane >> 2) ^ (i & 7)) << 4) + (lane & 3) * 4) = raw[i];
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (p.red) {
                        if (!p.swap) tma_red_add_3d(sl, &tmW, q0 + c * 32, prow0, 0);
                        else tma_red_add_3d(sl, &tmW, prow0, q0 + c * 32, 0);
                    } else if (!p.swap) tma_store_3d(sl, &tmW, q0 + c * 32, prow0, split);
                    else tma_store_3d(sl, &tmW, prow0, q0 + c * 32, split);
                    tma_store_commit();
                }
                nb ^= 1;
            }
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
    }
}

void launch_splitk_reduce(const MsuEpilogue& E, int64_t M, int64_t N, int splits, const float* ws, cudaStream_t st,
                          const float* bws = nullptr, int brows = 0, const float* sscale = nullptr, int splits_per_sample = 1);

// ================================================================================================
// 3x3 conv weight gradient on tcgen05:  dW[co, (tap, ci)] = sum_pix dZ[pix, co] * X[pix + off(tap), ci]
// Token slabs are WB consecutive pixels of one image row; the shifted X slabs come from the same 4-D NHWC
// tensor map with (x+dx, y+dy) coordinates (TMA zero fill = conv padding).  A CTA keeps the accumulators of
// up to G taps in TMEM (G*BN <= 512 columns) so dZ is fetched once per tap group.
// ================================================================================================
struct WcParams {
    int E, H, W, Bn;          // channels, image size, batch
    int WB, G, BN, n_groups, splits, boxes;   // slab width, taps per CTA, N per tap, tap groups, K splits, 64-ch boxes per operand
    int64_t slabs, slabs_per_split;
    int stages;
    int halo;                 // 1: one (WB+2)-pixel halo slab per dy, the three dx taps are row-shifted views of it
    int c32;                  // halo only: operands staged as 32-channel chunks (64 B rows, 64B swizzle) instead of 64-channel boxes:
                              // E = 96 is 3 chunks exactly, while a second 64-channel box is half out of range and made the TMA
                              // unit fetch 1.5x the bytes from L2 (ncu: 7.4 GB of sectors for 4.9 GB of operands)
    int onebox;               // c32 only: the three chunks of an operand arrive as ONE 5-D TMA box (chunk = outermost box dimension).
                              // The TMA unit serves a box of <= 8 KB in ~222 clk whatever its size (tools/tma_load_bench.cu), so six
                              // 4 KB boxes per K step (1330 clk) were the bound of this kernel, not bytes.
    int coll;                 // 1: K step outermost, the taps of a slice share dZ through the A collector (MSU_WGRAD_COLL=0: off)
    float* ws;
    float* bws;               // bias-gradient partials [split][E] (column sums of dZ), or nullptr
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_conv_tc_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmX, const WcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int BOX = p.WB * 128;
    const int grp = blockIdx.x % p.n_groups;
    const int split = blockIdx.x / p.n_groups;
    const int tap0 = grp * p.G;
    const int ntap = (9 - tap0) < p.G ? (9 - tap0) : p.G;
    const int nch = p.E / 32;                                                  // c32: chunks per operand
    const int ACH = p.WB * 64;                                                 // c32: chunk bytes (512 B = swizzle period)
    const int XCH = p.onebox ? (p.WB + 2) * 64 : ((p.WB + 2) * 64 + 511) / 512 * 512;
    const int A_BYTES = p.c32 ? nch * ACH : p.boxes * BOX;
    const int XBOX = p.halo ? ((p.WB + 2) * 128 + 1023) / 1024 * 1024 : BOX;   // halo slab box, 1 KB aligned
    const int B_BYTES = p.c32 ? (nch * XCH + 511) / 512 * 512 : (p.halo ? p.boxes * XBOX : p.G * p.boxes * BOX);
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)p.stages * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * B_BYTES);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    float2* sComb = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // [4 warps][2 boxes][32 pairs]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // conv bias gradient = column sums of dZ: added up from the dZ tiles by the epilogue warps while the MMAs run
    // the last tap group has the fewest taps; c32 (three groups, three chunks): every group sums its own chunk
    const bool do_bias = (p.bws != nullptr) && (grp == p.n_groups - 1 || (p.c32 && nch == p.n_groups));
    const int64_t s0 = (int64_t)split * p.slabs_per_split;
    int64_t s1 = s0 + p.slabs_per_split;
    if (s1 > p.slabs) s1 = p.slabs;
    const int KB = (int)(s1 - s0);
    const int slabs_per_row = p.W / p.WB;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmZ);
        prefetch_tmap(&tmX);
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], do_bias ? 5 : 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    pdl_wait();           // PDL: barrier init / TMEM allocation above overlapped the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < KB; kb++) {
                const int64_t slab = s0 + kb;
                const int x0 = (int)(slab % slabs_per_row) * p.WB;
                const int64_t row = slab / slabs_per_row;
                const int y = (int)(row % p.H), b = (int)(row / p.H);
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a_dst = sA + (size_t)stage * A_BYTES;
                uint8_t* b_dst = sB + (size_t)stage * B_BYTES;
                if (p.c32) {
                    mbar_arrive_expect_tx(&full[stage], nch * (ACH + (p.WB + 2) * 64));
                    if (p.onebox) {
                        tma_load_5d(a_dst, &tmZ, &full[stage], 0, x0, y, b, 0);
                        tma_load_5d(b_dst, &tmX, &full[stage], 0, x0 - 1, y + grp - 1, b, 0);
                    } else {
                        for (int c = 0; c < nch; c++) tma_load_4d(a_dst + c * ACH, &tmZ, &full[stage], c * 32, x0, y, b);
                        for (int c = 0; c < nch; c++) tma_load_4d(b_dst + c * XCH, &tmX, &full[stage], c * 32, x0 - 1, y + grp - 1, b);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    continue;
                }
                mbar_arrive_expect_tx(&full[stage], p.halo ? A_BYTES + p.boxes * (p.WB + 2) * 128 : A_BYTES + ntap * p.boxes * BOX);
                for (int bx = 0; bx < p.boxes; bx++) tma_load_4d(a_dst + bx * BOX, &tmZ, &full[stage], bx * 64, x0, y, b);
                if (p.halo) {   // group = dy; pixels x0-1 .. x0+WB of row y+dy-1 (zero fill outside the image)
                    for (int bx = 0; bx < p.boxes; bx++)
                        tma_load_4d(b_dst + bx * XBOX, &tmX, &full[stage], bx * 64, x0 - 1, y + grp - 1, b);
                }
                for (int t = 0; t < (p.halo ? 0 : ntap); t++) {
                    const int tap = tap0 + t;
                    for (int bx = 0; bx < p.boxes; bx++)
                        tma_load_4d(b_dst + (t * p.boxes + bx) * BOX, &tmX, &full[stage], bx * 64, x0 + tap % 3 - 1, y + tap / 3 - 1, b);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, p.BN, 1, 1);
            int stage = 0; uint32_t phase = 0;
            // descriptors built once, advanced through the 14-bit address field (the issuing thread is the critical
            // resource: ~21 instructions per MMA when they were rebuilt): +128 per 16-pixel K step, +tap_step per tap
            // (halo: tap t = dx is the slab shifted by t pixel rows = +128 B; the swizzle is address based)
            // c32: 64 B rows, so a 16-pixel K step is +1024 B and a one-pixel tap shift +64 B; M = 128 reads a fourth chunk of dZ
            // (whatever follows in shared memory: it only reaches accumulator rows >= E, which nobody reads)
            const uint64_t ad0 = p.c32 ? make_desc_mnmajor_sw64_lbo(smem_u32(sA), ACH) : make_desc_mnmajor_sw128(smem_u32(sA), BOX);
            const uint64_t bd0 = p.c32 ? make_desc_mnmajor_sw64_lbo(smem_u32(sB), XCH)
                                       : make_desc_mnmajor_sw128(smem_u32(sB), p.halo ? XBOX : BOX);
            const uint32_t a_step = (uint32_t)(A_BYTES >> 4), b_step = (uint32_t)(B_BYTES >> 4);
            const uint32_t tap_step = p.c32 ? (64 >> 4) : (p.halo ? (128 >> 4) : (uint32_t)((p.boxes * BOX) >> 4));
            const uint32_t kst = p.c32 ? 64 : 128;                                   // K step of 16 pixels, in 16 B units
            const int ksteps = p.WB / 16;
            const bool coll_on = p.coll != 0;
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t ad = ad0 + (uint64_t)(stage * a_step), bd = bd0 + (uint64_t)(stage * b_step);
                const uint32_t acc0 = kb != 0;
                if (ntap == 3 && ksteps == 4 && coll_on) {
                    // K step outermost: the three taps' MMAs of one 16-pixel slice share their dZ tile through the A collector
                    // (one shared-memory read instead of three: the loop is bound by shared-memory bandwidth, not by the tensor pipe)
                    const uint64_t b1 = bd + (uint64_t)tap_step, b2 = bd + (uint64_t)(2 * tap_step);
                    const uint32_t d1 = tmem_base + p.BN, d2 = tmem_base + 2 * p.BN;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t en = k == 0 ? acc0 : 1u;
                        tc_mma_bf16_coll<TC_COLL_FILL>(tmem_base, ad + (uint64_t)(k * kst), bd + (uint64_t)(k * kst), idesc, en);
                        tc_mma_bf16_coll<TC_COLL_USE>(d1, ad + (uint64_t)(k * kst), b1 + (uint64_t)(k * kst), idesc, en);
                        tc_mma_bf16_coll<TC_COLL_LASTUSE>(d2, ad + (uint64_t)(k * kst), b2 + (uint64_t)(k * kst), idesc, en);
                    }
                } else
                for (int t = 0; t < ntap; t++) {
                    const uint64_t bt = bd + (uint64_t)(t * tap_step);
                    const uint32_t d = tmem_base + t * p.BN;
                    if (ksteps == 4) {
                        tc_mma_bf16(d, ad, bt, idesc, acc0);
                        tc_mma_bf16(d, ad + kst, bt + kst, idesc, 1);
                        tc_mma_bf16(d, ad + 2 * kst, bt + 2 * kst, idesc, 1);
                        tc_mma_bf16(d, ad + 3 * kst, bt + 3 * kst, idesc, 1);
                    } else {
                        for (int k = 0; k < ksteps; k++) tc_mma_bf16(d, ad + (uint64_t)(k * kst), bt + (uint64_t)(k * kst), idesc, (kb | k) != 0);
                    }
                }
                tc_commit(&empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            tc_commit(tfull);
        }
    } else {
        const int quad = warp & 3;
        if (do_bias && p.c32) {
            // Three dy groups, three 32-channel chunks: group g adds up chunk g, so every CTA carries a third of the column-sum
            // work (the kernel ends with its slowest CTA; one group doing all of it cost +150 us).  64 B rows: a warp covers two
            // pixels per load (half-warp = pixel parity, lane & 15 = channel pair of the chunk).  Fixed order => deterministic.
            const int ew = warp - 2, tpw = p.WB / 4;
            const int hp = lane >> 4, pr = lane & 15;
            float2 acc = make_float2(0.f, 0.f);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&full[stage], phase);
                const uint8_t* src = sA + (size_t)stage * A_BYTES + grp * ACH;
#pragma unroll 4
                for (int tt = 0; tt < tpw; tt += 2) {
                    const int t = ew * tpw + tt + hp;
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(src + t * 64 + (((pr >> 2) ^ ((t >> 1) & 3)) << 4) + (pr & 3) * 4);
                    acc.x += bf16lo_f(u);
                    acc.y += bf16hi_f(u);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
            if (lane < 16) sComb[ew * 16 + lane] = acc;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (ew == 0 && lane < 16) {
                float2 t = sComb[lane];
                for (int w = 1; w < 4; w++) { const float2 u = sComb[w * 16 + lane]; t.x += u.x; t.y += u.y; }
                const int ch = grp * 32 + lane * 2;
                p.bws[(int64_t)split * p.E + ch] = t.x;
                p.bws[(int64_t)split * p.E + ch + 1] = t.y;
            }
        } else if (do_bias) {
            const int ew = warp - 2, tpw = p.WB / 4;               // tokens of a slab per warp
            float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < KB; kb++) {
                mbar_wait(&full[stage], phase);
                const uint8_t* src = sA + (size_t)stage * A_BYTES;
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    if (b < p.boxes) {
                        for (int tt = 0; tt < tpw; tt++) {
                            const int t = ew * tpw + tt;
                            const uint32_t u = *reinterpret_cast<const uint32_t*>(src + b * BOX + t * 128 + (((lane >> 2) ^ (t & 7)) << 4) + (lane & 3) * 4);
                            acc[b].x += bf16lo_f(u);
                            acc[b].y += bf16hi_f(u);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            sComb[(ew * 2 + 0) * 32 + lane] = acc[0];
            sComb[(ew * 2 + 1) * 32 + lane] = acc[1];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int idx = ew * 32 + lane; idx < p.boxes * 32; idx += 128) {
                const int b = idx >> 5, l = idx & 31;
                float2 t = sComb[(0 * 2 + b) * 32 + l];
                for (int w = 1; w < 4; w++) { const float2 u = sComb[(w * 2 + b) * 32 + l]; t.x += u.x; t.y += u.y; }
                const int c = b * 64 + l * 2;
                if (c < p.E) p.bws[(int64_t)split * p.E + c] = t.x;
                if (c + 1 < p.E) p.bws[(int64_t)split * p.E + c + 1] = t.y;
            }
        }
        mbar_wait(tfull, 0);
        tc_fence_after();
        const int J = 9 * p.E;
        float* wsz = p.ws + (int64_t)split * p.E * J;
        const int co = quad * 32 + lane;
        for (int t = 0; t < ntap; t++) {
            for (int c = 0; c < p.BN / 16; c++) {
                float v[16];
                tc_ld16(tmem_base + t * p.BN + c * 16 + ((uint32_t)(quad * 32) << 16), v);
                if (co >= p.E) continue;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int ci = c * 16 + i;
                    if (ci >= p.E) break;
                    wsz[(int64_t)co * J + (tap0 + t) * p.E + ci] = KB > 0 ? v[i] : 0.f;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
    }
}

// A = dZ (orient 1, plain [pix, E]); B = X (orient 1, MAP_CONV3 over NHWC [Bn,H,W,E]); out fp32 [E, 9E]
static int wgrad_conv_tc(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t I, int64_t J, int64_t T,
                         float* ws, int64_t ws_elems, cudaStream_t st, int* fused_colsum) {
    const int H = B->geo[0], W = B->geo[1], Ec = B->geo[2];
    if (I != Ec || J != 9 * (int64_t)Ec || Ec > 128 || Ec % 8 != 0 || A->ld != Ec || B->ld != Ec) return 1;
    if (T % ((int64_t)H * W) != 0) return 1;
    WcParams p{};
    p.E = Ec; p.H = H; p.W = W; p.Bn = (int)(T / ((int64_t)H * W)); p.ws = ws;
    p.WB = (W % 64 == 0) ? 64 : (W % 32 == 0 ? 32 : (W % 16 == 0 ? 16 : 0));
    if (p.WB == 0) return 1;
    p.BN = (Ec + 15) / 16 * 16;
    p.boxes = (Ec + 63) / 64;
    const bool want_bias = E->colsum != nullptr && Ec <= 128;
    p.G = TC_TMEM_COLS / p.BN;
    if (p.G > 9) p.G = 9;
    const int BOX = p.WB * 128;
    while (p.G > 1 && 2 * (p.boxes + p.G * p.boxes) * BOX > 200 * 1024) p.G--;
    p.n_groups = (9 + p.G - 1) / p.G;
    int stage_bytes = (p.boxes + p.G * p.boxes) * BOX;
    // halo slabs (taps grouped by dy) halve the L2 -> SM operand traffic of the 5+4 tap grouping: 1025 vs 1230 us at 16 x 512^2 x 96
    // once the MMA issue loop was made cheap (MSU_WGRAD_HALO=0 restores the tap grouping)
    static const int halo_on = getenv("MSU_WGRAD_HALO") ? atoi(getenv("MSU_WGRAD_HALO")) : 1;
    p.halo = (halo_on && p.WB == 64 && 3 * p.BN <= TC_TMEM_COLS) ? 1 : 0;
    static const int c32_on = getenv("MSU_WGRAD_C32") ? atoi(getenv("MSU_WGRAD_C32")) : 1;
    p.c32 = (p.halo && c32_on && Ec == 96) ? 1 : 0;   // three chunks = three dy groups (the bias sums are split that way)
    if (p.halo) {   // taps grouped by dy: 3 accumulators per CTA, dZ read 3x but X read once per dy (1.8x less L2 traffic)
        p.G = 3;
        p.n_groups = 3;
        stage_bytes = p.boxes * BOX + p.boxes * (((p.WB + 2) * 128 + 1023) / 1024 * 1024);
        static const int onebox_on = getenv("MSU_WGRAD_1BOX") ? atoi(getenv("MSU_WGRAD_1BOX")) : 1;
        p.onebox = (p.c32 && onebox_on) ? 1 : 0;
        if (p.c32) stage_bytes = (Ec / 32) * (p.WB * 64 + ((p.WB + 2) * 64 + 511) / 512 * 512);
        if (p.onebox) stage_bytes = (Ec / 32) * p.WB * 64 + ((Ec / 32) * (p.WB + 2) * 64 + 511) / 512 * 512;
    }
    static const int coll_on = getenv("MSU_WGRAD_COLL") ? atoi(getenv("MSU_WGRAD_COLL")) : 1;
    p.coll = coll_on;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 6) p.stages = 6;
    static const int env_wstages = getenv("MSU_WGRAD_STAGES") ? atoi(getenv("MSU_WGRAD_STAGES")) : 0;
    if (env_wstages && p.stages > env_wstages) p.stages = env_wstages;
    if (p.stages < 2) return 1;
    p.slabs = (int64_t)p.Bn * H * (W / p.WB);
    // whole waves: resident CTAs per SM follow from the shared memory of a CTA (TMEM: G * BN columns each)
    const int smem_cta = p.stages * stage_bytes + 4096;
    int cps = smem_cta <= 113 * 1024 ? 2 : 1;
    if (cps * p.G * p.BN > TC_TMEM_COLS) cps = 1;
    int splits = (cps == 1 ? 2 : 1) * cps * num_sms() / p.n_groups;   // two waves of single CTAs, or one wave of pairs
    if (p.halo) splits = cps * num_sms() / p.n_groups;
    if (splits < 1) splits = 1;
    if (splits > p.slabs) splits = (int)p.slabs;
    while (splits > 1 && (int64_t)splits * I * (J + 1) > ws_elems) splits--;
    if ((int64_t)splits * I * (J + 1) > ws_elems) return 1;
    p.slabs_per_split = (p.slabs + splits - 1) / splits;
    splits = (int)((p.slabs + p.slabs_per_split - 1) / p.slabs_per_split);
    p.splits = splits;
    p.bws = want_bias ? ws + (int64_t)splits * I * J : nullptr;
    CUtensorMap tmZ, tmX;
    static const int l2promo = getenv("MSU_WGRAD_L2PROMO") ? atoi(getenv("MSU_WGRAD_L2PROMO")) : (int)CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)Ec, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)p.Bn};
        cuuint64_t gstr[3] = {(cuuint64_t)Ec * 2, (cuuint64_t)W * Ec * 2, (cuuint64_t)H * W * Ec * 2};
        cuuint32_t box[4] = {(cuuint32_t)(p.c32 ? 32 : 64), (cuuint32_t)p.WB, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (p.onebox) {   // 5-D view: the 32-channel chunk index is the outermost dimension (64 B apart in memory)
            cuuint64_t gdim5[5] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)p.Bn, (cuuint64_t)(Ec / 32)};
            cuuint64_t gstr5[4] = {(cuuint64_t)Ec * 2, (cuuint64_t)W * Ec * 2, (cuuint64_t)H * W * Ec * 2, 64};
            cuuint32_t estr5[5] = {1, 1, 1, 1, 1};
            for (int k = 0; k < 2 && p.onebox; k++) {
                cuuint32_t box5[5] = {32, (cuuint32_t)(k ? p.WB + 2 : p.WB), 1, 1, (cuuint32_t)(Ec / 32)};
                if (get_encode()(k ? &tmX : &tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(k ? B->ptr : A->ptr), gdim5, gstr5,
                                 box5, estr5, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                 (CUtensorMapL2promotion)l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return 1;
            }
        }
        for (int k = 0; k < 2 && !p.onebox; k++) {
            box[1] = (cuuint32_t)((k == 1 && p.halo) ? p.WB + 2 : p.WB);
            if (get_encode()(k ? &tmX : &tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(k ? B->ptr : A->ptr), gdim, gstr,
                             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, p.c32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                             (CUtensorMapL2promotion)l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return 1;
        }
    }
    const int smem = p.stages * stage_bytes + (2 * p.stages + 2) * 8 + 48 + 4 * 2 * 32 * 8 + 1024;   // + column-sum combine buffer
    static PerDeviceOnce attr;
    if (attr.need()) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("wgrad_conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr.set();
    }
    {
        const cudaError_t e = launch_pdl<2>(wgrad_conv_tc_kernel, dim3((unsigned)(p.n_groups * splits)), dim3(WG_THREADS), (size_t)smem, st, 1, tmZ, tmX, p);
        if (e != cudaSuccess) { set_error("wgrad_conv_tc: launch: %s", cudaGetErrorString(e)); return (int)e; }
    }
    count_launch();
    launch_splitk_reduce(*E, I, J, splits, ws, st, p.bws, splits);
    if (fused_colsum) *fused_colsum = want_bias ? 1 : 0;
    return check_launch("wgrad_conv_tc");
}

static bool make_map_2d_box64(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
    return make_map_2d(tm, ptr, rows, cols, ld, 64);
}

// returns 0 = launched, 1 = unsupported (SIMT fallback), other = error.  A(i, t) and B(j, t) both orient=1.
int wgrad_tc(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t I, int64_t J, int64_t T,
             float* ws, int64_t ws_elems, cudaStream_t st, int* fused_colsum) {
    if (fused_colsum) *fused_colsum = 0;
    if (A->dtype != MSU_BF16 || B->dtype != MSU_BF16 || !E->out_f32) return 1;
    if (A->orient == 1 && B->orient == 1 && A->map == MSU_MAP_NONE && B->map == MSU_MAP_CONV3) {
        if (A->ptr2 || B->ptr2 || A->rowscale || B->rowscale || !aligned16(A->ptr) || !aligned16(B->ptr)) return 1;
        if (E->map != MSU_MAP_NONE || E->bias || E->R || E->H || E->Cpre || E->act || E->rowscale) return 1;
        if (ws == nullptr || get_encode() == nullptr) return 1;
        return wgrad_conv_tc(A, B, E, I, J, T, ws, ws_elems, st, fused_colsum);
    }
    if (A->orient != 1 || B->orient != 1 || A->map != MSU_MAP_NONE || B->map != MSU_MAP_NONE) return 1;
    if (A->ptr2 || B->ptr2 || B->rowscale) return 1;
    // per-sample scale of A's token rows (stochastic depth on gradient rows): every split is kept inside one sample and
    // the reduce scales whole partials, so the scaled rows are never materialised
    const int64_t rps = A->rowscale ? A->rows_per_sample : 0;
    if (A->rowscale && (rps < WG_BK || rps % WG_BK != 0 || T % rps != 0)) return 1;
    if (E->map != MSU_MAP_NONE || E->bias || E->R || E->H || E->Cpre || E->act || E->rowscale) return 1;
    if ((A->ld % 8) || (B->ld % 8) || !aligned16(A->ptr) || !aligned16(B->ptr) || (I % 8) || (J % 8)) return 1;
    if (T < 512 || ws == nullptr || get_encode() == nullptr) return 1;
    WgParams p{};
    p.T = T; p.I = (int)I; p.J = (int)J; p.ws = ws;
    p.swap = J > I ? 1 : 0;
    const MsuOperand* P = p.swap ? B : A;
    const MsuOperand* Q = p.swap ? A : B;
    p.Pn = p.swap ? (int)J : (int)I;
    p.Qn = p.swap ? (int)I : (int)J;
    const int m_tiles = (p.Pn + 127) / 128;
    // Long token axis (stage 0/1): activation reads dominate -> wide N tile and several M accumulators per CTA so
    // that each activation element is fetched once.  Short token axis with a large output (stage 2/3): many
    // output tiles, few splits -> small deterministic split-K partial traffic.
    // (only where the partials make a round trip through the workspace: with the TMA reduce-add the wide tiling wins everywhere,
    //  4096 x 3072 x 768: 32 -> 24 us)
    const bool out_heavy = deterministic_mode() && (int64_t)I * J * 8 > T * (I + J);
    if (out_heavy) {
        p.BN = p.Qn <= 128 ? (p.Qn + 15) / 16 * 16 : pick_bn(p.Qn, 128, 32);   // several q tiles: whole 32-column store boxes
        p.MT = 1;
    } else {
        p.BN = p.Qn <= 256 ? (p.Qn + 15) / 16 * 16 : pick_bn(p.Qn, 256, 32);
        p.MT = m_tiles < 4 ? m_tiles : 4;
    }
    static const int env_mt = getenv("MSU_WG_MT") ? atoi(getenv("MSU_WG_MT")) : 0;       // tuning overrides (tools/wgrad_case.py)
    static const int env_wbn = getenv("MSU_WG_BN") ? atoi(getenv("MSU_WG_BN")) : 0;
    static const int env_splits = getenv("MSU_WG_SPLITS") ? atoi(getenv("MSU_WG_SPLITS")) : 0;
    if (env_wbn) p.BN = p.Qn <= env_wbn ? (p.Qn + 15) / 16 * 16 : pick_bn(p.Qn, env_wbn, 32);
    if (env_mt) p.MT = m_tiles < env_mt ? m_tiles : env_mt;
    p.n_q_tiles = (p.Qn + p.BN - 1) / p.BN;
    p.qboxes = (p.BN + 63) / 64;
    const bool want_bias = (E->colsum != nullptr);   // column sums of the call's A operand ride along (either side)
    while (p.MT > 1 && p.MT * p.BN > TC_TMEM_COLS) p.MT--;
    // keep at least 2 pipeline stages in 200 KB
    while (p.MT > 1 && 2 * (p.MT * 2 + p.qboxes) * WG_BOX_BYTES > 200 * 1024) p.MT--;
    p.n_super = (m_tiles + p.MT - 1) / p.MT;
    const int stage_bytes = (p.MT * 2 + p.qboxes) * WG_BOX_BYTES;
    p.stages = (193 * 1024) / stage_bytes;
    if (p.stages > 6) p.stages = 6;
    if (p.stages < 2) return 1;
    const int base_ctas = p.n_super * p.n_q_tiles;
    // whole waves: one CTA per SM is resident (shared memory), so splits * base_ctas must not spill a few CTAs into a second
    // wave (13 splits x 12 tiles = 156 CTAs cost 43 us where 8 x 12 took 32 us at 16384 x 1536 x 384)
    int splits = num_sms() / base_ctas;
    if (splits < 1) splits = 1;
    if (env_splits) splits = env_splits;
    const int64_t max_splits_t = (T + 511) / 512;
    if (splits > max_splits_t) splits = (int)max_splits_t;
    const int nq = 1;   // bias partial rows per split
    while (splits > 1 && (int64_t)splits * I * (J + nq) > ws_elems) splits--;
    if ((int64_t)splits * I * (J + nq) > ws_elems) return 1;
    int64_t tps = (T + splits - 1) / splits;
    tps = (tps + WG_BK - 1) / WG_BK * WG_BK;
    int sps = 1;                                  // splits per sample
    if (rps) {
        // tokens per split = rps / sps with sps a power of two (rps = H*W is one), as close to the balanced choice as possible
        while (sps * 2 <= rps / WG_BK && rps / (sps * 2) >= tps && (rps / (sps * 2)) % WG_BK == 0) sps *= 2;
        if (rps % sps != 0 || (rps / sps) % WG_BK != 0) return 1;
        tps = rps / sps;
        if ((T / tps) * I * (J + nq) > ws_elems) return 1;
    }
    splits = (int)((T + tps - 1) / tps);
    p.splits = splits; p.tok_per_split = tps;
    p.bws = want_bias ? ws + (int64_t)splits * I * J : nullptr;
    // msu_set_deterministic(1): fp32 partials in the workspace + splitk_reduce_kernel, fixed order of additions
    const int env_red = !deterministic_mode();
    const int64_t ldc = E->ldc > 0 ? E->ldc : J;
    p.red = (env_red && aligned16(E->C) && ldc % 4 == 0) ? 1 : 0;  // 16-byte rows (J % 8 == 0 already)
    if (p.red) {
        p.bws = nullptr;
        p.colsum = want_bias ? E->colsum : nullptr;
        p.sscale = A->rowscale; p.sps = sps;
        cudaError_t e = cudaSuccess;
        if (!E->accumulate) e = cudaMemset2DAsync(E->C, (size_t)ldc * sizeof(float), 0, (size_t)J * sizeof(float), (size_t)I, st);
        if (e == cudaSuccess && want_bias) e = cudaMemsetAsync(E->colsum, 0, (size_t)I * sizeof(float), st);
        if (e != cudaSuccess) { set_error("wgrad_tc: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    }
    CUtensorMap tmP, tmQ;
    if (!make_map_2d_box64(&tmP, P->ptr, T, p.Pn, P->ld)) return 1;
    if (!make_map_2d_box64(&tmQ, Q->ptr, T, p.Qn, Q->ld)) return 1;
    // partials ws[split][I][J] fp32 as a 3-D map, box = [32 cols, 32 rows, 1], 128B swizzle
    CUtensorMap tmW;
    {
        cuuint64_t gdim[3] = {(cuuint64_t)J, (cuuint64_t)I, (cuuint64_t)(p.red ? 1 : splits)};
        cuuint64_t gstr[2] = {(cuuint64_t)(p.red ? ldc : J) * 4, (cuuint64_t)I * (p.red ? ldc : J) * 4};
        cuuint32_t box[3] = {32, 32, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        if (get_encode()(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.red ? (float*)E->C : ws, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 1;
    }
    const int smem = p.stages * stage_bytes + 1024 + 4 * 2 * 4096 + 1024;
    static PerDeviceOnce attr;
    if (attr.need()) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr.set();
    }
    {
        const cudaError_t e = launch_pdl<2>(wgrad_tc_kernel, dim3((unsigned)(base_ctas * splits)), dim3(WG_THREADS), (size_t)smem, st, 1, tmP, tmQ, tmW, p);
        if (e != cudaSuccess) { set_error("wgrad_tc: launch: %s", cudaGetErrorString(e)); return (int)e; }
    }
    count_launch();
    if (!p.red) launch_splitk_reduce(*E, I, J, splits, ws, st, p.bws, splits * nq, A->rowscale, sps);
    if (fused_colsum) *fused_colsum = want_bias ? 1 : 0;
    return check_launch("wgrad_tc");
}

}  // namespace msu

