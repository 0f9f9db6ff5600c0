/*
 * msunet_b200.h — C ABI of libmsunet_sm100.so: hand-written sm_100a kernels for the MS-UNet hot path
 * (forward/backward of the Swin-UNet, DynamicLoss, Dice/IoU counting).
 *
 * The reference (Sara-H-dev/Semantic_Segmentation_Of_StyleGAN2_Artifacts) is pure Python/PyTorch and has
 * no FFI layer; each entry below names the reference code whose arithmetic it replaces (paths relative to
 * the reference root; "TV:" = torchvision 0.26 `torchvision/`).  The Python host in
 * semantic_segmentation_of_stylegan2_artifacts_b200/ binds these through ctypes.
 *
 * Conventions
 *   - every entry returns int: 0 ok, >0 a cudaError_t, <0 an argument/shape error; nothing throws or
 *     aborts; msu_last_error_string() describes the last failure on the calling thread.
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch); the library allocates nothing
 *     persistent except cached TMA descriptors / function attributes, and keeps no pointer after return.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *   - activation dtype codes: 0 = float32, 1 = bfloat16, 2 = float16 (loss/metrics inputs only).
 *   - activations are token-major, channels-last, contiguous: [rows, C].
 */
#ifndef MSUNET_B200_H
#define MSUNET_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSU_F32 0
#define MSU_BF16 1
#define MSU_F16 2

/* Row maps: how a logical (row, col) of a GEMM operand / output / LayerNorm row is found in memory. */
#define MSU_MAP_NONE 0     /* memory row = row                                                             */
#define MSU_MAP_WINDOW 1   /* row is a window-order index (b, window, i<49) of the padded+rolled map;       \
                              memory row = source pixel, or absent (zero on load, skipped on store).       \
                              geo = {H, W, Ph, Pw, shift_h, shift_w}. TV:models/swin_transformer.py:152-172,219-227 */
#define MSU_MAP_SHUFFLE 2  /* depth-to-space: logical [token(b,h,w), (p1 p2 c)] <-> memory [(b,h*p+p1,w*p+p2), c]; \
                              geo = {H, W, p, c}. network/model_parts.py:402, 463-464 (einops rearrange)   */
#define MSU_MAP_CONV3 3    /* implicit 3x3 im2col: logical [pixel(b,y,x), (tap ci)] -> memory [(b,y+dy,x+dx), ci], \
                              zero outside the image; geo = {H, W, C}. network/model_parts.py:468-471       */
#define MSU_MAP_MERGE 4    /* 2x2 neighbourhood concat: logical [(b,h/2,w/2), (q c)] -> memory [(b,2h'+q%2,2w'+q/2), c]; \
                              geo = {H, W, C}. network/model_parts.py:87-92                                  */

#define MSU_MAP_UNSHUFFLE 5 /* inverse depth-to-space (backward of MSU_MAP_SHUFFLE): logical [(b,h*p+p1,w*p+p2), c] ->      \
                              memory [token(b,h,w), (p1 p2 c)]; geo = {H, W, p, c}                                 */

typedef struct {
    const void* ptr;        /* base pointer                                                              */
    const void* ptr2;       /* second source for logical columns >= k_split (skip concat), or NULL       */
    int64_t ld, ld2;        /* leading dimensions in elements                                            */
    int32_t k_split;        /* network/model_parts.py:792,804,823 torch.cat([x, skip], -1) folded here    */
    int32_t orient;         /* 0: memory is [i, k] (k contiguous); 1: memory is [k, i] (i contiguous)      */
    int32_t map;            /* MSU_MAP_* applied to the memory row                                        */
    int32_t dtype;          /* MSU_F32 / MSU_BF16                                                        */
    int32_t geo[6];
    const float* rowscale;  /* optional per-sample scale (stochastic depth) on memory rows, or NULL      */
    int32_t rows_per_sample;
    int32_t _pad;
} MsuOperand;

typedef struct {
    void* C;                /* output [M(mapped), N(mapped)], dtype `dtype`                               */
    void* Cpre;             /* optional pre-activation copy (same mapping), or NULL                       */
    const float* bias;      /* [N] fp32 or NULL                                                           */
    const void* R;          /* residual, indexed like C, or NULL                                          */
    const void* H;          /* if set: multiply by gelu'(H[m,n]) (unmapped, ld = ldh) — MLP backward       */
    int64_t ldc, ldr, ldh;
    const float* rowscale;  /* per-sample scale of the branch before the residual add                     */
    int32_t rows_per_sample;
    int32_t act;            /* 0 none, 1 exact (erf) GELU, 2 GELU with Cpre := GELU'(pre) (the MLP forward stores  \
                               the derivative, not the pre-activation), 3 (with H) H already holds GELU'(pre)   */
    int32_t map;            /* MSU_MAP_NONE / WINDOW / SHUFFLE / UNSHUFFLE on the output                  */
    int32_t dtype;          /* dtype of C, Cpre, R, H                                                     */
    int32_t geo[6];
    int32_t out_f32;        /* 1: C is fp32 regardless of dtype (weight gradients)                        */
    int32_t accumulate;     /* 1: C += result (shared weights / gradient accumulation)                    */
    float* colsum;          /* weight-gradient GEMMs (A, B orient 1) only, or NULL: also writes            \
                               colsum[m] = sum_k A(m,k), the bias gradient of the same layer (autograd's   \
                               grad_output.sum(0)); fused into the tensor-core kernel as one extra N=16     \
                               MMA against a tile of ones where the tiling allows, else a column-sum pass */
    /* Fused head LayerNorm + 1x1 conv (network/model_parts.py:475, 846 behind the second 3x3 conv, :471): when lnd_w != NULL the
     * GEMM (N = the LayerNorm width, one N tile, no other fused operand) also normalises every output row (after bias, rounded
     * to `dtype` as stored) and writes lnd_logits[m] = dot(LN(row), lnd_w) (`dtype`) with the statistics msu_ln_bwd(dotw) needs:
     * lnd_mean / lnd_rstd / lnd_m2 [M] fp32.  C may then be NULL (inference: the row itself is not stored).
     * Supported by the tcgen05 TMA-store path only: msu_gemm fails otherwise (no silent fallback). */
    const float* lnd_gamma;
    const float* lnd_beta;
    const float* lnd_w;
    void* lnd_logits;
    float* lnd_mean;
    float* lnd_rstd;
    float* lnd_m2;
} MsuEpilogue;

/* C[m,n] = epilogue( sum_k A(m,k) * B(n,k) ), fp32 accumulation.
 * Replaces every nn.Linear / F.linear / Conv2d on MSUNetSys.forward and their autograd backward:
 * TV:models/swin_transformer.py:179,215; TV:ops/misc.py:292-303; network/model_parts.py:95, 395, 459,
 * 468-471, 793, 805, 824.  `splitk_ws` (fp32, >= splits*M*N, or NULL) enables deterministic split-K.
 * `backend`: 0 auto (tcgen05 when the operand pattern is supported, else SIMT), 1 force SIMT fp32-accumulate. */
int msu_gemm(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t M, int64_t N, int64_t K,
             float* splitk_ws, int64_t splitk_ws_elems, int backend, void* stream);

/* Column sums: out[n] (+)= sum_m X(m,n) — bias gradients. `ws` fp32 >= 256*N. */
int msu_colsum(const MsuOperand* X, int64_t M, int64_t N, float* out, int accumulate, float* ws,
               int64_t ws_elems, void* stream);

/* LayerNorm forward over `rows` output rows of width C (eps 1e-5, fp32 statistics).
 * in_map: NONE | MERGE (gathers 2x2 neighbours, C = 4*Cin);  out_map: NONE | WINDOW (rows are window-order
 * indices; absent rows are written as zeros — the unmasked zero padding of TV:...:152-156).
 * mean/rstd are indexed by LayerNorm row (source pixel for WINDOW).  If `dotw` != NULL the kernel writes
 * logits[row] = dot(LN(x), dotw) instead of Y (head LN + 1x1 conv, network/model_parts.py:475, 846) and
 * dot_m2[row] = mean_c(gamma_c dotw_c x-hat_c), the one row reduction msu_ln_bwd then needs (C <= 768 bf16 / 384 fp32).
 * Replaces nn.LayerNorm at TV:...:453-454, network/model_parts.py:94, 224, 404, 475, 813, 827. */
int msu_ln_fwd(int dtype, const void* X, const float* gamma, const float* beta, void* Y, float* mean,
               float* rstd, int64_t rows, int32_t C, int32_t in_map, int32_t out_map, const int32_t* geo,
               const float* dotw, float* dot_m2, void* stream);

/* LayerNorm backward.  dY is read through `dy_map` (NONE, or WINDOW: gradient rows live in window order),
 * dX written through `dx_map` (NONE, MERGE scatter, or UNSHUFFLE: inverse depth-to-space).  dX = LN'(dY) + dRes (dRes optional).
 * partial: fp32 workspace [msu_ln_bwd_partial_rows(dtype,rows,C), 3, C] for deterministic dgamma/dbeta/(ddotw) reduction, finished by
 * msu_ln_param_reduce.  If dotw != NULL, dY is a per-row scalar (d logits) times dotw and dot_m2 is the forward's output. */
int msu_ln_bwd_partial_rows(int dtype, int64_t rows, int32_t C);
int msu_ln_bwd(int dtype, const void* dY, const void* X, const float* gamma, const float* beta,
               const float* mean, const float* rstd, const void* dRes, void* dX, int64_t rows, int32_t C,
               int32_t dy_map, int32_t dx_map, const int32_t* geo, const float* dotw, const float* dot_m2,
               float* partial, void* stream);
/* As msu_ln_bwd (no row maps) and additionally dXw[pix_to_win(row)] = rowscale[row / rows_per_sample] * dX[row]: the gradient rows
 * of the attention projection in window order (backward of TV:models/swin_transformer.py:219-227 + stochastic depth) come out of
 * the LayerNorm backward that produces them.  Padding rows of dXw are never written: the caller zeroes them once. */
int msu_ln_bwd_dual(int dtype, const void* dY, const void* X, const float* gamma, const float* beta, const float* mean,
                    const float* rstd, const void* dRes, void* dX, void* dXw, int64_t rows, int32_t C, const int32_t* wgeo,
                    const float* rowscale, int32_t rows_per_sample, float* partial, void* stream);
int msu_ln_param_reduce(const float* partial, int32_t partial_rows, int32_t C, float* dgamma, float* dbeta,
                        float* ddotw, int accumulate, void* stream);

/* Window attention core on window-ordered qkv [nWinTotal*49, 3C] -> O [nWinTotal*49, C]; head dim 32.
 * S = (q*32^-1/2) k^T + bias[h] + mask, softmax fp32, O = P v.  `bias` is the expanded [nH,49,49] table.
 * geo = {H, W, Ph, Pw, shift_h, shift_w}: the -100 shift mask is derived from the window index.
 * Replaces TV:models/swin_transformer.py:181-214. */
int msu_winattn_fwd(int dtype, const void* qkv, const float* bias, void* O, int64_t n_windows, int32_t nH,
                    const int32_t* geo, float p_drop, const uint32_t* seed, float* lse, void* stream);
/* `lse` (optional, fp32 [n_windows * 49, nH]): the log2-domain log-sum-exp of every softmax row, log2(sum_j 2^(l_j)) with
 * l_j = log2(e) * (q.k / sqrt(32) + bias + mask).  Handed to msu_winattn_bwd it lets the backward form P = 2^(l - lse) directly:
 * no row maximum, no row sum, no normalisation (a quarter of the instructions of its softmax part). */
/* Attention dropout (TV:...:205): p_drop > 0 with `seed` = device pointer to two 32-bit words drops softmax outputs with a
 * counter-based mask (hash of window, head, query, key and the seed; kept values scaled by 1/(1-p)); the backward must get the
 * same p_drop and seed words.  p_drop = 0 or seed = NULL: no dropout. */
/* 0 = auto (tcgen05 kernels for bf16), 1 = force the SIMT fp32-FMA kernels (parity checks of the tensor-core path). */
int msu_set_attn_backend(int backend);
/* Backward (recomputes P from qkv; O is the forward output): dqkv [.,3C];
 * dbias_partial fp32 [msu_winattn_bwd_grid(dtype,n_windows,nH), nH, 2401], reduced by msu_relbias_reduce.
 * When msu_winattn_bwd_direct(dtype) returns 1 (tcgen05 backend, not msu_set_deterministic(1)) the caller may instead pass
 * `dtable` = the bias-table gradient [169, nH] fp32 itself, zeroed (or holding a gradient to add to): the kernel adds every CTA's
 * contribution with red.global.add, dbias_partial may be NULL and no msu_relbias_reduce is needed.  Otherwise dtable must be NULL. */
int msu_winattn_bwd_direct(int dtype);
int msu_winattn_bwd_grid(int dtype, int64_t n_windows, int32_t nH);
int msu_winattn_bwd(int dtype, const void* qkv, const float* bias, const void* O, const void* dO, void* dqkv,
                    float* dbias_partial, float* dtable, int64_t n_windows, int32_t nH, const int32_t* geo, float p_drop,
                    const uint32_t* seed, const float* lse, void* stream);   /* lse: the forward's, or NULL (recomputed) */
/* bias[h,i,j] = table[index(i,j), h]  (TV:...:49-56) and its deterministic transpose-reduction. */
int msu_relbias_expand(const float* table, float* bias, int32_t nH, void* stream);
int msu_relbias_reduce(const float* dbias_partial, int32_t grid, int32_t nH, float* dtable, int accumulate,
                       void* stream);

/* Weight / layout preparation (fp32 master -> compute dtype):
 * mode 0: cast [R,C]; 1: transpose+cast -> [C,R]; 2: conv [co,ci,3,3] -> [co,(tap ci)];
 * 3: conv -> flipped-transposed [ci,(tap' co)] for dgrad; 4: conv-grad [co,(tap ci)] fp32 -> [co,ci,3,3] fp32
 * 5: patch-embed conv [E,3,4,4] -> [E, 64] zero-padded K (48 -> 64); 6: inverse of 5 for the gradient. */
int msu_prep_weight(int mode, int dst_dtype, const float* src, void* dst, int64_t R, int64_t C, void* stream);

/* im2col of non-overlapping 4x4 patches: image [B,3,S,S] fp32 NCHW -> [B*(S/4)^2, 64] (cols >=48 zero).
 * network/model_parts.py:222 (Conv2d k=4 s=4 as a GEMM). */
int msu_patchify4(int dst_dtype, const float* img, void* out, int32_t B, int32_t S, void* stream);

/* Fused BCE-with-logits + Tversky loss (loss/DynamicLoss.py:82-111), no host sync.
 * logits [B, N] (dtype f32/bf16/f16), target [B, N] fp32 ({0,1} or {0,255}: binarised at 127.5 when the
 * global max exceeds 1).  stats fp32 [B, 8] (per-sample backward coefficients),
 * flag int32[1] (global max>1), loss fp32[1].  msu_loss_bwd writes dlogits (same dtype) scaled by *gscale. */
int msu_loss_fwd(int dtype, const void* logits, const float* target, int32_t B, int64_t N, float alpha,
                 float beta, float mix, float* ws /* fp32 >= B*64*16 */, float* stats /* [B,8] */, int32_t* flag,
                 float* loss, void* stream);
int msu_loss_bwd(int dtype, const void* logits, const float* target, int32_t B, int64_t N, float alpha,
                 float beta, float mix, const float* stats, const int32_t* flag, const float* gscale,
                 void* dlogits, void* stream);

/* Per-image DynamicLoss values for batched validation: image b is evaluated as a batch of one, including ITS OWN {0,255}
 * label decision (the reference calls the loss once per image, scripts/validation_functions.py:89-104 with
 * loss/DynamicLoss.py:87-88).  stats[b*8 + 4] = loss of image b (other slots as msu_loss_fwd, with B = 1 scaling not applied). */
int msu_loss_per_sample(int dtype, const void* logits, const float* target, int32_t B, int64_t N, float alpha, float beta,
                        float mix, float* ws /* >= B*64*16 floats */, float* stats /* [B,8] */, void* stream);

/* Dice/IoU counting (scripts/validation_functions.py:106-108, 219-227, 267-292).
 * from_logits=1: pred = sigmoid(logit) rounded to `dtype`, pred_bin = pred > thr, gt = label > 0.
 * from_logits=2: pred = sigmoid(logit) kept in fp32 whatever `dtype` is (the reference thresholds an fp16 / fp32 sigmoid,
 *                validation_functions.py:78, 106-107; a bf16-rounded one would move the 0.5 boundary by 2^-9); pred_out is float.
 * from_logits=0: `in` is pred (dtype), pred_bin uint8 given, gt uint8 given.
 * counts int64 [B,4] = tp, fp, fn, tn (bit exact); soft fp64 [B,8] = TP, FP, FN, TN, sum p^2, sum g^2, sum p, sum g.
 * pred_out (optional; dtype, or float for from_logits=2) receives the probabilities. */
int msu_metrics(int dtype, int from_logits, const void* in, const void* label_or_gt, const uint8_t* pred_bin,
                int32_t B, int64_t N, float thr, long long* ws_counts /* >= B*64*4 */, double* ws_soft /* >= B*64*8 */,
                long long* counts, double* soft, void* pred_out, void* stream);

/* dst[m, n] = src(m, n) read through the operand's row map and per-sample scale (absent rows -> 0):
 * materialises the window-ordered gradient rows (backward of TV:models/swin_transformer.py:219-227) or the
 * inverse depth-to-space view (backward of network/model_parts.py:402, 463) as a dense [M, N] matrix. */
int msu_gather_rows(const MsuOperand* src, void* dst, int64_t M, int64_t N, void* stream);

/* Elementwise helpers used by the host (cast fp32 <-> compute dtype, y = a + b). */
int msu_cast(int src_dtype, int dst_dtype, const void* src, void* dst, int64_t n, void* stream);
int msu_add(int dtype, const void* a, const void* b, void* y, int64_t n, void* stream);

/* Fused multi-tensor AdamW (torch.optim.AdamW semantics, amsgrad off; trainer.py:143-152, 315), fp32 parameters / moments.
 * `table` is a DEVICE array of per-tensor records; block b of the launch updates elements
 * [blk_chunk[b] * msu_adamw_chunk(), ...) of tensor blk_tensor[b].  Per-tensor scalars are pre-folded by the host:
 * decay = 1 - lr * weight_decay, step_size = lr / (1 - beta1^t), inv_bias2_sqrt = 1 / sqrt(1 - beta2^t).
 * inv_scale (optional device scalar) multiplies every gradient (GradScaler unscale); found_inf (optional device scalar):
 * a non-zero value skips the whole step. */
typedef struct {
    void* p;
    const void* g;
    void* m;
    void* v;
    int64_t n;
    float decay, step_size, inv_bias2_sqrt, beta1, beta2, eps;
} MsuAdamTensor;
int msu_adamw_chunk(void);
int msu_adamw_step(const MsuAdamTensor* table, const int32_t* blk_tensor, const int32_t* blk_chunk, int32_t n_blocks,
                   const float* inv_scale, const float* found_inf, void* stream);

/* Every weight shadow of a model in one launch (the per-tensor form is msu_prep_weight).  Replaces the per-layer weight casts that
 * CUDA autocast performs inside every forward of the reference (trainer.py:308, scripts/validation_functions.py:78, 324).  A job of mode 0 reads the fp32 master
 * [R, C] once and writes the plain cast `dst` [R, C] and / or the transposed cast `dst_t` [C, R] (either may be NULL); modes 2, 3, 5
 * are msu_prep_weight's conv / patch-embed re-layouts into `dst`.  blk_job / blk_tile map each block to (job, tile); a job owns
 * msu_shadow_blocks(mode, R, C) consecutive tiles 0..n-1.  All pointers are device pointers. */
typedef struct {
    const float* src;
    void* dst;
    void* dst_t;
    int64_t R, C;
    int32_t mode;
    int32_t dtype;      /* MSU_F32 / MSU_BF16 of dst and dst_t */
} MsuShadowJob;
int msu_shadow_blocks(int mode, int64_t R, int64_t C);
int msu_refresh_shadows(const MsuShadowJob* jobs, const int32_t* blk_job, const int32_t* blk_tile, int32_t n_blocks, void* stream);

/* Device-side input staging = the tail of the reference's per-sample transform for a whole batch,
 * dataset/dataset.py:13-16 (horizontal flip of image and label) and :49-63 (uint8 HWC -> float32 / 255 in CHW; label -> (label > 127)
 * as float32).  img_hwc [B,H,W,3] u8, label [B,H,W] u8 or NULL (then label_out NULL), flip [B] u8 (non-zero = flip that sample) or
 * NULL; image_out [B,3,H,W] f32, label_out [B,H,W] f32.  Bit-exact against the numpy formulas. */
int msu_stage_u8(const uint8_t* img_hwc, const uint8_t* label, const uint8_t* flip, float* image_out, float* label_out,
                 int32_t B, int32_t H, int32_t W, void* stream);

/* Weight-gradient GEMMs split the token axis over the SMs.  0 (default; MSU_DETERMINISTIC=1 in the environment flips it): every
 * split adds its fp32 tile into the output with a TMA reduce (L2 atomics) - no workspace pass, but the order of the additions varies
 * from run to run (as it does for cuBLAS / cuDNN split-K under the reference, which sets no torch.use_deterministic_algorithms).
 * 1: partial tiles go to the workspace and a second kernel adds them in a fixed order (bit-reproducible gradients).
 * Returns the previous setting. */
int msu_set_deterministic(int on);
int msu_version(void);
/* sizeof(MsuOperand) (which=0) / sizeof(MsuEpilogue) (which=1): lets a binding verify its struct layout. */
int msu_struct_size(int which);
const char* msu_last_error_string(void);
/* Number of kernels this library has launched on the calling process (bench.py "gpu_launches"). */
long long msu_launch_count(void);
/* 1 if the tcgen05 GEMM path was used by the last msu_gemm call on this thread, else 0. */
int msu_last_gemm_backend(void);

#ifdef __cplusplus
}
#endif
#endif
