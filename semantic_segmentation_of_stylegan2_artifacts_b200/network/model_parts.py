"""MS-UNet module tree with the reference's parameter names, running on the sm_100a kernels.

Mirrors the *interface* of the reference's network/model_parts.py (class names, constructor
arguments, attribute/parameter names and therefore state_dict keys — network/model_parts.py:59-97,
109-173, 187-225, 374-407, 437-541, 543-893) so checkpoints, the name-based weight-decay split of
trainer.py:133-140 and MSUNet.freeze_encoder keep working.  nn.Linear / nn.LayerNorm / nn.Conv2d are
used purely as parameter containers; every forward below dispatches to functional.py, i.e. to
hand-written CUDA kernels through the C ABI.  There is no PyTorch-op fallback.
"""
from __future__ import annotations

import os
import sys
import weakref

import torch
import torch.nn as nn

from .. import functional as Fn

_SHADOW_STATE = weakref.WeakKeyDictionary()     # model -> cached weight-shadow refresh plan (functional.refresh_shadows)

_PREC = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp32": torch.float32, "float32": torch.float32}


def default_compute_dtype() -> torch.dtype:
    """bf16 (tcgen05 path) unless MSUNET_B200_PRECISION=fp32 selects the fp32 parity mode."""
    return _PREC[os.environ.get("MSUNET_B200_PRECISION", "bf16").lower()]


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class ShiftedWindowAttention(nn.Module):
    """Parameter holder named like torchvision's (TV:models/swin_transformer.py:234-284)."""

    def __init__(self, dim, window_size, shift_size, num_heads, attention_dropout=0.0, dropout=0.0):
        super().__init__()
        if dim % num_heads != 0 or dim // num_heads != 32:
            raise ValueError(f"head dim must be 32 (got dim={dim}, heads={num_heads}); every reference config uses 32")
        if list(window_size) != [7, 7]:
            raise ValueError("window size must be 7 (config.yaml:40)")
        self.window_size, self.shift_size, self.num_heads = list(window_size), list(shift_size), num_heads
        self.attention_dropout, self.dropout = attention_dropout, dropout
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(169, num_heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        c = torch.arange(7)
        cy, cx = torch.meshgrid(c, c, indexing="ij")
        cy, cx = cy.flatten(), cx.flatten()
        idx = (cy[:, None] - cy[None, :] + 6) * 13 + (cx[:, None] - cx[None, :] + 6)
        self.register_buffer("relative_position_index", idx.flatten())


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio=4.0, dropout=0.0,
                 attention_dropout=0.0, stochastic_depth_prob=0.0, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim, self.num_heads, self.shift = dim, num_heads, int(shift_size[0])
        self.sd_prob = float(stochastic_depth_prob)
        self.norm1 = nn.LayerNorm(dim)
        self.attn = ShiftedWindowAttention(dim, window_size, shift_size, num_heads, attention_dropout, dropout)
        self.norm2 = nn.LayerNorm(dim)
        hidden = int(dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden, dim),
                                 nn.Dropout(dropout))
        self.attn_p = float(attention_dropout)
        if dropout > 0:
            raise NotImplementedError("MLP / projection dropout (MODEL.DROP_RATE > 0; config.yaml:22 uses 0.0) is not built into the "
                                      "fc1 / fc2 / proj epilogues: refusing to train a different model silently "
                                      "(attention dropout and stochastic depth are supported)")

    def forward(self, x):  # x [B, H, W, C]
        B, H, W, _ = x.shape
        pool = getattr(self, "_sd_pool", None)      # rows pre-drawn for the whole forward by MSUNetSys (3 launches instead of 88)
        if pool is not None and pool[0].shape[0] == B and self.training:
            sd1, sd2 = pool
            self._sd_pool = None
        else:
            sd1 = Fn.drop_path_noise(self.sd_prob, self.training, B, x.device)
            sd2 = Fn.drop_path_noise(self.sd_prob, self.training, B, x.device)
        a, m = self.attn, self.mlp
        # attention dropout (TV:models/swin_transformer.py:205): two fresh 32-bit words per call from PyTorch's generator
        # (a device tensor, so CUDA-graph replays draw new masks); the kernels hash (seed, window, head, query, key)
        drop_seed = (torch.randint(-2 ** 31, 2 ** 31 - 1, (2,), dtype=torch.int32, device=x.device)
                     if (self.training and self.attn_p > 0.0) else None)
        return Fn.SwinBlockFn.apply(x, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight,
                                    a.proj.bias, a.relative_position_bias_table, self.norm2.weight, self.norm2.bias,
                                    m[0].weight, m[0].bias, m[3].weight, m[3].bias, sd1, sd2, B, H, W,
                                    self.num_heads, self.shift, self.attn_p if drop_seed is not None else 0.0, drop_seed)


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)

    def forward(self, x):  # [B,H,W,C] -> [B, HW/4, 2C]
        B, H, W, _ = x.shape
        assert (H, W) == tuple(self.input_resolution), "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        return Fn.PatchMergeFn.apply(x, self.norm.weight, self.norm.bias, self.reduction.weight, B, H, W)


class PatchExpand(nn.Module):
    def __init__(self, input_resolution, dim, dim_scale=2, norm_layer=nn.LayerNorm):
        super().__init__()
        if dim_scale != 2:
            raise ValueError("only dim_scale=2 is used by the reference")
        self.input_resolution, self.dim = input_resolution, dim
        self.expand = nn.Linear(dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(dim // dim_scale)

    def forward(self, x):  # [B,L,C] or [B,H,W,C] -> [B,4L,C/2]
        if x.dim() == 4:
            B, H, W, _ = x.shape
        elif x.dim() == 3:
            B, L, _ = x.shape
            H, W = self.input_resolution
            assert L == H * W, "input feature has wrong size"
        else:
            raise ValueError(f"Unexpected dimensionality: x.dim()={x.dim()}")
        return Fn.PatchExpandFn.apply(x, self.expand.weight, self.norm.weight, self.norm.bias, B, H, W)


class _Stage(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio, drop, attn_drop, drop_path,
                 use_checkpoint):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.use_checkpoint = use_checkpoint  # accepted; activations fit in 180 GB so nothing is recomputed
        ws = [window_size, window_size]
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, num_heads, ws, [0 if i % 2 == 0 else w // 2 for w in ws], mlp_ratio, drop,
                                 attn_drop, drop_path[i] if isinstance(drop_path, list) else drop_path)
            for i in range(depth)])

    def _run_blocks(self, x):
        B, N, C = x.shape
        H, W = self.input_resolution
        assert H * W == N, f"{N=} passt nicht zu {H}x{W}"
        x = x.view(B, H, W, C)
        for blk in self.blocks:
            x = blk(x)
        return x


class BasicLayer(_Stage):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0.1, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 fused_window_process=False):
        super().__init__(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, drop, attn_drop, drop_path,
                         use_checkpoint)
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        x = self._run_blocks(x)
        return self.downsample(x) if self.downsample is not None else x


class BasicLayer_up(_Stage):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, upsample=None, use_checkpoint=False):
        super().__init__(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, drop, attn_drop, drop_path,
                         use_checkpoint)
        self.upsample = PatchExpand(input_resolution, dim=dim, dim_scale=2) if upsample is not None else None

    def forward(self, x):
        x = self._run_blocks(x)
        return self.upsample(x) if self.upsample is not None else x


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        img_size, patch_size = _pair(img_size), _pair(patch_size)
        if patch_size != (4, 4) or in_chans != 3:
            raise ValueError("the fused patch embed supports patch 4 / 3 channels (config.yaml:34-35)")
        if norm_layer is None:
            raise ValueError("patch_norm=False is not supported (config.yaml:45 uses True)")
        self.img_size, self.patch_size = img_size, patch_size
        self.patches_resolution = [img_size[0] // 4, img_size[1] // 4]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim)
        self.compute_dtype = default_compute_dtype()

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        return Fn.PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias,
                                     self.compute_dtype)


class FinalPatchExpand_X4_V2(nn.Module):
    def __init__(self, input_resolution, dim, dim_scale=4, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim, self.dim_scale, self.output_dim = input_resolution, dim, dim_scale, dim
        self.expand = nn.Linear(dim, 16 * dim, bias=False)
        self.act = nn.GELU()
        self.refine1 = nn.Conv2d(dim, dim, kernel_size=3, padding=1, bias=True)
        self.refine2 = nn.Conv2d(dim, dim, kernel_size=3, padding=1, bias=True)
        self.norm = nn.LayerNorm(dim)


class MSUNetSys(nn.Module):
    def __init__(self, img_size=1024, patch_size=4, in_chans=3, num_classes=1, embed_dim=128, depths=[2, 2, 18, 2],
                 depths_decoder=[2, 2, 6, 2], num_heads=[4, 8, 16, 32], window_size=7, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False,
                 patch_norm=True, use_checkpoint=False, final_upsample="expand_first", run_dead_branches=False,
                 **kwargs):
        super().__init__()
        print("SwinTransformerSys expand initial---- \n depths:{}; \n depths_decoder:{}; \n drop_path_rate:{};\n "
              "num_classes:{}".format(depths, depths_decoder, drop_path_rate, num_classes), file=sys.stderr)
        if len(depths) != 4:
            raise ValueError("MS-UNet wiring assumes 4 stages (network/model_parts.py:783-810)")
        if ape:
            raise NotImplementedError("absolute position embedding (APE) is off in every reference config")
        if num_classes != 1 or final_upsample != "expand_first":
            raise NotImplementedError("binary head with final_upsample='expand_first' only (config.yaml:20,46)")
        if img_size % 32 != 0:
            raise ValueError("img_size must be a multiple of 32")
        self.num_classes, self.num_layers, self.embed_dim = num_classes, 4, embed_dim
        self.ape, self.patch_norm, self.mlp_ratio, self.final_upsample = ape, patch_norm, mlp_ratio, final_upsample
        self.num_features, self.num_features_up = embed_dim * 8, embed_dim * 2
        # the reference evaluates two decoder stacks whose outputs it discards (model_parts.py:794-795, 806-807)
        self.run_dead_branches = run_dead_branches
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim, norm_layer if patch_norm else None)
        pr = self.patch_embed.patches_resolution
        self.patches_resolution = pr
        if drop_rate > 0:
            raise NotImplementedError("MODEL.DROP_RATE > 0 (pos_drop / MLP / projection dropout, network/model_parts.py:606, 779) is "
                                      "not built: config.yaml:22 uses 0.0; attention dropout and stochastic depth are supported")
        self.pos_drop = nn.Dropout(p=drop_rate)          # p == 0: the identity (kept for the module tree)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        common = dict(window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                      attn_drop=attn_drop_rate, norm_layer=norm_layer, use_checkpoint=use_checkpoint)

        def res(k):
            return (pr[0] // (2 ** k), pr[1] // (2 ** k))

        def dp(k):
            return dpr[sum(depths[:k]):sum(depths[:k + 1])]

        self.layers = nn.ModuleList([
            BasicLayer(dim=embed_dim << k, input_resolution=res(k), depth=depths[k], num_heads=num_heads[k],
                       drop_path=dp(k), downsample=PatchMerging if k < 3 else None, **common) for k in range(4)])
        self.layers_up = nn.ModuleList()
        self.concat_back_dim = nn.ModuleList()
        for i in range(4):
            k = 3 - i
            self.concat_back_dim.append(nn.Linear(2 * (embed_dim << k), embed_dim << k) if i > 0 else nn.Identity())
            if i == 0:
                self.layers_up.append(PatchExpand(res(k), dim=embed_dim << k, dim_scale=2))
            else:
                self.layers_up.append(BasicLayer_up(dim=embed_dim << k, input_resolution=res(k), depth=depths[k],
                                                    num_heads=num_heads[k], drop_path=dp(k),
                                                    upsample=PatchExpand if i < 3 else None, **common))
        self.layers_cent1 = nn.ModuleList()
        for i in range(3):
            k = 2 - i
            if i == 0:
                self.layers_cent1.append(PatchExpand(res(k), dim=embed_dim << k, dim_scale=2))
            else:
                self.layers_cent1.append(BasicLayer_up(dim=embed_dim << k, input_resolution=res(k), depth=depths[k],
                                                       num_heads=num_heads[k], drop_path=dp(k),
                                                       upsample=PatchExpand if i < 2 else None, **common))
        self.layers_cent2 = nn.ModuleList()
        for i in range(2):
            k = 1 - i
            if i == 0:
                self.layers_cent2.append(PatchExpand(res(k), dim=embed_dim << k, dim_scale=2))
            else:
                self.layers_cent2.append(BasicLayer_up(dim=embed_dim << k, input_resolution=res(k), depth=depths[k],
                                                       num_heads=num_heads[k], drop_path=dp(k), upsample=None, **common))
        self.norm = nn.LayerNorm(self.num_features)
        self.norm_up = nn.LayerNorm(embed_dim)
        print("---final upsample expand_first---", file=sys.stderr)
        self.up = FinalPatchExpand_X4_V2((img_size // patch_size, img_size // patch_size), dim=embed_dim, dim_scale=4)
        self.output = nn.Conv2d(embed_dim, num_classes, kernel_size=1, bias=False)
        self.apply(self._init_weights)
        print("Finished MSUNet Construktor", file=sys.stderr)

    def _init_weights(self, m):  # same policy as network/model_parts.py:757-764
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'absolute_pos_embed'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table'}

    @property
    def compute_dtype(self):
        return self.patch_embed.compute_dtype

    def set_precision(self, dtype):
        self.patch_embed.compute_dtype = _PREC[dtype] if isinstance(dtype, str) else dtype
        return self

    def _cbd(self, i, x, skip):
        l = self.concat_back_dim[i]
        return Fn.ConcatLinearFn.apply(x, skip, l.weight, l.bias)

    def forward_features(self, x):
        """Encoder + the two central decoders (network/model_parts.py:775-815)."""
        x = self.patch_embed(x)
        P = x
        A1 = self.layers[0](P)
        F0 = self._cbd(3, self.layers_cent2[0](A1), P)
        if self.run_dead_branches:
            self.layers_cent2[1](F0)
        A2 = self.layers[1](A1)
        F1 = self._cbd(2, self.layers_cent1[0](A2), A1)
        F0 = self._cbd(3, self.layers_cent1[1](F1), F0)
        if self.run_dead_branches:
            self.layers_cent1[2](F0)
        A3 = self.layers[2](A2)
        A4 = self.layers[3](A3)
        x = Fn.LayerNormFn.apply(A4, self.norm.weight, self.norm.bias)
        return x, [F0, F1, A2, A3]

    def forward_up_features(self, x, x_downsample):
        for inx, layer_up in enumerate(self.layers_up):
            if inx == 0:
                x = layer_up(x)
            else:
                x = layer_up(self._cbd(inx, x, x_downsample[3 - inx]))
        return Fn.LayerNormFn.apply(x, self.norm_up.weight, self.norm_up.bias)

    def up_x4(self, x):
        H, W = self.patches_resolution
        if x.dim() == 4:
            B, H, W, C = x.shape
            x = x.reshape(B, H * W, C)
        else:
            B, L, C = x.shape
            assert L == H * W, "input features has wrong size"
        if H != W:
            raise ValueError("square inputs only")
        u = self.up
        return Fn.HeadFn.apply(x, u.expand.weight, u.refine1.weight, u.refine1.bias, u.refine2.weight, u.refine2.bias,
                               u.norm.weight, u.norm.bias, self.output.weight, B, H)

    def _draw_drop_path(self, B, device):
        """Stochastic-depth noise Bernoulli(1-p)/(1-p) per sample for every block of this forward in one shot
        (TV:ops/stochastic_depth.py:35-44 draws it block by block: ~90 tiny launches per step)."""
        inject = getattr(self, "_sd_inject", None)
        if inject is not None:
            for name, m in self.named_modules():
                if isinstance(m, SwinTransformerBlock) and m.sd_prob > 0.0:
                    n1, n2 = inject[name]           # KeyError: the injected set must cover every block that drops
                    m._sd_pool = (n1.to(device=device, dtype=torch.float32).contiguous(),
                                  n2.to(device=device, dtype=torch.float32).contiguous())
            return
        blocks = [m for m in self.modules() if isinstance(m, SwinTransformerBlock) and m.sd_prob > 0.0]
        if not blocks:
            return
        keep = getattr(self, "_sd_keep", None)
        if keep is None or keep.device != device or keep.numel() != 2 * len(blocks):
            keep = torch.tensor([1.0 - b.sd_prob for b in blocks for _ in range(2)], dtype=torch.float32, device=device)
            self._sd_keep = keep
        noise = (torch.rand(2 * len(blocks), B, device=device) < keep[:, None]).float() / keep[:, None]
        for i, b in enumerate(blocks):
            b._sd_pool = (noise[2 * i], noise[2 * i + 1])

    def inject_drop_path_noise(self, noise):
        """Parity hook: `noise` = {block module name: (attention-branch noise [B], MLP-branch noise [B])} replaces the
        Bernoulli draws of the following training forwards (the values torchvision's StochasticDepth would have drawn,
        TV:ops/stochastic_depth.py:38-42); None goes back to drawing."""
        self._sd_inject = noise
        return self

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("MSUNet (B200) runs on CUDA only: there is no CPU fallback for the hot path")
        if self.training:
            self._draw_drop_path(x.shape[0], x.device)
        if not getattr(self, "_is_replica", False):        # nn.DataParallel replicas get new parameter tensors every forward
            st = _SHADOW_STATE.get(self)
            if st is None:
                st = _SHADOW_STATE[self] = {"params": list(self.parameters())}
            # a training forward re-derives the bf16 weight shadows unconditionally (one launch): optimizers may update the
            # masters without touching their version counters; the first forward after training does the same once
            train = self.training and torch.is_grad_enabled()
            Fn.refresh_shadows(st["params"], st, force=train or st.get("dirty", False))
            st["dirty"] = train
        x, x_downsample = self.forward_features(x)
        x = self.forward_up_features(x, x_downsample)
        return self.up_x4(x)

    def freeze_encoder(self, freeze=True):
        for p in self.patch_embed.parameters():
            p.requires_grad = not freeze
        for layer in self.layers:
            for p in layer.parameters():
                p.requires_grad = not freeze

    def unfreeze_encoder(self, num_stage: int):
        n_stages = len(self.layers)
        if not (0 <= num_stage < n_stages):
            raise ValueError(f"num_stage={num_stage} out of range [0, {n_stages-1}]")
        for p in self.layers[num_stage].parameters():
            if not p.requires_grad:
                p.requires_grad_(True)
        if num_stage == 0:
            for p in self.patch_embed.parameters():
                if not p.requires_grad:
                    p.requires_grad_(True)
