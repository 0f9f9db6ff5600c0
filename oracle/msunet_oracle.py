"""CPU oracle for the MS-UNet hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32) restatement of the reference's algorithm for
the hot path named in BASELINE.json: MSUNetSys.forward, DynamicLoss and the Dice/IoU
counting.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package never does.

Parity pinning: the reference has no golden vectors of its own, so this
restatement is pinned against the reference *itself*, imported in the build container by
oracle/make_golden.py (which writes tests/golden/*.npz).  tests/test_oracle_golden.py
replays those fixtures on any machine.

Reference citations (relative to the reference repo root; `TV:` = torchvision 0.26.0
`torchvision/`, the third-party dependency holding the Swin block arithmetic, see
network/model_parts.py:36):

* patch embed ............ network/model_parts.py:187-225
* Swin block ............. TV:models/swin_transformer.py:401-455
* shifted window attn .... TV:models/swin_transformer.py:116-228 (bias :49-56, index :272-284)
* MLP .................... TV:ops/misc.py:264-305
* PatchMerging ........... network/model_parts.py:59-97
* PatchExpand ............ network/model_parts.py:374-407
* head (X4_V2 + output) .. network/model_parts.py:437-476, 746-751, 832-848
* wiring ................. network/model_parts.py:775-855
* DynamicLoss ............ loss/DynamicLoss.py:6-52, 73-111
* metrics ................ scripts/validation_functions.py:106-108, 214-309
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

WS = 7  # window side; the reference hard-wires 7 through config.yaml:40
LN_EPS = 1e-5


@dataclass
class Cfg:
    img_size: int = 224
    embed_dim: int = 96
    depths: Tuple[int, ...] = (2, 2, 6, 2)
    num_heads: Tuple[int, ...] = (3, 6, 12, 24)
    mlp_ratio: float = 4.0
    patch_size: int = 4
    in_chans: int = 3
    num_classes: int = 1
    window: int = WS

    @property
    def res0(self) -> int:
        return self.img_size // self.patch_size


T96 = dict(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24))
B128 = dict(embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32))
T32 = dict(embed_dim=32, depths=(2, 2, 2, 2), num_heads=(1, 2, 4, 8))


# ----------------------------------------------------------------------------------------------
# parameter tree (names + shapes of MSUNetSys.state_dict(); network/model_parts.py:569-755)
# ----------------------------------------------------------------------------------------------
def _block_spec(prefix: str, C: int, nH: int, hidden: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [
        (f"{prefix}.norm1.weight", (C,), "ln_w"),
        (f"{prefix}.norm1.bias", (C,), "ln_b"),
        (f"{prefix}.attn.relative_position_bias_table", ((2 * WS - 1) ** 2, nH), "w"),
        (f"{prefix}.attn.relative_position_index", (WS ** 4,), "index"),
        (f"{prefix}.attn.qkv.weight", (3 * C, C), "w"),
        (f"{prefix}.attn.qkv.bias", (3 * C,), "b"),
        (f"{prefix}.attn.proj.weight", (C, C), "w"),
        (f"{prefix}.attn.proj.bias", (C,), "b"),
        (f"{prefix}.norm2.weight", (C,), "ln_w"),
        (f"{prefix}.norm2.bias", (C,), "ln_b"),
        (f"{prefix}.mlp.0.weight", (hidden, C), "w"),
        (f"{prefix}.mlp.0.bias", (hidden,), "b"),
        (f"{prefix}.mlp.3.weight", (C, hidden), "w"),
        (f"{prefix}.mlp.3.bias", (C,), "b"),
    ]


def _expand_spec(prefix: str, C: int):
    return [
        (f"{prefix}.expand.weight", (2 * C, C), "w"),
        (f"{prefix}.norm.weight", (C // 2,), "ln_w"),
        (f"{prefix}.norm.bias", (C // 2,), "ln_b"),
    ]


def param_spec(cfg: Cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) in the reference's registration order."""
    E, d, nh = cfg.embed_dim, cfg.depths, cfg.num_heads
    hid = lambda C: int(C * cfg.mlp_ratio)
    s: List[Tuple[str, Tuple[int, ...], str]] = []
    s += [("patch_embed.proj.weight", (E, cfg.in_chans, 4, 4), "conv"),
          ("patch_embed.proj.bias", (E,), "b"),
          ("patch_embed.norm.weight", (E,), "ln_w"),
          ("patch_embed.norm.bias", (E,), "ln_b")]
    for i in range(4):
        C = E << i
        for j in range(d[i]):
            s += _block_spec(f"layers.{i}.blocks.{j}", C, nh[i], hid(C))
        if i < 3:
            s += [(f"layers.{i}.downsample.reduction.weight", (2 * C, 4 * C), "w"),
                  (f"layers.{i}.downsample.norm.weight", (4 * C,), "ln_w"),
                  (f"layers.{i}.downsample.norm.bias", (4 * C,), "ln_b")]
    # layers_up / concat_back_dim are appended alternately (model_parts.py:627-663) but
    # live in different ModuleLists, so state_dict groups them per list.
    s += _expand_spec("layers_up.0", E << 3)
    for i in (1, 2, 3):
        k = 3 - i
        C = E << k
        for j in range(d[k]):
            s += _block_spec(f"layers_up.{i}.blocks.{j}", C, nh[k], hid(C))
        if i < 3:
            s += _expand_spec(f"layers_up.{i}.upsample", C)
    for i in (1, 2, 3):
        C = E << (3 - i)
        s += [(f"concat_back_dim.{i}.weight", (C, 2 * C), "w"),
              (f"concat_back_dim.{i}.bias", (C,), "b")]
    s += _expand_spec("layers_cent1.0", E << 2)
    for i in (1, 2):
        k = 2 - i
        C = E << k
        for j in range(d[k]):
            s += _block_spec(f"layers_cent1.{i}.blocks.{j}", C, nh[k], hid(C))
        if i < 2:
            s += _expand_spec(f"layers_cent1.{i}.upsample", C)
    s += _expand_spec("layers_cent2.0", E << 1)
    for j in range(d[0]):
        s += _block_spec(f"layers_cent2.1.blocks.{j}", E, nh[0], hid(E))
    s += [("norm.weight", (E << 3,), "ln_w"), ("norm.bias", (E << 3,), "ln_b"),
          ("norm_up.weight", (E,), "ln_w"), ("norm_up.bias", (E,), "ln_b"),
          ("up.expand.weight", (16 * E, E), "w"),
          ("up.refine1.weight", (E, E, 3, 3), "conv"), ("up.refine1.bias", (E,), "b"),
          ("up.refine2.weight", (E, E, 3, 3), "conv"), ("up.refine2.bias", (E,), "b"),
          ("up.norm.weight", (E,), "ln_w"), ("up.norm.bias", (E,), "ln_b"),
          ("output.weight", (cfg.num_classes, E, 1, 1), "conv1")]
    return s


def relative_position_index() -> torch.Tensor:
    """TV:models/swin_transformer.py:272-284 for a 7x7 window -> int64[2401]."""
    idx = torch.empty(WS * WS, WS * WS, dtype=torch.int64)
    for i in range(WS * WS):
        for j in range(WS * WS):
            dy = i // WS - j // WS + WS - 1
            dx = i % WS - j % WS + WS - 1
            idx[i, j] = dy * (2 * WS - 1) + dx
    return idx.flatten()


def make_weights(cfg: Cfg, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Deterministic, init-order-independent weights for parity tests.

    Each tensor is drawn from its own CPU generator seeded by crc32(name)+seed, so the same
    state_dict can be rebuilt on any machine without the reference.  Unlike the reference's
    init (biases 0, LN weight 1: model_parts.py:757-764) every parameter is perturbed so
    that no term of the arithmetic is silently multiplied by 0 or 1.
    """
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in param_spec(cfg):
        if kind == "index":
            sd[name] = relative_position_index()
            continue
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if kind == "ln_w":
            t = 1.0 + 0.1 * r
        elif kind in ("b", "ln_b"):
            t = 0.02 * r
        elif kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = r / math.sqrt(3.0 * fan_in)
        elif kind == "conv1":
            t = r / math.sqrt(shape[1])
        else:
            t = 0.02 * r
            if "relative_position_bias_table" in name:
                t = 0.5 * r  # make the bias matter
            elif len(shape) == 2:
                t = r / math.sqrt(shape[1]) * 0.7
        sd[name] = t.contiguous()
    return sd


def make_inputs(cfg: Cfg, batch: int, seed: int = 4321, real_last: bool = True):
    """Golden recipe inputs: x~U[0,1), sparse masks, last sample all-zero."""
    g = torch.Generator().manual_seed(seed)
    S = cfg.img_size
    x = torch.rand(batch, cfg.in_chans, S, S, generator=g)
    y = (torch.rand(batch, S, S, generator=g) > 0.9).float()
    if real_last and batch > 1:
        y[batch - 1] = 0
    return x, y


# ----------------------------------------------------------------------------------------------
# window geometry (index-math restatement of TV:models/swin_transformer.py:152-172, 193-209)
# ----------------------------------------------------------------------------------------------
def window_geometry(H: int, W: int, shift: int):
    """Returns (Ph, Pw, sh, sw, src[nW*49] flat source index or -1 for pad, region[nW*49])."""
    Ph = WS * ((H + WS - 1) // WS)
    Pw = WS * ((W + WS - 1) // WS)
    sh = 0 if WS >= Ph else shift
    sw = 0 if WS >= Pw else shift
    nwx = Pw // WS
    nW = (Ph // WS) * nwx
    src = torch.full((nW * WS * WS,), -1, dtype=torch.int64)
    region = torch.zeros(nW * WS * WS, dtype=torch.int64)

    def rho(t, P, s):
        return 0 if t < P - WS else (1 if t < P - s else 2)

    for w in range(nW):
        for i in range(WS * WS):
            ry = (w // nwx) * WS + i // WS
            rx = (w % nwx) * WS + i % WS
            py = (ry + sh) % Ph
            px = (rx + sw) % Pw
            if py < H and px < W:
                src[w * 49 + i] = py * W + px
            if sh + sw > 0:
                region[w * 49 + i] = 3 * rho(ry, Ph, sh) + rho(rx, Pw, sw)
    return Ph, Pw, sh, sw, src, region


def window_attention(x: torch.Tensor, sd, p: str, nH: int, shift: int) -> torch.Tensor:
    """x = LN1 output [B,H,W,C] -> attention branch output [B,H,W,C] (pre-residual)."""
    B, H, W, C = x.shape
    Ph, Pw, sh, sw, src, region = window_geometry(H, W, shift)
    nW = (Ph // WS) * (Pw // WS)
    xf = x.reshape(B, H * W, C)
    pad_row = torch.zeros(B, 1, C, dtype=x.dtype)
    xg = torch.cat([xf, pad_row], 1)[:, torch.where(src < 0, H * W, src)]  # pad tokens are ZERO, not masked
    xg = xg.reshape(B * nW, 49, C)
    qkv = xg @ sd[p + ".attn.qkv.weight"].t() + sd[p + ".attn.qkv.bias"]
    qkv = qkv.reshape(B * nW, 49, 3, nH, C // nH).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // nH) ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + ".attn.relative_position_bias_table"][sd[p + ".attn.relative_position_index"]]
    attn = attn + bias.view(49, 49, nH).permute(2, 0, 1).unsqueeze(0)
    if sh + sw > 0:
        reg = region.view(nW, 49)
        mask = torch.where(reg[:, :, None] != reg[:, None, :], -100.0, 0.0).to(x.dtype)
        attn = (attn.view(B, nW, nH, 49, 49) + mask[None, :, None]).view(B * nW, nH, 49, 49)
    attn = attn.softmax(-1)
    o = (attn @ v).transpose(1, 2).reshape(B * nW, 49, C)
    o = o @ sd[p + ".attn.proj.weight"].t() + sd[p + ".attn.proj.bias"]
    o = o.reshape(B, nW * 49, C)
    out = torch.zeros(B, H * W + 1, C, dtype=x.dtype)
    out = out.index_copy(1, torch.where(src < 0, H * W, src), o)  # pad outputs land in the dump row
    return out[:, : H * W].reshape(B, H, W, C)


def _lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return x


def attn_drop_keep(n_windows: int, nH: int, p: float, seed0: int, seed1: int) -> torch.Tensor:
    """Keep mask [n_windows, nH, 49, 49] (bool) of this repo's attention dropout — a restatement of the counter hash in
    csrc/common.cuh (attn_drop_hash / attn_drop_keep): one 32-bit hash per (window, head, query row, key pair), its two
    16-bit halves decide the two keys, keep iff half >= round(p * 65536).  The reference draws its mask from torch's Philox
    stream (TV:models/swin_transformer.py:205), which no other implementation can reproduce bit for bit; parity tests apply
    THIS mask to the PyTorch restatement of the attention so that forward and backward are compared element for element."""
    m = np.uint64(0xFFFFFFFF)
    thr = int(p * 65536.0 + 0.5)
    rowkey = np.arange(n_windows * nH * 49, dtype=np.uint64).reshape(n_windows, nH, 49, 1)
    jp = np.arange(25, dtype=np.uint64).reshape(1, 1, 1, 25)
    key = ((rowkey << np.uint64(5)) & m) + jp
    x = (((key + np.uint64(seed0 & 0xFFFFFFFF)) & m) * np.uint64(0x9E3779B1)) & m
    h = _lowbias32(x ^ np.uint64(seed1 & 0xFFFFFFFF))
    lo, hi = h & np.uint64(0xFFFF), h >> np.uint64(16)
    keep = np.stack([lo >= thr, hi >= thr], -1).reshape(n_windows, nH, 49, 50)[..., :49]
    return torch.from_numpy(np.ascontiguousarray(keep))


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], LN_EPS)


def _sd(branch, noise):
    """StochasticDepth(p, "row") with the noise handed in (TV:ops/stochastic_depth.py:35-44: `noise` is one
    Bernoulli(1-p)/(1-p) value per sample, broadcast over [1,1,1]; None = eval mode or p == 0)."""
    return branch if noise is None else branch * noise.to(branch.dtype).view(-1, 1, 1, 1)


def swin_block(x, sd, p: str, nH: int, shift: int, noise=None):
    """TV:models/swin_transformer.py:452-455.  `noise` = (attention-branch noise [B], MLP-branch noise [B]) or None."""
    n1, n2 = noise if noise is not None else (None, None)
    x = x + _sd(window_attention(_ln(x, sd, p + ".norm1"), sd, p, nH, shift), n1)
    h = _ln(x, sd, p + ".norm2") @ sd[p + ".mlp.0.weight"].t() + sd[p + ".mlp.0.bias"]
    h = F.gelu(h)  # exact erf GELU, TV:ops/misc.py:264-305 with nn.GELU
    return x + _sd(h @ sd[p + ".mlp.3.weight"].t() + sd[p + ".mlp.3.bias"], n2)


def stage(x, sd, p: str, depth: int, nH: int, res: int, sd_noise=None):
    B, L, C = x.shape
    x = x.view(B, res, res, C)
    for j in range(depth):
        blk = f"{p}.blocks.{j}"
        x = swin_block(x, sd, blk, nH, 0 if j % 2 == 0 else WS // 2, sd_noise.get(blk) if sd_noise else None)
    return x


def block_drop_probs(cfg: Cfg, drop_path_rate: float) -> Dict[str, float]:
    """Stochastic-depth probability of every Swin block by module name: dpr = linspace(0, rate, sum(depths))
    (network/model_parts.py:610); encoder stage k takes dpr[sum(depths[:k]) : sum(depths[:k+1])] (:618-632) and every decoder
    stack working at stage k re-uses the same slice (:648-660, :676-690, :706-720)."""
    d = cfg.depths
    dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(d))]
    out: Dict[str, float] = {}
    for prefix, k in (("layers.0", 0), ("layers.1", 1), ("layers.2", 2), ("layers.3", 3),
                      ("layers_up.1", 2), ("layers_up.2", 1), ("layers_up.3", 0),
                      ("layers_cent1.1", 1), ("layers_cent1.2", 0), ("layers_cent2.1", 0)):
        for j in range(d[k]):
            out[f"{prefix}.blocks.{j}"] = dpr[sum(d[:k]) + j]
    return out


def draw_sd_noise(cfg: Cfg, batch: int, drop_path_rate: float, seed: int = 99) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """One (attention, MLP) pair of row-mode noise vectors per block, from a private CPU generator: the values a training
    forward would draw (TV:ops/stochastic_depth.py:38-42), made reproducible so that the CUDA path and the oracle see the same."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, p in block_drop_probs(cfg, drop_path_rate).items():
        keep = 1.0 - p
        pair = []
        for _ in range(2):
            n = torch.empty(batch).bernoulli_(keep, generator=g)
            pair.append(n / keep if keep > 0 else n)
        out[name] = tuple(pair) if p > 0 else None
    return {k: v for k, v in out.items() if v is not None}


def patch_merging(x, sd, p: str):
    """x [B,H,W,C] -> [B,HW/4,2C]; neighbour order (0,0),(1,0),(0,1),(1,1); LN then Linear."""
    B, H, W, C = x.shape
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = _ln(x.reshape(B, -1, 4 * C), sd, p + ".norm")
    return x @ sd[p + ".reduction.weight"].t()


def patch_expand(x, sd, p: str, res: int):
    """x [B,L,C] or [B,H,W,C] -> [B,4L,C/2]; Linear(C,2C) -> pixel-shuffle(2) -> LN(C/2)."""
    if x.dim() == 4:
        x = x.reshape(x.shape[0], -1, x.shape[-1])
    B, L, C = x.shape
    x = x @ sd[p + ".expand.weight"].t()
    x = x.view(B, res, res, 2, 2, C // 2).permute(0, 1, 3, 2, 4, 5).reshape(B, 4 * L, C // 2)
    return _ln(x, sd, p + ".norm")


def _cbd(a, b, sd, i):
    return torch.cat([a, b], -1) @ sd[f"concat_back_dim.{i}.weight"].t() + sd[f"concat_back_dim.{i}.bias"]


def head(x, sd, cfg: Cfg):
    """norm_up output [B,L0,E] -> logits [B,1,S,S] (model_parts.py:451-476, 832-848)."""
    B, L, E = x.shape
    r = cfg.res0
    x = F.gelu(x @ sd["up.expand.weight"].t())
    x = x.view(B, r, r, 4, 4, E).permute(0, 1, 3, 2, 4, 5).reshape(B, 4 * r, 4 * r, E)
    x = x.permute(0, 3, 1, 2)
    x = F.gelu(F.conv2d(x, sd["up.refine1.weight"], sd["up.refine1.bias"], padding=1))
    x = F.conv2d(x, sd["up.refine2.weight"], sd["up.refine2.bias"], padding=1)
    x = _ln(x.permute(0, 2, 3, 1), sd, "up.norm")
    return F.conv2d(x.permute(0, 3, 1, 2), sd["output.weight"])


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: Cfg, run_dead: bool = False, sd_noise=None):
    """MSUNetSys.forward (model_parts.py:775-855).  `run_dead` also evaluates the two decoder
    stacks whose outputs the reference discards (:794-795, :806-807); logits are identical.
    `sd_noise`: {block module name: (noise1 [B], noise2 [B])} = the stochastic-depth noise of a training forward (see
    `draw_sd_noise`); None = eval mode / drop_path_rate 0."""
    if x.size(1) != 3:
        raise ValueError(f"Expected 3 channels, but got {x.size(1)}")  # network/MSUNet.py:48-51
    assert x.shape[2] == cfg.img_size and x.shape[3] == cfg.img_size
    E, d, nh, r = cfg.embed_dim, cfg.depths, cfg.num_heads, cfg.res0
    B = x.shape[0]
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=4)
    P = _ln(t.flatten(2).transpose(1, 2), sd, "patch_embed.norm")
    # encoder stage 0
    A1 = patch_merging(stage(P, sd, "layers.0", d[0], nh[0], r, sd_noise), sd, "layers.0.downsample")
    # central decoder 2
    U = patch_expand(A1, sd, "layers_cent2.0", r // 2)
    F0 = _cbd(U, P, sd, 3)
    if run_dead:
        stage(F0, sd, "layers_cent2.1", d[0], nh[0], r, sd_noise)
    A2 = patch_merging(stage(A1, sd, "layers.1", d[1], nh[1], r // 2, sd_noise), sd, "layers.1.downsample")
    # central decoder 1
    V = patch_expand(A2, sd, "layers_cent1.0", r // 4)
    F1 = _cbd(V, A1, sd, 2)
    Wd = patch_expand(stage(F1, sd, "layers_cent1.1", d[1], nh[1], r // 2, sd_noise), sd,
                      "layers_cent1.1.upsample", r // 2)
    F0b = _cbd(Wd, F0, sd, 3)
    if run_dead:
        stage(F0b, sd, "layers_cent1.2", d[0], nh[0], r, sd_noise)
    A3 = patch_merging(stage(A2, sd, "layers.2", d[2], nh[2], r // 4, sd_noise), sd, "layers.2.downsample")
    A4 = stage(A3, sd, "layers.3", d[3], nh[3], r // 8, sd_noise)
    bott = _ln(A4, sd, "norm")
    # decoder
    D0 = patch_expand(bott, sd, "layers_up.0", r // 8)
    D1 = patch_expand(stage(_cbd(D0, A2, sd, 1), sd, "layers_up.1", d[2], nh[2], r // 4, sd_noise), sd,
                      "layers_up.1.upsample", r // 4)
    D2 = patch_expand(stage(_cbd(D1, F1, sd, 2), sd, "layers_up.2", d[1], nh[1], r // 2, sd_noise), sd,
                      "layers_up.2.upsample", r // 2)
    D3 = stage(_cbd(D2, F0b, sd, 3), sd, "layers_up.3", d[0], nh[0], r, sd_noise)
    up = _ln(D3, sd, "norm_up").reshape(B, r * r, E)
    return head(up, sd, cfg)


DEAD_PREFIXES = ("layers_cent1.2.", "layers_cent2.1.")


# ----------------------------------------------------------------------------------------------
# DynamicLoss closed form (loss/DynamicLoss.py:73-111)
# ----------------------------------------------------------------------------------------------
def dynamic_loss(output: torch.Tensor, target: torch.Tensor, alpha=0.4, beta=0.6, mix=0.5,
                 smooth=1e-6) -> torch.Tensor:
    if target.dim() == 3:
        target = target.unsqueeze(1)
    target = target.float()
    if target.max() > 1:
        target = (target > 127.5).float()
    if output.size(0) != target.size(0):
        raise ValueError("batch mismatch")
    B = output.size(0)
    x = output.float().reshape(B, -1)
    t = target.reshape(B, -1)
    bce = (x.clamp_min(0) - x * t + torch.log1p(torch.exp(-x.abs()))).mean(1)
    p = torch.sigmoid(x)
    tp = (p * t).sum(1)
    fp = (p * (1 - t)).sum(1)
    fn = ((1 - p) * t).sum(1)
    tv = 1 - (tp + smooth) / (tp + alpha * fp + beta * fn + smooth)
    has_pos = t.sum(1) != 0
    per = torch.where(has_pos, (1 - mix) * bce + mix * tv, bce)
    return per.mean()


# ----------------------------------------------------------------------------------------------
# metrics (scripts/validation_functions.py:214-309; medpy.metric.binary dc/jc/precision/recall
# restated from medpy 0.4 — package absent here: bool inputs, count_nonzero, ZeroDivision -> 0.0
# for dc/precision/recall, unguarded division for jc)
# ----------------------------------------------------------------------------------------------
def confusion_counts(pred_bin: np.ndarray, gt: np.ndarray):
    pb = pred_bin.astype(bool)
    g = gt.astype(bool)
    tp = int(np.count_nonzero(pb & g))
    fp = int(np.count_nonzero(pb & ~g))
    fn = int(np.count_nonzero(~pb & g))
    tn = int(np.count_nonzero(~pb & ~g))
    return tp, fp, fn, tn


def soft_sums(pred: np.ndarray, gt: np.ndarray):
    """(TP, FP, FN, TN, sum p^2, sum g^2, sum p, sum g) in float64 (exact reference for fp32 sums)."""
    p = pred.astype(np.float64).ravel()
    g = gt.astype(bool).astype(np.float64).ravel()
    return (float((p * g).sum()), float(((1 - g) * p).sum()), float((g * (1 - p)).sum()),
            float(((1 - p) * (1 - g)).sum()), float((p * p).sum()), float((g * g).sum()),
            float(p.sum()), float(g.sum()))


def metrics_fake(pred_bin, pred, gt):
    smooth = 1e-8
    tp, fp, fn, tn = confusion_counts(pred_bin, gt)
    dice = 2.0 * tp / float(2 * tp + fp + fn) if (2 * tp + fp + fn) else 0.0
    recall = tp / float(tp + fn) if (tp + fn) else 0.0
    precision = tp / float(tp + fp) if (tp + fp) else 0.0
    iou = tp / float(tp + fp + fn)
    f1 = 2 * (precision * recall) / (precision + recall + smooth)
    acc = (tp + tn) / (tp + tn + fp + fn)
    s = soft_sums(pred, gt)
    soft_dice = (2.0 * s[0] + smooth) / (s[4] + s[5] + smooth)
    soft_iou = (s[0] + smooth) / (s[6] + s[7] - s[0] + smooth)
    return (acc, recall, precision, iou, dice, f1, [[tp, fp], [fn, tn]],
            [[s[0], s[1]], [s[2], s[3]]], soft_dice, soft_iou)


def metrics_real(pred_bin, pred, gt):
    tp, fp, fn, tn = confusion_counts(pred_bin, gt)
    s = soft_sums(pred, gt)
    return [[tp, fp], [fn, tn]], [[s[0], s[1]], [s[2], s[3]]], (tp + tn) / (tp + tn + fp + fn), fp / (fp + tn)


# ----------------------------------------------------------------------------------------------
# convenience: one fwd+loss+bwd step on CPU (used by bench.py cpu_baseline and the parity tests)
# ----------------------------------------------------------------------------------------------
def train_step(sd, x, y, cfg: Cfg, alpha=0.2, beta=0.8, mix=0.45, run_dead=False, sd_noise=None):
    leaves = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v)
              for k, v in sd.items()}
    logits = forward(leaves, x, cfg, run_dead=run_dead, sd_noise=sd_noise)
    loss = dynamic_loss(logits, y, alpha, beta, mix)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.is_floating_point()}
    return logits.detach(), loss.detach(), grads


# ----------------------------------------------------------------------------------------------
# Input staging: the tail of the per-sample transform, for a batch
# ----------------------------------------------------------------------------------------------
def stage_batch(images_u8: np.ndarray, labels_u8: np.ndarray, flips: np.ndarray):
    """dataset/dataset.py:13-16 (flip of image and label along W), :62 (`astype(float32) / 255.0`), :63 (`label > 127`),
    :83-84 (HWC -> CHW, label as float32), stacked over the batch like the DataLoader's default collate.
    images_u8 [B,H,W,3] u8, labels_u8 [B,H,W] u8, flips [B] -> float32 [B,3,H,W], float32 [B,H,W]."""
    imgs, labs = [], []
    for im, lb, f in zip(images_u8, labels_u8, flips):
        if f:
            im, lb = np.flip(im, axis=1), np.flip(lb, axis=1)
        imgs.append(np.transpose(im.astype(np.float32) / 255.0, (2, 0, 1)))
        labs.append((lb > 127).astype(np.uint8).astype(np.float32))
    return np.stack(imgs), np.stack(labs)


# ----------------------------------------------------------------------------------------------
# Pretrained-encoder interchange: synthetic SegFace / torchvision-style checkpoints
# ----------------------------------------------------------------------------------------------
def encoder_checkpoint(state_dict, root: str):
    """A checkpoint in the naming network/MSUNet.py:83-146 (SegFace, root 'backbone.0.') and :167-227 (ImageNet, root 'features.')
    expect, covering every encoder entry of `state_dict` (patch_embed.*, layers.*): stage L blocks live under `{2L+1}.{i}`, its
    downsample under `{2L+2}`, the patch embedding under `0.0` / `0.2`.  Entry number n is filled with the value n + 1, so a loaded
    model tells which checkpoint tensor ended up where.  Returns (checkpoint dict, {model key: checkpoint key})."""
    ckpt, where = {}, {}
    for k, v in state_dict.items():
        parts = k.split(".")
        if parts[0] == "patch_embed":
            nk = root + {"proj": "0.0", "norm": "0.2"}[parts[1]] + "." + ".".join(parts[2:])
        elif parts[0] == "layers" and parts[2] == "blocks":
            nk = root + f"{2 * int(parts[1]) + 1}.{parts[3]}." + ".".join(parts[4:])
        elif parts[0] == "layers" and parts[2] == "downsample":
            nk = root + f"{2 * int(parts[1]) + 2}." + ".".join(parts[3:])
        else:
            continue
        ckpt[nk] = torch.full(tuple(v.shape), float(len(ckpt) + 1), dtype=v.dtype)
        where[k] = nk
    return ckpt, where
