"""tcgen05 GEMM path vs the fp32-accumulate SIMT engine on identical bf16 inputs (both through msu_gemm),
and vs an fp64 CPU product.  Every operand mode / output map the tensor-core kernel claims is exercised."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from semantic_segmentation_of_stylegan2_artifacts_b200 import ops as o
    from semantic_segmentation_of_stylegan2_artifacts_b200 import _lib
    _lib.lib()
    return o


def run_both(ops, fn):
    """fn(backend) -> tensor; returns (tc_result, simt_result) and asserts the TC path really ran."""
    from semantic_segmentation_of_stylegan2_artifacts_b200 import _lib
    ops.GEMM_BACKEND = 0
    try:
        a = fn()
        torch.cuda.synchronize()
        assert _lib.lib().msu_last_gemm_backend() == 1, "tcgen05 path was not taken"
        ops.GEMM_BACKEND = 1
        b = fn()
        torch.cuda.synchronize()
        assert _lib.lib().msu_last_gemm_backend() == 0
    finally:
        ops.GEMM_BACKEND = 0
    return a, b


def relmax(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(128, 96, 96), (1000, 288, 96), (4096, 384, 96), (777, 96, 384), (300, 1152, 384),
                                   (5000, 2304, 768), (256, 768, 3072), (130, 16, 64), (2048, 1536, 96)])
def test_plain_bias_gelu_residual(ops, shape):
    M, N, K = shape
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K) * 0.5).bfloat16().to(DEV)
    w = (torch.randn(N, K) * 0.1).bfloat16().to(DEV)
    bias = (torch.randn(N) * 0.1).to(DEV)
    res = torch.randn(M, N).bfloat16().to(DEV)

    def go():
        y = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        pre = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, Cpre=pre, bias=bias, act=1, R=res), M, N, K, a.device)
        return torch.stack([y.float(), pre.float()])

    tc, simt = run_both(ops, go)
    ref_pre = a.double().cpu() @ w.double().cpu().t() + bias.double().cpu()
    ref = torch.nn.functional.gelu(ref_pre) + res.double().cpu()
    assert relmax(tc[1], ref_pre) < 1e-2
    assert relmax(tc[0], ref) < 1e-2
    assert relmax(tc, simt) < 8e-3   # both round an fp32 accumulator to bf16


def test_dual_source_concat(ops):
    for (M, C) in [(1000, 96), (640, 192), (300, 32)]:
        torch.manual_seed(C)
        x = torch.randn(M, C).bfloat16().to(DEV)
        s = torch.randn(M, C).bfloat16().to(DEV)
        w = (torch.randn(C, 2 * C) * 0.1).bfloat16().to(DEV)
        b = torch.randn(C).to(DEV) * 0.1

        def go():
            y = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
            ops.gemm(ops.operand(x, t2=s, ld2=C, k_split=C), ops.operand(w), ops.epilogue(y, bias=b), M, C, 2 * C, x.device)
            return y.float()

        tc, simt = run_both(ops, go)
        ref = torch.cat([x, s], -1).double().cpu() @ w.double().cpu().t() + b.double().cpu()
        assert relmax(tc, ref) < 1e-2 and relmax(tc, simt) < 8e-3


@pytest.mark.parametrize("geom", [(2, 64, 96), (1, 128, 96), (2, 32, 32), (1, 224, 96), (1, 64, 128), (2, 256, 96),
                                  (1, 128, 64), (1, 128, 128)])
def test_conv3x3_implicit(ops, geom):
    B, S, E = geom
    torch.manual_seed(S + E)
    x = torch.randn(B, S, S, E).bfloat16().to(DEV)
    wt = (torch.randn(E, E, 3, 3) * 0.05)
    bias = (torch.randn(E) * 0.1).to(DEV)
    wr = ops.prep_weight(2, wt.to(DEV), E, E, (E, 9 * E), torch.bfloat16)
    Mp = B * S * S

    def go():
        y = torch.empty(Mp, E, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(x.view(Mp, E), ld=E, map=ops.MAP_CONV3, geo=[S, S, E]), ops.operand(wr),
                 ops.epilogue(y, bias=bias), Mp, E, 9 * E, x.device)
        return y.float()

    tc, simt = run_both(ops, go)
    ref = torch.nn.functional.conv2d(x.float().cpu().permute(0, 3, 1, 2).double(), wt.bfloat16().double(),
                                     bias.double().cpu(), padding=1).permute(0, 2, 3, 1).reshape(Mp, E)
    assert relmax(tc, ref) < 1e-2 and relmax(tc, simt) < 8e-3


def test_output_maps_window_and_shuffle(ops):
    from semantic_segmentation_of_stylegan2_artifacts_b200.functional import window_geo
    torch.manual_seed(5)
    # window-reverse scatter + residual + per-sample scale (proj GEMM epilogue)
    B, H, W, C = 3, 16, 16, 96
    geo = window_geo(H, W, 3)
    nW = (geo[2] // 7) * (geo[3] // 7)
    Tw, T = B * nW * 49, B * H * W
    o = torch.randn(Tw, C).bfloat16().to(DEV)
    w = (torch.randn(C, C) * 0.1).bfloat16().to(DEV)
    res = torch.randn(T, C).bfloat16().to(DEV)
    sd = torch.tensor([2.0, 0.0, 2.0], device=DEV)

    def go():
        y = torch.zeros(T, C, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(o), ops.operand(w), ops.epilogue(y, R=res, map=ops.MAP_WINDOW, geo=geo, rowscale=sd, rps=H * W),
                 Tw, C, C, o.device)
        return y.float()

    tc, simt = run_both(ops, go)
    assert relmax(tc, simt) < 8e-3
    # depth-to-space x2 and x4 (PatchExpand / head expand) with GELU and pre-activation copy
    # (Hh = 32 / 64 with 32-aligned channel groups take the TMA-store epilogue over a 5-D view of the output)
    for (p, Cin, cc, Hh) in [(2, 192, 96, 8), (4, 96, 96, 8), (2, 96, 48, 8), (2, 192, 96, 32), (4, 96, 96, 32), (2, 128, 64, 64),
                             (4, 96, 96, 24)]:
        T = 2 * Hh * Hh
        x = torch.randn(T, Cin).bfloat16().to(DEV)
        N = p * p * cc
        w = (torch.randn(N, Cin) * 0.1).bfloat16().to(DEV)

        def go2():
            y = torch.zeros(T * p * p, cc, dtype=torch.bfloat16, device=DEV)
            pre = torch.zeros(T * p * p, cc, dtype=torch.bfloat16, device=DEV)
            ops.gemm(ops.operand(x), ops.operand(w), ops.epilogue(y, ldc=cc, Cpre=pre, act=1, map=ops.MAP_SHUFFLE, geo=[Hh, Hh, p, cc]),
                     T, N, Cin, x.device)
            return torch.stack([y.float(), pre.float()])

        tc, simt = run_both(ops, go2)
        assert relmax(tc, simt) < 8e-3
        ref = (x.double().cpu() @ w.double().cpu().t()).view(2, Hh, Hh, p, p, cc).permute(0, 1, 3, 2, 4, 5).reshape(-1, cc)
        assert relmax(tc[1], ref) < 1e-2


@pytest.mark.parametrize("shape", [(5000, 384, 96), (20000, 288, 96), (4096, 96, 384), (3000, 96, 96), (2048, 1152, 384),
                                   (1024, 768, 3072), (7777, 64, 32), (600, 192, 96), (4096, 96, 864)])
def test_wgrad_mn_major(ops, shape):
    """dW[i, j] = sum_t dY[t, i] X[t, j]: both operands MN-major for the tensor core, deterministic split-K."""
    T, I, J = shape
    torch.manual_seed(T + I)
    dy = torch.randn(T, I).bfloat16().to(DEV)
    x = torch.randn(T, J).bfloat16().to(DEV)

    def go():
        dw = torch.empty(I, J, dtype=torch.float32, device=DEV)
        ops.gemm(ops.operand(dy, orient=1), ops.operand(x, orient=1), ops.epilogue(dw, out_f32=True), I, J, T, dy.device)
        return dw

    def go_bias():   # bias gradient (column sums of dY) riding along: MsuEpilogue.colsum
        dw = torch.empty(I, J, dtype=torch.float32, device=DEV)
        db = torch.full((I,), float("nan"), dtype=torch.float32, device=DEV)
        ops.gemm(ops.operand(dy, orient=1), ops.operand(x, orient=1), ops.epilogue(dw, out_f32=True, colsum=db), I, J, T, dy.device)
        return torch.cat([dw.flatten(), db])

    ref = dy.double().cpu().t() @ x.double().cpu()
    refb = dy.double().cpu().sum(0)
    from semantic_segmentation_of_stylegan2_artifacts_b200 import _lib
    lib = _lib.lib()
    for det in (0, 1):           # 0: every split adds its tile with a TMA reduce (default); 1: workspace + fixed-order reduce kernel
        prev = lib.msu_set_deterministic(det)
        try:
            tc, simt = run_both(ops, go)
            assert relmax(tc, ref) < 2e-5
            assert relmax(tc, simt) < 2e-5
            tc2, _ = run_both(ops, go)
            tcb, simtb = run_both(ops, go_bias)
            tcb2 = run_both(ops, go_bias)[0]
        finally:
            lib.msu_set_deterministic(prev)
        assert relmax(tcb[I * J:], refb) < 2e-5 and relmax(simtb[I * J:], refb) < 2e-5
        if det:
            assert torch.equal(tc, tc2)  # fixed reduction order
            assert torch.equal(tcb[:I * J].view(I, J), tc)
            assert torch.equal(tcb, tcb2)
        else:
            assert relmax(tc, tc2) < 1e-6 and relmax(tcb, tcb2) < 1e-6 and relmax(tcb[:I * J].view(I, J), tc) < 1e-6


@pytest.mark.parametrize("shape", [(4, 4096, 96, 384), (3, 1024, 384, 96), (16, 256, 192, 192), (2, 16384, 96, 96)])
def test_wgrad_per_sample_rowscale(ops, shape):
    """dW = sum_t s[sample(t)] dY[t]^T X[t] (stochastic depth on gradient rows): split-K slabs stay inside a sample and the
    reduce scales whole partials; the fused bias gradient is scaled the same way."""
    Bn, HW, I, J = shape
    T = Bn * HW
    torch.manual_seed(HW + I)
    dy = torch.randn(T, I).bfloat16().to(DEV)
    x = torch.randn(T, J).bfloat16().to(DEV)
    sd = (torch.rand(Bn) > 0.3).float().div(0.7).to(DEV)

    def go():
        dw = torch.empty(I, J, dtype=torch.float32, device=DEV)
        db = torch.full((I,), float("nan"), dtype=torch.float32, device=DEV)
        ops.gemm(ops.operand(dy, orient=1, rowscale=sd, rps=HW), ops.operand(x, orient=1), ops.epilogue(dw, out_f32=True, colsum=db),
                 I, J, T, dy.device)
        return torch.cat([dw.flatten(), db])

    tc, simt = run_both(ops, go)
    dys = dy.double().cpu() * sd.double().cpu().repeat_interleave(HW)[:, None]
    ref = torch.cat([(dys.t() @ x.double().cpu()).flatten(), dys.sum(0)])
    assert relmax(tc, ref) < 2e-5 and relmax(simt, ref) < 2e-5
    assert relmax(tc, run_both(ops, go)[0]) < 1e-6
    from semantic_segmentation_of_stylegan2_artifacts_b200 import _lib
    prev = _lib.lib().msu_set_deterministic(1)
    try:
        d1, d2 = run_both(ops, go)[0], run_both(ops, go)[0]
    finally:
        _lib.lib().msu_set_deterministic(prev)
    assert relmax(d1, ref) < 2e-5 and torch.equal(d1, d2)


@pytest.mark.parametrize("geom", [(2, 64, 96), (1, 128, 96), (2, 32, 32), (1, 224, 96), (1, 64, 128), (3, 48, 64)])
def test_conv3x3_wgrad(ops, geom):
    """dW[co,(tap ci)] = sum_pix dZ[pix,co] X[pix+off(tap),ci] with shifted NHWC slabs straight from TMA."""
    B, S, E = geom
    torch.manual_seed(S * 7 + E)
    x = torch.randn(B, S, S, E).bfloat16().to(DEV)
    dz = torch.randn(B, S, S, E).bfloat16().to(DEV)
    Mp = B * S * S

    def go():
        dw = torch.empty(E, 9 * E, dtype=torch.float32, device=DEV)
        db = torch.full((E,), float("nan"), dtype=torch.float32, device=DEV)
        ops.gemm(ops.operand(dz.view(Mp, E), orient=1), ops.operand(x.view(Mp, E), ld=E, orient=1, map=ops.MAP_CONV3, geo=[S, S, E]),
                 ops.epilogue(dw, out_f32=True, colsum=db), E, 9 * E, Mp, x.device)
        return torch.cat([dw.flatten(), db])

    tc, simt = run_both(ops, go)
    refb = dz.double().cpu().view(Mp, E).sum(0)                       # conv bias gradient from the same kernel
    assert relmax(tc[E * 9 * E:], refb) < 2e-5 and relmax(simt[E * 9 * E:], refb) < 2e-5
    tc, simt = tc[:E * 9 * E].view(E, 9 * E), simt[:E * 9 * E].view(E, 9 * E)
    xr = x.float().cpu().permute(0, 3, 1, 2).double().requires_grad_(False)
    w = torch.zeros(E, E, 3, 3, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(xr, w, padding=1)
    y.backward(dz.float().cpu().permute(0, 3, 1, 2).double())
    ref = w.grad.permute(0, 2, 3, 1).reshape(E, 9 * E)   # [co, (tap ci)]
    assert relmax(tc, ref) < 2e-5
    assert relmax(tc, simt) < 2e-5


def test_wgrad_into_column_slice(ops):
    """concat_back_dim weight gradient: two products written into the halves of one [C, 2C] buffer."""
    T, C = 3000, 96
    torch.manual_seed(1)
    dy = torch.randn(T, C).bfloat16().to(DEV)
    x = torch.randn(T, C).bfloat16().to(DEV)
    s = torch.randn(T, C).bfloat16().to(DEV)

    def go():
        dw = torch.zeros(C, 2 * C, dtype=torch.float32, device=DEV)
        ops.gemm(ops.operand(dy, orient=1), ops.operand(x, orient=1), ops.epilogue(dw, ldc=2 * C, out_f32=True), C, C, T, dy.device)
        ops.gemm(ops.operand(dy, orient=1), ops.operand(s, orient=1), ops.epilogue(dw, ldc=2 * C, out_f32=True, offset=C), C, C, T, dy.device)
        return dw

    tc, simt = run_both(ops, go)
    ref = dy.double().cpu().t() @ torch.cat([x, s], -1).double().cpu()
    assert relmax(tc, ref) < 2e-5 and relmax(tc, simt) < 2e-5


@pytest.mark.parametrize("geom", [(2, 14, 14, 3, 1), (3, 16, 16, 3, 2), (1, 16, 16, 0, 3), (5, 7, 7, 3, 1), (2, 35, 35, 3, 12),
                                  (1, 10, 12, 3, 6), (16, 21, 21, 3, 24)])
def test_window_attention_fwd_tcgen05(ops, geom):
    """tcgen05 attention core (two windows per M=128 MMA, P staged block-diagonally) vs the fp32-FMA kernel and
    vs an fp64 softmax(QK^T*s + bias + mask) V on the same bf16 inputs."""
    from semantic_segmentation_of_stylegan2_artifacts_b200 import _lib
    from semantic_segmentation_of_stylegan2_artifacts_b200.functional import window_geo
    B, H, W, shift, nH = geom
    C = nH * 32
    geo = window_geo(H, W, shift)
    nW = (geo[2] // 7) * (geo[3] // 7)
    nwin = B * nW
    torch.manual_seed(H * 13 + nH)
    qkv = (torch.randn(nwin * 49, 3 * C) * 1.5).bfloat16().to(DEV)
    table = (torch.randn(169, nH) * 0.5).to(DEV)
    bias = ops.relbias_expand(table, nH)
    lib = _lib.lib()
    lib.msu_set_attn_backend(0)
    o_tc = ops.winattn_fwd(qkv, bias, nwin, nH, geo).float().cpu()
    lib.msu_set_attn_backend(1)
    o_simt = ops.winattn_fwd(qkv, bias, nwin, nH, geo).float().cpu()
    lib.msu_set_attn_backend(0)
    # fp64 reference with the mask rebuilt from the geometry (oracle.window_geometry)
    from oracle import msunet_oracle as O
    _, _, sh, sw, _, region = O.window_geometry(H, W, shift)
    x = qkv.double().cpu().view(B, nW, 49, 3, nH, 32).permute(3, 0, 1, 4, 2, 5)
    q, k, v = x[0] * 32 ** -0.5, x[1], x[2]
    att = q @ k.transpose(-1, -2) + bias.double().cpu()[None, None]
    if sh + sw > 0:
        reg = region.view(nW, 49)
        att = att + torch.where(reg[:, :, None] != reg[:, None, :], -100.0, 0.0).double()[None, :, None]
    ref = (att.softmax(-1) @ v).permute(0, 1, 3, 2, 4).reshape(nwin * 49, C)
    assert relmax(o_simt, ref) < 1.2e-2
    assert relmax(o_tc, ref) < 1.5e-2
    assert float((o_tc.double() - ref).norm() / ref.norm()) < 6e-3


@pytest.mark.parametrize("geom", [(2, 14, 14, 3, 1, 0.05), (3, 16, 16, 3, 2, 0.2), (2, 35, 35, 3, 12, 0.05), (1, 10, 12, 0, 6, 0.5)])
def test_window_attention_dropout(ops, geom):
    """Attention dropout (TV:models/swin_transformer.py:205) fused into both attention backends: forward and backward vs an
    fp64 autograd reference that applies the identical counter-hash mask (oracle.attn_drop_keep), plus the drop rate."""
    from semantic_segmentation_of_stylegan2_artifacts_b200 import _lib
    from semantic_segmentation_of_stylegan2_artifacts_b200.functional import window_geo
    from oracle import msunet_oracle as O
    B, H, W, shift, nH, p = geom
    C = nH * 32
    geo = window_geo(H, W, shift)
    nW = (geo[2] // 7) * (geo[3] // 7)
    nwin = B * nW
    torch.manual_seed(H * 19 + nH)
    qkv = (torch.randn(nwin * 49, 3 * C) * 1.2).bfloat16().to(DEV)
    do = torch.randn(nwin * 49, C).bfloat16().to(DEV)
    table = (torch.randn(169, nH) * 0.5).to(DEV)
    bias = ops.relbias_expand(table, nH)
    seed = torch.tensor([123456789, -987654321], dtype=torch.int32, device=DEV)
    keep = O.attn_drop_keep(nwin, nH, p, 123456789, -987654321)
    assert abs(1.0 - keep.double().mean().item() - p) < 0.02
    lib = _lib.lib()
    res = {}
    for name, be in (("tc", 0), ("simt", 1)):
        lib.msu_set_attn_backend(be)
        o = ops.winattn_fwd(qkv, bias, nwin, nH, geo, p, seed)
        dqkv, dtable = ops.winattn_bwd(qkv, bias, o, do, nwin, nH, geo, p, seed)
        res[name] = (o.float().cpu(), dqkv.float().cpu(), dtable.cpu())
    lib.msu_set_attn_backend(0)
    o_nodrop = ops.winattn_fwd(qkv, bias, nwin, nH, geo).float().cpu()
    _, _, sh, sw, _, region = O.window_geometry(H, W, shift)
    x = qkv.double().cpu().requires_grad_(True)
    tb = table.double().cpu().requires_grad_(True)
    xx = x.view(B, nW, 49, 3, nH, 32).permute(3, 0, 1, 4, 2, 5)
    q, k, v = xx[0] * 32 ** -0.5, xx[1], xx[2]
    bexp = tb[O.relative_position_index()].view(49, 49, nH).permute(2, 0, 1)
    att = q @ k.transpose(-1, -2) + bexp[None, None]
    if sh + sw > 0:
        reg = region.view(nW, 49)
        att = att + torch.where(reg[:, :, None] != reg[:, None, :], -100.0, 0.0).double()[None, :, None]
    pm = att.softmax(-1) * keep.view(B, nW, nH, 49, 49).double() / (1.0 - p)
    out = (pm @ v).permute(0, 1, 3, 2, 4).reshape(nwin * 49, C)
    out.backward(do.double().cpu())

    def rl2(a, b):
        return float((a.double() - b).norm() / b.norm())

    ref = out.detach()
    assert rl2(o_nodrop, ref) > 0.05                                  # the mask really changes the output
    assert rl2(res["simt"][0], ref) < 6e-3 and rl2(res["tc"][0], ref) < 8e-3
    assert rl2(res["simt"][1], x.grad) < 6e-3 and rl2(res["simt"][2], tb.grad) < 2e-3
    assert rl2(res["tc"][1], x.grad) < 1.2e-2 and rl2(res["tc"][2], tb.grad) < 1.2e-2


@pytest.mark.parametrize("geom", [(2, 14, 14, 3, 1), (3, 16, 16, 3, 2), (1, 16, 16, 0, 3), (5, 7, 7, 3, 1), (2, 35, 35, 3, 12),
                                  (1, 10, 12, 3, 6), (16, 21, 21, 3, 24)])
def test_window_attention_bwd_tcgen05(ops, geom):
    """tcgen05 attention backward (5 tensor-core products, P recomputed) vs autograd through an fp64 reference."""
    from semantic_segmentation_of_stylegan2_artifacts_b200 import _lib
    from semantic_segmentation_of_stylegan2_artifacts_b200.functional import window_geo
    from oracle import msunet_oracle as O
    B, H, W, shift, nH = geom
    C = nH * 32
    geo = window_geo(H, W, shift)
    nW = (geo[2] // 7) * (geo[3] // 7)
    nwin = B * nW
    torch.manual_seed(H * 17 + nH)
    qkv = (torch.randn(nwin * 49, 3 * C) * 1.2).bfloat16().to(DEV)
    do = torch.randn(nwin * 49, C).bfloat16().to(DEV)
    table = (torch.randn(169, nH) * 0.5).to(DEV)
    bias = ops.relbias_expand(table, nH)
    lib = _lib.lib()
    res = {}
    for name, be in (("tc", 0), ("simt", 1)):
        lib.msu_set_attn_backend(be)
        o = ops.winattn_fwd(qkv, bias, nwin, nH, geo)
        dqkv, dtable = ops.winattn_bwd(qkv, bias, o, do, nwin, nH, geo)
        res[name] = (dqkv.float().cpu(), dtable.cpu())
    lib.msu_set_attn_backend(0)
    # default: every CTA adds its share of the bias-table gradient into the table with atomics (order not fixed);
    # msu_set_deterministic(1): per-CTA partial slabs + fixed-order reduce kernels, bit-reproducible
    dq2, dt2 = ops.winattn_bwd(qkv, bias, ops.winattn_fwd(qkv, bias, nwin, nH, geo), do, nwin, nH, geo)
    assert torch.equal(dq2.float().cpu(), res["tc"][0])
    assert float((dt2.cpu() - res["tc"][1]).abs().max()) <= 1e-5 * float(res["tc"][1].abs().max())
    assert lib.msu_winattn_bwd_direct(_lib.dt(qkv)) == 1
    prev = lib.msu_set_deterministic(1)
    try:
        assert lib.msu_winattn_bwd_direct(_lib.dt(qkv)) == 0
        o_d = ops.winattn_fwd(qkv, bias, nwin, nH, geo)
        dq3, dt3 = ops.winattn_bwd(qkv, bias, o_d, do, nwin, nH, geo)
        dq4, dt4 = ops.winattn_bwd(qkv, bias, o_d, do, nwin, nH, geo)
    finally:
        lib.msu_set_deterministic(prev)
    assert torch.equal(dt3, dt4) and torch.equal(dq3, dq4)
    assert float((dt3.cpu() - res["tc"][1]).abs().max()) <= 1e-5 * float(res["tc"][1].abs().max())
    # fp64 autograd reference
    _, _, sh, sw, _, region = O.window_geometry(H, W, shift)
    x = qkv.double().cpu().requires_grad_(True)
    tb = table.double().cpu().requires_grad_(True)
    xx = x.view(B, nW, 49, 3, nH, 32).permute(3, 0, 1, 4, 2, 5)
    q, k, v = xx[0] * 32 ** -0.5, xx[1], xx[2]
    bexp = tb[O.relative_position_index()].view(49, 49, nH).permute(2, 0, 1)
    att = q @ k.transpose(-1, -2) + bexp[None, None]
    if sh + sw > 0:
        reg = region.view(nW, 49)
        att = att + torch.where(reg[:, :, None] != reg[:, None, :], -100.0, 0.0).double()[None, :, None]
    out = (att.softmax(-1) @ v).permute(0, 1, 3, 2, 4).reshape(nwin * 49, C)
    out.backward(do.double().cpu())

    def rl2(a, b):
        return float((a.double() - b).norm() / b.norm())

    assert rl2(res["simt"][0], x.grad) < 6e-3 and rl2(res["simt"][1], tb.grad) < 2e-3
    assert rl2(res["tc"][0], x.grad) < 1.2e-2, rl2(res["tc"][0], x.grad)
    assert rl2(res["tc"][1], tb.grad) < 6e-3, rl2(res["tc"][1], tb.grad)
    assert relmax(res["tc"][0], x.grad) < 4e-2
    # the training path: the forward hands its rows' log-sum-exp to the backward (P = 2^(l - lse): no row maximum / sum there)
    o_l, lse = ops.winattn_fwd(qkv, bias, nwin, nH, geo, want_lse=True)
    lref = torch.logsumexp(att.detach(), -1).permute(0, 1, 3, 2).reshape(nwin * 49, nH) * 1.4426950408889634   # [rows, nH], log2
    assert float((lse.double().cpu() - lref).abs().max()) < 2e-2
    dq3, dt3 = ops.winattn_bwd(qkv, bias, o_l, do, nwin, nH, geo, lse=lse)
    assert rl2(dq3.float().cpu(), x.grad) < 1.2e-2 and rl2(dt3.cpu(), tb.grad) < 6e-3
    assert relmax(dq3.float().cpu(), x.grad) < 4e-2
    lib.msu_set_attn_backend(1)                # the fp32-FMA kernels write the same statistic
    try:
        _, lse_s = ops.winattn_fwd(qkv, bias, nwin, nH, geo, want_lse=True)
    finally:
        lib.msu_set_attn_backend(0)
    assert float((lse_s.double().cpu() - lref).abs().max()) < 2e-2


def test_gelu_grad_epilogue(ops):
    torch.manual_seed(9)
    M, N, K = 900, 384, 96
    dy = torch.randn(M, K).bfloat16().to(DEV)
    wT = (torch.randn(N, K) * 0.1).bfloat16().to(DEV)
    h = torch.randn(M, N).bfloat16().to(DEV)

    def go():
        y = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(dy), ops.operand(wT), ops.epilogue(y, H=h, ldh=N), M, N, K, dy.device)
        return y.float()

    tc, simt = run_both(ops, go)
    assert relmax(tc, simt) < 8e-3


@pytest.mark.parametrize("shape", [(900, 96, 384), (4096, 192, 768), (333, 288, 96), (70000, 96, 96)])
def test_tma_epilogue_residual_rowscale(ops, shape):
    """(acc + bias) * per-sample scale + residual through the TMA-slab epilogue (fc2 / unmapped proj form)."""
    M, N, K = shape
    torch.manual_seed(M + N)
    a = (torch.randn(M, K) * 0.5).bfloat16().to(DEV)
    w = (torch.randn(N, K) * 0.1).bfloat16().to(DEV)
    bias = (torch.randn(N) * 0.1).to(DEV)
    res = torch.randn(M, N).bfloat16().to(DEV)
    rps = 128
    sd = (torch.rand((M + rps - 1) // rps) > 0.3).float().to(DEV) * 1.25

    def go():
        y = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, bias=bias, R=res, rowscale=sd, rps=rps), M, N, K, a.device)
        return y.float()

    tc, simt = run_both(ops, go)
    rows = torch.arange(M) // rps
    ref = (a.double().cpu() @ w.double().cpu().t() + bias.double().cpu()) * sd.double().cpu()[rows][:, None] + res.double().cpu()
    assert relmax(tc, ref) < 1e-2 and relmax(tc, simt) < 8e-3


@pytest.mark.parametrize("shape", [(1000, 384, 96), (16384, 1536, 384), (300, 64, 64)])
def test_tma_epilogue_preact_gelu_and_gelu_grad(ops, shape):
    """fc1 form (bias + GELU with the pre-activation copy) and the dh form (x GELU'(h) x per-sample scale)."""
    M, N, K = shape
    torch.manual_seed(N + K)
    a = (torch.randn(M, K) * 0.5).bfloat16().to(DEV)
    w = (torch.randn(N, K) * 0.1).bfloat16().to(DEV)
    bias = (torch.randn(N) * 0.1).to(DEV)
    h = torch.randn(M, N).bfloat16().to(DEV)
    sd = (torch.rand((M + 63) // 64) > 0.3).float().to(DEV) * 1.25

    def fc1():
        y = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        pre = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, Cpre=pre, bias=bias, act=1), M, N, K, a.device)
        return torch.stack([y.float(), pre.float()])

    tc, simt = run_both(ops, fc1)
    ref_pre = a.double().cpu() @ w.double().cpu().t() + bias.double().cpu()
    assert relmax(tc[1], ref_pre) < 1e-2 and relmax(tc[0], torch.nn.functional.gelu(ref_pre)) < 1e-2
    assert relmax(tc, simt) < 8e-3
    # the activation is evaluated on the stored (bf16) pre-activation: GELU(pre) reproduces y to bf16 rounding
    assert relmax(tc[0], torch.nn.functional.gelu(tc[1].double())) < 5e-3

    def dh():
        y = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, H=h, ldh=N, rowscale=sd, rps=64), M, N, K, a.device)
        return y.float()

    tc, simt = run_both(ops, dh)
    hd = h.double().cpu()
    gp = 0.5 * (1 + torch.erf(hd / 2 ** 0.5)) + hd * torch.exp(-0.5 * hd * hd) / (2 * torch.pi) ** 0.5
    ref = (a.double().cpu() @ w.double().cpu().t()) * gp * sd.double().cpu()[torch.arange(M) // 64][:, None]
    assert relmax(tc, ref) < 1e-2 and relmax(tc, simt) < 8e-3


@pytest.mark.parametrize("geom", [(2, 128, 96), (1, 256, 96), (1, 128, 128)])
def test_conv3x3_two_row_tiles_fused_epilogues(ops, geom):
    """Whole-row conv tiles (two image rows per tile, two accumulators): bias + GELU + pre-activation copy through
    the TMA-slab epilogue, and x GELU'(h) with the inverse depth-to-space output map through the generic epilogue
    (the head's conv1 forward / conv1 dgrad forms, network/model_parts.py:468-471)."""
    B, S, E = geom
    torch.manual_seed(S * E)
    x = torch.randn(B, S, S, E).bfloat16().to(DEV)
    wt = torch.randn(E, E, 3, 3) * 0.05
    bias = (torch.randn(E) * 0.1).to(DEV)
    wr = ops.prep_weight(2, wt.to(DEV), E, E, (E, 9 * E), torch.bfloat16)
    Mp = B * S * S

    def fwd():
        y = torch.empty(Mp, E, dtype=torch.bfloat16, device=DEV)
        pre = torch.empty(Mp, E, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(x.view(Mp, E), ld=E, map=ops.MAP_CONV3, geo=[S, S, E]), ops.operand(wr),
                 ops.epilogue(y, Cpre=pre, bias=bias, act=1), Mp, E, 9 * E, x.device)
        return torch.stack([y.float(), pre.float()])

    tc, simt = run_both(ops, fwd)
    ref = torch.nn.functional.conv2d(x.float().cpu().permute(0, 3, 1, 2).double(), wt.bfloat16().double(),
                                     bias.double().cpu(), padding=1).permute(0, 2, 3, 1).reshape(Mp, E)
    assert relmax(tc[1], ref) < 1e-2 and relmax(tc[0], torch.nn.functional.gelu(ref)) < 1e-2
    assert relmax(tc, simt) < 8e-3

    r = S // 4
    T = B * r * r
    h0 = torch.randn(Mp, E).bfloat16().to(DEV)

    def dgrad():
        out = torch.zeros(T, 16 * E, dtype=torch.bfloat16, device=DEV)
        ops.gemm(ops.operand(x.view(Mp, E), ld=E, map=ops.MAP_CONV3, geo=[S, S, E]), ops.operand(wr),
                 ops.epilogue(out, ldc=16 * E, H=h0, ldh=E, map=ops.MAP_UNSHUFFLE, geo=[r, r, 4, E]), Mp, E, 9 * E, x.device)
        return out.float()

    tc, simt = run_both(ops, dgrad)
    assert relmax(tc, simt) < 8e-3


@pytest.mark.parametrize("E,r,B", [(96, 32, 2), (128, 32, 1), (64, 64, 1)])
def test_head_layernorm_dot_fused_into_conv_epilogue(ops, E, r, B):
    """Head: second 3x3 conv with the LayerNorm + 1x1 conv fused into its epilogue (MsuEpilogue.lnd_*, per-row statistics in the
    epilogue registers, CTA pairs at these shapes) against the separate conv -> msu_ln_fwd(dotw) pair: logits, the saved
    statistics and every gradient of the head (the backward consumes mean / rstd / dot statistic of either path).
    network/model_parts.py:468-476, 842-846."""
    from semantic_segmentation_of_stylegan2_artifacts_b200 import functional as Fn
    torch.manual_seed(5)
    bf = torch.bfloat16
    T = B * r * r
    x = (torch.randn(B, r * r, E, device=DEV) * 0.7).to(bf).requires_grad_(True)
    prm = [torch.randn(16 * E, E, device=DEV) * 0.08,                       # expand
           torch.randn(E, E, 3, 3, device=DEV) * 0.04, torch.randn(E, device=DEV) * 0.1,
           torch.randn(E, E, 3, 3, device=DEV) * 0.04, torch.randn(E, device=DEV) * 0.1,
           1.0 + 0.2 * torch.randn(E, device=DEV), 0.1 * torch.randn(E, device=DEV),
           torch.randn(1, E, 1, 1, device=DEV) * 0.2]
    prm = [p.requires_grad_(True) for p in prm]
    g = torch.randn(B, 1, 4 * r, 4 * r, device=DEV).to(bf)

    def run(fused):
        old = Fn._FUSED_HEAD_LN
        Fn._FUSED_HEAD_LN = fused
        try:
            for t in [x] + prm:
                t.grad = None
            out = Fn.HeadFn.apply(x, *prm, B, r)
            out.backward(g)
            torch.cuda.synchronize()
            return out.detach().float(), [t.grad.detach().float().clone() for t in [x] + prm]
        finally:
            Fn._FUSED_HEAD_LN = old

    lo_f, gr_f = run(True)
    lo_u, gr_u = run(False)
    # same bf16 rows in, fp32 statistics either way: the logits differ by fp32 summation order and one bf16 rounding
    assert relmax(lo_f, lo_u) < 1.0e-2
    assert float((lo_f - lo_u).abs().mean() / lo_u.abs().mean()) < 1.5e-3
    for a, b in zip(gr_f, gr_u):
        assert float((a - b).norm() / b.norm().clamp_min(1e-20)) < 5e-3
    # inference: no gradient wanted -> the rows are not stored at all, the logits are the same
    old = Fn._FUSED_HEAD_LN
    Fn._FUSED_HEAD_LN = True
    try:
        with torch.no_grad():
            lo_i = Fn.HeadFn.apply(x.detach(), *[p.detach() for p in prm], B, r).float()
    finally:
        Fn._FUSED_HEAD_LN = old
    assert torch.equal(lo_i, lo_f)


@pytest.mark.parametrize("M,N,K", [(1024, 384, 96), (640, 200, 72)])
def test_gelu_with_stored_derivative(ops, M, N, K):
    """MsuEpilogue.act = 2 (forward: C = GELU(pre), Cpre = GELU'(pre) of the dtype-rounded pre-activation) and act = 3 (backward:
    multiply by the stored derivative) on the TMA-store path (N % 32 == 0) and the generic path, tcgen05 vs the SIMT engine
    and vs fp64 (TV:ops/misc.py:292-303 forward / its autograd backward)."""
    torch.manual_seed(3)
    bf = torch.bfloat16
    a = torch.randn(M, K, device=DEV).to(bf)
    w = (torch.randn(N, K, device=DEV) * 0.2).to(bf)
    b = torch.randn(N, device=DEV) * 0.5

    def fwd():
        y = torch.empty(M, N, dtype=bf, device=DEV)
        g = torch.empty(M, N, dtype=bf, device=DEV)
        ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, Cpre=g, bias=b, act=2), M, N, K, torch.device(DEV))
        return torch.stack([y.float(), g.float()])
    tc, simt = run_both(ops, fwd)
    pre = (a.double() @ w.double().t() + b.double()).to(bf).double()       # the activation sees the rounded pre-activation
    cdf = 0.5 * (1 + torch.erf(pre / 2 ** 0.5))
    ref = torch.stack([pre * cdf, cdf + pre * torch.exp(-pre * pre / 2) / (2 * torch.pi) ** 0.5])
    # a product that lands on a bf16 rounding boundary may round the other way on the two engines: compare against fp64 with a
    # bound that allows one such flip of the pre-activation (its effect on GELU / GELU' is <= 1 ulp(bf16) of |pre|)
    assert float((tc.double().cpu() - ref.cpu()).abs().max()) < 4e-2
    assert float((tc.double().cpu() - ref.cpu()).abs().mean()) < 2.5e-3
    assert float((simt.double().cpu() - ref.cpu()).abs().mean()) < 2.5e-3
    g = tc[1].to(bf)
    dy = torch.randn(M, K, device=DEV).to(bf)                               # backward-like product: [M, K] x [N, K]^T ⊙ g

    def bwd():
        d = torch.empty(M, N, dtype=bf, device=DEV)
        ops.gemm(ops.operand(dy), ops.operand(w), ops.epilogue(d, H=g, ldh=N, act=3), M, N, K, torch.device(DEV))
        return d.float()
    tb, sb = run_both(ops, bwd)
    refb = (dy.double() @ w.double().t()) * g.double()
    assert relmax(tb, refb) < 1.5e-2 and relmax(sb, refb) < 1.5e-2


@pytest.mark.parametrize("M,N,K", [(640, 384, 768), (1000, 200, 1024), (129, 96, 1536)])
def test_cta_pair_gemm_with_odd_tile_count(ops, M, N, K):
    """Plain GEMMs at K >= 768 run as CTA pairs (cta_group::2): an odd number of 128-row tiles is rounded up to whole pairs (the
    extra tile loads zeros and stores nothing), M tails are clipped; bias + residual, GELU + pre-activation copy and the
    unmapped generic epilogue (N not a multiple of 32).  tcgen05 vs the SIMT engine vs fp64."""
    torch.manual_seed(11)
    bf = torch.bfloat16
    a = torch.randn(M, K, device=DEV).to(bf)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(bf)
    b = torch.randn(N, device=DEV)
    r = torch.randn(M, N, device=DEV).to(bf)
    guard = torch.full((256, N), 7.0, device=DEV).to(bf)      # rows past M must stay untouched

    def resid():
        y = torch.cat([torch.empty(M, N, dtype=bf, device=DEV), guard])
        ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y[:M], bias=b, R=r), M, N, K, torch.device(DEV))
        assert torch.equal(y[M:], guard)
        return y[:M].float()
    t, s_ = run_both(ops, resid)
    ref = a.double() @ w.double().t() + b.double() + r.double()
    assert relmax(t, ref) < 1.5e-2 and relmax(s_, ref) < 1.5e-2 and relmax(t, s_) < 1.5e-2

    def act():
        y = torch.empty(M, N, dtype=bf, device=DEV)
        pre = torch.empty(M, N, dtype=bf, device=DEV)
        ops.gemm(ops.operand(a), ops.operand(w), ops.epilogue(y, Cpre=pre, bias=b, act=1), M, N, K, torch.device(DEV))
        return torch.stack([y.float(), pre.float()])
    t, s_ = run_both(ops, act)
    pre = (a.double() @ w.double().t() + b.double())
    assert relmax(t[1], pre) < 1.5e-2 and relmax(t[0], s_[0]) < 3e-2
