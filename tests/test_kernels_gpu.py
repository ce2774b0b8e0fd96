"""GPU parity of the helper / attention / LayerNorm kernels (through the C ABI) against fp32 PyTorch."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _r(shape, seed, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(dtype).cuda()


def _close(got, ref, tol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    assert err <= tol * (ref.abs().max().item() + 1e-6), f"max err {err}, ref max {ref.abs().max().item()}"


def test_stem_im2col_and_maxpool():
    from tubedetr_b200 import kernels as K
    from tubedetr_b200.gemm import gemm
    N, H, W = 2, 64, 96
    x = _r((N, 3, H, W), 1, torch.float32)
    w = _r((64, 3, 7, 7), 2, torch.float32) * 0.1
    Ho, Wo = H // 2, W // 2
    col = torch.empty(N * Ho * Wo, 192, dtype=torch.bfloat16, device="cuda")
    K.stem_im2col(x, col, N, H, W)
    wk = torch.empty(64, 192, dtype=torch.bfloat16, device="cuda")
    K.prep_weight(w, wk, None, None, 64, 3, 49, 192)
    y = torch.empty(N * Ho * Wo, 64, dtype=torch.bfloat16, device="cuda")
    gemm(col, wk, y, N * Ho * Wo, 64, 192, relu=True)
    ref = F.relu(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), stride=2, padding=3))
    _close(y.view(N, Ho, Wo, 64).permute(0, 3, 1, 2), ref, 1e-2)
    Hp, Wp = (Ho - 1) // 2 + 1, (Wo - 1) // 2 + 1
    z = torch.empty(N, Hp, Wp, 64, dtype=torch.bfloat16, device="cuda")
    K.maxpool3x3s2(y, z, N, Ho, Wo, 64)
    refp = F.max_pool2d(y.view(N, Ho, Wo, 64).permute(0, 3, 1, 2).float(), 3, 2, 1)
    assert torch.equal(z.permute(0, 3, 1, 2).float(), refp)


@pytest.mark.parametrize("N,H,W", [(2, 96, 96), (1, 352, 352), (3, 64, 128), (1, 75, 101), (2, 224, 224), (1, 37, 33)])
def test_stem_fused_matches_torch_and_unfused(N, H, W):
    """conv 7x7/2 + FrozenBN + ReLU + maxpool 3x3/2 in ONE kernel (tdb_stem.cu: implicit im2col in shared memory, tcgen05, pooled
    epilogue) vs fp32 torch on the same bf16-rounded inputs / weights, and vs the unfused im2col + GEMM + maxpool kernels"""
    from tubedetr_b200 import kernels as K
    from tubedetr_b200.gemm import gemm
    x = _r((N, 3, H, W), 11, torch.float32)
    w = _r((64, 3, 7, 7), 12, torch.float32) * 0.1
    scale = torch.rand(64, device="cuda") + 0.5
    shift = torch.randn(64, device="cuda") * 0.2
    H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    H2, W2 = (H1 - 1) // 2 + 1, (W1 - 1) // 2 + 1
    wk = torch.zeros(64, 192, dtype=torch.bfloat16, device="cuda")
    wk[:, :168] = F.pad(w, (0, 1)).reshape(64, 168).to(torch.bfloat16)
    out = torch.full((N * H2 * W2, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    K.stem_fused(x, wk, scale, shift, out, N, H, W)
    torch.cuda.synchronize()
    conv = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), stride=2, padding=3)
    act = F.relu(conv * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).bfloat16().float()      # the conv output is bf16 before the pool
    ref = F.max_pool2d(act, 3, 2, 1)
    got = out.view(N, H2, W2, 64).permute(0, 3, 1, 2).float()
    assert torch.isfinite(got).all()
    _close(got, ref, 1e-2)
    assert (got - ref).abs().mean().item() <= 1e-3 * ref.abs().mean().item()
    # unfused kernels: same math, different accumulation order -> equal up to one bf16 rounding of single elements
    col = torch.empty(N * H1 * W1, 192, dtype=torch.bfloat16, device="cuda")
    K.stem_im2col(x, col, N, H, W)
    wk0 = torch.empty(64, 192, dtype=torch.bfloat16, device="cuda")
    K.prep_weight(w, wk0, None, None, 64, 3, 49, 192)
    y = torch.empty(N * H1 * W1, 64, dtype=torch.bfloat16, device="cuda")
    gemm(col, wk0, y, N * H1 * W1, 64, 192, scale=scale, bias=shift, relu=True)
    z = torch.empty(N * H2 * W2, 64, dtype=torch.bfloat16, device="cuda")
    K.maxpool3x3s2(y, z, N, H1, W1, 64)
    torch.cuda.synchronize()
    _close(out, z, 1e-2)
    assert (out.float() != z.float()).float().mean().item() < 0.02
    # the pooled rows written into a column slice of a wider matrix (layer1.0's [conv2 output | block input] operand): same values
    wide = torch.full((N * H2 * W2, 128), 3.0, dtype=torch.bfloat16, device="cuda")
    K.stem_fused(x, wk, scale, shift, wide[:, 64:], N, H, W)
    torch.cuda.synchronize()
    assert torch.equal(wide[:, 64:], out) and bool((wide[:, :64] == 3).all())


@pytest.mark.parametrize("H,W", [(22, 22), (11, 13)])
def test_stride2_conv_paths(H, W):
    from tubedetr_b200 import kernels as K
    from tubedetr_b200.gemm import gemm
    N, C, Co = 3, 64, 128
    x = _r((N, H, W, C), 3)
    w = _r((Co, C, 3, 3), 4, torch.float32) * 0.1
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    col = torch.empty(N * Ho * Wo, 9 * C, dtype=torch.bfloat16, device="cuda")
    K.im2col3x3s2(x, col, N, H, W, C)
    wk = torch.empty(Co, 9 * C, dtype=torch.bfloat16, device="cuda")
    K.prep_weight(w, wk, None, None, Co, C, 9, 9 * C)
    y = torch.empty(N * Ho * Wo, Co, dtype=torch.bfloat16, device="cuda")
    gemm(col, wk, y, N * Ho * Wo, Co, 9 * C)
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.conv2d(xf, w.bfloat16().float(), stride=2, padding=1)
    _close(y.view(N, Ho, Wo, Co).permute(0, 3, 1, 2), ref, 1e-2)
    # backward data path: dcol = g @ W (MN-major B), gather back, masked
    g = _r((N * Ho * Wo, Co), 5)
    dcol = torch.empty(N * Ho * Wo, 9 * C, dtype=torch.bfloat16, device="cuda")
    gemm(g, wk, dcol, N * Ho * Wo, 9 * C, Co, b_major=1)
    dx = torch.empty(N, H, W, C, dtype=torch.bfloat16, device="cuda")
    K.col2im3x3s2_mask(dcol, x, dx, N, H, W, C)
    ref.backward(g.float().view(N, Ho, Wo, Co).permute(0, 3, 1, 2))
    _close(dx, xf.grad.permute(0, 2, 3, 1) * (x.float() > 0), 2e-2)
    # subsample / zero-upsample are exact transposes
    xs = torch.empty(N, Ho, Wo, C, dtype=torch.bfloat16, device="cuda")
    K.subsample2(x, xs, N, H, W, C)
    assert torch.equal(xs, x[:, ::2, ::2].contiguous())
    wide = torch.full((N * Ho * Wo, 3 * C), 5.0, dtype=torch.bfloat16, device="cuda")    # ... into a column slice of a wider matrix
    K.subsample2(x, wide[:, C:2 * C], N, H, W, C)
    assert torch.equal(wide[:, C:2 * C], xs.view(-1, C)) and bool((wide[:, :C] == 5).all()) and bool((wide[:, 2 * C:] == 5).all())
    up = torch.empty(N, H, W, C, dtype=torch.bfloat16, device="cuda")
    K.upsample2_zero(xs, up, N, H, W, C)
    refu = torch.zeros_like(x)
    refu[:, ::2, ::2] = xs
    assert torch.equal(up, refu)


def test_prep_weight_layouts():
    from tubedetr_b200 import kernels as K
    w = _r((64, 32, 3, 3), 6, torch.float32)
    s = torch.rand(64, device="cuda") + 0.5
    o = torch.empty(64, 9 * 32, dtype=torch.bfloat16, device="cuda")
    os_ = torch.empty_like(o)
    K.prep_weight(w, o, os_, s, 64, 32, 9, 9 * 32)
    ref = w.permute(0, 2, 3, 1).reshape(64, -1)
    assert torch.equal(o, ref.bfloat16())
    assert torch.equal(os_, (ref * s[:, None]).bfloat16())


def test_layernorm_fwd_bwd():
    from tubedetr_b200 import kernels as K
    rows, D = 777, 256
    x, r = _r((rows, D), 7, torch.float32), _r((rows, D), 8, torch.float32)
    pos = _r((rows, D), 9, torch.float32)
    gm, bt = _r((D,), 10, torch.float32) * 0.1 + 1, _r((D,), 11, torch.float32) * 0.1
    y = torch.empty(rows, D, device="cuda")
    yb = torch.empty(rows, D, dtype=torch.bfloat16, device="cuda")
    yp = torch.empty_like(yb)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    K.layernorm_fwd(x, r, gm, bt, pos, y, yb, yp, mean, rstd, rows, D, 1e-5)
    z = (x + r).requires_grad_(True)
    gmr, btr = gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
    ref = F.layer_norm(z, (D,), gmr, btr, 1e-5)
    _close(y, ref, 1e-5)
    _close(yb, ref, 1e-2)
    _close(yp, ref + pos, 1e-2)
    dy = _r((rows, D), 12, torch.float32)
    ref.backward(dy)
    dz = torch.empty(rows, D, device="cuda")
    dg, db = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
    K.layernorm_bwd(dy, x, r, gm, mean, rstd, dz, dg, db, rows, D)
    _close(dz, z.grad, 1e-4)
    _close(dg, gmr.grad, 1e-4)
    _close(db, btr.grad, 1e-4)


def _ref_attn(q, k, v, kpm, H, scale):
    B, Lq, d = q.shape
    Lk = k.shape[1]
    qh = (q * scale).view(B, Lq, H, 32).transpose(1, 2)
    kh = k.view(B, Lk, H, 32).transpose(1, 2)
    vh = v.view(B, Lk, H, 32).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    p = s.softmax(-1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, d), p


@pytest.mark.parametrize("B,Lq,Lk", [(5, 59, 59), (2, 100, 100), (12, 1, 141), (1, 200, 200), (1, 7, 180), (1, 5, 300), (1, 3, 500)])
def test_mha_fwd_bwd(B, Lq, Lk):
    from tubedetr_b200 import kernels as K
    H, d = 8, 256
    scale = 1 / math.sqrt(32)
    q, k, v = _r((B, Lq, d), 20), _r((B, Lk, d), 21), _r((B, Lk, d), 22)
    kpm = torch.zeros(B, Lk, dtype=torch.uint8, device="cuda")
    kpm[:, Lk - Lk // 5:] = 1
    kpm[0] = 0
    o = torch.empty(B * Lq, d, dtype=torch.bfloat16, device="cuda")
    p = torch.empty(B, H, Lq, Lk, device="cuda")
    pbar = torch.empty(B, Lq, Lk, device="cuda")
    K.mha_fwd_cuda_core(q.view(-1, d), k.view(-1, d), v.view(-1, d), kpm, o, p, pbar, B, H, Lq, Lk, scale)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ro, rp = _ref_attn(qf, kf, vf, kpm, H, scale)
    _close(o.view(B, Lq, d), ro, 1e-2)
    _close(p, rp, 1e-4)
    _close(pbar, rp.mean(1), 1e-4)
    do = _r((B, Lq, d), 23)
    dpbar = _r((B, Lq, Lk), 24, torch.float32)
    (ro * do.float()).sum().add((rp.mean(1) * dpbar).sum()).backward()
    ds = torch.empty_like(p)
    dq, dk, dv = (torch.empty(t.shape[0] * t.shape[1], d, dtype=torch.bfloat16, device="cuda") for t in (q, k, v))
    K.mha_bwd_cuda_core(q.view(-1, d), k.view(-1, d), v.view(-1, d), do.view(-1, d), p, dpbar, ds, dq, dk, dv, B, H, Lq, Lk, scale)
    _close(dq.view_as(q), qf.grad, 2e-2)
    _close(dk.view_as(k), kf.grad, 2e-2)
    _close(dv.view_as(v), vf.grad, 2e-2)
    for got, ref in ((dq.view_as(q), qf.grad), (dk.view_as(k), kf.grad), (dv.view_as(v), vf.grad)):
        a = (got.float() * ref).sum() / (ref * ref).sum()      # no systematic scaling of the gradient
        assert abs(a.item() - 1) < 5e-3, a.item()


def test_colsum_and_cast():
    from tubedetr_b200 import kernels as K
    x = _r((1000, 256), 30)
    out = torch.zeros(256, device="cuda")
    K.colsum_bf16(x, out)
    _close(out, x.float().sum(0), 1e-4)
    a, b = _r((300, 256), 31, torch.float32), _r((300, 256), 32, torch.float32)
    y = torch.empty(300, 256, dtype=torch.bfloat16, device="cuda")
    K.cast_add_bf16(a, b, y)
    assert torch.equal(y, (a + b).bfloat16())


@pytest.mark.parametrize("F,S", [(100, 141), (8, 59), (3, 43), (37, 200)])
def test_xattn_fused_matches_unfused_math(F, S):
    """fused KV-projection + cross-attention (tcgen05/TMEM) vs fp32 torch on the same bf16 operands."""
    from tubedetr_b200 import kernels as K
    d, H = 256, 8
    scale = 1 / math.sqrt(32)
    q = _r((F, d), 40)
    mem, pos = _r((F * S, d), 41), _r((F * S, d), 42) * 0.5
    memb, mempb = mem, (mem.float() + pos.float()).bfloat16()
    W = (_r((3 * d, d), 43, torch.float32) / 16).bfloat16()
    b = _r((3 * d,), 44, torch.float32) * 0.1
    kpm = torch.zeros(F, S, dtype=torch.uint8, device="cuda")
    kpm[:, S - S // 4:] = 1
    kpm[0] = 0
    kpm[1, 1:] = 1          # a frame with a single visible key
    o = torch.empty(F, d, dtype=torch.bfloat16, device="cuda")
    p = torch.empty(F, H, 1, S, device="cuda")
    pbar = torch.empty(F, 1, S, device="cuda")
    K.xattn_fused_fwd(q, mempb, memb, W[d:], b[2 * d:], kpm, o, p, pbar, F, S, scale)
    kk = (mempb.float() @ W[d:2 * d].float().t() + b[d:2 * d]).view(F, S, d)
    vv = (memb.float() @ W[2 * d:].float().t() + b[2 * d:]).view(F, S, d)
    ro, rp = _ref_attn(q.float().view(F, 1, d), kk, vv, kpm, H, scale)
    _close(p, rp, 2e-2)
    _close(pbar, rp.mean(1), 2e-2)
    _close(o.view(F, 1, d), ro, 2e-2)
    assert torch.isfinite(o.float()).all()
    # attention dropout inside the fused kernel: same probabilities, context and head-mean use the dropped ones
    pdrop = 0.1
    keep = (torch.rand(F, H, 1, S, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9)) >= pdrop).to(torch.uint8)
    K.xattn_fused_fwd(q, mempb, memb, W[d:], b[2 * d:], kpm, o, p, pbar, F, S, scale, keep=keep, keep_scale=1 / (1 - pdrop))
    rpd = rp * keep.float() / (1 - pdrop)
    vh = vv.view(F, S, H, 32).transpose(1, 2)
    _close(p, rp, 2e-2)
    _close(pbar, rpd.mean(1), 2e-2)
    _close(o.view(F, 1, d), (rpd @ vh).transpose(1, 2).reshape(F, 1, d), 2e-2)


def test_mha_attention_dropout_fwd_bwd():
    """attention-probability dropout with a given keep mask: forward, head-mean of the DROPPED probabilities, backward"""
    from tubedetr_b200 import kernels as K
    B, H, Lq, Lk, d, pd = 3, 8, 37, 59, 256, 0.1
    scale = 1 / math.sqrt(32)
    q, k, v = _r((B, Lq, d), 50), _r((B, Lk, d), 51), _r((B, Lk, d), 52)
    kpm = torch.zeros(B, Lk, dtype=torch.uint8, device="cuda")
    kpm[:, Lk - 7:] = 1
    keep = (torch.rand(B, H, Lq, Lk, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) >= pd).to(torch.uint8)
    o = torch.empty(B * Lq, d, dtype=torch.bfloat16, device="cuda")
    p, pdrop = torch.empty(B, H, Lq, Lk, device="cuda"), torch.empty(B, H, Lq, Lk, device="cuda")
    pbar = torch.empty(B, Lq, Lk, device="cuda")
    K.mha_fwd_cuda_core(q.view(-1, d), k.view(-1, d), v.view(-1, d), kpm, o, p, pbar, B, H, Lq, Lk, scale, keep=keep, pdrop=pdrop,
              keep_scale=1 / (1 - pd))
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = (qf * scale).view(B, Lq, H, 32).transpose(1, 2)
    kh, vh = kf.view(B, Lk, H, 32).transpose(1, 2), vf.view(B, Lk, H, 32).transpose(1, 2)
    sc = (qh @ kh.transpose(-1, -2)).masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    rp = sc.softmax(-1)
    rpd = rp * keep.float() / (1 - pd)
    ro = (rpd @ vh).transpose(1, 2).reshape(B, Lq, d)
    _close(p, rp, 1e-4)
    _close(pdrop, rpd, 1e-4)
    _close(pbar, rpd.mean(1), 1e-4)
    _close(o.view(B, Lq, d), ro, 1e-2)
    do = _r((B, Lq, d), 53)
    dpbar = _r((B, Lq, Lk), 54, torch.float32)
    (ro * do.float()).sum().add((rpd.mean(1) * dpbar).sum()).backward()
    ds, pds = torch.empty_like(p), torch.empty_like(p)
    dq, dk, dv = (torch.empty(t.shape[0] * t.shape[1], d, dtype=torch.bfloat16, device="cuda") for t in (q, k, v))
    K.mha_bwd_cuda_core(q.view(-1, d), k.view(-1, d), v.view(-1, d), do.view(-1, d), p, dpbar, ds, dq, dk, dv, B, H, Lq, Lk, scale,
              keep=keep, keep_scale=1 / (1 - pd), pd_scratch=pds)
    for got, ref in ((dq.view_as(q), qf.grad), (dk.view_as(k), kf.grad), (dv.view_as(v), vf.grad)):
        _close(got, ref, 2e-2)
        a = (got.float() * ref).sum() / (ref * ref).sum()
        assert abs(a.item() - 1) < 5e-3, a.item()


@pytest.mark.parametrize("F,S,drop", [(100, 141, 0.0), (7, 59, 0.1), (3, 200, 0.1), (400, 141, 0.0)])
def test_xattn_bwd_streaming_kernel(F, S, drop):
    """one-query-per-frame attention backward (tdb_xattn_bwd) vs autograd through fp32 torch on the same bf16 operands,
    with key padding, the head-mean-probability gradient and (optionally) attention dropout"""
    from tubedetr_b200 import kernels as K
    H, d = 8, 256
    scale = 1 / math.sqrt(32)
    q, k, v = _r((F, 1, d), 60), _r((F, S, d), 61), _r((F, S, d), 62)
    kpm = torch.zeros(F, S, dtype=torch.uint8, device="cuda")
    kpm[:, S - S // 5:] = 1
    kpm[0] = 0
    keep = None
    if drop > 0:
        keep = (torch.rand(F, H, 1, S, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6)) >= drop).to(torch.uint8)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = (qf * scale).view(F, 1, H, 32).transpose(1, 2)
    kh, vh = kf.view(F, S, H, 32).transpose(1, 2), vf.view(F, S, H, 32).transpose(1, 2)
    rp = (qh @ kh.transpose(-1, -2)).masked_fill(kpm[:, None, None, :].bool(), float("-inf")).softmax(-1)
    rpd = rp * keep.float() / (1 - drop) if keep is not None else rp
    ro = (rpd @ vh).transpose(1, 2).reshape(F, 1, d)
    do = _r((F, 1, d), 63)
    dpbar = _r((F, 1, S), 64, torch.float32)
    (ro * do.float()).sum().add((rpd.mean(1) * dpbar).sum()).backward()
    p = rp.detach().contiguous()
    dq = torch.empty(F, d, dtype=torch.bfloat16, device="cuda")
    dk, dv = torch.empty(F * S, d, dtype=torch.bfloat16, device="cuda"), torch.empty(F * S, d, dtype=torch.bfloat16, device="cuda")
    K.xattn_bwd(q.view(F, d), k.view(-1, d), v.view(-1, d), do.view(F, d), p, dpbar, dq, dk, dv, F, S, scale, keep=keep,
                keep_scale=1 / (1 - drop) if drop > 0 else 1.0)
    for got, ref in ((dq.view_as(q), qf.grad), (dk.view_as(k), kf.grad), (dv.view_as(v), vf.grad)):
        _close(got, ref, 2e-2)
        a = (got.float() * ref).sum() / (ref * ref).sum()
        assert abs(a.item() - 1) < 5e-3, a.item()
    # padded keys receive exactly zero key / value gradient; no dpbar, second run bit-identical (deterministic reductions)
    assert (dk.view(F, S, d)[1:, S - S // 5:] == 0).all() and (dv.view(F, S, d)[1:, S - S // 5:] == 0).all()
    dq2, dk2, dv2 = torch.empty_like(dq), torch.empty_like(dk), torch.empty_like(dv)
    K.xattn_bwd(q.view(F, d), k.view(-1, d), v.view(-1, d), do.view(F, d), p, dpbar, dq2, dk2, dv2, F, S, scale, keep=keep,
                keep_scale=1 / (1 - drop) if drop > 0 else 1.0)
    assert torch.equal(dq, dq2) and torch.equal(dk, dk2) and torch.equal(dv, dv2)


@pytest.mark.parametrize("F,S,drop,layer", [(100, 141, 0.0, 4), (7, 59, 0.1, 0), (3, 20, 0.1, 2), (400, 141, 0.0, 5), (50, 69, 0.0, 1)])
def test_xattn_core_strided_fwd_bwd(F, S, drop, layer):
    """the decoder's default cross-attention core: one query per frame against the 256-column slice `layer` of [F*S, 6*256] key / value
    buffers (row stride 1536), forward and backward (gradients written into the same slice of strided buffers) vs fp32 torch"""
    from tubedetr_b200 import kernels as K
    H, d, nl = 8, 256, 6
    scale = 1 / math.sqrt(32)
    q = _r((F, d), 160)
    K_all, V_all = _r((F * S, nl * d), 161), _r((F * S, nl * d), 162)
    sl = slice(layer * d, (layer + 1) * d)
    kpm = torch.zeros(F, S, dtype=torch.uint8, device="cuda")
    kpm[:, S - S // 5:] = 1
    kpm[0] = 0
    keep, ks = None, 1.0
    if drop > 0:
        keep = (torch.rand(F, H, 1, S, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6)) >= drop).to(torch.uint8)
        ks = 1 / (1 - drop)
    o = torch.full((F, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    p = torch.full((F, H, 1, S), float("nan"), device="cuda")
    pbar = torch.full((F, 1, S), float("nan"), device="cuda")
    K.xattn_core_fwd(q, K_all[:, sl], V_all[:, sl], kpm, o, p, pbar, F, S, scale, keep=keep, keep_scale=ks)
    qf = q.float().view(F, 1, d).requires_grad_(True)
    kf = K_all[:, sl].float().reshape(F, S, d).requires_grad_(True)
    vf = V_all[:, sl].float().reshape(F, S, d).requires_grad_(True)
    qh = (qf * scale).view(F, 1, H, 32).transpose(1, 2)
    kh, vh = kf.view(F, S, H, 32).transpose(1, 2), vf.view(F, S, H, 32).transpose(1, 2)
    rp = (qh @ kh.transpose(-1, -2)).masked_fill(kpm[:, None, None, :].bool(), float("-inf")).softmax(-1)
    rpd = rp * keep.float() * ks if keep is not None else rp
    ro = (rpd @ vh).transpose(1, 2).reshape(F, d)
    _close(p, rp, 1e-4)
    _close(pbar, rpd.mean(1), 1e-4)
    _close(o, ro, 1e-2)
    do = _r((F, d), 163)
    dpbar = _r((F, 1, S), 164, torch.float32)
    (ro * do.float()).sum().add((rpd.mean(1) * dpbar).sum()).backward()
    dq = torch.empty(F, d, dtype=torch.bfloat16, device="cuda")
    dK_all, dV_all = torch.zeros_like(K_all), torch.zeros_like(V_all)
    K.xattn_core_bwd(q, K_all[:, sl], V_all[:, sl], do, p, dpbar, dq, dK_all[:, sl], dV_all[:, sl], F, S, scale, keep=keep, keep_scale=ks)
    for got, ref in ((dq, qf.grad.view(F, d)), (dK_all[:, sl].reshape(F, S, d), kf.grad), (dV_all[:, sl].reshape(F, S, d), vf.grad)):
        _close(got, ref, 2e-2)
        a = (got.float() * ref).sum() / (ref * ref).sum()
        assert abs(a.item() - 1) < 5e-3, a.item()
    other = torch.ones(nl * d, dtype=torch.bool, device="cuda")
    other[sl] = False
    assert (dK_all[:, other] == 0).all() and (dV_all[:, other] == 0).all()       # only the layer's slice is written


@pytest.mark.parametrize("rows,N", [(3525, 2048), (14100, 256), (1000, 768), (20, 256), (37, 64)])
def test_colsum_vector_path(rows, N):
    from tubedetr_b200 import kernels as K
    x = _r((rows, N), 70)
    out = torch.zeros(N, device="cuda")
    K.colsum_bf16(x, out)
    _close(out, x.float().sum(0), 1e-4)
    out2 = torch.zeros(N, device="cuda")
    K.colsum_bf16(x, out2)
    assert torch.equal(out, out2)


def test_dropout_mask_kernel_statistics_and_streams():
    """keep masks from the counter-based hash: keep rate 1 - p, independent across sites and seeds, reproducible for the same
    (seed, site); odd lengths (tail path)"""
    from tubedetr_b200 import kernels as K
    n, p = 1_000_003, 0.1
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    a = K.dropout_mask(torch.empty(n, dtype=torch.uint8, device="cuda"), seed, 1, p)
    a2 = K.dropout_mask(torch.empty(n, dtype=torch.uint8, device="cuda"), seed, 1, p)
    b = K.dropout_mask(torch.empty(n, dtype=torch.uint8, device="cuda"), seed, 2, p)
    seed.add_(1)
    c = K.dropout_mask(torch.empty(n, dtype=torch.uint8, device="cuda"), seed, 1, p)
    assert a.max().item() == 1 and a.min().item() == 0
    assert torch.equal(a, a2)
    sigma = math.sqrt(p * (1 - p) / n)
    for m in (a, b, c):
        assert abs(m.float().mean().item() - (1 - p)) < 5 * sigma
    # independence: P(both dropped) = p^2
    for x, y in ((a, b), (a, c)):
        both = ((x == 0) & (y == 0)).float().mean().item()
        assert abs(both - p * p) < 5 * math.sqrt(p * p * (1 - p * p) / n)
    # no short-range structure: lag-1 autocorrelation of the drop indicator ~ 0
    d = (a == 0).float()
    ac = ((d[1:] - p) * (d[:-1] - p)).mean().item() / (p * (1 - p))
    assert abs(ac) < 5 / math.sqrt(n)


def test_layernorm_fused_residual_dropout():
    """y = LN(x + dropout(r)) with the keep bits generated inside the LayerNorm kernels == the same computation with the explicit
    mask of tdb_dropout_mask(seed, site) (same hash stream), forward and backward (dz for x, dr = keep * dz / (1-p) for r)"""
    from tubedetr_b200 import kernels as K
    rows, D, p = 777, 256, 0.1
    x, r = _r((rows, D), 80, torch.float32), _r((rows, D), 81, torch.float32)
    gm, bt = _r((D,), 82, torch.float32) * 0.1 + 1, _r((D,), 83, torch.float32) * 0.1
    seed = torch.tensor([987654321], dtype=torch.int64, device="cuda")
    site = 17
    keep = K.dropout_mask(torch.empty(rows * D, dtype=torch.uint8, device="cuda"), seed, site, p).view(rows, D).float()
    assert 0.88 < keep.mean().item() < 0.92
    y, yb = torch.empty_like(x), torch.empty(rows, D, dtype=torch.bfloat16, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    K.layernorm_fwd(x, r, gm, bt, None, y, yb, None, mean, rstd, rows, D, 1e-5, drop=(seed, site, p))
    xr, rr = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    ref = F.layer_norm(xr + rr * keep / (1 - p), (D,), gm, bt, 1e-5)
    _close(y, ref, 1e-5)
    dy = _r((rows, D), 84, torch.float32)
    ref.backward(dy)
    dz, dr = torch.empty_like(x), torch.empty_like(x)
    drb = torch.empty(rows, D, dtype=torch.bfloat16, device="cuda")
    dgb = torch.empty(3 * D, device="cuda")
    K.layernorm_bwd(dy, x, r, gm, mean, rstd, dz, dgb[:D], dgb[D:2 * D], rows, D, drop=(seed, site, p), dr=dr, dr_bf=drb,
                    dbias=dgb[2 * D:])
    _close(dgb[2 * D:], rr.grad.sum(0), 1e-4)            # bias gradient of the layer that produced r, for free
    _close(dgb[D:2 * D], dy.sum(0), 1e-4)
    _close(dz, xr.grad, 1e-4)
    _close(dr, rr.grad, 1e-4)
    _close(drb, rr.grad, 1e-2)
    assert ((dr == 0) == (keep == 0)).all()              # exactly the dropped positions carry no gradient
    # a different site draws a different mask
    y2 = torch.empty_like(x)
    K.layernorm_fwd(x, r, gm, bt, None, y2, yb, None, mean, rstd, rows, D, 1e-5, drop=(seed, site + 1, p))
    assert not torch.equal(y, y2)


def test_ffn_hidden_dropout_fused_backward():
    """linear1(relu, masked_by_consumer) -> hidden_dropout -> linear2(mask_dx, dx_scale): forward equals the explicit-mask
    computation, and the gradients equal autograd through relu -> mask/(1-p) -> linear (bf16 operands, fp32 reference)"""
    from tubedetr_b200 import kernels as K
    from tubedetr_b200 import ops
    R, D, Hd, p = 384, 256, 2048, 0.1
    x = _r((R, D), 90)
    W1, b1 = (_r((Hd, D), 91, torch.float32) / 16).requires_grad_(True), (_r((Hd,), 92, torch.float32) * 0.1).requires_grad_(True)
    W2, b2 = (_r((D, Hd), 93, torch.float32) / 45).requires_grad_(True), (_r((D,), 94, torch.float32) * 0.1).requires_grad_(True)
    st = ops._drop_state(x.device)
    site = st[1] + 1                                     # the site hidden_dropout will draw
    xg = x.clone().requires_grad_(True)
    h = ops.linear(xg, W1, b1, relu=True, masked_by_consumer=True)
    hd = ops.hidden_dropout(h, p)
    out = ops.linear(hd, W2, b2, out_fp32=True, mask_dx=True, dx_scale=1 / (1 - p))
    keep = K.dropout_mask(torch.empty(R * Hd, dtype=torch.uint8, device="cuda"), st[0], site, p).view(R, Hd).float()
    s32 = (1.0 / (torch.tensor(1.0) - torch.tensor(p, dtype=torch.float32))).item()        # the kernel's fp32 1 / (1 - p)
    assert torch.equal(hd, (h.float() * (keep * s32)).bfloat16())
    dy = _r((R, D), 95, torch.float32)
    out.backward(dy)
    # fp32 reference on the same bf16-rounded ACTIVATIONS; the forward GEMMs read the weights in split precision (bf16 hi + lo =
    # the fp32 weights to 2^-16), so the reference uses the unrounded weights (its ReLU mask must agree with the kernel's)
    xr = x.float().requires_grad_(True)
    W1r, b1r, W2r, b2r = (t.detach().clone().requires_grad_(True) for t in (W1, b1, W2, b2))
    wq = (lambda w: w) if ops.WSPLIT else (lambda w: w.bfloat16().float())
    hr = torch.relu(xr @ wq(W1r).t() + b1r).bfloat16().float()
    hdr = (hr * keep / (1 - p)).bfloat16().float()
    outr = hdr @ wq(W2r).t() + b2r
    _close(out, outr, 2e-2)
    # gradient reference: straight-through the bf16 roundings
    hr2 = torch.relu(xr @ wq(W1r).t() + b1r)
    (((hr2 * keep / (1 - p)) @ wq(W2r).t() + b2r) * dy).sum().backward()
    for got, ref in ((xg.grad, xr.grad), (W1.grad, W1r.grad), (b1.grad, b1r.grad), (W2.grad, W2r.grad), (b2.grad, b2r.grad)):
        _close(got, ref, 3e-2)
        a = (got.float() * ref).sum() / (ref * ref).sum()
        assert abs(a.item() - 1) < 1e-2, a.item()


@pytest.mark.parametrize("F,S", [(100, 141), (8, 59), (3, 43), (37, 200), (1, 141)])
def test_xattn_pair_variant_is_bit_identical(F, S):
    """CTA-pair (cta_group::2) variant of the fused cross-attention kernel == the 1-CTA kernel, bit for bit (same k-block
    order into the same fp32 accumulators); odd and single tile counts exercise the idle-peer path"""
    from tubedetr_b200 import kernels as K
    from tubedetr_b200._lib import lib
    d, H = 256, 8
    scale = 1 / math.sqrt(32)
    q = _r((F, d), 40)
    mem, pos = _r((F * S, d), 41), _r((F * S, d), 42) * 0.5
    memb, mempb = mem, (mem.float() + pos.float()).bfloat16()
    W = (_r((3 * d, d), 43, torch.float32) / 16).bfloat16()
    b = _r((3 * d,), 44, torch.float32) * 0.1
    kpm = torch.zeros(F, S, dtype=torch.uint8, device="cuda")
    kpm[:, S - S // 4:] = 1
    kpm[0] = 0
    outs = []
    try:
        for pair in (0, 1):
            lib().tdb_xattn_set_pair(pair)
            o = torch.zeros(F, d, dtype=torch.bfloat16, device="cuda")
            p = torch.zeros(F, H, 1, S, device="cuda")
            pbar = torch.zeros(F, 1, S, device="cuda")
            K.xattn_fused_fwd(q, mempb, memb, W[d:], b[2 * d:], kpm, o, p, pbar, F, S, scale)
            torch.cuda.synchronize()
            outs.append((o, p, pbar))
    finally:
        lib().tdb_xattn_set_pair(0)
    assert torch.isfinite(outs[1][0].float()).all()
    for a, c in zip(outs[0], outs[1]):
        assert torch.equal(a, c)


def _ref_attn_drop(q, k, v, kpm, H, scale, keep, ks):
    """fp32 torch attention with an explicit keep mask on the probabilities (torch's attention dropout): returns o, p, p_dropped"""
    B, Lq, d = q.shape
    Lk = k.shape[1]
    qh = (q * scale).view(B, Lq, H, 32).transpose(1, 2)
    kh = k.view(B, Lk, H, 32).transpose(1, 2)
    vh = v.view(B, Lk, H, 32).transpose(1, 2)
    s_ = qh @ kh.transpose(-1, -2)
    if kpm is not None:
        s_ = s_.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    p = s_.softmax(-1)
    pd = p * keep.float() * ks if keep is not None else p
    return (pd @ vh).transpose(1, 2).reshape(B, Lq, d), p, pd


TC_SHAPES = [(25, 141, 141, 0.0), (25, 141, 141, 0.1), (1, 100, 100, 0.1), (2, 200, 200, 0.0), (2, 200, 200, 0.1), (3, 59, 59, 0.1),
             (2, 7, 33, 0.0), (50, 69, 69, 0.1), (1, 8, 8, 0.0), (2, 256, 256, 0.1), (1, 130, 192, 0.0)]


@pytest.mark.parametrize("B,Lq,Lk,drop", TC_SHAPES)
def test_mha_tc_forward_matches_fp32_torch(B, Lq, Lk, drop):
    """tcgen05 self-attention forward (QK^T and PV on tensor cores, S and O in TMEM) vs fp32 torch math on the same bf16
    operands, and vs the CUDA-core kernel"""
    from tubedetr_b200 import kernels as K
    H, d = 8, 256
    scale = 1 / math.sqrt(32)
    if Lq == Lk:                        # packed self-attention operands as the model passes them ([R, 512] = q | k)
        qk, v = _r((B * Lq, 512), 120), _r((B * Lk, d), 121)
        qv, kv = qk[:, :256], qk[:, 256:]
    else:
        qv, kv, v = _r((B * Lq, d), 120), _r((B * Lk, d), 122), _r((B * Lk, d), 121)
    kpm = torch.zeros(B, Lk, dtype=torch.uint8, device="cuda")
    kpm[:, Lk - Lk // 5:] = 1
    kpm[0] = 0
    keep, ks, dropt = None, 1.0, None
    if drop > 0:      # the kernel draws its keep bits from the hash stream (seed, site); the explicit mask of the same stream feeds the references
        seed = torch.tensor([77], dtype=torch.int64, device="cuda")
        keep = K.dropout_mask(torch.empty(B, H, Lq, Lk, dtype=torch.uint8, device="cuda"), seed, 5, drop)
        ks, dropt = 1 / (1 - drop), (seed, 5, drop)
        assert 0.8 * drop < 1 - keep.float().mean().item() < 1.2 * drop
    res = []
    for tc in (False, True):
        o = torch.zeros(B * Lq, d, dtype=torch.bfloat16, device="cuda")
        p, pd = torch.zeros(B, H, Lq, Lk, device="cuda"), (torch.zeros(B, H, Lq, Lk, device="cuda") if keep is not None else None)
        pbar = torch.zeros(B, Lq, Lk, device="cuda")
        if tc:
            K.mha_tc_fwd(qv, kv, v, kpm, o, p, pbar, B, H, Lq, Lk, scale, drop=dropt, pdrop=pd)
        else:
            K.mha_fwd_cuda_core(qv, kv, v, kpm, o, p, pbar, B, H, Lq, Lk, scale, keep=keep, pdrop=pd, keep_scale=ks)
        torch.cuda.synchronize()
        res.append((o, p, pbar))
    ro, rp, rpd = _ref_attn_drop(qv.float().reshape(B, Lq, d), kv.float().reshape(B, Lk, d), v.float().view(B, Lk, d), kpm, H, scale, keep, ks)
    _close(res[1][1], rp, 1e-4)
    _close(res[1][2], rpd.mean(1), 1e-4)
    _close(res[1][0].view(B, Lq, d), ro, 1e-2)
    _close(res[1][1], res[0][1], 1e-4)
    _close(res[1][0], res[0][0], 1e-2)


@pytest.mark.parametrize("B,Lq,Lk,drop", TC_SHAPES)
def test_mha_tc_backward_matches_fp32_autograd(B, Lq, Lk, drop):
    """tcgen05 attention backward (dP = dO V^T, dQ = dS K, dK = dS^T Q, dV = P^T dO on tensor cores) vs torch autograd in fp32 on
    the same bf16 operands, including the gradient that arrives through the head-averaged weights (guided-attention loss)"""
    from tubedetr_b200 import kernels as K
    H, d = 8, 256
    scale = 1 / math.sqrt(32)
    if Lq == Lk:
        qk, v = _r((B * Lq, 512), 130), _r((B * Lk, d), 131)
        qv, kv = qk[:, :256], qk[:, 256:]
    else:
        qv, kv, v = _r((B * Lq, d), 130), _r((B * Lk, d), 134), _r((B * Lk, d), 131)
    do = _r((B * Lq, d), 132)
    kpm = torch.zeros(B, Lk, dtype=torch.uint8, device="cuda")
    kpm[:, Lk - Lk // 5:] = 1
    kpm[0] = 0
    keep, ks, dropt = None, 1.0, None
    if drop > 0:
        seed = torch.tensor([78], dtype=torch.int64, device="cuda")
        keep = K.dropout_mask(torch.empty(B, H, Lq, Lk, dtype=torch.uint8, device="cuda"), seed, 9, drop)
        ks, dropt = 1 / (1 - drop), (seed, 9, drop)
    dpbar = _r((B, Lq, Lk), 133, torch.float32)
    qf, kf, vf = (t.float().reshape(B, -1, d).clone().requires_grad_(True) for t in (qv, kv, v))
    ro, rp, rpd = _ref_attn_drop(qf, kf, vf, kpm, H, scale, keep, ks)
    (ro * do.float().view(B, Lq, d)).sum().add((rpd.mean(1) * dpbar).sum()).backward()
    p = rp.detach().contiguous()
    if Lq == Lk:
        dqk, dv = torch.zeros_like(qk), torch.zeros_like(v)
        dq, dk = dqk[:, :256], dqk[:, 256:]
    else:
        dq, dk, dv = torch.zeros_like(qv), torch.zeros_like(kv), torch.zeros_like(v)
    K.mha_tc_bwd(qv, kv, v, do, p, dpbar, dq, dk, dv, B, H, Lq, Lk, scale, drop=dropt)
    torch.cuda.synchronize()
    for got, ref in ((dq.reshape(B, Lq, d), qf.grad), (dk.reshape(B, Lk, d), kf.grad), (dv.reshape(B, Lk, d), vf.grad)):
        _close(got, ref, 2e-2)
        a = (got.float() * ref).sum() / (ref * ref).sum()      # no systematic scaling of the gradient
        assert abs(a.item() - 1) < 5e-3, a.item()


@pytest.mark.parametrize("losses", [["boxes", "sted", "guided_attn"], ["boxes", "sted"]])
def test_fused_criterion_matches_torch_criterion(losses):
    """tdb_loss.cu (one forward kernel for all loss terms of all decoder layers, one backward kernel for all input gradients) vs the
    torch expressions of SetCriterion (pinned to the reference's own loss values on CPU, tests/test_boundary_cpu.py) on random
    ragged batches: values, and gradients under random per-term weights (a weight of exactly 0 / an unused term included)"""
    from tubedetr_b200.model import SetCriterion
    g = torch.Generator().manual_seed(23)
    fused, plain = SetCriterion(losses, sigma=1), SetCriterion(losses, sigma=1)
    fused.fused, plain.fused = True, False
    for trial in range(3):
        B, T = (3, 12 + trial) if trial < 2 else (1, 100)
        durs = [T, T - 3, T - 5][:B]
        time_mask = torch.zeros(B, T, dtype=torch.bool)
        for b, d in enumerate(durs):
            time_mask[b, :d] = True
        inter = [[1, 4], [2, durs[1] - 2 if B > 1 else 3], [0, 0]][:B] if trial < 2 else [[T // 4, 3 * T // 4]]
        keep = torch.tensor([b * T + t for b, (s, e) in enumerate(inter) for t in range(s, e + 1)]).cuda()
        K_ = keep.numel()
        tb = torch.cat([torch.rand(K_, 2, generator=g) * 0.5 + 0.25, torch.rand(K_, 2, generator=g) * 0.3 + 0.1], 1).cuda()

        def layer():
            w = torch.rand(B, T, T, generator=g).softmax(-1)
            pb = torch.cat([torch.rand(B * T, 2, generator=g) * 0.5 + 0.25, torch.rand(B * T, 2, generator=g) * 0.3 + 0.1], 1)
            return {"pred_boxes": pb.cuda().requires_grad_(True), "pred_sted": torch.randn(B, T, 2, generator=g).cuda().requires_grad_(True),
                    "weights": w.cuda().requires_grad_(True)}
        layers = [layer() for _ in range(6)]
        layers[2]["pred_boxes"].data[keep[0]] = tb[0]          # an exact hit: |x|' = 0, IoU = 1
        out = dict(layers[0], aux_outputs=layers[1:])
        o = dict(out, pred_boxes=out["pred_boxes"][keep], aux_outputs=[dict(a, pred_boxes=a["pred_boxes"][keep]) for a in out["aux_outputs"]])
        tg = [{"boxes": tb[i:i + 1]} for i in range(K_)]
        got = fused(o, tg, inter, time_mask.cuda())
        ref = plain(o, tg, inter, time_mask.cuda())
        assert set(got) == set(ref) and len(got) == (4 if "guided_attn" in losses else 3) * 6
        for k_ in ref:
            torch.testing.assert_close(got[k_], ref[k_], atol=2e-5, rtol=2e-5, msg=k_)
        wts = {k_: float(torch.rand((), generator=g) * 3) for k_ in ref}
        wts["loss_giou_1"] = 0.0
        drop = "loss_sted_3"
        leaves = [t for l in layers for n_, t in l.items() if n_ != "weights" or "guided_attn" in losses]
        ga = torch.autograd.grad(sum(got[k_] * wts[k_] for k_ in got if k_ != drop), leaves, allow_unused=True, retain_graph=True)
        gb = torch.autograd.grad(sum(ref[k_] * wts[k_] for k_ in ref if k_ != drop), leaves, allow_unused=True)
        for a, c, t in zip(ga, gb, leaves):
            a = torch.zeros_like(t) if a is None else a
            c = torch.zeros_like(t) if c is None else c
            torch.testing.assert_close(a, c, atol=2e-5 * (c.abs().max().item() + 1e-6) + 1e-7, rtol=1e-4)


def test_pos_sine_kernel_matches_reference_math():
    """tdb_pos_sine (one kernel) == PositionEmbeddingSine(128, normalize=True) of reference models/position_encoding.py:71-94 (restated in
    the oracle), with ragged padding masks"""
    from oracle import tubedetr_oracle as O
    from tubedetr_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    for N, h, w in ((5, 11, 11), (3, 7, 9), (2, 14, 5)):
        mask = torch.zeros(N, h, w, dtype=torch.bool)
        for n in range(1, N):
            mask[n, int(torch.randint(2, h + 1, (1,), generator=g)):, :] = True
            mask[n, :, int(torch.randint(2, w + 1, (1,), generator=g)):] = True
        ref = O.pos_sine(mask).permute(0, 2, 3, 1).reshape(N, h * w, 256)
        out = torch.empty(N, h * w, 256, device="cuda")
        K.pos_sine(mask.cuda().view(torch.uint8), out, N, h, w)
        assert (out.cpu() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("B,T,k,HW,L", [(1, 100, 4, 121, 20), (2, 7, 3, 25, 6), (3, 5, 5, 9, 4), (2, 8, 2, 49, 10)])
def test_encoder_glue_kernels_match_torch(B, T, k, HW, L):
    """enc_assemble / fast_mix / aggregate (tdb_glue.cu) vs the torch expressions they replace (gather, repeat_interleave, cat, add, casts),
    forward values and gradients"""
    from tubedetr_b200 import ops
    n_clips = (T + k - 1) // k
    n, S, D = B * n_clips, HW + L, 256
    g = torch.Generator().manual_seed(41)
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()
    src, txt, pos = rnd(n, HW, D).requires_grad_(True), rnd(B, L, D).requires_grad_(True), rnd(n, HW, D)
    x32, xb, xpb, pe = ops.enc_assemble(src, txt, pos, n_clips)
    r_x = torch.cat([src, txt.repeat_interleave(n_clips, 0)], 1).reshape(n * S, D)
    r_pe = torch.cat([pos, torch.zeros(n, L, D, device="cuda")], 1).reshape(n * S, D)
    assert torch.equal(x32, r_x) and torch.equal(pe, r_pe)
    assert torch.equal(xb, r_x.bfloat16()) and torch.equal(xpb, (r_x + r_pe).bfloat16())
    w1, w2, w3 = rnd(n * S, D), rnd(n * S, D).bfloat16(), rnd(n * S, D).bfloat16()
    ga = torch.autograd.grad((x32 * w1).sum() + (xb.float() * w2.float()).sum() + (xpb.float() * w3.float()).sum(), [src, txt])
    tot = (w1 + w2.float() + w3.float()).view(n, S, D)
    torch.testing.assert_close(ga[0], tot[:, :HW], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(ga[1], tot[:, HW:].reshape(B, n_clips, L, D).sum(1), atol=1e-4, rtol=1e-5)
    # fast mix + aggregate
    clip = (torch.arange(B)[:, None] * n_clips + torch.arange(T)[None] // k).flatten().cuda()
    enc = rnd(n * S, D).requires_grad_(True)
    fm = rnd(B * T * HW, D).bfloat16().requires_grad_(True)
    z = ops.fast_mix(enc, fm, B, T, k, HW, S)
    r_z = (enc.view(n, S, D)[clip][:, :HW].reshape(B * T * HW, D) + fm.float()).bfloat16()
    assert torch.equal(z, r_z)
    wz = rnd(B * T * HW, D).bfloat16()
    gz = torch.autograd.grad((z.float() * wz.float()).sum(), [enc, fm])
    rz = torch.autograd.grad(((enc.view(n, S, D)[clip][:, :HW].reshape(B * T * HW, D) + fm.float()) * wz.float()).sum(), [enc, fm])
    torch.testing.assert_close(gz[0], rz[0], atol=1e-4, rtol=1e-5)
    torch.testing.assert_close(gz[1].float(), rz[1].float(), atol=1e-6, rtol=1e-6)
    for has_upd in (True, False):
        upd = rnd(B * T * HW, D).requires_grad_(True) if has_upd else None
        mem, mem_pos, memb, mempb = ops.aggregate(enc, r_pe, upd, B, T, k, HW, S)
        r_mem = enc.view(n, S, D)[clip]
        if has_upd:
            r_mem = torch.cat([r_mem[:, :HW] + upd.view(B * T, HW, D), r_mem[:, HW:]], 1)
        r_pos = r_pe.view(n, S, D)[clip]
        assert torch.equal(mem.view(B * T, S, D), r_mem) and torch.equal(mem_pos.view(B * T, S, D), r_pos)
        assert torch.equal(memb.view(B * T, S, D), r_mem.bfloat16()) and torch.equal(mempb.view(B * T, S, D), (r_mem + r_pos).bfloat16())
        v1, v2, v3 = rnd(B * T * S, D), rnd(B * T * S, D).bfloat16(), rnd(B * T * S, D).bfloat16()
        leaves = [enc, upd] if has_upd else [enc]
        gg = torch.autograd.grad((mem * v1).sum() + (memb.float() * v2.float()).sum() + (mempb.float() * v3.float()).sum(), leaves, retain_graph=True)
        rr = torch.autograd.grad((r_mem.reshape(B * T * S, D) * (v1 + v2.float() + v3.float())).sum(), leaves, retain_graph=True)
        for a_, c_ in zip(gg, rr):
            torch.testing.assert_close(a_, c_, atol=1e-4, rtol=1e-5)
        # only the bf16 operands carry gradient (the decoder's case): fp32 gradient slot is None inside the kernel
        gg2 = torch.autograd.grad((memb.float() * v2.float()).sum() + (mempb.float() * v3.float()).sum(), leaves)
        rr2 = torch.autograd.grad((r_mem.reshape(B * T * S, D) * (v2.float() + v3.float())).sum(), leaves)
        for a_, c_ in zip(gg2, rr2):
            torch.testing.assert_close(a_, c_, atol=1e-4, rtol=1e-5)


def test_frames_preprocess_matches_torch_pipeline():
    """GPU input pipeline (tubedetr_b200/preprocess.py: resize + /255 + normalise + pad-and-pack in one kernel per clip) vs the same steps
    in torch (bilinear, half-pixel centres), ragged clips, stride-2 slow frames"""
    from tubedetr_b200 import preprocess as P
    g = torch.Generator().manual_seed(9)
    clips = [torch.randint(0, 256, (5, 72, 128, 3), generator=g, dtype=torch.uint8), torch.randint(0, 256, (4, 90, 60, 3), generator=g, dtype=torch.uint8)]
    fast, slow = P.clips_to_nested(clips, size=48, max_size=96, stride=2)
    sizes = [P.target_size(c.shape[1], c.shape[2], 48, 96) for c in clips]
    assert sizes == [(48, 85), (72, 48)]
    Hp, Wp = 72, 85
    assert fast.tensors.shape == (9, 3, Hp, Wp) and slow.tensors.shape == (5, 3, Hp, Wp)
    mean, std = torch.tensor(P.MEAN).view(1, 3, 1, 1), torch.tensor(P.STD).view(1, 3, 1, 1)
    o = 0
    for c, (h, w) in zip(clips, sizes):
        x = c.permute(0, 3, 1, 2).float()
        ref = (F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False) / 255.0 - mean) / std
        got = fast.tensors[o:o + c.shape[0]].cpu()
        assert (got[:, :, :h, :w] - ref).abs().max().item() < 2e-4
        assert (got[:, :, h:, :] == 0).all() and (got[:, :, :, w:] == 0).all()
        m = fast.mask[o:o + c.shape[0]].cpu()
        assert not m[:, :h, :w].any() and m[:, h:, :].all() and m[:, :, w:].all()
        o += c.shape[0]
    assert torch.equal(slow.tensors[:3], fast.tensors[0:5:2]) and torch.equal(slow.tensors[3:], fast.tensors[5:9:2])
