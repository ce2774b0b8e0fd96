"""GPU parity of the tcgen05 GEMM (through the C ABI) against plain fp32 PyTorch on the same bf16 operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(torch.bfloat16).cuda()


def _close(got, ref, tol=2e-2):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, f"max err {err} vs scale {scale}"


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (300, 128, 192, 0), (1000, 256, 512, 256), (4096, 512, 1024, 0),
                                        (77, 768, 256, 128), (20000, 64, 64, 0)])
def test_gemm_nt_plain(M, N, K, bn):
    from tubedetr_b200.gemm import gemm
    A, B = _rand((M, K), 1), _rand((N, K), 2)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(A, B, out, M, N, K, block_n=bn)
    torch.cuda.synchronize()
    _close(out, A.float() @ B.float().t())


def test_gemm_epilogue_scale_bias_residual_relu_f32():
    from tubedetr_b200.gemm import gemm
    M, N, K = 1500, 256, 320
    A, B, R = _rand((M, K), 3), _rand((N, K), 4), _rand((M, N), 5)
    scale = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    ref = torch.relu((A.float() @ B.float().t()) * scale + bias + R.float())
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    gemm(A, B, out, M, N, K, scale=scale, bias=bias, residual=R, relu=True)
    _close(out, ref)
    out32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    gemm(A, B, out32, M, N, K, scale=scale, bias=bias, residual=R, relu=True)
    _close(out32, ref, tol=1e-3)
    # mask epilogue (ReLU backward): keep where mask > 0
    mk = _rand((M, N), 6)
    outm = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    gemm(A, B, outm, M, N, K, mask=mk)
    _close(outm, (A.float() @ B.float().t()) * (mk.float() > 0))


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_b_mn_major_dgrad_form(bn):
    """dX[M,Cin] = dY[M,Cout] @ W[Cout,Cin]: W is read MN-major (reduction dim = rows)."""
    from tubedetr_b200.gemm import gemm
    M, Cout, Cin = 700, 192, 256
    dY, W = _rand((M, Cout), 7), _rand((Cout, Cin), 8)
    out = torch.empty(M, Cin, dtype=torch.bfloat16, device="cuda")
    gemm(dY, W, out, M, Cin, Cout, b_major=1, block_n=bn)
    _close(out, dY.float() @ W.float())


@pytest.mark.parametrize("splits", [1, 5])
def test_gemm_mn_mn_wgrad_form(splits):
    """dW[Cout,Cin] = dY[R,Cout]^T @ X[R,Cin]: both operands MN-major, R not a multiple of 64, split reduction."""
    from tubedetr_b200.gemm import effective_splits, gemm, splitk_reduce
    R, Cout, Cin = 1000, 256, 128
    dY, X = _rand((R, Cout), 9), _rand((R, Cin), 10)
    ref = dY.float().t() @ X.float()
    s = effective_splits(R, splits)
    part = torch.empty(s, Cout, Cin, dtype=torch.float32, device="cuda")
    gemm(dY, X, part, Cout, Cin, R, a_major=1, b_major=1, splits=splits)
    out = torch.empty(Cout, Cin, dtype=torch.float32, device="cuda")
    rs = torch.rand(Cout, device="cuda") + 0.5
    splitk_reduce(part, s, Cout, Cin, out, rowscale=rs)
    _close(out, ref * rs[:, None], tol=2e-3)


def _pad_rows(x):  # (N,H,W,C) -> zero-haloed rows [(N*(H+2)*(W+2)), C]
    N, H, W, C = x.shape
    xp = torch.zeros(N, H + 2, W + 2, C, dtype=x.dtype, device=x.device)
    xp[:, 1:-1, 1:-1] = x
    return xp.view(-1, C)


def test_gemm_implicit_conv3x3_padded_grid():
    """3x3/s1/p1 convolution as 9 row-shifted taps over the zero-haloed activation matrix, output compacted."""
    from tubedetr_b200.gemm import REMAP_P2C, gemm
    Nimg, H, W, Cin, Cout = 5, 11, 13, 64, 128
    x = _rand((Nimg, H, W, Cin), 11)
    w = _rand((Cout, Cin, 3, 3), 12) * 0.1
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1)
    xp = _pad_rows(x)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()  # [Cout][tap][Cin]
    Wp = W + 2
    taps = [(kh - 1) * Wp + (kw - 1) for kh in range(3) for kw in range(3)]
    out = torch.full((Nimg * H * W, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(xp, wk, out, xp.shape[0], Cout, Cin, ntaps=9, a_off1=taps, b_off0=[t * Cin for t in range(9)],
         remap=REMAP_P2C, img_hw=(H, W))
    _close(out.view(Nimg, H, W, Cout), ref)


def test_gemm_conv3x3_dgrad_and_wgrad_padded_grid():
    from tubedetr_b200.gemm import REMAP_C2P, REMAP_P2C, effective_splits, gemm, splitk_reduce
    Nimg, H, W, Cin, Cout = 3, 9, 10, 128, 64
    x = _rand((Nimg, H, W, Cin), 13)
    w = _rand((Cout, Cin, 3, 3), 14) * 0.1
    g = _rand((Nimg, H, W, Cout), 15)
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wf = w.float().requires_grad_(True)
    y = torch.nn.functional.conv2d(xf, wf, padding=1)
    y.backward(g.float().permute(0, 3, 1, 2))
    Wp = W + 2
    taps = [(kh - 1) * Wp + (kw - 1) for kh in range(3) for kw in range(3)]
    gp, xp = _pad_rows(g), _pad_rows(x)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    # dgrad: dX[r] = sum_tap g[r - off_tap] @ W_tap  (W read MN-major: rows = Cout = reduction)
    dx = torch.full((Nimg * H * W, Cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(gp, wk, dx, gp.shape[0], Cin, Cout, b_major=1, ntaps=9, a_off1=[-t for t in taps],
         b_off0=[t * Cin for t in range(9)], remap=REMAP_P2C, img_hw=(H, W))
    _close(dx.view(Nimg, H, W, Cin), xf.grad.permute(0, 2, 3, 1))
    # wgrad: dW[:, tap, :] = sum_r g[r]^T x[r + off_tap]
    R = gp.shape[0]
    s = effective_splits(R, 3)
    part = torch.empty(s, Cout, 9 * Cin, dtype=torch.float32, device="cuda")
    gemm(gp, xp, part, Cout, Cin, R, a_major=1, b_major=1, nz=9, z_b_off1=taps, z_out_col=[t * Cin for t in range(9)],
         splits=3)
    dw = torch.empty(Cout, Cin, 3, 3, dtype=torch.float32, device="cuda")
    splitk_reduce(part, s, Cout, 9 * Cin, dw, taps=9)
    _close(dw, wf.grad, tol=5e-3)
    # compact -> padded remap (1x1 conv writing into a zero-haloed buffer)
    w1 = _rand((64, Cin), 16)
    outp = torch.zeros(Nimg * (H + 2) * (W + 2), 64, dtype=torch.bfloat16, device="cuda")
    gemm(x.view(-1, Cin), w1, outp, Nimg * H * W, 64, Cin, remap=REMAP_C2P, img_hw=(H, W))
    ref = _pad_rows((x.view(-1, Cin).float() @ w1.float().t()).view(Nimg, H, W, 64))
    _close(outp, ref)


@pytest.mark.parametrize("mode", [0, 2, 3, 4, 5])
@pytest.mark.parametrize("M,N,K", [(1500, 256, 320), (12100, 1024, 256), (130, 128, 64)])
def test_gemm_epilogue_variants(mode, M, N, K):
    """every epilogue implementation (direct, register-prefetch, pipelined, TMA-fed residual) gives the same result"""
    from tubedetr_b200.gemm import gemm
    A, B, R, Mk = _rand((M, K), 21), _rand((N, K), 22), _rand((M, N), 23), _rand((M, N), 24)
    scale = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    ref = torch.relu((A.float() @ B.float().t()) * scale + bias + R.float()) * (Mk.float() > 0)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(A, B, out, M, N, K, scale=scale, bias=bias, residual=R, relu=True, mask=Mk, debug_flags=(mode + 1) << 1)
    torch.cuda.synchronize()
    _close(out, ref)


TWO_CTA = 64  # debug flag: force the cta_group::2 kernel


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1000, 256, 512), (4096, 512, 1024), (300, 768, 256)])
def test_gemm_2cta_plain_and_epilogue(M, N, K):
    from tubedetr_b200.gemm import gemm
    A, B, R, Mk = _rand((M, K), 31), _rand((N, K), 32), _rand((M, N), 33), _rand((M, N), 34)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(A, B, out, M, N, K, debug_flags=TWO_CTA)
    torch.cuda.synchronize()
    _close(out, A.float() @ B.float().t())
    scale = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    ref = torch.relu((A.float() @ B.float().t()) * scale + bias + R.float()) * (Mk.float() > 0)
    gemm(A, B, out, M, N, K, scale=scale, bias=bias, residual=R, relu=True, mask=Mk, debug_flags=TWO_CTA)
    torch.cuda.synchronize()
    _close(out, ref)


def test_gemm_2cta_dgrad_and_implicit_conv():
    from tubedetr_b200.gemm import REMAP_P2C, gemm
    M, Cout, Cin = 700, 512, 256
    dY, W = _rand((M, Cout), 35), _rand((Cout, Cin), 36)
    out = torch.empty(M, Cin, dtype=torch.bfloat16, device="cuda")
    gemm(dY, W, out, M, Cin, Cout, b_major=1, debug_flags=TWO_CTA)
    _close(out, dY.float() @ W.float())
    Nimg, H, Wd, C = 4, 11, 13, 256
    x = _rand((Nimg, H, Wd, C), 37)
    w = _rand((C, C, 3, 3), 38) * 0.05
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1)
    xp = _pad_rows(x)
    wk = w.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
    taps = [(kh - 1) * (Wd + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    o = torch.full((Nimg * H * Wd, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(xp, wk, o, xp.shape[0], C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], remap=REMAP_P2C, img_hw=(H, Wd),
         debug_flags=TWO_CTA)
    _close(o.view(Nimg, H, Wd, C), ref)


@pytest.mark.parametrize("splits", [1, 4])
def test_gemm_2cta_wgrad_forms(splits):
    from tubedetr_b200.gemm import effective_splits, gemm, splitk_reduce
    R, Cout, Cin = 3000, 512, 256
    dY, X = _rand((R, Cout), 41), _rand((R, Cin), 42)
    s = effective_splits(R, splits)
    part = torch.empty(s, Cout, Cin, dtype=torch.float32, device="cuda")
    gemm(dY, X, part, Cout, Cin, R, a_major=1, b_major=1, splits=splits, debug_flags=TWO_CTA)
    out = torch.empty(Cout, Cin, dtype=torch.float32, device="cuda")
    splitk_reduce(part, s, Cout, Cin, out)
    _close(out, dY.float().t() @ X.float(), tol=2e-3)
    # z-batched 3x3 wgrad over the haloed grid
    Nimg, H, W, C = 3, 9, 10, 256
    x, g = _rand((Nimg, H, W, C), 43), _rand((Nimg, H, W, C), 44)
    w = (_rand((C, C, 3, 3), 45) * 0.05).float().requires_grad_(True)
    y = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1)
    y.backward(g.float().permute(0, 3, 1, 2))
    taps = [(kh - 1) * (W + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    gp, xp = _pad_rows(g), _pad_rows(x)
    Rr = gp.shape[0]
    s = effective_splits(Rr, splits)
    part = torch.empty(s, C, 9 * C, dtype=torch.float32, device="cuda")
    gemm(gp, xp, part, C, C, Rr, a_major=1, b_major=1, nz=9, z_b_off1=taps, z_out_col=[t * C for t in range(9)], splits=splits,
         debug_flags=TWO_CTA)
    dw = torch.empty(C, C, 3, 3, dtype=torch.float32, device="cuda")
    splitk_reduce(part, s, C, 9 * C, dw, taps=9)
    _close(dw, w.grad, tol=5e-3)


@pytest.mark.parametrize("M,N,K,relu,f32", [(3525, 512, 256, False, False), (3525, 256, 256, False, True), (3525, 2048, 256, True, False),
                                             (3525, 256, 2048, False, True), (100, 256, 256, False, False), (100, 2048, 256, True, False),
                                             (15125, 256, 2048, False, True), (12100, 256, 256, False, False), (20, 256, 768, False, True)])
def test_gemm_split_precision_weights_two_taps(M, N, K, relu, f32):
    """the transformer's forward form (ops.gemm_fwd_w): fp32 weights as bf16 [hi | lo] read through two reduction taps over the same
    activation tile.  Against fp32 matmul with the UNROUNDED weights the error must be far below what bf16 weights give."""
    from tubedetr_b200 import kernels as Kn
    from tubedetr_b200.gemm import gemm
    g = torch.Generator().manual_seed(31)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda() * 0.1
    Ws = Kn.split_bf16(W, torch.empty(N, 2 * K, dtype=torch.bfloat16, device="cuda"))
    assert torch.equal(Ws[:, :K], W.to(torch.bfloat16))
    assert (Ws[:, :K].float() + Ws[:, K:].float() - W).abs().max().item() <= 2 ** -16 * W.abs().max().item()
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
    gemm(A, Ws, out, M, N, K, ntaps=2, a_off0=(0, 0), b_off0=(0, K), bias=bias, relu=relu)
    one = torch.empty_like(out)
    gemm(A, Ws[:, :K], one, M, N, K, bias=bias, relu=relu)        # plain bf16 weights
    torch.cuda.synchronize()
    ref = A.float() @ W.t() + bias
    if relu:
        ref = ref.relu()
    if f32:
        e2, e1 = (out - ref).abs().max().item(), (one - ref).abs().max().item()
        assert e2 <= 2e-5 * ref.abs().max().item() + 1e-5, (e2, e1)       # fp32-grade: only accumulation-order noise is left
        assert e2 < 0.1 * e1
    else:
        _close(out, ref, 6e-3)                                              # bf16 output rounding only
