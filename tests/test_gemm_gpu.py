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


NO_TMA_EPI = 512  # debug flag: row-per-thread global epilogue instead of the TMA boxes (tdb_gemm2_kernel<0>)


@pytest.mark.parametrize("flags", [TWO_CTA, TWO_CTA | NO_TMA_EPI])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1000, 256, 512), (4096, 512, 1024), (300, 768, 256)])
def test_gemm_2cta_plain_and_epilogue(M, N, K, flags):
    from tubedetr_b200.gemm import gemm
    A, B, R, Mk = _rand((M, K), 31), _rand((N, K), 32), _rand((M, N), 33), _rand((M, N), 34)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(A, B, out, M, N, K, debug_flags=flags)
    torch.cuda.synchronize()
    _close(out, A.float() @ B.float().t())
    scale = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    ref = torch.relu((A.float() @ B.float().t()) * scale + bias + R.float()) * (Mk.float() > 0)
    gemm(A, B, out, M, N, K, scale=scale, bias=bias, residual=R, relu=True, mask=Mk, debug_flags=flags)
    torch.cuda.synchronize()
    _close(out, ref)


@pytest.mark.parametrize("M,N,K,res,relu,mask", [(60500, 1024, 256, True, True, False), (30001, 512, 128, True, True, False),
                                                  (40000, 256, 64, True, True, False), (14100, 1536, 256, False, False, False),
                                                  (12100, 1024, 256, True, False, True), (9999, 256, 1024, False, True, False),
                                                  (2500, 2048, 256, False, True, False)])
def test_gemm_2cta_tma_epilogue_is_bit_identical_to_the_row_epilogue(M, N, K, res, relu, mask):
    """the TMA-box epilogue of the pair kernel (the backbone's 1x1 conv + FrozenBN + residual + ReLU shapes, the decoder's K/V projection,
    the dgrad with residual + ReLU mask): many tiles per CTA pair (box recycling), ragged last tile, output written into a column
    slice of a wider matrix, residual read from one; compared with fp32 torch and, bit for bit, with the row-per-thread epilogue"""
    from tubedetr_b200.gemm import gemm
    A, B = _rand((M, K), 51), _rand((N, K), 52)
    Rw = _rand((M, N + 64), 53)
    R = Rw[:, 64:] if res else None
    Mk = _rand((M, N), 54) if mask else None
    scale = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    wide = torch.full((M + 3, 2 * N), 7.0, dtype=torch.bfloat16, device="cuda")
    out = wide[:M, N:]
    out2 = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    gemm(A, B, out, M, N, K, scale=scale, bias=bias, residual=R, relu=relu, mask=Mk, debug_flags=TWO_CTA)
    gemm(A, B, out2, M, N, K, scale=scale, bias=bias, residual=R, relu=relu, mask=Mk, debug_flags=TWO_CTA | NO_TMA_EPI)
    torch.cuda.synchronize()
    ref = (A.float() @ B.float().t()) * scale + bias
    if res:
        ref = ref + R.float()
    if relu:
        ref = torch.relu(ref)
    if mask:
        ref = ref * (Mk.float() > 0)
    _close(out, ref)
    assert torch.equal(out, out2)
    assert bool((wide[:M, :N] == 7.0).all()) and bool((wide[M:] == 7.0).all()), "the TMA store wrote outside its column slice / row range"
    # the default routing (no debug flag) must agree as well, whichever kernel it picks
    out3 = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    gemm(A, B, out3, M, N, K, scale=scale, bias=bias, residual=R, relu=relu, mask=Mk)
    torch.cuda.synchronize()
    _close(out3, ref)


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


@pytest.mark.parametrize("frames,H,W,C,N,bmaj,mask", [(20, 22, 22, 1024, 256, 0, False), (9, 11, 11, 2048, 512, 0, False),
                                                      (7, 22, 22, 1024, 256, 1, True)])
def test_gemm_2cta_copy_out_compact_to_padded(frames, H, W, C, N, bmaj, mask):
    """1x1 conv (or its dgrad with ReLU mask) whose rows go into the zero-haloed pixel grid of the following 3x3 conv: the pair
    kernel's shared-memory-box epilogue with the coalesced copy-out == the row-per-thread epilogue, halo rows untouched"""
    from tubedetr_b200.gemm import REMAP_C2P, gemm
    M = frames * H * W
    A = _rand((M, C), 61)
    B = (_rand((C, N), 62) if bmaj else _rand((N, C), 62)) * 0.05
    Mk = _rand((M, N), 63) if mask else None
    scale = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    Rp = frames * (H + 2) * (W + 2)
    outs = []
    for fl in (TWO_CTA, TWO_CTA | NO_TMA_EPI, 0):
        o = torch.zeros(Rp, N, dtype=torch.bfloat16, device="cuda")
        gemm(A, B, o, M, N, C, b_major=bmaj, scale=scale, bias=bias, relu=not mask, mask=Mk, remap=REMAP_C2P, img_hw=(H, W), debug_flags=fl)
        outs.append(o)
    torch.cuda.synchronize()
    ref = (A.float() @ (B.float() if bmaj else B.float().t())) * scale + bias
    ref = ref * (Mk.float() > 0) if mask else torch.relu(ref)
    got = outs[0].view(frames, H + 2, W + 2, N)
    _close(got[:, 1:-1, 1:-1].reshape(M, N), ref)
    halo = got.clone()
    halo[:, 1:-1, 1:-1] = 0
    assert not bool(halo.any()), "halo rows of the padded grid were written"
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("frames,H,W,C,bmaj,mask", [(12, 22, 22, 256, 0, False), (10, 11, 11, 512, 0, False), (5, 22, 22, 256, 1, True)])
def test_gemm_2cta_copy_out_implicit_conv_padded_to_compact(frames, H, W, C, bmaj, mask):
    """3x3 conv over the zero-haloed grid (halo A tile, 9 taps) -> compact rows, fprop and dgrad form, box epilogue == row epilogue"""
    from tubedetr_b200.gemm import REMAP_P2C, gemm
    x = _rand((frames, H, W, C), 71)
    xp = _pad_rows(x)
    Rp = xp.shape[0]
    Bm = _rand((C, 9 * C), 72) * 0.05
    Mk = _rand((Rp, C), 73) if mask else None
    taps = [(kh - 1) * (W + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    outs = []
    for fl in (TWO_CTA, TWO_CTA | NO_TMA_EPI, 0):
        o = torch.full((frames * H * W, C), float("nan"), dtype=torch.bfloat16, device="cuda")
        gemm(xp, Bm, o, Rp, C, C, b_major=bmaj, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], relu=not mask, mask=Mk,
             remap=REMAP_P2C, img_hw=(H, W), debug_flags=fl)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    _close(outs[2], outs[0])          # the default route may be the 1-CTA kernel (other tap order in the fp32 sum): close, not equal
    assert bool(torch.isfinite(outs[0].float()).all())
    if not bmaj and not mask:
        w = Bm.view(C, 3, 3, C).permute(0, 3, 1, 2).float()
        ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1)).permute(0, 2, 3, 1).reshape(-1, C)
        _close(outs[0], ref)


DYNAMIC_TILES = 2048  # TDB_GEMM_FLAG_DYNAMIC_TILES: cluster launch control (work stealing) instead of the static schedule (tile += grid)


@pytest.mark.parametrize("M,N,K,res,f32,flags", [(20000, 64, 64, False, False, 0), (60500, 1024, 256, True, False, 0), (60500, 256, 1024, False, False, 0),
                                                   (3525, 256, 2048, False, True, 0), (9000, 512, 512, True, False, TWO_CTA), (300, 128, 192, False, False, 0),
                                                   (129, 64, 64, False, False, 0)])
def test_gemm_work_stealing_schedule_equals_static_schedule(M, N, K, res, f32, flags):
    """cluster launch control (one CTA / cluster per tile, resident ones steal the rest) must produce exactly what the static
    persistent schedule produces: every tile computed once, by whichever CTA"""
    from tubedetr_b200.gemm import gemm
    A, B = _rand((M, K), 81), _rand((N, K), 82)
    R = _rand((M, N), 83) if res else None
    scale = torch.rand(N, device="cuda") + 0.5
    bias = torch.randn(N, device="cuda")
    dt = torch.float32 if f32 else torch.bfloat16
    o1 = torch.full((M, N), float("nan"), dtype=dt, device="cuda")
    o2 = torch.full((M, N), float("nan"), dtype=dt, device="cuda")
    for _ in range(3):      # several launches back to back: the response ring / barriers of one launch must not leak into the next
        gemm(A, B, o1, M, N, K, scale=scale, bias=bias, residual=R, relu=True, debug_flags=flags | DYNAMIC_TILES)
    gemm(A, B, o2, M, N, K, scale=scale, bias=bias, residual=R, relu=True, debug_flags=flags)
    torch.cuda.synchronize()
    ref = (A.float() @ B.float().t()) * scale + bias
    if res:
        ref = ref + R.float()
    _close(o1, torch.relu(ref), tol=2e-2 if not f32 else 2e-3)
    assert torch.equal(o1, o2)


def test_gemm_work_stealing_with_a_foreign_kernel_holding_sms():
    """a spin kernel occupies 24 SMs on another stream while the GEMM runs: results unchanged (CTAs that start late take what is left)"""
    import ctypes as C
    from tubedetr_b200 import _lib
    from tubedetr_b200.gemm import gemm
    M, N, K = 60500, 256, 1024
    A, B = _rand((M, K), 91), _rand((N, K), 92)
    o1 = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    o2 = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    gemm(A, B, o2, M, N, K)
    torch.cuda.synchronize()
    lib = _lib.lib()
    lib.tdb_debug_spin.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_void_p]
    side = torch.cuda.Stream()
    _lib.check(lib.tdb_debug_spin(24, 200 * 1024, 1_000_000, C.c_void_p(side.cuda_stream)), "spin")
    for _ in range(4):
        gemm(A, B, o1, M, N, K, debug_flags=DYNAMIC_TILES)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("frames,H,W,Cin,C", [(6, 44, 44, 512, 256), (5, 22, 22, 256, 128), (3, 11, 9, 128, 128), (40, 22, 22, 1024, 512),
                                              (2, 7, 12, 64, 64)])
def test_stride2_conv_as_implicit_gemm_over_space_to_depth(frames, H, W, Cin, C):
    """1x1 conv + ReLU written space-to-depth (TDB_REMAP_COMPACT_TO_S2D), then the 3x3 / stride 2 / pad 1 conv as ONE 9-tap implicit
    GEMM over that matrix (TDB_REMAP_S2D_TO_COMPACT) == torch conv2d of the bf16-rounded intermediate; even and odd image sizes,
    1-CTA and pair kernels; equal to the im2col route bit for bit (same tap-major reduction order)"""
    from tubedetr_b200 import kernels as K
    from tubedetr_b200.gemm import REMAP_C2S, REMAP_S2C, gemm
    x = _rand((frames * H * W, Cin), 101)
    w1 = _rand((C, Cin), 102) * 0.05
    w2 = _rand((C, C, 3, 3), 103) * 0.05
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    ohp, owp = Ho + 1, Wo + 1
    y1s = torch.zeros(frames * ohp * owp, 4 * C, dtype=torch.bfloat16, device="cuda")
    gemm(x, w1, y1s, frames * H * W, C, Cin, relu=True, remap=REMAP_C2S, img_hw=(H, W))
    wk = w2.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
    a0 = [(((kh - 1) & 1) * 2 + ((kw - 1) & 1)) * C for kh in range(3) for kw in range(3)]
    a1 = [(-1 if kh == 0 else 0) * owp + (-1 if kw == 0 else 0) for kh in range(3) for kw in range(3)]
    y2 = torch.full((frames * Ho * Wo, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(y1s, wk, y2, frames * ohp * owp, C, C, ntaps=9, a_off0=a0, a_off1=a1, b_off0=[t * C for t in range(9)], relu=True,
         remap=REMAP_S2C, img_hw=(Ho, Wo))
    # the im2col route on the same operands
    y1 = torch.empty(frames * H * W, C, dtype=torch.bfloat16, device="cuda")
    gemm(x, w1, y1, frames * H * W, C, Cin, relu=True)
    col = torch.empty(frames * Ho * Wo, 9 * C, dtype=torch.bfloat16, device="cuda")
    K.im2col3x3s2(y1, col, frames, H, W, C)
    y2b = torch.empty(frames * Ho * Wo, C, dtype=torch.bfloat16, device="cuda")
    gemm(col, wk, y2b, frames * Ho * Wo, C, 9 * C, relu=True)
    torch.cuda.synchronize()
    ref = torch.relu(torch.nn.functional.conv2d(y1.float().view(frames, H, W, C).permute(0, 3, 1, 2), w2.float(), stride=2, padding=1))
    _close(y2, ref.permute(0, 2, 3, 1).reshape(-1, C))
    _close(y2, y2b, tol=1e-2)
    # the space-to-depth matrix holds exactly the compact conv1 output, halo positions zero
    s = y1s.view(frames, ohp, owp, 2, 2, C)
    assert not bool(s[:, 0].any()) and not bool(s[:, :, 0].any())
    full = torch.zeros(frames, 2 * Ho, 2 * Wo, C, dtype=torch.bfloat16, device="cuda")
    full[:, :H, :W] = y1.view(frames, H, W, C)
    assert torch.equal(s[:, 1:, 1:].permute(0, 1, 3, 2, 4, 5).reshape(frames, 2 * Ho, 2 * Wo, C), full)


@pytest.mark.parametrize("frames,H,W,Cin,C", [(12, 22, 22, 1024, 256), (9, 11, 11, 2048, 512), (3, 44, 44, 512, 128), (2, 9, 13, 256, 64)])
def test_conv3x3_over_the_shared_halo_grid(frames, H, W, Cin, C):
    """1x1 conv + ReLU written into the (H + 1) x (W + 1) grid with ONE shared zero halo row / column per image
    (TDB_REMAP_COMPACT_TO_PADDED1), then the 3x3 / pad 1 conv as a 9-tap implicit GEMM over it (way back = TDB_REMAP_S2D_TO_COMPACT)
    == torch conv2d; also the last image's bottom halo, which lies beyond the matrix (TMA zero fill)"""
    from tubedetr_b200.gemm import REMAP_C2P1, REMAP_S2C, gemm
    x = _rand((frames * H * W, Cin), 111)
    w1 = _rand((C, Cin), 112) * 0.05
    w2 = _rand((C, C, 3, 3), 113) * 0.05
    Rp = frames * (H + 1) * (W + 1)
    y1p = torch.zeros(Rp, C, dtype=torch.bfloat16, device="cuda")
    gemm(x, w1, y1p, frames * H * W, C, Cin, relu=True, remap=REMAP_C2P1, img_hw=(H, W))
    wk = w2.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
    taps = [(kh - 1) * (W + 1) + (kw - 1) for kh in range(3) for kw in range(3)]
    y2 = torch.full((frames * H * W, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    gemm(y1p, wk, y2, Rp, C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], relu=True, remap=REMAP_S2C, img_hw=(H, W))
    torch.cuda.synchronize()
    g = y1p.view(frames, H + 1, W + 1, C)
    assert not bool(g[:, 0].any()) and not bool(g[:, :, 0].any())
    y1 = g[:, 1:, 1:].float()
    ref = torch.relu(torch.nn.functional.conv2d(y1.permute(0, 3, 1, 2), w2.float(), padding=1)).permute(0, 2, 3, 1).reshape(-1, C)
    _close(y2, ref)
    # dgrad form over the same grid: dX = sum_t dY[r - shift_t] W_t, masked by the ReLU of the grid's own values
    dy = torch.zeros(Rp, C, dtype=torch.bfloat16, device="cuda")
    dy.view(frames, H + 1, W + 1, C)[:, 1:, 1:] = _rand((frames, H, W, C), 114)
    dx = torch.full((frames * H * W, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    wt = w2.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()          # [Cout(K), tap * Cin(N)] read MN-major
    gemm(dy, wt, dx, Rp, C, C, b_major=1, ntaps=9, a_off1=[-t for t in taps], b_off0=[t * C for t in range(9)], mask=y1p,
         remap=REMAP_S2C, img_hw=(H, W))
    torch.cuda.synchronize()
    dyc = dy.view(frames, H + 1, W + 1, C)[:, 1:, 1:].float().permute(0, 3, 1, 2)
    refdx = torch.nn.functional.conv_transpose2d(dyc, w2.float(), padding=1).permute(0, 2, 3, 1) * (y1 > 0)
    _close(dx, refdx.reshape(-1, C))
