"""Host-side contract of the pixel-row layouts the backbone engine feeds to tdb_gemm (include/tubedetr_b200.h, TDB_REMAP_*), checked on the
CPU against torch convolutions: a tiny pure-torch emulator of the multi-tap implicit GEMM (row shifts a_off1, column offsets a_off0,
out-of-range rows read as zero, row remaps of the epilogue) runs the exact offset lists `tubedetr_b200/resnet.py` builds.
  * shared-halo grid ((H + 1) x (W + 1), TDB_REMAP_COMPACT_TO_PADDED1 / S2D_TO_COMPACT) for the stride-1 3x3 convs,
  * space-to-depth matrix (TDB_REMAP_COMPACT_TO_S2D) for the stride-2 3x3 convs without an im2col matrix,
  * conv3 + downsample of a stage's first block as one reduction over [y2 | xs] with folded FrozenBN scales.
Reference semantics: torchvision Bottleneck as used by models/backbone.py:93-122."""
import pytest
import torch
import torch.nn.functional as F


def _rows_shifted(A, shift):
    """A[r + shift] for every row r, zero outside the matrix (what the TMA zero fill provides)"""
    R = A.shape[0]
    out = torch.zeros_like(A)
    lo, hi = max(0, -shift), min(R, R - shift)
    if hi > lo:
        out[lo:hi] = A[lo + shift:hi + shift]
    return out


def implicit_gemm(A, Wk, a_off0, a_off1, C):
    """sum_t A[r + a_off1[t], a_off0[t] : a_off0[t] + C] @ Wk[:, t*C:(t+1)*C]^T  (tap-major K, as the kernel reduces)"""
    acc = torch.zeros(A.shape[0], Wk.shape[0], dtype=torch.float64)
    for t, (c0, r0) in enumerate(zip(a_off0, a_off1)):
        acc += _rows_shifted(A, r0)[:, c0:c0 + C].double() @ Wk[:, t * C:(t + 1) * C].double().t()
    return acc


def compact_to_padded1(x, N, H, W):
    """TDB_REMAP_COMPACT_TO_PADDED1: row (n, h, w) -> ((n (H + 1) + h + 1) (W + 1) + w + 1); everything else stays zero"""
    C = x.shape[1]
    g = torch.zeros(N, H + 1, W + 1, C, dtype=x.dtype)
    g[:, 1:, 1:] = x.view(N, H, W, C)
    return g.view(-1, C)


def padded1_to_compact(y, N, H, W):
    """TDB_REMAP_S2D_TO_COMPACT with img = (H, W): keep positions i >= 1 and j >= 1 of the (H + 1) x (W + 1) grid"""
    return y.view(N, H + 1, W + 1, -1)[:, 1:, 1:].reshape(N * H * W, -1)


def compact_to_s2d(x, N, H, W):
    """TDB_REMAP_COMPACT_TO_S2D: row (n, h, w) -> row (n, h/2 + 1, w/2 + 1), columns [plane(h&1, w&1) * C, +C)"""
    C = x.shape[1]
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    g = torch.zeros(N, Ho + 1, Wo + 1, 4, C, dtype=x.dtype)
    xv = x.view(N, H, W, C)
    for a in range(2):
        for b in range(2):
            sub = xv[:, a::2, b::2]
            g[:, 1:1 + sub.shape[1], 1:1 + sub.shape[2], a * 2 + b] = sub
    return g.view(-1, 4 * C)


@pytest.mark.parametrize("N,H,W,C", [(2, 6, 5, 8), (1, 3, 3, 4), (3, 11, 11, 8), (2, 1, 4, 4)])
def test_shared_halo_grid_conv3x3(N, H, W, C):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N * H * W, C, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g)
    wk = w.permute(0, 2, 3, 1).reshape(C, 9 * C)
    wp = W + 1                                                   # resnet.py: wp = w + PADH, taps = (kh - 1) * wp + (kw - 1)
    taps = [(kh - 1) * wp + (kw - 1) for kh in range(3) for kw in range(3)]
    y = implicit_gemm(compact_to_padded1(x, N, H, W), wk, [0] * 9, taps, C)
    got = padded1_to_compact(y, N, H, W)
    ref = F.conv2d(x.view(N, H, W, C).permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    torch.testing.assert_close(got, ref, rtol=1e-9, atol=1e-9)
    # data-gradient form (resnet.py backward: a_off1 = [-t for t in taps], weights read as [Cout, tap * Cin])
    dy = torch.randn(N * H * W, C, generator=g)
    wt = w.permute(0, 2, 3, 1).reshape(C, 9 * C)                # [Cout (K), tap * Cin (N)], MN-major B operand
    dyp = compact_to_padded1(dy, N, H, W)
    acc = torch.zeros(dyp.shape[0], C, dtype=torch.float64)
    for t, s in enumerate(taps):
        acc += _rows_shifted(dyp, -s).double() @ wt[:, t * C:(t + 1) * C].double()
    refdx = F.conv_transpose2d(dy.view(N, H, W, C).permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    torch.testing.assert_close(padded1_to_compact(acc, N, H, W), refdx, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("N,H,W,C", [(2, 6, 4, 8), (1, 7, 5, 4), (3, 11, 9, 4), (1, 2, 2, 4), (2, 1, 3, 4)])
def test_space_to_depth_stride2_conv3x3(N, H, W, C):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(N * H * W, C, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g)
    wk = w.permute(0, 2, 3, 1).reshape(C, 9 * C)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1                  # conv_out(h, 3, 2, 1)
    owp = Wo + 1
    a0, a1 = [], []                                              # exactly the lists of resnet.py's S2_IMPLICIT branch
    for kh in range(3):
        for kw in range(3):
            a0.append((((kh - 1) & 1) * 2 + ((kw - 1) & 1)) * C)
            a1.append((-1 if kh == 0 else 0) * owp + (-1 if kw == 0 else 0))
    y = implicit_gemm(compact_to_s2d(x, N, H, W), wk, a0, a1, C)
    got = padded1_to_compact(y, N, Ho, Wo)                       # TDB_REMAP_S2D_TO_COMPACT with img = (Ho, Wo)
    ref = F.conv2d(x.view(N, H, W, C).permute(0, 3, 1, 2).double(), w.double(), stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    torch.testing.assert_close(got, ref, rtol=1e-9, atol=1e-9)


def test_first_block_conv3_plus_downsample_as_one_reduction():
    """out = relu(bn3(conv3(y2)) + bnd(convd(xs))) == relu([y2 | xs] @ [w3 * s3 | wd * sd]^T + (b3 + bd)) (resnet.py DS_JOINT)"""
    g = torch.Generator().manual_seed(3)
    R, width, cin, cout = 50, 8, 16, 32
    y2, xs = torch.randn(R, width, generator=g).double(), torch.randn(R, cin, generator=g).double()
    w3, wd = torch.randn(cout, width, generator=g).double(), torch.randn(cout, cin, generator=g).double()
    s3, b3, sd_, bd = (torch.randn(cout, generator=g).double() for _ in range(4))
    ref = torch.relu((y2 @ w3.t()) * s3 + b3 + (xs @ wd.t()) * sd_ + bd)
    J = torch.cat([y2, xs], 1)
    Wj = torch.cat([w3 * s3[:, None], wd * sd_[:, None]], 1)
    torch.testing.assert_close(torch.relu(J @ Wj.t() + (b3 + bd)), ref, rtol=1e-12, atol=1e-12)
