"""Python face of tdb_gemm: thin argument marshalling, no math."""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmDesc, OUT_BF16, OUT_F32, REMAP_C2P, REMAP_C2P1, REMAP_C2S, REMAP_NONE, REMAP_P2C, REMAP_S2C  # noqa: F401


DYNAMIC_TILES = 1 << 11     # TDB_GEMM_FLAG_DYNAMIC_TILES (include/tubedetr_b200.h)


def _mat(t):
    assert t.dim() == 2 and t.dtype == torch.bfloat16 and t.is_cuda and t.stride(1) == 1, (t.shape, t.dtype, t.stride())
    return t.data_ptr(), t.shape[0], t.shape[1], t.stride(0)


def gemm(A, B, out, M, N, K, a_major=0, b_major=0, ntaps=1, a_off0=None, a_off1=None, b_off0=None, b_off1=None,
         nz=0, z_b_off1=None, z_out_col=None, splits=1, scale=None, bias=None, residual=None, mask=None, relu=False,
         remap=REMAP_NONE, img_hw=(0, 0), block_n=0, max_ctas=0, debug_flags=0):
    d = GemmDesc()
    d.A, d.a_rows, d.a_cols, d.lda = _mat(A)
    d.B, d.b_rows, d.b_cols, d.ldb = _mat(B)
    d.a_major, d.b_major = a_major, b_major
    d.M, d.N, d.K, d.ntaps = M, N, K, ntaps
    for name, val in (("a_off0", a_off0), ("a_off1", a_off1), ("b_off0", b_off0), ("b_off1", b_off1),
                      ("z_b_off1", z_b_off1), ("z_out_col", z_out_col)):
        if val is not None:
            arr = getattr(d, name)
            for i, v in enumerate(val):
                arr[i] = int(v)
    d.nz, d.splits = nz, splits
    for name, t in (("scale", scale), ("bias", bias)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() >= N
            setattr(d, name, t.data_ptr())
    if residual is not None:
        d.residual, _, _, d.ldr = _mat(residual)
    if mask is not None:
        d.mask, _, _, d.ldmask = _mat(mask)
    d.relu = int(relu)
    assert out.is_cuda and out.stride(-1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    d.out, d.out_dtype, d.ldo = out.data_ptr(), (OUT_F32 if out.dtype == torch.float32 else OUT_BF16), out.stride(-2)
    d.remap, d.img_h, d.img_w = remap, img_hw[0], img_hw[1]
    d.block_n, d.max_ctas, d.debug_flags = block_n, max_ctas, debug_flags
    _lib.check(_lib.lib().tdb_gemm(C.byref(d), _lib.stream_ptr()), "tdb_gemm")
    return out


def effective_splits(K, splits):
    return int(_lib.lib().tdb_gemm_effective_splits(int(K), int(splits)))


def splitk_reduce(part, splits, M, N, out, rowscale=None, taps=1, accumulate=False):
    _lib.check(_lib.lib().tdb_splitk_reduce(_lib.ptr(part), int(splits), int(M), int(N), _lib.ptr(rowscale), _lib.ptr(out),
                                           int(taps), int(accumulate), _lib.stream_ptr()), "tdb_splitk_reduce")
    return out
