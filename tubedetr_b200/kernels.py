"""Python marshalling for the non-GEMM C-ABI entry points (no math here)."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

_i64 = C.c_int64
_f = C.c_float


def stem_im2col(x, col, N, H, W):
    check(lib().tdb_stem_im2col(ptr(x), ptr(col), N, H, W, stream_ptr()), "stem_im2col")


def stem_fused(x, wk, scale, shift, out, N, H, W):
    """out: [N * H2 * W2, 64] bf16 rows, possibly a column slice of a wider matrix (row stride = out.stride(0))"""
    assert out.stride(-1) == 1
    check(lib().tdb_stem_fused_ld(ptr(x), ptr(wk), ptr(scale), ptr(shift), ptr(out), C.c_int64(out.stride(-2)), N, H, W, stream_ptr()), "stem_fused")


def maxpool3x3s2(x, y, N, H, W, Cc):
    check(lib().tdb_maxpool3x3s2(ptr(x), ptr(y), N, H, W, Cc, stream_ptr()), "maxpool")


def im2col3x3s2(x, col, N, H, W, Cc):
    check(lib().tdb_im2col3x3s2(ptr(x), ptr(col), N, H, W, Cc, stream_ptr()), "im2col3x3s2")


def col2im3x3s2_mask(dcol, ymask, dx, N, H, W, Cc):
    check(lib().tdb_col2im3x3s2_mask(ptr(dcol), ptr(ymask), ptr(dx), N, H, W, Cc, stream_ptr()), "col2im3x3s2")


def subsample2(x, y, N, H, W, Cc):
    """y: [N * Ho * Wo, Cc] rows, possibly a column slice of a wider matrix (row stride = y.stride(0))"""
    assert y.stride(-1) == 1 and x.is_contiguous()
    check(lib().tdb_subsample2_ld(ptr(x), ptr(y), C.c_int64(y.stride(-2)), N, H, W, Cc, stream_ptr()), "subsample2")


def upsample2_zero(y, x, N, H, W, Cc):
    check(lib().tdb_upsample2_zero(ptr(y), ptr(x), N, H, W, Cc, stream_ptr()), "upsample2_zero")


def prep_weight(w, out, out_scaled, rowscale, Cout, Cin, taps, Kpad):
    check(lib().tdb_prep_weight(ptr(w), ptr(out), ptr(out_scaled), ptr(rowscale), Cout, Cin, taps, Kpad, stream_ptr()),
          "prep_weight")


def cast_add_bf16(x, add, y):
    check(lib().tdb_cast_add_bf16(ptr(x), ptr(add), ptr(y), _i64(x.numel()), stream_ptr()), "cast_add_bf16")


def split_bf16(x, y):
    """x fp32 [rows, K] -> y bf16 [rows, 2K] = [hi | lo]"""
    rows, Kd = x.shape
    check(lib().tdb_split_bf16(ptr(x), ptr(y), _i64(rows), Kd, stream_ptr()), "split_bf16")
    return y


def layernorm_fwd(x, r, gamma, beta, pos, y, y_bf, ypos_bf, mean, rstd, rows, D, eps, drop=None):
    """drop = (seed tensor, site, p): residual dropout on r inside the kernel (no mask tensor)"""
    seed, site, p = drop if drop is not None else (None, 0, 0.0)
    check(lib().tdb_layernorm_fwd(ptr(x), ptr(r), ptr(gamma), ptr(beta), ptr(pos), ptr(y), ptr(y_bf), ptr(ypos_bf),
                                  ptr(mean), ptr(rstd), rows, D, _f(eps), ptr(seed), _i64(site), _f(p), stream_ptr()), "layernorm_fwd")


def layernorm_bwd(dy, x, r, gamma, mean, rstd, dz, dgamma, dbeta, rows, D, accumulate=False, dy2=None, dy3=None, dz_bf=None,
                  drop=None, dr=None, dr_bf=None, dbias=None):
    nb = int(lib().tdb_layernorm_bwd_blocks(rows))
    partial = torch.empty(nb * 3 * D, dtype=torch.float32, device=x.device)
    seed, site, p = drop if drop is not None else (None, 0, 0.0)
    check(lib().tdb_layernorm_bwd(ptr(dy), ptr(dy2), ptr(dy3), ptr(x), ptr(r), ptr(gamma), ptr(mean), ptr(rstd), ptr(dz), ptr(dz_bf), ptr(dgamma),
                                  ptr(dbeta), ptr(partial), rows, D, int(accumulate), ptr(seed), _i64(site), _f(p), ptr(dr), ptr(dr_bf), ptr(dbias),
                                  stream_ptr()), "layernorm_bwd")


def colsum_bf16(x, out, accumulate=False):
    rows, N = x.shape
    nparts = max(1, min(296, rows // 32))
    partial = torch.empty(nparts * N, dtype=torch.float32, device=x.device)
    check(lib().tdb_colsum_bf16(ptr(x), _i64(x.stride(0)), rows, N, ptr(partial), nparts, ptr(out), int(accumulate),
                                stream_ptr()), "colsum_bf16")
    return out


def mha_uses_tc(H, Lq, Lk, level=1):
    """tcgen05 self-attention (tdb_attn_tc.cu) is the default for every shape it supports (even H, 2 <= Lq <= 256, Lk <= 256);
    tdb_mha_set_tc(0|1|2) / env TDB_MHA_TC select CUDA cores | tcgen05 forward | tcgen05 forward + backward (default 2)"""
    return lib().tdb_mha_tc_enabled() >= level and bool(lib().tdb_mha_tc_supported(H, Lq, Lk))


def mha_fwd_cuda_core(q, k, v, kpm, o, p, pbar, B, H, Lq, Lk, scale, keep=None, pdrop=None, keep_scale=1.0):
    """the CUDA-core row kernel (tdb_attn.cu): one-query sequences and sequences longer than 256.
    q,k,v: 2-D bf16 views [B*L, >=H*32] (row stride = .stride(0)); kpm uint8 [B,Lk] or None.
    keep (uint8 [B,H,Lq,Lk]) + pdrop enable attention-probability dropout."""
    check(lib().tdb_mha_fwd(ptr(q), _i64(q.stride(0)), ptr(k), _i64(k.stride(0)), ptr(v), _i64(v.stride(0)), ptr(kpm),
                            ptr(o), _i64(o.stride(0)), ptr(p), ptr(pbar), ptr(keep), ptr(pdrop), _f(keep_scale),
                            B, H, Lq, Lk, _f(scale), stream_ptr()), "mha_fwd")


def mha_tc_fwd(q, k, v, kpm, o, p, pbar, B, H, Lq, Lk, scale, drop=None, pdrop=None):
    """QK^T and PV on tcgen05 (S / O accumulators in TMEM).  drop = (seed tensor, site, p): attention dropout from the hash
    stream inside the kernel (no mask tensor); pbar = head mean of p (of pdrop under dropout)"""
    seed, site, dp = drop if drop is not None else (None, 0, 0.0)
    check(lib().tdb_mha_tc_fwd(ptr(q), _i64(q.stride(0)), ptr(k), _i64(k.stride(0)), ptr(v), _i64(v.stride(0)), ptr(kpm),
                               ptr(o), _i64(o.stride(0)), ptr(p), ptr(pdrop), ptr(seed), _i64(site), _f(dp),
                               B, H, Lq, Lk, _f(scale), stream_ptr()), "mha_tc_fwd")
    if pbar is not None:
        assert drop is None or pdrop is not None
        check(lib().tdb_head_mean(ptr(pdrop if drop is not None else p), ptr(pbar), B, H, Lq, Lk, stream_ptr()), "head_mean")


def mha_tc_bwd(q, k, v, dout, p, dpbar, dq, dk, dv, B, H, Lq, Lk, scale, drop=None):
    seed, site, dp = drop if drop is not None else (None, 0, 0.0)
    check(lib().tdb_mha_tc_bwd(ptr(q), _i64(q.stride(0)), ptr(k), _i64(k.stride(0)), ptr(v), _i64(v.stride(0)),
                               ptr(dout), _i64(dout.stride(0)), ptr(p), ptr(seed), _i64(site), _f(dp), ptr(dpbar),
                               ptr(dq), _i64(dq.stride(0)), ptr(dk), _i64(dk.stride(0)), ptr(dv), _i64(dv.stride(0)),
                               B, H, Lq, Lk, _f(scale), stream_ptr()), "mha_tc_bwd")


def mha_bwd_cuda_core(q, k, v, dout, p, dpbar, ds, dq, dk, dv, B, H, Lq, Lk, scale, keep=None, keep_scale=1.0, pd_scratch=None):
    """ds / pd_scratch: fp32 [B,H,Lq,Lk] scratch"""
    check(lib().tdb_mha_bwd(ptr(q), _i64(q.stride(0)), ptr(k), _i64(k.stride(0)), ptr(v), _i64(v.stride(0)),
                            ptr(dout), _i64(dout.stride(0)), ptr(p), ptr(keep), _f(keep_scale), ptr(pd_scratch), ptr(dpbar),
                            ptr(ds), ptr(dq), _i64(dq.stride(0)), ptr(dk), _i64(dk.stride(0)), ptr(dv), _i64(dv.stride(0)),
                            B, H, Lq, Lk, _f(scale), stream_ptr()), "mha_bwd")


def xattn_fused_fwd(q, mempb, memb, wkv, bv, kpm, o, p, pbar, F, S, scale, keep=None, keep_scale=1.0):
    lib().tdb_xattn_workspace_bytes.restype = C.c_int64
    nbytes = int(lib().tdb_xattn_workspace_bytes(F, S))
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=q.device)
    check(lib().tdb_xattn_fused_fwd(ptr(q), ptr(mempb), ptr(memb), ptr(wkv), ptr(bv), ptr(kpm), ptr(keep), _f(keep_scale), ptr(o), ptr(p), ptr(pbar),
                                    ptr(ws), _i64(nbytes), F, S, _f(scale), stream_ptr()), "xattn_fused_fwd")


def xattn_bwd(q, kp, vp, dout, p, dpbar, dq, dk, dv, F, S, scale, keep=None, keep_scale=1.0):
    """one-query-per-frame attention backward: q, dout [F,256]; kp, vp [F*S,256] (contiguous bf16); p [F,8,1,S] fp32"""
    check(lib().tdb_xattn_bwd(ptr(q), ptr(kp), ptr(vp), ptr(dout), ptr(p), ptr(keep), _f(keep_scale), ptr(dpbar), ptr(dq), ptr(dk),
                              ptr(dv), F, S, _f(scale), stream_ptr()), "xattn_bwd")


def xattn_core_fwd(q, k, v, kpm, o, p, pbar, F, S, scale, keep=None, keep_scale=1.0):
    """one-query-per-frame attention on projected k, v (2-D bf16 views [F*S, 256] with any row stride); q [F,256]"""
    check(lib().tdb_xattn_core_fwd(ptr(q), ptr(k), _i64(k.stride(0)), ptr(v), _i64(v.stride(0)), ptr(kpm), ptr(keep), _f(keep_scale),
                                   ptr(o), ptr(p), ptr(pbar), F, S, _f(scale), stream_ptr()), "xattn_core_fwd")


def xattn_core_bwd(q, k, v, dout, p, dpbar, dq, dk, dv, F, S, scale, keep=None, keep_scale=1.0):
    check(lib().tdb_xattn_core_bwd(ptr(q), ptr(k), _i64(k.stride(0)), ptr(v), _i64(v.stride(0)), ptr(dout), ptr(p), ptr(keep),
                                   _f(keep_scale), ptr(dpbar), ptr(dq), ptr(dk), _i64(dk.stride(0)), ptr(dv), _i64(dv.stride(0)),
                                   F, S, _f(scale), stream_ptr()), "xattn_core_bwd")


def dropout_mask(keep, seed, site, p):
    check(lib().tdb_dropout_mask(ptr(keep), _i64(keep.numel()), ptr(seed), _i64(site), _f(p), stream_ptr()), "dropout_mask")
    return keep


def dropout_bf16(x, y, seed, site, p):
    check(lib().tdb_dropout_bf16(ptr(x), ptr(y), _i64(x.numel()), ptr(seed), _i64(site), _f(p), stream_ptr()), "dropout_bf16")
    return y


def loss_desc(pbs, sts, ws, tgt_boxes, num_boxes, gauss, time_mask, neg, nneg, K_, B, T):
    """tdb_loss_desc from per-layer tensor lists (None entries / lists = that loss family is off)"""
    d = _lib.LossDesc()
    nl = max(len(x) for x in (pbs or [], sts or [], ws or []))
    d.nlayers, d.K, d.B, d.T = nl, int(K_), int(B), int(T)
    for name, lst in (("pred_boxes", pbs), ("pred_sted", sts), ("weights", ws)):
        arr = getattr(d, name)
        for i in range(nl):
            t = lst[i] if lst else None
            if t is not None:
                assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
            arr[i] = t.data_ptr() if t is not None else None
    for name, t in (("tgt_boxes", tgt_boxes), ("num_boxes", num_boxes), ("gauss", gauss), ("time_mask", time_mask), ("neg", neg),
                    ("nneg", nneg)):
        if t is not None:
            assert t.is_contiguous() and t.is_cuda
            setattr(d, name, t.data_ptr())
    return d


def criterion_fwd(desc, losses):
    check(lib().tdb_criterion_fwd(C.byref(desc), ptr(losses), stream_ptr()), "criterion_fwd")
    return losses


def criterion_bwd(desc, grad_losses, d_boxes, d_sted, d_weights):
    check(lib().tdb_criterion_bwd(C.byref(desc), ptr(grad_losses), ptr(d_boxes), ptr(d_sted), ptr(d_weights), stream_ptr()), "criterion_bwd")


def pos_sine(mask_u8, out, N, h, w):
    check(lib().tdb_pos_sine(ptr(mask_u8), ptr(out), N, h, w, stream_ptr()), "pos_sine")
    return out


def enc_assemble_fwd(src, txt, pos, x32, xb, xpb, pe, n, HW, L, n_clips):
    check(lib().tdb_enc_assemble_fwd(ptr(src), ptr(txt), ptr(pos), ptr(x32), ptr(xb), ptr(xpb), ptr(pe), n, HW, L, n_clips, stream_ptr()), "enc_assemble_fwd")


def enc_assemble_bwd(g32, gb, gpb, dsrc, dtxt, n, HW, L, n_clips):
    check(lib().tdb_enc_assemble_bwd(ptr(g32), ptr(gb), ptr(gpb), ptr(dsrc), ptr(dtxt), n, HW, L, n_clips, stream_ptr()), "enc_assemble_bwd")


def fast_mix_fwd(enc, fm, z, B, T, k, HW, S):
    check(lib().tdb_fast_mix_fwd(ptr(enc), ptr(fm), ptr(z), B, T, k, HW, S, stream_ptr()), "fast_mix_fwd")


def fast_mix_bwd(dz, denc, B, T, k, HW, S):
    check(lib().tdb_fast_mix_bwd(ptr(dz), ptr(denc), B, T, k, HW, S, stream_ptr()), "fast_mix_bwd")


def aggregate_fwd(enc, pe, upd, mem, mem_pos, memb, mempb, B, T, k, HW, S):
    check(lib().tdb_aggregate_fwd(ptr(enc), ptr(pe), ptr(upd), ptr(mem), ptr(mem_pos), ptr(memb), ptr(mempb), B, T, k, HW, S, stream_ptr()), "aggregate_fwd")


def aggregate_bwd(gmem, gmemb, gmempb, denc, dupd, dupd_b, B, T, k, HW, S):
    check(lib().tdb_aggregate_bwd(ptr(gmem), ptr(gmemb), ptr(gmempb), ptr(denc), ptr(dupd), ptr(dupd_b), B, T, k, HW, S, stream_ptr()), "aggregate_bwd")


def head_out_fwd(x, W, b, y, act, drop=None):
    seed, site, p = drop if drop is not None else (None, 0, 0.0)
    R, J = y.shape
    check(lib().tdb_head_out_fwd(ptr(x), ptr(W), ptr(b), ptr(y), R, J, int(act), ptr(seed), _i64(site), _f(p), stream_ptr()), "head_out_fwd")
    return y


def head_out_bwd(dy, y, x, W, dpre, dx, dW, db, act, mask_dx, dx_scale, drop=None):
    seed, site, p = drop if drop is not None else (None, 0, 0.0)
    R, J = y.shape
    check(lib().tdb_head_out_bwd(ptr(dy), ptr(y), ptr(x), ptr(W), ptr(dpre), ptr(dx), ptr(dW), ptr(db), R, J, int(act), int(mask_dx),
                                 _f(dx_scale), ptr(seed), _i64(site), _f(p), stream_ptr()), "head_out_bwd")


def gelu_fwd(x, y):
    check(lib().tdb_gelu_fwd(ptr(x), ptr(y), _i64(x.numel()), stream_ptr()), "gelu_fwd")
    return y


def gelu_bwd(dy, x, dx):
    check(lib().tdb_gelu_bwd(ptr(dy), ptr(x), ptr(dx), _i64(x.numel()), stream_ptr()), "gelu_bwd")
    return dx


def skinny_wgrad(dy, x, dW, db):
    R, N = dy.shape
    check(lib().tdb_skinny_wgrad(ptr(dy), _i64(dy.stride(0)), ptr(x), _i64(x.stride(0)), ptr(dW), ptr(db), R, N, x.shape[1], stream_ptr()), "skinny_wgrad")


def text_attn_fwd(q, k, v, kpm, o, p, B, H, L, scale, drop=None):
    seed, site, dp = drop if drop is not None else (None, 0, 0.0)
    check(lib().tdb_text_attn_fwd(ptr(q), _i64(q.stride(0)), ptr(k), _i64(k.stride(0)), ptr(v), _i64(v.stride(0)), ptr(kpm), ptr(o),
                                  _i64(o.stride(0)), ptr(p), ptr(seed), _i64(site), _f(dp), B, H, L, _f(scale), stream_ptr()), "text_attn_fwd")


def text_attn_bwd(q, k, v, dout, p, dq, dk, dv, B, H, L, scale, drop=None):
    seed, site, dp = drop if drop is not None else (None, 0, 0.0)
    check(lib().tdb_text_attn_bwd(ptr(q), _i64(q.stride(0)), ptr(k), _i64(k.stride(0)), ptr(v), _i64(v.stride(0)), ptr(dout),
                                  _i64(dout.stride(0)), ptr(p), ptr(seed), _i64(site), _f(dp), ptr(dq), _i64(dq.stride(0)), ptr(dk),
                                  _i64(dk.stride(0)), ptr(dv), _i64(dv.stride(0)), B, H, L, _f(scale), stream_ptr()), "text_attn_bwd")


def frames_preprocess(src_u8, dst, mask, T, H0, W0, H, W, Hp, Wp, mean, std):
    m = (C.c_float * 3)(*[float(v) for v in mean])
    sd = (C.c_float * 3)(*[float(v) for v in std])
    check(lib().tdb_frames_preprocess(ptr(src_u8), ptr(dst), ptr(mask), T, H0, W0, H, W, Hp, Wp, m, sd, stream_ptr()), "frames_preprocess")
