"""Optimizer-side step: the CPU oracle against the reference's own call sequence (golden), and the CUDA
kernels of tdb_optim.cu against the oracle / golden through the C ABI (FusedAdamWEMA)."""
import copy
import os

import pytest
import torch

from helpers import load_gold

TOL = dict(atol=2e-6, rtol=2e-5)        # fp32 arithmetic with a different association / FMA contraction


def _group_lr(name, h):
    return h["lr_backbone"] if "backbone" in name else h["text_encoder_lr"] if "text_encoder" in name else h["lr"]


def test_oracle_matches_reference_sequence():
    from oracle import optim_oracle as OO
    g = load_gold("optim")
    h = g["hyper"]
    names = [n for n, _ in g["shapes"]]
    params = [p.clone() for p in g["init"]]
    ema = [p.clone() for p in g["init"]]
    m, v = [torch.zeros_like(p) for p in params], [torch.zeros_like(p) for p in params]
    lrs = [_group_lr(n, h) for n in names]
    for step, grads in enumerate(g["grads"], start=1):
        norm = OO.adamw_ema_step(params, [x.clone() for x in grads], m, v, ema, lrs, h["weight_decay"], step,
                                 max_norm=h["max_norm"], ema_decay=h["ema_decay"])
        torch.testing.assert_close(norm, g["norm"][step - 1], rtol=1e-6, atol=0)
        for a, b in zip(params, g["params"][step - 1]):
            torch.testing.assert_close(a, b, atol=1e-7, rtol=1e-6)
        for a, b in zip(ema, g["ema"][step - 1]):
            torch.testing.assert_close(a, b, atol=1e-7, rtol=1e-6)


class _Holder(torch.nn.Module):
    def __init__(self, shapes, init):
        super().__init__()
        self._names = [n for n, _ in shapes]
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(p.clone()) for p in init])

    def named_parameters(self, *a, **k):          # reference-style dotted names select the LR groups
        return iter(zip(self._names, self.ps))


@pytest.mark.gpu
def test_fused_step_matches_golden_gpu():
    from tubedetr_b200.optim import FusedAdamWEMA
    g = load_gold("optim")
    h = g["hyper"]
    model = _Holder(g["shapes"], g["init"]).cuda()
    ema = copy.deepcopy(model)
    opt = FusedAdamWEMA(model, lr=h["lr"], lr_backbone=h["lr_backbone"], text_encoder_lr=h["text_encoder_lr"],
                        weight_decay=h["weight_decay"], max_norm=h["max_norm"], model_ema=ema, ema_decay=h["ema_decay"])
    assert [len(pg["params"]) for pg in opt.param_groups] == [5, 2, 2]          # default / backbone / text_encoder
    for step, grads in enumerate(g["grads"], start=1):
        opt.zero_grad(set_to_none=True)
        for p, gr in zip(model.ps, grads):
            p.grad = gr.cuda()
        v0 = [p._version for p in model.ps]
        norm = opt.step()
        assert all(p._version > a for p, a in zip(model.ps, v0))                  # caches keyed on _version must see the update
        torch.testing.assert_close(norm.cpu(), g["norm"][step - 1], rtol=1e-6, atol=0)
        for a, b in zip(model.ps, g["params"][step - 1]):
            torch.testing.assert_close(a.detach().cpu(), b, **TOL)
        for a, b in zip(ema.ps, g["ema"][step - 1]):
            torch.testing.assert_close(a.detach().cpu(), b, **TOL)
    # bf16 operand mirror of the transformer group = bf16(new fp32 weights), bit exact
    for p, mv in opt._mirror_views:
        assert torch.equal(mv, p.detach().to(torch.bfloat16))


@pytest.mark.gpu
@pytest.mark.parametrize("n_rest,n_back,n_text", [(1 << 20, 3 << 18, (1 << 21) + 40), (7, 0, 13)])
def test_fused_step_matches_oracle_random_sizes_gpu(n_rest, n_back, n_text):
    """odd sizes (tails, empty group), no clipping / no EMA variants, 3 steps; the reference sequence (torch AdamW on the GPU
    as the library computes it) is the second witness"""
    from oracle import optim_oracle as OO
    from tubedetr_b200.optim import FusedAdamWEMA
    gen = torch.Generator().manual_seed(3)
    shapes = [("transformer.decoder.w", (n_rest // 2 or 1, 2)), ("transformer.decoder.b", (max(n_rest % 5, 1),)),
              ("transformer.text_encoder.w", (n_text,))]
    if n_back:
        shapes.insert(2, ("backbone.0.body.w", (n_back // 9, 1, 3, 3)))
    init = [torch.randn(s, generator=gen) for _, s in shapes]
    for max_norm, with_ema in ((0.1, True), (0.0, False)):
        model = _Holder(shapes, init).cuda()
        ema = copy.deepcopy(model) if with_ema else None
        opt = FusedAdamWEMA(model, lr=3e-3, lr_backbone=1e-3, text_encoder_lr=2e-3, weight_decay=1e-2, max_norm=max_norm,
                            model_ema=ema, ema_decay=0.9)
        params = [p.clone() for p in init]
        emas = [p.clone() for p in init] if with_ema else None
        m, v = [torch.zeros_like(p) for p in params], [torch.zeros_like(p) for p in params]
        lrs = [1e-3 if "backbone" in n else 2e-3 if "text_encoder" in n else 3e-3 for n, _ in shapes]
        for step in range(1, 4):
            grads = [torch.randn(s, generator=gen) * (5.0 if step == 2 else 1e-4) for _, s in shapes]
            opt.zero_grad(set_to_none=True)
            for p, gr in zip(model.ps, grads):
                p.grad = gr.cuda()
            opt.step()
            OO.adamw_ema_step(params, grads, m, v, emas, lrs, 1e-2, step, max_norm=max_norm, ema_decay=0.9)
            for a, b in zip(model.ps, params):
                torch.testing.assert_close(a.detach().cpu(), b, **TOL)
            if with_ema:
                for a, b in zip(ema.ps, emas):
                    torch.testing.assert_close(a.detach().cpu(), b, **TOL)


@pytest.mark.gpu
def test_grad_norm_is_deterministic_and_accurate_gpu():
    import ctypes as C
    from tubedetr_b200._lib import check, lib, ptr, stream_ptr
    n = 184_640_000 // 8 + 3
    g = torch.randn(n, device="cuda")
    lib().tdb_optim_workspace_bytes.restype = C.c_int64
    wsb = int(lib().tdb_optim_workspace_bytes())
    ws = torch.empty(wsb // 8, dtype=torch.float64, device="cuda")
    out = torch.zeros(2, device="cuda")
    for i in range(2):
        check(lib().tdb_grad_sqnorm(ptr(g), C.c_int64(n), ptr(ws), C.c_int64(wsb), C.c_void_p(out.data_ptr() + 4 * i), stream_ptr()))
    torch.cuda.synchronize()
    assert out[0].item() == out[1].item()
    ref = g.double().norm().item()
    assert abs(out[0].item() - ref) / ref < 1e-6


@pytest.mark.gpu
def test_state_dict_round_trip_and_layout_check_gpu():
    """moments + step survive a state_dict round trip into a fresh optimizer of the same model; a checkpoint whose flat-buffer
    layout (names / offsets / sizes) differs is refused instead of being copied into the wrong parameters' moments"""
    from tubedetr_b200.optim import FusedAdamWEMA
    gen = torch.Generator().manual_seed(5)
    shapes = [("transformer.decoder.w", (33, 2)), ("transformer.decoder.b", (5,)), ("backbone.0.body.w", (4, 1, 3, 3)),
              ("transformer.text_encoder.w", (77,))]
    init = [torch.randn(s, generator=gen) for _, s in shapes]
    a = _Holder(shapes, init).cuda()
    oa = FusedAdamWEMA(a, lr=3e-3, lr_backbone=1e-3, text_encoder_lr=2e-3, weight_decay=1e-2, max_norm=0.1)
    for _ in range(2):
        oa.zero_grad(set_to_none=True)
        for p in a.ps:
            p.grad = torch.randn(p.shape, generator=gen).cuda()
        oa.step()
    sd = oa.state_dict()
    b = _Holder(shapes, [p.detach().cpu().clone() for p in a.ps]).cuda()
    ob = FusedAdamWEMA(b, lr=3e-3, lr_backbone=1e-3, text_encoder_lr=2e-3, weight_decay=1e-2, max_norm=0.1)
    ob.load_state_dict(sd)
    assert ob.step_count == 2 and torch.equal(ob.exp_avg, oa.exp_avg) and torch.equal(ob.exp_avg_sq, oa.exp_avg_sq)
    g = [torch.randn(p.shape, generator=gen).cuda() for p in a.ps]
    for o, mdl in ((oa, a), (ob, b)):
        o.zero_grad(set_to_none=True)
        for p, gr in zip(mdl.ps, g):
            p.grad = gr.clone()
        o.step()
    for pa, pb in zip(a.ps, b.ps):
        assert torch.equal(pa, pb)
    other = [("transformer.decoder.w", (33, 2)), ("transformer.decoder.b", (6,)), ("backbone.0.body.w", (4, 1, 3, 3)),
             ("transformer.text_encoder.w", (76,))]
    c = _Holder(other, [torch.randn(s, generator=gen) for _, s in other]).cuda()
    oc = FusedAdamWEMA(c, max_norm=0.1)
    with pytest.raises(ValueError, match="layout"):
        oc.load_state_dict(sd)
