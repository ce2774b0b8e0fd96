"""Isolated ResNet-101 backward (dgrad/wgrad chain on the tcgen05 GEMM) vs autograd through the CPU oracle.

The oracle runs with emulate_bf16=True (weights / stored activations rounded to bf16 where the CUDA path rounds them):
the backward pass is only comparable when the ~100 ReLU masks of the two forwards agree.  Against the pure-fp32 network
the bf16 forward flips the sign of a fraction of a percent of near-zero pre-activations per layer, which removes that
share of the reference gradient at every ReLU (measured on B200: projection coefficient 0.94-0.96 with cosine 0.98-0.99
at layer2, i.e. an apparent 5 % shrink that is a property of bf16 forward numerics, not of the backward kernels)."""
import pytest
import torch

from helpers import state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,H,W", [(2, 96, 96), (3, 64, 128)])
def test_backbone_weight_gradients_match_oracle(N, H, W):
    from oracle import tubedetr_oracle as O
    from tubedetr_b200 import ops
    from tubedetr_b200.resnet import ResNet101Engine
    sd = state_dict()
    pre = "backbone.0.body."
    names = [k for k in sd if k.startswith(pre) and k.endswith("weight") and sd[k].dim() == 4
             and any(f"layer{i}" in k for i in (2, 3, 4))]
    g = torch.Generator().manual_seed(3)
    frames = torch.randn(N, 3, H, W, generator=g)
    # oracle
    osd = {k: v.clone() for k, v in sd.items() if k.startswith(pre)}
    for k in names:
        osd[k].requires_grad_(True)
    feat = O.resnet101_layer4(frames, osd, emulate_bf16=True)    # (N,2048,h,w)
    gout = torch.randn(feat.shape, generator=g)
    ref = torch.autograd.grad((feat * gout).sum(), [osd[k] for k in names])
    # ours
    eng = ResNet101Engine()
    dsd = {k: v.cuda() for k, v in sd.items() if k.startswith(pre)}
    params = [dsd[k].requires_grad_(True) for k in names]
    Wt = eng.prepare(dsd)
    f = ops.BackboneFn.apply(frames.cuda(), eng, Wt, names, "t", *params)
    h, w = eng.last_hw
    ref_rows = feat.detach().permute(0, 2, 3, 1).reshape(N * h * w, 2048)
    err = (f.float().cpu() - ref_rows).abs().max().item()
    assert err <= 2e-2 * ref_rows.abs().max().item(), err
    gr = gout.permute(0, 2, 3, 1).reshape(N * h * w, 2048).cuda()
    (f.float() * gr).sum().backward()
    worst = []
    for k, p, r in zip(names, params, ref):
        m = p.grad.float().cpu()
        a = ((m * r).sum() / (r * r).sum()).item()
        cos = torch.nn.functional.cosine_similarity(m.flatten(), r.flatten(), dim=0).item()
        worst.append((abs(a - 1), cos, k, a))
    worst.sort(reverse=True)
    print("worst projection coefficients:", [(k, round(a, 4), round(c, 4)) for _, c, k, a in worst[:6]])
    # measured on B200: worst projection 0.965, worst cosine 0.992 (residual ReLU-mask flips from accumulation order)
    assert worst[0][0] < 0.05, worst[:5]
    assert min(c for _, c, _, _ in worst) > 0.985


def test_engine_memory_is_bounded_under_varying_shapes():
    """The reference's data pipeline changes frame count and resolution nearly every iteration (durations up to
    video_max_len, RandomResize).  Engine buffers are ONE allocation per tag sized to the largest shape seen: after the
    largest shape has run, further shapes allocate nothing, and results do not depend on what ran before (halo buffers are
    re-zeroed on a shape change)."""
    from tubedetr_b200.resnet import ResNet101Engine
    sd = state_dict()
    pre = "backbone.0.body."
    eng = ResNet101Engine()
    dsd = {k: v.cuda() for k, v in sd.items() if k.startswith(pre)}
    g = torch.Generator().manual_seed(11)
    shapes = [(4, 128, 160), (2, 96, 96), (3, 64, 128), (1, 128, 96), (4, 128, 160), (2, 96, 96)]
    frames = {s: torch.randn(s[0], 3, s[1], s[2], generator=g).cuda() for s in set(shapes)}
    with torch.no_grad():
        Wt = eng.prepare(dsd)
        fresh = ResNet101Engine()
        Wf = fresh.prepare(dsd)
        ref, _, _, _ = fresh.forward(frames[(2, 96, 96)], Wf, save=False, tag="m")
        ref = ref.clone()
        sizes = []
        for s in shapes:
            out, h, w, _ = eng.forward(frames[s], Wt, save=False, tag="m")
            torch.cuda.synchronize()
            sizes.append(eng.allocated_bytes())
    assert sizes[0] == max(sizes) == sizes[-1], sizes          # the largest shape ran first: nothing grows afterwards
    assert torch.equal(out, ref)                               # (2, 96, 96) after bigger shapes == on a fresh engine


def test_backward_after_a_later_forward_raises():
    """saved activations are engine-owned buffers: a second forward with the same tag before the backward must raise, not
    return silently wrong gradients"""
    from tubedetr_b200 import ops
    from tubedetr_b200.resnet import ResNet101Engine
    sd = state_dict()
    pre = "backbone.0.body."
    names = [k for k in sd if k.startswith(pre) and k.endswith("weight") and sd[k].dim() == 4
             and any(f"layer{i}" in k for i in (2, 3, 4))]
    eng = ResNet101Engine()
    dsd = {k: v.cuda() for k, v in sd.items() if k.startswith(pre)}
    params = [dsd[k].requires_grad_(True) for k in names]
    Wt = eng.prepare(dsd)
    x = torch.randn(1, 3, 64, 64, device="cuda")
    f1 = ops.BackboneFn.apply(x, eng, Wt, names, "t", *params)
    f2 = ops.BackboneFn.apply(x + 1, eng, Wt, names, "t", *params)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        f1.float().sum().backward()
    f2.float().sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
