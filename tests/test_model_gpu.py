"""End-to-end GPU parity of the drop-in TubeDETR / SetCriterion against the reference fixtures (tests/golden/*.pt)
and against the CPU oracle.  Tolerances: BASELINE.json north_star -- 2e-2 for the bf16 path on pred_boxes / pred_sted."""
import argparse
import os

import pytest
import torch

from helpers import batch_for, load_gold, state_dict

pytestmark = pytest.mark.gpu
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _args(cfg):
    fl = cfg["flags"]
    return argparse.Namespace(
        num_queries=1, aux_loss=True, video_max_len_train=200, stride=cfg["stride"], guided_attn="--no_guided_attn" not in fl,
        fast="--no_fast" not in fl, fast_mode="", sted=True, no_tsa="--no_tsa" in fl, enc_layers=6, dec_layers=6,
        lr_backbone=1e-5, bbox_loss_coef=5, giou_loss_coef=2, sted_loss_coef=10, guided_attn_loss_coef=1, sigma=1,
        device="cuda", hidden_dim=256, nheads=8, dim_feedforward=2048, backbone="resnet101", dilation=False,
        position_embedding="sine", offline_text_encoder=True)


_MODELS = {}
# per-tensor gradient norm / sampled direction vs the fp32 reference backward.  At BASELINE.json's measured sizes (cfg2 / cfg5
# fixtures: 100 frames, thousands of tokens per weight gradient) the measured worst case is 3 % / 0.98 (median 0.3-0.5 %).  The toy
# fixtures (4-8 frames at res <= 224, as few as 25 tokens per frame) average the bf16 ReLU-mask flips over ~100x fewer samples:
# measured up to 13 % / 0.94 on single tensors there, so they keep the wider bound.
GRAD_NORM_TOL, GRAD_COS_TOL = 0.08, 0.97
GRAD_NORM_TOL_TOY, GRAD_COS_TOL_TOY = 0.15, 0.9


def _model(cfg):
    from tubedetr_b200 import build_model
    key = (cfg["stride"], tuple(cfg["flags"]))
    if key not in _MODELS:
        model, crit, wd = build_model(_args(cfg))
        sd = state_dict()
        model.load_state_dict({k: sd[k] for k in model.state_dict()}, strict=True)  # --no_fast drops fast_* keys
        _MODELS[key] = (model.cuda().eval(), crit, wd)
    return _MODELS[key]


def _run(cfg):
    from tubedetr_b200 import NestedTensor
    model, crit, wd = _model(cfg)
    b = batch_for(cfg)
    samples = NestedTensor(b["frames_slow"].cuda(), b["mask_slow"].cuda())
    fast = NestedTensor(b["frames_fast"].cuda(), b["mask_fast"].cuda())
    caps = (b["input_ids"].cuda(), b["attention_mask"].cuda())
    mc = model(samples, cfg["durations"], caps, encode_and_save=True, samples_fast=fast if model.fast else None)
    out = model(samples, cfg["durations"], caps, encode_and_save=False, memory_cache=mc)
    return model, crit, wd, b, mc, out


def _err(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item()


def _log(msg):
    os.makedirs(REPORT, exist_ok=True)
    with open(os.path.join(REPORT, "model_parity.txt"), "a") as f:
        f.write(msg + "\n")
    print(msg)


def test_backbone_features_match_reference():
    g = load_gold("cfg1")
    model, *_ = _model(g["cfg"])
    b = batch_for(g["cfg"])
    with torch.no_grad():
        W = model._engine.prepare(model._backbone_tensors())
        feat, h, w, _ = model._engine.forward(b["frames_slow"][:1].cuda(), W, save=False, tag="t")
    ref = g["feat_slow0"][0].permute(1, 2, 0).reshape(h * w, 2048)
    e = _err(feat, ref)
    _log(f"backbone feat: max err {e:.4g} (ref absmax {ref.abs().max():.3f}, mean {ref.abs().mean():.3f})")
    assert e <= 3e-2 * ref.abs().max().item()


@pytest.mark.parametrize("name", ["cfg1", "cfg1b", "nofast", "notsa"])
def test_forward_matches_reference(name):
    g = load_gold(name)
    with torch.no_grad():
        model, crit, wd, b, mc, out = _run(g["cfg"])
    for k in ("mask", "query_mask", "text_attention_mask"):
        assert torch.equal(mc[k].cpu(), g[k]), k
    for k in ("pos_embed", "query_embed"):
        assert _err(mc[k], g[k]) < 1e-4, k
    e_mem = _err(mc["img_memory"], g["img_memory"])
    e_box = _err(out["pred_boxes"], g["pred_boxes"])
    e_sted = _err(out["pred_sted"], g["pred_sted"])
    e_aux = _err(torch.stack([a["pred_boxes"] for a in out["aux_outputs"]]), g["aux_pred_boxes"])
    _log(f"{name}: img_memory err {e_mem:.4g} (absmax {g['img_memory'].abs().max():.3f}) pred_boxes {e_box:.4g} "
         f"pred_sted {e_sted:.4g} (absmax {g['pred_sted'].abs().max():.3f}) aux_boxes {e_aux:.4g}")
    assert out["pred_boxes"].shape == g["pred_boxes"].shape and out["pred_sted"].shape == g["pred_sted"].shape
    assert e_box <= 2e-2 and e_aux <= 2e-2
    assert e_sted <= 2e-2                               # BASELINE.json north_star: 2e-2 absolute on the bf16 path
    assert e_mem <= 3e-2 * g["img_memory"].abs().max().item()
    if "weights" in g and name != "notsa":
        e_w, e_cw = _err(out["weights"], g["weights"]), _err(out["ca_weights"], g["ca_weights"])
        _log(f"{name}: weights err {e_w:.4g} ca_weights err {e_cw:.4g}")
        assert e_w <= 2e-2 and e_cw <= 2e-2


@pytest.mark.parametrize("name", ["cfg1b", "cfg1"])
def test_losses_and_gradients_match_reference(name):
    g = load_gold(name)
    model, crit, wd, b, mc, out = _run(g["cfg"])
    keep = b["keep"].cuda()
    o = dict(out, pred_boxes=out["pred_boxes"][keep],
             aux_outputs=[dict(a, pred_boxes=a["pred_boxes"][keep]) for a in out["aux_outputs"]])
    targets = [{"boxes": bx[None].cuda()} for bx in b["target_boxes"]]
    losses = crit(o, targets, b["inter_idx"], b["time_mask"].cuda())
    total = sum(losses[k] * wd[k] for k in losses if k in wd)
    _log(f"{name}: total loss {total.item():.5f} vs reference {g['loss_total'].item():.5f}")
    assert abs(total.item() - g["loss_total"].item()) <= 2e-2 * abs(g["loss_total"].item())
    for k, v in g["losses"].items():
        assert abs(losses[k].item() - v.item()) <= 3e-2 * max(abs(v.item()), 0.05), (k, losses[k].item(), v.item())
    model.zero_grad()
    total.backward()
    worst = []
    checked = 0
    for k, p in model.named_parameters():
        if k not in g["grad_norm"]:
            assert p.grad is None or not p.requires_grad or p.grad.abs().max() == 0 or "pooler" in k, k
            continue
        assert p.grad is not None, k
        ref_n = g["grad_norm"][k]
        got_n = p.grad.float().norm().item()
        f = p.grad.flatten()
        step = max(f.numel() // 64, 1)
        samp = f[::step][:64].float().cpu()
        ref_s = g["grad_sample"][k]
        rel = abs(got_n - ref_n) / max(ref_n, 1e-6)
        cos = torch.nn.functional.cosine_similarity(samp, ref_s, dim=0).item() if ref_s.norm() > 0 else 1.0
        worst.append((rel, cos, k, got_n, ref_n))
        checked += 1
    worst.sort(reverse=True)
    for rel, cos, k, got_n, ref_n in worst[:12]:
        _log(f"{name}: grad {k}: norm {got_n:.4g} vs {ref_n:.4g} (rel {rel:.3f}) sample-cos {cos:.4f}")
    assert checked > 300
    # Reference gradients come from the fp32 network.  The bf16 forward flips a fraction of a percent of near-zero ReLU
    # pre-activations per layer, which removes that share of the fp32 gradient at each of the ~100 ReLUs: backbone weight
    # gradients come out 5-8 % smaller in norm with cosine >= 0.98 (tests/test_backbone_gpu.py validates the backward
    # kernels themselves to 2 % against a bf16-faithful oracle).  Hence: direction must agree, norms within 15 %.
    bad = [w for w in worst if (w[0] > GRAD_NORM_TOL_TOY or w[1] < GRAD_COS_TOL_TOY) and w[4] > 1e-4]
    assert len(bad) <= 0.02 * checked, bad[:10]
    med = sorted(w[0] for w in worst)[len(worst) // 2]
    _log(f"{name}: grad-norm rel err median {med:.4f}, params checked {checked}")
    assert med < 0.08


def test_train_mode_applies_dropout_and_backpropagates():
    """train(): every dropout of the reference is live (two passes differ, eval is unchanged), gradients stay finite."""
    g = load_gold("cfg1")
    model, crit, wd, b, mc, out_eval = _run(g["cfg"])
    model.train()
    try:
        torch.manual_seed(1)
        _, _, _, _, _, o1 = _run(g["cfg"])
        _, _, _, _, _, o2 = _run(g["cfg"])
        assert torch.isfinite(o1["pred_boxes"]).all() and torch.isfinite(o2["pred_sted"]).all()
        assert (o1["pred_boxes"] - o2["pred_boxes"]).abs().max() > 1e-4            # stochastic
        assert (o1["pred_boxes"] - out_eval["pred_boxes"]).abs().max() < 0.5       # same network
        keep = b["keep"].cuda()
        # backward belongs to the LATEST forward (o2): engine-owned activations and the dropout seed are per step
        o = dict(o2, pred_boxes=o2["pred_boxes"][keep], aux_outputs=[dict(a, pred_boxes=a["pred_boxes"][keep]) for a in o2["aux_outputs"]])
        stale = (o1["pred_boxes"].float().square().sum() + o1["pred_sted"].float().square().sum())
        with pytest.raises(RuntimeError, match="one forward per backward|overwritten by a later forward"):
            stale.backward()                                                       # o1's saved state is gone: loud, not wrong
        losses = crit(o, [{"boxes": bx[None].cuda()} for bx in b["target_boxes"]], b["inter_idx"], b["time_mask"].cuda())
        model.zero_grad()
        sum(losses[k] * wd[k] for k in losses if k in wd).backward()
        for k, p in model.named_parameters():
            if p.grad is not None:
                assert torch.isfinite(p.grad).all(), k
    finally:
        model.eval()
    _, _, _, _, _, o3 = _run(g["cfg"])
    assert torch.equal(o3["pred_boxes"], out_eval["pred_boxes"])


@pytest.mark.parametrize("name", ["cfg1b", "notsa"])
def test_decoder_cross_attention_variants_agree(name):
    """hoisted K/V projections (default) vs the per-layer fused tcgen05 kernel vs the unfused path: same outputs to bf16 noise, same
    gradients for the cross-attention parameters and for the encoder (which receives the memory gradient of all six layers)"""
    g = load_gold(name)
    cfg = g["cfg"]
    res = {}
    model = _model(cfg)[0]
    try:
        for mode in ("hoist", "fused", "unfused"):
            model.transformer.xattn_mode = mode
            model.zero_grad(set_to_none=True)
            _, _, _, _, mc, out = _run(cfg)
            (out["pred_boxes"].float().square().sum() + out["pred_sted"].float().square().sum() + out["ca_weights"].square().sum()
             if "ca_weights" in out else out["pred_boxes"].float().square().sum() + out["pred_sted"].float().square().sum()).backward()
            res[mode] = (out["pred_boxes"].detach().clone(), out["pred_sted"].detach().clone(),
                         {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None and "text_encoder" not in n})
    finally:
        model.transformer.xattn_mode = "hoist"
    for mode in ("fused", "unfused"):
        assert _err(res["hoist"][0], res[mode][0]) <= 5e-3 and _err(res["hoist"][1], res[mode][1]) <= 1e-2, mode
        assert res["hoist"][2].keys() == res[mode][2].keys()
        for n in ("transformer.decoder.layers.0.cross_attn_image.in_proj_weight", "transformer.decoder.layers.5.cross_attn_image.in_proj_bias",
                  "transformer.decoder.layers.3.cross_attn_image.out_proj.weight", "transformer.encoder.layers.5.linear2.weight",
                  "input_proj.weight"):
            a, c = res["hoist"][2][n].float(), res[mode][2][n].float()
            cos = (a * c).sum() / (a.norm() * c.norm() + 1e-20)
            assert cos > 0.995 and abs(a.norm().item() / (c.norm().item() + 1e-20) - 1) < 0.03, (mode, n, cos.item(), a.norm().item(), c.norm().item())


def test_l2_chunked_fast_pass_is_equivalent():
    """the L2-resident chunked schedule of stem+layer1+layer2 (no-grad pass) gives the same features as one big batch"""
    g = load_gold("cfg1")
    model, *_ = _model(g["cfg"])
    b = batch_for(g["cfg"])
    frames = b["frames_fast"].cuda()
    with torch.no_grad():
        W = model._engine.prepare(model._backbone_tensors())
        f0, h, w, _ = model._engine.forward(frames, W, save=False, tag="eq0")
        f0 = f0.clone()
        f1, h1, w1, _ = model._engine.forward(frames, W, save=False, tag="eq1", l2_chunk=3)
    assert (h, w) == (h1, w1) and f0.shape == f1.shape
    assert (f0.float() - f1.float()).abs().max().item() <= 1e-2 * f0.float().abs().max().item()


def test_joint_backbone_batch_equals_two_passes():
    """slow (with grad) + fast (no grad) frames as ONE backbone batch give the outputs and the backbone weight gradients of
    the reference's two separate backbone calls (models/tubedetr.py:120-131): rows of a GEMM are independent."""
    from tubedetr_b200 import NestedTensor
    g = load_gold("cfg1")
    cfg = g["cfg"]
    model, crit, wd = _model(cfg)
    b = batch_for(cfg)
    samples = NestedTensor(b["frames_slow"].cuda(), b["mask_slow"].cuda())
    fast = NestedTensor(b["frames_fast"].cuda(), b["mask_fast"].cuda())
    caps = (b["input_ids"].cuda(), b["attention_mask"].cuda())
    res = {}
    for joint in (True, False):
        model.joint_backbone = joint
        model.zero_grad(set_to_none=True)
        mc = model(samples, cfg["durations"], caps, encode_and_save=True, samples_fast=fast)
        out = model(samples, cfg["durations"], caps, encode_and_save=False, memory_cache=mc)
        (out["pred_boxes"].float().square().sum() + out["pred_sted"].float().square().sum()).backward()
        gr = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None and "backbone" in n}
        res[joint] = (out["pred_boxes"].detach().clone(), out["pred_sted"].detach().clone(), gr)
    model.joint_backbone = True
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    assert res[True][2].keys() == res[False][2].keys() and len(res[True][2]) > 50
    for n in res[True][2]:
        a, c = res[True][2][n], res[False][2][n]
        assert (a - c).abs().max().item() <= 1e-5 * (c.abs().max().item() + 1e-12), n


@pytest.mark.parametrize("name", ["cfg1", "notsa"])
def test_slow_frames_alias_fast_dedup(name):
    """opt-in contract `slow_frames_alias_fast` (the slow frames ARE fast frames [::k], reference datasets/vidstg.py:250-251):
    the backbone runs once per distinct frame.  Same parity with the reference fixtures as the two-pass path, and the two
    paths agree with each other to bf16 noise (the batch size selects different tile schedules)."""
    from tubedetr_b200 import NestedTensor
    g = load_gold(name)
    cfg = g["cfg"]
    model, crit, wd = _model(cfg)
    b = batch_for(cfg)
    assert torch.equal(b["frames_slow"], b["frames_fast"][::cfg["stride"]])          # the contract holds for these inputs
    samples = NestedTensor(b["frames_slow"].cuda(), b["mask_slow"].cuda())
    fast = NestedTensor(b["frames_fast"].cuda(), b["mask_fast"].cuda())
    caps = (b["input_ids"].cuda(), b["attention_mask"].cuda())
    res = {}
    for flag in (False, True):
        model.slow_frames_alias_fast = flag
        model.zero_grad(set_to_none=True)
        mc = model(samples, cfg["durations"], caps, encode_and_save=True, samples_fast=fast)
        out = model(samples, cfg["durations"], caps, encode_and_save=False, memory_cache=mc)
        (out["pred_boxes"].float().square().sum() + out["pred_sted"].float().square().sum()).backward()
        res[flag] = (out, mc, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    model.slow_frames_alias_fast = False
    out, mc, gr = res[True]
    e_box, e_sted = _err(out["pred_boxes"], g["pred_boxes"]), _err(out["pred_sted"], g["pred_sted"])
    e_mem = _err(mc["img_memory"], g["img_memory"])
    _log(f"{name} (dedup): img_memory err {e_mem:.4g} pred_boxes {e_box:.4g} pred_sted {e_sted:.4g}")
    assert e_box <= 2e-2 and e_sted <= 2e-2
    assert e_mem <= 3e-2 * g["img_memory"].abs().max().item()
    assert _err(out["pred_boxes"], res[False][0]["pred_boxes"]) <= 1e-2
    assert gr.keys() == res[False][2].keys()
    for n in ("input_proj.weight", "backbone.0.body.layer3.5.conv2.weight", "transformer.fast_encoder.weight"):
        a, c = gr[n].float(), res[False][2][n].float()
        cos = (a * c).sum() / (a.norm() * c.norm() + 1e-20)
        assert cos > 0.99, (n, cos.item())


# ----------------------------------------------------------------------------------------------- full-size reference fixtures
# tests/golden/{cfg2,cfg4,cfg5}.pt come from the UNMODIFIED reference at BASELINE.json's measured sizes (make_golden.py):
# cfg2 = configs[1] (B=1, T=100, k=4, res 352, L=20), cfg4 = the --no_fast T=200 k=2 shape of configs[3] at one clip,
# cfg5 = configs[4] (--no_tsa --no_guided_attn, T=100, res 224, k=2).  img_memory / pos_embed are stored as [::5, ::7] samples.
def _fullsize_outputs(name):
    g = load_gold(name)
    with torch.no_grad():
        model, crit, wd, b, mc, out = _run(g["cfg"])
    return g, model, crit, wd, b, mc, out


@pytest.mark.parametrize("name", ["cfg2", "cfg4", "cfg5"])
def test_fullsize_forward_matches_reference(name):
    g, model, crit, wd, b, mc, out = _fullsize_outputs(name)
    ts, fs = g["sample_steps"]
    for k in ("mask", "query_mask", "text_attention_mask"):
        assert torch.equal(mc[k].cpu(), g[k]), k
    assert _err(mc["pos_embed"][::ts, ::fs], g["pos_embed_sample"]) < 1e-4
    assert _err(mc["query_embed"], g["query_embed"]) < 1e-4
    e_mem = _err(mc["img_memory"][::ts, ::fs], g["img_memory_sample"])
    e_box = _err(out["pred_boxes"], g["pred_boxes"])
    e_sted = _err(out["pred_sted"], g["pred_sted"])
    e_aux = _err(torch.stack([a["pred_boxes"] for a in out["aux_outputs"]]), g["aux_pred_boxes"])
    e_auxs = _err(torch.stack([a["pred_sted"] for a in out["aux_outputs"]]), g["aux_pred_sted"])
    _log(f"{name} (full size): img_memory err {e_mem:.4g} (absmax {g['img_memory_sample'].abs().max():.3f}) pred_boxes {e_box:.4g} "
         f"pred_sted {e_sted:.4g} (absmax {g['pred_sted'].abs().max():.3f}) aux_boxes {e_aux:.4g} aux_sted {e_auxs:.4g}")
    assert out["pred_boxes"].shape == g["pred_boxes"].shape and out["pred_sted"].shape == g["pred_sted"].shape
    assert e_box <= 2e-2 and e_aux <= 2e-2 and e_sted <= 2e-2 and e_auxs <= 2e-2
    assert e_mem <= 3e-2 * g["img_memory_sample"].abs().max().item()
    if "weights" in g:
        e_w, e_cw = _err(out["weights"], g["weights"]), _err(out["ca_weights"], g["ca_weights"])
        e_aw = _err(torch.stack([a["weights"] for a in out["aux_outputs"]]), g["aux_weights"])
        _log(f"{name} (full size): weights err {e_w:.4g} aux weights {e_aw:.4g} ca_weights err {e_cw:.4g}")
        assert e_w <= 2e-2 and e_cw <= 2e-2 and e_aw <= 2e-2


@pytest.mark.parametrize("name", ["cfg2", "cfg5"])
def test_fullsize_losses_and_gradients_match_reference(name):
    g = load_gold(name)
    model, crit, wd, b, mc, out = _run(g["cfg"])
    keep = b["keep"].cuda()
    o = dict(out, pred_boxes=out["pred_boxes"][keep],
             aux_outputs=[dict(a, pred_boxes=a["pred_boxes"][keep]) for a in out["aux_outputs"]])
    targets = [{"boxes": bx[None].cuda()} for bx in b["target_boxes"]]
    losses = crit(o, targets, b["inter_idx"], b["time_mask"].cuda())
    total = sum(losses[k] * wd[k] for k in losses if k in wd)
    _log(f"{name} (full size): total loss {total.item():.5f} vs reference {g['loss_total'].item():.5f}")
    assert abs(total.item() - g["loss_total"].item()) <= 2e-2 * abs(g["loss_total"].item())
    for k, v in g["losses"].items():
        assert abs(losses[k].item() - v.item()) <= 3e-2 * max(abs(v.item()), 0.05), (k, losses[k].item(), v.item())
    model.zero_grad(set_to_none=True)
    total.backward()
    worst = []
    for k, p in model.named_parameters():
        if k not in g["grad_norm"]:
            continue
        assert p.grad is not None, k
        ref_n, got_n = g["grad_norm"][k], p.grad.float().norm().item()
        f = p.grad.flatten()
        samp = f[::max(f.numel() // 64, 1)][:64].float().cpu()
        ref_s = g["grad_sample"][k]
        cos = torch.nn.functional.cosine_similarity(samp, ref_s, dim=0).item() if ref_s.norm() > 0 else 1.0
        worst.append((abs(got_n - ref_n) / max(ref_n, 1e-6), cos, k, got_n, ref_n))
    worst.sort(reverse=True)
    for rel, cos, k, got_n, ref_n in worst[:8]:
        _log(f"{name} (full size): grad {k}: norm {got_n:.4g} vs {ref_n:.4g} (rel {rel:.3f}) sample-cos {cos:.4f}")
    assert len(worst) > 300
    bad = [w for w in worst if (w[0] > GRAD_NORM_TOL or w[1] < GRAD_COS_TOL) and w[4] > 1e-4]
    assert len(bad) <= 0.02 * len(worst), bad[:10]
    med = sorted(w[0] for w in worst)[len(worst) // 2]
    _log(f"{name} (full size): grad-norm rel err median {med:.4f}, params checked {len(worst)}")
    assert med < 0.08
