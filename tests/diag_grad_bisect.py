"""GPU-box diagnostic (a checker, hence under tests/: it runs the CPU oracle; not collected by pytest): per-parameter-group
gradient error (signed) vs the reference fixture, and intermediate-tensor gradients (feat, src, enc, mem) vs the CPU oracle.
  python tests/diag_grad_bisect.py        -> gpurun_out/grad_debug.txt"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import batch_for, load_gold, run_oracle, state_dict  # noqa: E402
import test_model_gpu as T  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/grad_debug.txt", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    log.write(s + "\n")
    log.flush()


def group(k):
    if k.startswith("backbone"):
        parts = k.split(".")
        return "backbone." + parts[3] + "." + parts[5].split(".")[0] if parts[3].startswith("layer") else "backbone.stem"
    if "text_encoder" in k:
        return "text_encoder"
    if "encoder.layers" in k:
        return "enc." + k.split(".")[3] + "." + ".".join(k.split(".")[4:6])
    if "decoder.layers" in k:
        return "dec." + k.split(".")[3] + "." + ".".join(k.split(".")[4:6])
    return k


def main(name="cfg1b"):
    g = load_gold(name)
    cfg = g["cfg"]
    # intermediate grads: hook my tensors
    from tubedetr_b200 import ops
    captured = {}
    orig_linear = ops.linear
    model0, *_ = T._model(cfg)
    tr = model0.transformer
    enc_io = []
    orig = tr._enc_layer

    def wrapped(l, x32, *a, **k):
        if not enc_io:
            x32.retain_grad()
            enc_io.append(x32)
        res = orig(l, x32, *a, **k)
        res[0].retain_grad()
        enc_io.append(res[0])
        return res
    tr._enc_layer = wrapped
    feat_g = {}
    orig_lin = ops.linear

    def lin_wrapped(x, W, b, *a, **k):
        if "feat" not in feat_g and x.shape[1] == 2048 and x.requires_grad:
            feat_g["feat"] = x
            x.register_hook(lambda g: feat_g.__setitem__("g", g.detach().float().cpu()))
        return orig_lin(x, W, b, *a, **k)
    ops.linear = lin_wrapped
    model, crit, wd, b, mc, out = T._run(cfg)
    tr._enc_layer = orig
    ops.linear = orig_lin
    keep = b["keep"].cuda()
    o = dict(out, pred_boxes=out["pred_boxes"][keep], aux_outputs=[dict(a, pred_boxes=a["pred_boxes"][keep]) for a in out["aux_outputs"]])
    targets = [{"boxes": bx[None].cuda()} for bx in b["target_boxes"]]
    losses = crit(o, targets, b["inter_idx"], b["time_mask"].cuda())
    total = sum(losses[k] * wd[k] for k in losses if k in wd)
    mem = mc["img_memory"]
    mem.retain_grad()
    model.zero_grad()
    total.backward()
    P("loss", total.item(), "ref", g["loss_total"].item())
    rows = collections.defaultdict(list)
    for k, p in model.named_parameters():
        if k in g["grad_norm"] and p.grad is not None and g["grad_norm"][k] > 1e-5:
            rows[group(k)].append(p.grad.float().norm().item() / g["grad_norm"][k] - 1.0)
    for grp in sorted(rows):
        v = rows[grp]
        P(f"{grp:50s} n={len(v):3d} mean signed rel err {sum(v) / len(v):+.4f}  min {min(v):+.4f} max {max(v):+.4f}")
    # oracle intermediate gradients on CPU
    from oracle import tubedetr_oracle as O
    sd = {k: v.clone() for k, v in state_dict().items()}
    if "--no_fast" in cfg["flags"]:
        sd = {k: v for k, v in sd.items() if "fast_" not in k}
    for k in ("input_proj.weight", "backbone.0.body.layer2.0.conv1.weight"):
        sd[k].requires_grad_(True)
    oout, cache, ob = run_oracle(cfg, sd=sd)
    ol = O.criterion(oout, ob["target_boxes"], ob["inter_idx"], ob["time_mask"], ob["keep"])
    owd = O.weight_dict()
    ototal = sum(ol[k] * owd[k] for k in ol)
    gl = torch.autograd.grad(ototal, cache["enc_layers"], retain_graph=True)
    for i, (mine, ref) in enumerate(zip(enc_io, gl)):
        m = mine.grad.float().cpu().view_as(ref)
        a = (m * ref).sum() / (ref * ref).sum()
        cos = torch.nn.functional.cosine_similarity(m.flatten(), ref.flatten(), dim=0)
        fm, fr = (mine.detach().float().cpu().view_as(ref) - cache["enc_layers"][i].detach()), cache["enc_layers"][i].detach()
        P(f"enc boundary {i}: grad proj a={a.item():.4f} cos={cos.item():.4f} norm mine {m.norm().item():.4f} ref {ref.norm().item():.4f}"
          f" | fwd act rel err {fm.norm().item() / fr.norm().item():.4f}")
    gfeat = torch.autograd.grad(ototal, cache["feat_slow"], retain_graph=True)[0]          # (n,2048,h,w)
    ref_f = gfeat.permute(0, 2, 3, 1).reshape(-1, 2048)
    mine_f = feat_g["g"].view_as(ref_f)
    a = (mine_f * ref_f).sum() / (ref_f * ref_f).sum()
    P(f"d/d feat (backbone output, before ReLU mask): proj a={a.item():.4f} cos={torch.nn.functional.cosine_similarity(mine_f.flatten(), ref_f.flatten(), dim=0).item():.4f}"
      f" norm mine {mine_f.norm().item():.5f} ref {ref_f.norm().item():.5f}")
    gsrc = torch.autograd.grad(ototal, cache["src"], retain_graph=True)[0]                 # (n,256,h,w)
    ref_s = gsrc.flatten(2).transpose(1, 2)
    HW = ref_s.shape[1]
    mine_s = enc_io[0].grad.float().cpu().view(ref_s.shape[0], -1, 256)[:, :HW]
    a = (mine_s * ref_s).sum() / (ref_s * ref_s).sum()
    P(f"d/d src (visual rows of encoder input): proj a={a.item():.4f} norm mine {mine_s.norm().item():.5f} ref {ref_s.norm().item():.5f}")
    # isolate LinearFn.backward: feed the oracle's exact d src and feat through our dgrad GEMM
    from tubedetr_b200.gemm import gemm
    Wv = model.input_proj.weight.detach().view(256, 2048)
    dyb = ref_s.reshape(-1, 256).to(torch.bfloat16).cuda().contiguous()
    dx = torch.empty(dyb.shape[0], 2048, dtype=torch.bfloat16, device="cuda")
    gemm(dyb, Wv.to(torch.bfloat16).contiguous(), dx, dyb.shape[0], 2048, 256, b_major=1)
    iso = dx.float().cpu()
    a = (iso * ref_f).sum() / (ref_f * ref_f).sum()
    P(f"isolated dgrad GEMM on oracle d src: proj a={a.item():.4f} norm {iso.norm().item():.5f} ref {ref_f.norm().item():.5f}")
    gs = torch.autograd.grad(ototal, [cache["src"], cache["enc"], cache["mem"]], allow_unused=True)
    P("oracle d/dsrc norm", gs[0].norm().item(), "d/denc", gs[1].norm().item(), "d/dmem", gs[2].norm().item())
    mg = mem.grad.transpose(0, 1).float().cpu()
    P("mine   d/dmem norm", mg.norm().item(), "cos", torch.nn.functional.cosine_similarity(mg.flatten(), gs[2].flatten(), dim=0).item())


if __name__ == "__main__":
    main(*sys.argv[1:])
