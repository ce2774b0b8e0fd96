"""Where does the step go?  CUDA-graph replays of sub-sets of the bench step (cfg-2), each timed with CUDA events.
  python tools/step_ablation.py > gpurun_out/step_ablation.txt
Variants: full train-mode step | eval-mode step (no dropout kernels) | text encoder frozen (no RoBERTa backward) |
forward + loss only | backbone forward only (125 frames) | backbone forward + backward (25 slow frames with grad)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
model, crit, wd = bench.build_everything(dev)


def timed_graph(body, reps=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


st = bench.Step(model, crit, wd, dev, 0, 1, use_graph=True)


class _Rows(list):
    def append(self, item):          # item = (name, thunk result) -- thunks are evaluated eagerly below; keep failures local
        super().append(item)


rows = _Rows()
_tg = timed_graph


def timed_graph(body, reps=10):     # noqa: F811  (wrap: one failing variant must not take the others down)
    try:
        return _tg(body, reps)
    except Exception as e:
        torch.cuda.synchronize()
        print(f"variant failed: {type(e).__name__}: {e}", file=sys.stderr)
        return float("nan")


rows.append(("full step, train mode (bench)", timed_graph(st.body)))
model.eval()
rows.append(("full step, eval mode (no dropout)", timed_graph(st.body)))
model.train()
tp = [p for n, p in model.named_parameters() if "text_encoder" in n and p.requires_grad]
for p in tp:
    p.requires_grad_(False)
st2 = bench.Step(model, crit, wd, dev, 0, 1, use_graph=True)
rows.append(("train mode, text encoder frozen (no RoBERTa backward)", timed_graph(st2.body)))
for p in tp:
    p.requires_grad_(True)


def fwd_only():
    mc = model(st.samples, [bench.T_FRAMES], st.caps, encode_and_save=True, samples_fast=st.fast)
    out = model(st.samples, [bench.T_FRAMES], st.caps, encode_and_save=False, memory_cache=mc)
    out = dict(out, pred_boxes=out["pred_boxes"][st.keep], aux_outputs=[dict(a, pred_boxes=a["pred_boxes"][st.keep]) for a in out["aux_outputs"]])
    losses = crit(out, st.targets, st.inter_idx, st.time_mask)
    st.loss.copy_(sum(losses[k] * wd[k] for k in losses if k in wd).detach())


rows.append(("forward + loss only (train mode, autograd graph built, no backward)", timed_graph(fwd_only)))


def fwd_nograd():
    with torch.no_grad():
        fwd_only()


rows.append(("forward + loss under no_grad (nothing saved)", timed_graph(fwd_nograd)))


def enc_nograd():
    with torch.no_grad():
        model(st.samples, [bench.T_FRAMES], st.caps, encode_and_save=True, samples_fast=st.fast)


def encdec_nograd():
    with torch.no_grad():
        mc = model(st.samples, [bench.T_FRAMES], st.caps, encode_and_save=True, samples_fast=st.fast)
        model(st.samples, [bench.T_FRAMES], st.caps, encode_and_save=False, memory_cache=mc)


rows.append(("  encode phase only under no_grad (backbone + text + encoder + fast branch)", timed_graph(enc_nograd)))
rows.append(("  encode + decode phases under no_grad (no criterion)", timed_graph(encdec_nograd)))


def step_trivial_loss():
    for p in st.fgb.params:
        p.grad = None
    mc = model(st.samples, [bench.T_FRAMES], st.caps, encode_and_save=True, samples_fast=st.fast)
    out = model(st.samples, [bench.T_FRAMES], st.caps, encode_and_save=False, memory_cache=mc)
    tot = torch.cat([out["pred_boxes"].flatten(), out["pred_sted"].flatten()] + [a["pred_boxes"].flatten() for a in out["aux_outputs"]]
                    + [a["pred_sted"].flatten() for a in out["aux_outputs"]]).square().sum()
    tot.backward()


rows.append(("full step with a 3-kernel stand-in loss instead of SetCriterion (criterion fwd+bwd = bench - this)", timed_graph(step_trivial_loss)))
W = model._engine.prepare(model._backbone_tensors())
fr_s, fr_f = st.samples.tensors.float(), st.fast.tensors.float()


def backbone_fwd():
    with torch.no_grad():
        model._engine.forward([fr_s, fr_f], W, save=False, tag="abl")


rows.append(("backbone forward only, 125 frames joint, nothing saved", timed_graph(backbone_fwd)))
sd = model._backbone_tensors()
names = [n for n, p in sd.items() if isinstance(p, torch.nn.Parameter) and p.requires_grad]
from tubedetr_b200 import ops  # noqa: E402


def backbone_fwd_bwd():
    for n in names:
        sd[n].grad = None
    feat, _ = ops.BackboneJointFn.apply(fr_s, fr_f, model._engine, W, names, "abl2", *[sd[n] for n in names])
    feat.float().sum().backward()


rows.append(("backbone forward (125 frames) + backward (25 slow frames)", timed_graph(backbone_fwd_bwd)))
eng = model._engine
N = fr_s.shape[0] + fr_f.shape[0]
H, Wd = fr_s.shape[2:]


def upto(li_hi):
    def f():
        with torch.no_grad():
            x = eng._stem([fr_s, fr_f], W, "abl3")
            if li_hi:
                eng._stages(x, N, W, False, "abl3", 1, li_hi, H, Wd)
    return f


prev = 0.0
for li, nm in enumerate(["stem (im2col + conv1 GEMM + maxpool)", "+ layer1", "+ layer2", "+ layer3", "+ layer4"]):
    t = timed_graph(upto(li))
    rows.append((f"backbone forward 125 frames up to: {nm}   [stage alone: {t - prev:.3f} ms]", t))
    prev = t
model.slow_frames_alias_fast = True
st3 = bench.Step(model, crit, wd, dev, 0, 1, use_graph=True)
rows.append(("full step, train mode, slow_frames_alias_fast (100 backbone frames)", timed_graph(st3.body)))
model.slow_frames_alias_fast = False
for k, v in rows:
    print(f"{v:8.3f} ms  {k}")
