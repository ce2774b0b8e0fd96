"""Generate golden fixtures by running the UNMODIFIED reference (antoyang/TubeDETR) on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Writes tests/golden/<cfg>.pt and tests/golden/state_dict_manifest.json.

The reference is imported with the shims of SURVEY.md section 8(c): stub modules for timm /
hostlist / ffmpeg, resnet101(pretrained=False), RobertaModel built from a RobertaConfig,
and a fake tokenizer returning the fixture's own token ids.  Weights are the seeded
synthetic state_dict of tubedetr_b200/weights.py loaded with strict=True, so nothing
but this script and the reference decides the numbers.
"""
import argparse
import importlib.machinery
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402
import torchvision  # noqa: E402
import transformers  # noqa: E402,F401
from transformers import BatchEncoding, RobertaConfig, RobertaModel, RobertaTokenizerFast  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_stub("timm").models = _stub("timm.models", create_model=None)
_stub("hostlist", expand_hostlist=lambda s: [s])
_stub("ffmpeg")
_r101 = torchvision.models.resnet101
torchvision.models.resnet101 = lambda **kw: _r101(**{**kw, "pretrained": False})

_TOK = {}


class FakeTok:
    def batch_encode_plus(self, text, padding="longest", return_tensors="pt"):
        be = BatchEncoding({"input_ids": _TOK["input_ids"].clone(), "attention_mask": _TOK["attention_mask"].clone()})
        be._encodings = [None] * len(text)
        return be


RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: FakeTok())
RobertaModel.from_pretrained = classmethod(lambda cls, *a, **k: RobertaModel(RobertaConfig(
    vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1, bos_token_id=0,
    eos_token_id=2, layer_norm_eps=1e-5)))

from main import get_args_parser  # noqa: E402
from models import build_model  # noqa: E402
from util.misc import NestedTensor  # noqa: E402

from tubedetr_b200.weights import seeded_state_dict  # noqa: E402
from tubedetr_b200.synthetic import make_batch  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[0]: B=1 T=8 res=224 k=2, 10-token text
    "cfg1": dict(durations=[8], res=(224, 224), stride=2, ntok=[10], flags=[], seed=0),
    # ragged batch: two videos of different duration / frame size / caption length (pad masks everywhere)
    "cfg1b": dict(durations=[6, 5], res=[(160, 160), (128, 160)], stride=2, ntok=[6, 4], flags=[], seed=1),
    "nofast": dict(durations=[6], res=(160, 160), stride=2, ntok=[6], flags=["--no_fast"], seed=2),
    "notsa": dict(durations=[6], res=(160, 160), stride=3, ntok=[5], flags=["--no_tsa", "--no_guided_attn"], seed=3),
    # ---- full-size fixtures (BASELINE.json configs[1], [3], [4]); big tensors are stored as strided samples (`big`)
    # cfg-2: the measured configuration, B=1 T=100 k=4 res 352, 20 tokens
    "cfg2": dict(durations=[100], res=(352, 352), stride=4, ntok=[20], flags=[], seed=4, big=True),
    # cfg-4 shape at one clip: --no_fast, T=200 (the longest temporal self-attention), k=2, res 352
    "cfg4": dict(durations=[200], res=(352, 352), stride=2, ntok=[20], flags=["--no_fast"], seed=5, big=True),
    # cfg-5: --no_tsa --no_guided_attn, T=100, res 224, k=2
    "cfg5": dict(durations=[100], res=(224, 224), stride=2, ntok=[20], flags=["--no_tsa", "--no_guided_attn"], seed=6, big=True),
}
BIG_TOKEN_STEP, BIG_FRAME_STEP = 5, 7     # img_memory / pos_embed of the full-size fixtures keep [::5, ::7]


def sample_grad(g, n=64):
    f = g.flatten()
    step = max(f.numel() // n, 1)
    return f[::step][:n].clone()


def run(name, cfg, with_grad=True):
    batch = make_batch(cfg["durations"], cfg["res"], cfg["stride"], cfg["ntok"], seed=cfg["seed"])
    _TOK["input_ids"], _TOK["attention_mask"] = batch["input_ids"], batch["attention_mask"]
    args = argparse.ArgumentParser(parents=[get_args_parser()]).parse_args(
        ["--dataset_config", "x", "--combine_datasets", "vidstg", "--combine_datasets_val", "vidstg", "--device", "cpu",
         "--resolution", "224", "--stride", str(cfg["stride"])] + cfg["flags"])
    model, criterion, weight_dict = build_model(args)
    manifest = [(k, list(v.shape), str(v.dtype)) for k, v in model.state_dict().items()]
    sd = seeded_state_dict(manifest, seed=0)
    model.load_state_dict(sd, strict=True)
    model.eval()

    clips = batch["clips"]
    samples_fast = NestedTensor.from_tensor_list(clips)
    samples = NestedTensor.from_tensor_list([c[:, ::cfg["stride"]] for c in clips])
    durations = cfg["durations"]
    caps = ["x"] * len(durations)
    fast = "--no_fast" not in cfg["flags"]
    mc = model(samples, durations, caps, encode_and_save=True, samples_fast=samples_fast if fast else None)
    out = model(samples, durations, caps, encode_and_save=False, memory_cache=mc)

    gold = {"cfg": cfg, "pred_boxes": out["pred_boxes"], "pred_sted": out["pred_sted"],
            "aux_pred_boxes": torch.stack([a["pred_boxes"] for a in out["aux_outputs"]]),
            "aux_pred_sted": torch.stack([a["pred_sted"] for a in out["aux_outputs"]]),
            "img_memory": mc["img_memory"], "pos_embed": mc["pos_embed"], "mask": mc["mask"],
            "query_embed": mc["query_embed"], "query_mask": mc["query_mask"],
            "text_memory_resized": mc["text_memory_resized"], "text_attention_mask": mc["text_attention_mask"]}
    if "weights" in out:
        gold["weights"] = out["weights"]
        gold["ca_weights"] = out["ca_weights"]
        gold["aux_weights"] = torch.stack([a["weights"] for a in out["aux_outputs"]])
    with torch.no_grad():
        feat = model.backbone(NestedTensor(samples.tensors[:1], samples.mask[:1]))[0][-1].tensors
    gold["feat_slow0"] = feat  # layer4 features of the first slow frame

    # keep-index + criterion exactly as reference engine.py:83-122
    T = max(durations)
    inter_idx = batch["inter_idx"]
    keep = torch.tensor([e for i, it in enumerate(inter_idx) for e in range(i * T + it[0], i * T + it[1] + 1)]).long()
    outc = dict(out)
    outc["pred_boxes"] = out["pred_boxes"][keep]
    outc["aux_outputs"] = [dict(a, pred_boxes=a["pred_boxes"][keep]) for a in out["aux_outputs"]]
    targets = [{"boxes": b[None]} for b in batch["target_boxes"]]
    time_mask = batch["time_mask"]
    loss_dict = criterion(outc, targets, inter_idx, time_mask)
    gold["losses"] = {k: v.detach() for k, v in loss_dict.items()}
    total = sum(loss_dict[k] * weight_dict[k] for k in loss_dict if k in weight_dict)
    gold["loss_total"] = total.detach()
    gold["weight_dict"] = weight_dict
    if with_grad:
        model.zero_grad()
        total.backward()
        gn, gs = {}, {}
        for k, p in model.named_parameters():
            if p.grad is not None:
                gn[k] = p.grad.norm().item()
                gs[k] = sample_grad(p.grad)
        gold["grad_norm"], gold["grad_sample"] = gn, gs
        gold["requires_grad"] = {k: p.requires_grad for k, p in model.named_parameters()}
    if cfg.get("big"):
        for k in ("img_memory", "pos_embed"):
            gold[k + "_sample"] = gold.pop(k)[::BIG_TOKEN_STEP, ::BIG_FRAME_STEP]
        gold["sample_steps"] = (BIG_TOKEN_STEP, BIG_FRAME_STEP)
    gold = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in gold.items()}
    torch.save(gold, os.path.join(HERE, name + ".pt"))
    print(name, "pred_boxes", out["pred_boxes"][0].tolist(), "loss", float(total),
          "feat absmax %.3f mean %.3f" % (feat.abs().max(), feat.abs().mean()))
    return manifest


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or list(CONFIGS)
    man = None
    for nm in which:
        man = run(nm, CONFIGS[nm])
    if "cfg1" in which:
        json.dump(man, open(os.path.join(HERE, "state_dict_manifest.json"), "w"))
