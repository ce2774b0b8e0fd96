"""The reference's own training driver against the drop-in model, step for step (GPU).

`train_steps` below restates the body of reference engine.py:23-175 (train_one_epoch) and the set-up of main.py:363-414 in the same
order with the same calls: DistributedDataParallel wrap with find_unused_parameters=True (main.py:372-376), EMA copy by deepcopy
(main.py:370), AdamW over the three parameter groups selected by the "backbone" / "text_encoder" substrings (main.py:381-414), TWO
forward calls per backward (engine.py:67-80), captions as list[str], targets as list[dict], the keep-index on `pred_boxes` of the main
and the auxiliary outputs (engine.py:83-102), time_mask (engine.py:112-118), SetCriterion + weight_dict sum (engine.py:120-127),
optimizer.zero_grad / backward / clip_grad_norm_ / step (engine.py:147-152), update_ema over state_dict (util/optim.py:8-25).
Where the reference checkout is mounted (the build container) its `util.optim.update_ema` and `util.misc.NestedTensor` are imported and
used unmodified; on the GPU box their restatements are used.
"""
import argparse
import copy
import os
import sys

import pytest
import torch

from helpers import state_dict

pytestmark = pytest.mark.gpu
REF = "/root/reference"


def _args():
    return argparse.Namespace(
        num_queries=1, aux_loss=True, video_max_len_train=200, stride=2, guided_attn=True, fast=True, fast_mode="", sted=True,
        no_tsa=False, enc_layers=6, dec_layers=6, lr_backbone=1e-5, bbox_loss_coef=5, giou_loss_coef=2, sted_loss_coef=10,
        guided_attn_loss_coef=1, sigma=1, device="cuda", hidden_dim=256, nheads=8, dim_feedforward=2048, backbone="resnet101",
        dilation=False, position_embedding="sine", offline_text_encoder=True, lr=5e-5, text_encoder_lr=1e-5, weight_decay=1e-4,
        ema_decay=0.9998, clip_max_norm=0.1)


def _update_ema(model, model_ema, decay):
    """util/optim.py:8-25 restated (the reference function itself is used when the checkout is mounted)"""
    with torch.no_grad():
        if hasattr(model, "module"):
            model = model.module
        msd = model.state_dict()
        for k, ema_v in model_ema.state_dict().items():
            ema_v.copy_(ema_v * decay + (1.0 - decay) * msd[k].detach())


def _reference_pieces():
    if not os.path.isdir(os.path.join(REF, "util")):
        return None, None
    sys.path.insert(0, REF)
    try:
        import importlib.machinery
        import types
        for name in ("hostlist",):
            if name not in sys.modules:
                m = types.ModuleType(name)
                m.__spec__ = importlib.machinery.ModuleSpec(name, None)
                m.expand_hostlist = lambda s: [s]
                sys.modules[name] = m
        from util.misc import NestedTensor as RefNested
        from util.optim import update_ema as ref_update_ema
        return RefNested, ref_update_ema
    except Exception:
        return None, None
    finally:
        sys.path.remove(REF)


def _batches(nsteps, stride):
    from tubedetr_b200.synthetic import make_batch, pack_clips
    out = []
    for s in range(nsteps):
        durations = [6, 5] if s % 2 == 0 else [3, 4]      # per batch: equal clip counts ceil(T / k) (the one layout restriction of this model)
        b = make_batch(durations, [(96, 96), (64, 96)], stride, [5, 3], seed=40 + s)
        ff, fm = pack_clips(b["clips"])
        fs, ms = pack_clips([c[:, ::stride] for c in b["clips"]])
        T = max(durations)
        # one target dict per frame of every video; frames outside the annotated moment carry no box (engine.py:103-105 filters them)
        targets, kb = [], 0
        for i, (d, it) in enumerate(zip(durations, b["inter_idx"])):
            for t in range(T):
                if it[0] <= t <= it[1]:
                    targets.append({"boxes": b["target_boxes"][kb:kb + 1]})
                    kb += 1
                else:
                    targets.append({"boxes": torch.zeros(0, 4)})
        caps = ["a person walks towards the door", "the dog jumps"] if s % 2 == 0 else ["someone opens a window", "a child runs away fast"]
        out.append({"samples": (fs, ms), "samples_fast": (ff, fm), "durations": durations, "captions": caps, "targets": targets,
                    "inter_idx": b["inter_idx"]})
    return out


def train_steps(model, criterion, weight_dict, data, optimizer, device, args, max_norm, model_ema, Nested, update_ema):
    model.train()
    criterion.train()
    log = []
    for batch_dict in data:
        samples = Nested(*batch_dict["samples"]).to(device)
        samples_fast = Nested(*batch_dict["samples_fast"]).to(device)
        durations, captions = batch_dict["durations"], batch_dict["captions"]
        targets = [{k: v.to(device) for k, v in t.items()} for t in batch_dict["targets"]]
        memory_cache = model(samples, durations, captions, encode_and_save=True, samples_fast=samples_fast)
        outputs = model(samples, durations, captions, encode_and_save=False, memory_cache=memory_cache)
        max_duration = max(durations)
        inter_idx = batch_dict["inter_idx"]
        keep_list = []
        for i_dur, (duration, inter) in enumerate(zip(durations, inter_idx)):
            keep_list.extend(range(i_dur * max_duration + inter[0], i_dur * max_duration + inter[1] + 1))
        keep = torch.tensor(keep_list).long().to(device)
        outputs["pred_boxes"] = outputs["pred_boxes"][keep]
        for i_aux in range(len(outputs["aux_outputs"])):
            outputs["aux_outputs"][i_aux]["pred_boxes"] = outputs["aux_outputs"][i_aux]["pred_boxes"][keep]
        b = len(durations)
        targets = [x for x in targets if len(x["boxes"])]
        assert len(targets) == len(outputs["pred_boxes"])
        time_mask = torch.zeros(b, outputs["pred_sted"].shape[1]).bool().to(device)
        for i_dur, duration in enumerate(durations):
            time_mask[i_dur, :duration] = True
        loss_dict = criterion(outputs, targets, inter_idx, time_mask)
        losses = sum(loss_dict[k] * weight_dict[k] for k in loss_dict.keys() if k in weight_dict)
        loss_value = losses.item()
        assert torch.isfinite(losses), loss_dict
        optimizer.zero_grad()
        losses.backward()
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) if max_norm > 0 else None
        optimizer.step()
        update_ema(model, model_ema, args.ema_decay)
        log.append((loss_value, float(gn)))
    return log


def test_reference_training_loop_runs_unchanged_on_the_drop_in_model():
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel
    from tubedetr_b200 import NestedTensor, build_model
    args = _args()
    RefNested, ref_update_ema = _reference_pieces()
    own = not dist.is_initialized()
    if own:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(29700 + os.getpid() % 200))
        dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        device = torch.device("cuda")
        model, criterion, weight_dict = build_model(args)
        sd = state_dict()
        model.load_state_dict({k: sd[k] for k in model.state_dict()}, strict=True)
        model.to(device)
        model_ema = copy.deepcopy(model)                                   # main.py:370
        model_ddp = DistributedDataParallel(model, device_ids=[0], find_unused_parameters=True)   # main.py:372-376
        without_ddp = model_ddp.module
        param_dicts = [                                                    # main.py:381-405
            {"params": [p for n, p in without_ddp.named_parameters() if "backbone" not in n and "text_encoder" not in n and p.requires_grad]},
            {"params": [p for n, p in without_ddp.named_parameters() if "backbone" in n and p.requires_grad], "lr": args.lr_backbone},
            {"params": [p for n, p in without_ddp.named_parameters() if "text_encoder" in n and p.requires_grad], "lr": args.text_encoder_lr},
        ]
        assert all(len(g["params"]) > 0 for g in param_dicts)
        optimizer = torch.optim.AdamW(param_dicts, lr=args.lr, weight_decay=args.weight_decay)
        before = {n: p.detach().clone() for n, p in without_ddp.named_parameters()}
        ema_before = {k: v.detach().clone() for k, v in model_ema.state_dict().items()}
        log = train_steps(model_ddp, criterion, weight_dict, _batches(3, args.stride), optimizer, device, args, args.clip_max_norm, model_ema,
                          RefNested or NestedTensor, ref_update_ema or _update_ema)
        assert len(log) == 3 and all(l == l and g == g and g > 0 for l, g in log), log
        # every trainable parameter the loss reaches moved; the frozen stem / layer1 and all FrozenBN buffers did not
        moved = {n for n, p in without_ddp.named_parameters() if not torch.equal(p.detach(), before[n])}
        frozen = {n for n, p in without_ddp.named_parameters() if not p.requires_grad}
        assert not (moved & frozen)
        still = [n for n, p in without_ddp.named_parameters() if p.requires_grad and n not in moved and "pooler" not in n]
        assert not still, still[:5]
        # the EMA copy follows (decay 0.9998: small but non-zero change), incl. a deepcopy taken AFTER training forwards
        ema_sd = model_ema.state_dict()
        k_ = "transformer.encoder.layers.3.linear1.weight"
        assert not torch.equal(ema_sd[k_], ema_before[k_])
        assert (ema_sd[k_] - ema_before[k_]).abs().max() < (without_ddp.state_dict()[k_] - ema_before[k_]).abs().max()
        again = copy.deepcopy(without_ddp)
        assert again.state_dict().keys() == without_ddp.state_dict().keys()
        # eval after training: deterministic, finite
        model_ddp.eval()
        data = _batches(1, args.stride)[0]
        Nested = RefNested or NestedTensor
        with torch.no_grad():
            mc = model_ddp(Nested(*data["samples"]).to(device), data["durations"], data["captions"], encode_and_save=True,
                           samples_fast=Nested(*data["samples_fast"]).to(device))
            o1 = model_ddp(None, data["durations"], data["captions"], encode_and_save=False, memory_cache=mc)
            assert "text_memory_resized" in mc                               # engine.py:371 reads it during evaluation
            mc2 = model_ddp(Nested(*data["samples"]).to(device), data["durations"], data["captions"], encode_and_save=True,
                            samples_fast=Nested(*data["samples_fast"]).to(device))
            o2 = model_ddp(None, data["durations"], data["captions"], encode_and_save=False, memory_cache=mc2)
        assert torch.isfinite(o1["pred_boxes"]).all() and torch.equal(o1["pred_boxes"], o2["pred_boxes"])
    finally:
        if own:
            dist.destroy_process_group()
