"""Shared helpers for the parity tests: fixtures, seeded weights, oracle invocation."""
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def manifest():
    return [tuple(m) for m in json.load(open(os.path.join(GOLD, "state_dict_manifest.json")))]


_SD = {}


def state_dict(seed=0):
    from tubedetr_b200.weights import seeded_state_dict
    if seed not in _SD:
        _SD[seed] = seeded_state_dict(manifest(), seed)
    return _SD[seed]


def load_gold(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def batch_for(cfg):
    from tubedetr_b200.synthetic import make_batch, pack_clips
    b = make_batch(cfg["durations"], cfg["res"], cfg["stride"], cfg["ntok"], seed=cfg["seed"])
    b["frames_fast"], b["mask_fast"] = pack_clips(b["clips"])
    b["frames_slow"], b["mask_slow"] = pack_clips([c[:, ::cfg["stride"]] for c in b["clips"]])
    T = max(cfg["durations"])
    b["keep"] = torch.tensor([e for i, it in enumerate(b["inter_idx"]) for e in range(i * T + it[0], i * T + it[1] + 1)])
    return b


def run_oracle(cfg, sd=None, grad=False):
    from oracle import tubedetr_oracle as O
    sd = sd or state_dict()
    b = batch_for(cfg)
    fast = "--no_fast" not in cfg["flags"]
    no_tsa = "--no_tsa" in cfg["flags"]
    out, cache = O.forward(sd, b["frames_slow"], b["mask_slow"], b["frames_fast"], b["mask_fast"], cfg["durations"],
                           b["input_ids"], b["attention_mask"], cfg["stride"], fast=fast, no_tsa=no_tsa)
    return out, cache, b
