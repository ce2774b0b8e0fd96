"""Golden vectors for the optimizer-side step, produced by the REFERENCE's own call sequence (engine.py:147-161):
torch.nn.utils.clip_grad_norm_ -> torch.optim.AdamW(param_dicts of main.py:381-414).step() -> util.optim.update_ema
(imported unmodified from /root/reference).  Run in the build container:  python tests/golden/make_optim_golden.py
Small tensors whose names carry the LR-group substrings; 4 steps with a new seeded gradient each step."""
import copy
import os
import sys

import torch

sys.path.insert(0, "/root/reference")
from util.optim import update_ema  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = [("transformer.encoder.layers.0.linear1.weight", (24, 16)), ("transformer.encoder.layers.0.linear1.bias", (24,)),
          ("bbox_embed.layers.2.bias", (4,)), ("sted_embed.layers.1.bias", (2,)), ("query_embed.weight", (1, 16)),
          ("backbone.0.body.layer2.0.conv2.weight", (8, 8, 3, 3)), ("backbone.0.body.layer3.0.conv1.weight", (16, 8, 1, 1)),
          ("transformer.text_encoder.embeddings.word_embeddings.weight", (37, 12)),
          ("transformer.text_encoder.encoder.layer.0.output.dense.bias", (13,))]
LR, LR_BACKBONE, LR_TEXT, WD, MAX_NORM, DECAY, STEPS = 5e-3, 1e-3, 2e-3, 1e-2, 0.1, 0.98, 4


class Holder(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(11)
        self.names = [n for n, _ in SHAPES]
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, generator=g)) for _, s in SHAPES])

    def named(self):
        return list(zip(self.names, self.ps))


def main():
    model = Holder()
    ema = copy.deepcopy(model)
    named = model.named()
    groups = [{"params": [p for n, p in named if "backbone" not in n and "text_encoder" not in n]},
              {"params": [p for n, p in named if "backbone" in n], "lr": LR_BACKBONE},
              {"params": [p for n, p in named if "text_encoder" in n], "lr": LR_TEXT}]
    opt = torch.optim.AdamW(groups, lr=LR, weight_decay=WD)
    out = {"shapes": SHAPES, "hyper": dict(lr=LR, lr_backbone=LR_BACKBONE, text_encoder_lr=LR_TEXT, weight_decay=WD,
                                           max_norm=MAX_NORM, ema_decay=DECAY),
           "init": [p.detach().clone() for p in model.ps], "grads": [], "params": [], "ema": [], "norm": []}
    g = torch.Generator().manual_seed(12)
    for step in range(STEPS):
        scale = 10.0 if step % 2 == 0 else 1e-3            # one clipped step, one unclipped step, ...
        grads = [torch.randn(s, generator=g) * scale for _, s in SHAPES]
        opt.zero_grad()
        for p, gr in zip(model.ps, grads):
            p.grad = gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(model.parameters(), MAX_NORM)
        opt.step()
        update_ema(model, ema, DECAY)
        out["grads"].append(grads)
        out["norm"].append(norm.clone())
        out["params"].append([p.detach().clone() for p in model.ps])
        out["ema"].append([p.detach().clone() for p in ema.ps])
    torch.save(out, os.path.join(HERE, "optim.pt"))
    print("wrote optim.pt", [float(n) for n in out["norm"]])


if __name__ == "__main__":
    main()
