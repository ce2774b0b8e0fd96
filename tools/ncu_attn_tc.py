"""Driver for `ncu --set full -k regex:mha_tc`: the tcgen05 self-attention kernels at the encoder shape of the bench (25 x 141, 8 heads,
train-mode hash dropout) and the temporal self-attention shape (1 x 100 with the head-mean gradient)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200 import kernels as K  # noqa: E402

d, H, scale = 256, 8, 1 / math.sqrt(32)
seed = torch.tensor([5], dtype=torch.int64, device="cuda")


def r(*shape, dtype=torch.bfloat16):
    return torch.randn(*shape, device="cuda").to(dtype)


for B, L, tsa in ((25, 141, False), (1, 100, True)):
    qk, v, do = r(B * L, 512), r(B * L, d), r(B * L, d)
    o = torch.empty(B * L, d, dtype=torch.bfloat16, device="cuda")
    p = torch.empty(B, H, L, L, device="cuda")
    pd = torch.empty_like(p) if tsa else None
    pbar = torch.empty(B, L, L, device="cuda") if tsa else None
    kpm = torch.zeros(B, L, dtype=torch.uint8, device="cuda")
    dpbar = r(B, L, L, dtype=torch.float32) if tsa else None
    dqk, dv = torch.empty_like(qk), torch.empty_like(v)
    for _ in range(3):
        K.mha_tc_fwd(qk[:, :256], qk[:, 256:], v, kpm, o, p, pbar, B, H, L, L, scale, drop=(seed, 3, 0.1), pdrop=pd)
        K.mha_tc_bwd(qk[:, :256], qk[:, 256:], v, do, p, dpbar, dqk[:, :256], dqk[:, 256:], dv, B, H, L, L, scale, drop=(seed, 3, 0.1))
torch.cuda.synchronize()
print("done")
