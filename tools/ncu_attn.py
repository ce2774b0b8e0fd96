"""Driver for `ncu --set full -k regex:"mha_|xattn_bwd"`: the CUDA-core attention kernels at the bench shapes:
encoder self-attention (25 sequences x 141 tokens, 8 heads), temporal self-attention (1 x 100), one-query backward (100 x 141)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200 import kernels as K  # noqa: E402

d, H, scale = 256, 8, 1 / math.sqrt(32)


def r(*shape, dtype=torch.bfloat16):
    return torch.randn(*shape, device="cuda").to(dtype)


for B, L in ((25, 141), (1, 100)):
    qk, v = r(B * L, 512), r(B * L, d)
    o = torch.empty(B * L, d, dtype=torch.bfloat16, device="cuda")
    p, pbar = torch.empty(B, H, L, L, device="cuda"), torch.empty(B, L, L, device="cuda")
    kpm = torch.zeros(B, L, dtype=torch.uint8, device="cuda")
    do, dpbar = r(B * L, d), r(B, L, L, dtype=torch.float32)
    ds = torch.empty_like(p)
    dqk, dv = torch.empty_like(qk), torch.empty_like(v)
    for _ in range(3):
        K.mha_fwd(qk[:, :256], qk[:, 256:], v, kpm, o, p, pbar, B, H, L, L, scale)
        K.mha_bwd(qk[:, :256], qk[:, 256:], v, do, p, dpbar, ds, dqk[:, :256], dqk[:, 256:], dv, B, H, L, L, scale)
F, S = 100, 141
q, kp, vp, do = r(F, d), r(F * S, d), r(F * S, d), r(F, d)
p = torch.softmax(torch.randn(F, H, 1, S, device="cuda"), -1)
dq, dk, dv = torch.empty_like(q), torch.empty_like(kp), torch.empty_like(vp)
for _ in range(3):
    K.xattn_bwd(q, kp, vp, do, p, r(F, 1, S, dtype=torch.float32), dq, dk, dv, F, S, scale)
torch.cuda.synchronize()
print("done")
