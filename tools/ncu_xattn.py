"""Driver for `ncu --set full -k regex:xattn_fused`: the fused KV-projection + cross-attention kernel at the bench shape
(F = 100 frames, S = 141 tokens)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200 import kernels as K  # noqa: E402

F, S, d = 100, 141, 256
q = torch.randn(F, d, device="cuda").to(torch.bfloat16)
mem = torch.randn(F * S, d, device="cuda").to(torch.bfloat16)
mempb = (mem.float() + 0.5 * torch.randn(F * S, d, device="cuda")).to(torch.bfloat16)
W = (torch.randn(3 * d, d, device="cuda") / 16).to(torch.bfloat16)
b = torch.randn(3 * d, device="cuda") * 0.1
kpm = torch.zeros(F, S, dtype=torch.uint8, device="cuda")
o = torch.empty(F, d, dtype=torch.bfloat16, device="cuda")
p = torch.empty(F, 8, 1, S, device="cuda")
pbar = torch.empty(F, 1, S, device="cuda")
for _ in range(3):
    K.xattn_fused_fwd(q, mempb, mem, W[d:], b[2 * d:], kpm, o, p, pbar, F, S, 1 / math.sqrt(32))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    K.xattn_fused_fwd(q, mempb, mem, W[d:], b[2 * d:], kpm, o, p, pbar, F, S, 1 / math.sqrt(32))
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
fl = 2.0 * F * S * d * 2 * d
print(f"xattn fused (fused kernel + merge): {us:.1f} us per layer, KV-projection {fl / us / 1e6:.1f} TFLOP/s")
