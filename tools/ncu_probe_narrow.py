"""Driver for `ncu --set full`: the narrow 3x3 convolutions of layer1 (C=64, 88x88) and layer2 (C=128, 44x44), 100 frames."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200.gemm import REMAP_P2C, gemm  # noqa: E402

for (Nimg, H, W, C) in [(100, 88, 88, 64), (100, 44, 44, 128)]:
    xp = torch.randn(Nimg * (H + 2) * (W + 2), C, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(C, 9 * C, device="cuda") * 0.02).to(torch.bfloat16)
    o = torch.empty(Nimg * H * W, C, dtype=torch.bfloat16, device="cuda")
    sc, sh = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    taps = [(kh - 1) * (W + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    for _ in range(2):
        gemm(xp, wk, o, xp.shape[0], C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], scale=sc, bias=sh, relu=True,
             remap=REMAP_P2C, img_hw=(H, W))
torch.cuda.synchronize()
print("done")
