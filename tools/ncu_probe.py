"""Tiny driver for `ncu --set full`: the two dominant GEMM flavours of the TubeDETR step on the 125-frame (25 slow + 100 fast) batch.
  A: layer3 3x3 conv as implicit GEMM (M=72000 haloed rows, N=256, K=9x256), FrozenBN+ReLU epilogue, halo rows dropped
  B: layer3 conv3 1x1 (M=60500, N=1024, K=256) with FrozenBN + residual + ReLU epilogue
Usage: ncu --set full --clock-control none --import-source on -k regex:tdb_gemm -o gpurun_out/prof_gemm python tools/ncu_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200.gemm import REMAP_P2C, gemm  # noqa: E402

mode = int(os.environ.get("TDB_EPI_MODE", "0"))
N_, h, w, C = 125, 22, 22, 256
Rp = N_ * (h + 2) * (w + 2)
x = torch.randn(Rp, C, device="cuda").to(torch.bfloat16)
wk = (torch.randn(C, 9 * C, device="cuda") * 0.02).to(torch.bfloat16)
y = torch.empty(N_ * h * w, C, dtype=torch.bfloat16, device="cuda")
sc, sh = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
taps = [(kh - 1) * (w + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
for _ in range(2):
    gemm(x, wk, y, Rp, C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], scale=sc, bias=sh, relu=True,
         remap=REMAP_P2C, img_hw=(h, w))
M, N, K = 60500, 1024, 256
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
B = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
R = torch.randn(M, N, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
s2, b2 = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
for _ in range(2):
    gemm(A, B, out, M, N, K, scale=s2, bias=b2, residual=R, relu=True)
torch.cuda.synchronize()
print("probe done")
