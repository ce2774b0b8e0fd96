"""Driver for `ncu --set full` (round 2): the kernels that are new or dominant this round, each alone at the bench shapes (cfg-2).
  A  tdb_gemm2_kernel      layer3 3x3 conv, 125 frames (the bench's roofline kernel)
  B  tdb_gemm_kernel<128,5> layer3 conv3 1x1 + FrozenBN + residual + ReLU, 125 frames
  C  stem_fused_kernel     conv 7x7/2 + FrozenBN + ReLU + maxpool, 125 frames of 352 x 352
  D  K / V projection of all 6 decoder layers (two-tap split-precision GEMM [14100, 256] x [1536, 256]^T)
  E  xattn_core_fwd_kernel / xattn_bwd_kernel on a 256-column slice (100 frames x 141 tokens)
Usage: ncu --set full --clock-control none --import-source on -k regex:"tdb_gemm|stem_fused|xattn_core|xattn_bwd" -o gpurun_out/prof_r02 python tools/ncu_probe_r02.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200 import kernels as K  # noqa: E402
from tubedetr_b200.gemm import REMAP_P2C, gemm  # noqa: E402

dev = "cuda"
N_, h, w, C = 125, 22, 22, 256
Rp = N_ * (h + 2) * (w + 2)
x = torch.randn(Rp, C, device=dev).bfloat16()
wk = (torch.randn(C, 9 * C, device=dev) * 0.02).bfloat16()
y = torch.empty(N_ * h * w, C, dtype=torch.bfloat16, device=dev)
sc, sh = torch.ones(C, device=dev), torch.zeros(C, device=dev)
taps = [(kh - 1) * (w + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
for _ in range(2):      # A
    gemm(x, wk, y, Rp, C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], scale=sc, bias=sh, relu=True, remap=REMAP_P2C, img_hw=(h, w))
M, N, Kd = 60500, 1024, 256
A = torch.randn(M, Kd, device=dev).bfloat16()
B = (torch.randn(N, Kd, device=dev) * 0.05).bfloat16()
R = torch.randn(M, N, device=dev).bfloat16()
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
s2, b2 = torch.ones(N, device=dev), torch.zeros(N, device=dev)
for _ in range(2):      # B
    gemm(A, B, out, M, N, Kd, scale=s2, bias=b2, residual=R, relu=True)
fr = torch.randn(125, 3, 352, 352, device=dev)
wst = torch.zeros(64, 192, dtype=torch.bfloat16, device=dev)
wst[:, :168] = (torch.randn(64, 168, device=dev) * 0.1).bfloat16()
pooled = torch.empty(125 * 88 * 88, 64, dtype=torch.bfloat16, device=dev)
s1, b1 = torch.ones(64, device=dev), torch.zeros(64, device=dev)
for _ in range(2):      # C
    K.stem_fused(fr, wst, s1, b1, pooled, 125, 352, 352)
F_, S, nl, d = 100, 141, 6, 256
mem = torch.randn(F_ * S, d, device=dev).bfloat16()
Wk = (torch.randn(nl * d, 2 * d, device=dev) / 16).bfloat16()
Kall = torch.empty(F_ * S, nl * d, dtype=torch.bfloat16, device=dev)
Vall = torch.randn(F_ * S, nl * d, device=dev).bfloat16()
bias = torch.zeros(nl * d, device=dev)
for _ in range(2):      # D
    gemm(mem, Wk, Kall, F_ * S, nl * d, d, ntaps=2, a_off0=(0, 0), b_off0=(0, d), bias=bias)
q = torch.randn(F_, d, device=dev).bfloat16()
kpm = torch.zeros(F_, S, dtype=torch.uint8, device=dev)
o = torch.empty(F_, d, dtype=torch.bfloat16, device=dev)
p = torch.empty(F_, 8, 1, S, device=dev)
pbar = torch.empty(F_, 1, S, device=dev)
dq = torch.empty_like(q)
dK, dV = torch.empty_like(Kall), torch.empty_like(Vall)
sl = slice(2 * d, 3 * d)
for _ in range(2):      # E
    K.xattn_core_fwd(q, Kall[:, sl], Vall[:, sl], kpm, o, p, pbar, F_, S, 1 / math.sqrt(32))
    K.xattn_core_bwd(q, Kall[:, sl], Vall[:, sl], o, p, pbar, dq, dK[:, sl], dV[:, sl], F_, S, 1 / math.sqrt(32))
torch.cuda.synchronize()
print("probe done")
