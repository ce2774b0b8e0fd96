"""Decoder cross-attention kernel (xattn_fused_kernel): tensor-pipe utilisation OVER ITS MMA PHASE from
in-kernel SM clock stamps, plus the streaming view (ii): algorithmic bytes / kernel time against the measured HBM peak.
  python tools/xattn_phase.py > gpurun_out/xattn_phase.txt"""
import ctypes as C
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tubedetr_b200 import kernels as K  # noqa: E402
from tubedetr_b200._lib import lib  # noqa: E402

MMA_CYCLES = 2 * 16 * 128      # K and V: 16 tcgen05.mma (M128 N256 K16) each, 128 tensor-pipe cycles per instruction
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6555.2
for pair, F, S in ((0, 100, 141), (1, 100, 141), (0, 400, 141), (1, 400, 141), (0, 800, 141), (1, 800, 141)):
    lib().tdb_xattn_set_pair(pair)
    d = 256
    g = torch.Generator().manual_seed(1)
    q = torch.randn(F, d, generator=g).bfloat16().cuda()
    nb = 8                       # rotate memory buffers so the timed launches read HBM, not L2 (8 x 2 x 7.2 MB x F/100)
    mems = [(torch.randn(F * S, d, generator=g).bfloat16().cuda(), torch.randn(F * S, d, generator=g).bfloat16().cuda()) for _ in range(nb)]
    W = (torch.randn(2 * d, d, generator=g) / 16).bfloat16().cuda()
    bv = torch.zeros(d).cuda()
    kpm = torch.zeros(F, S, dtype=torch.uint8).cuda()
    o = torch.empty(F, d, dtype=torch.bfloat16).cuda()
    p = torch.empty(F, 8, 1, S).cuda()
    pbar = torch.empty(F, 1, S).cuda()
    tiles = (F * S + 127) // 128
    stamps = torch.zeros(tiles, 4, dtype=torch.int64).cuda()

    def run(i):
        K.xattn_fused_fwd(q, mems[i % nb][0], mems[i % nb][1], W, bv, kpm, o, p, pbar, F, S, 1 / math.sqrt(32))
    for i in range(4):
        run(i)
    torch.cuda.synchronize()
    reps = 24
    gr = torch.cuda.CUDAGraph()          # replay a graph: the eager ctypes launch path (~10 us per call) would hide the kernels
    with torch.cuda.graph(gr):
        for i in range(reps):
            run(i)
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * reps) * 1e3        # fused + merge kernels, back to back, rotating HBM-resident inputs
    lib().tdb_xattn_set_timing_buffer(C.c_void_p(stamps.data_ptr()))
    run(5)
    torch.cuda.synchronize()
    lib().tdb_xattn_set_timing_buffer(None)
    st = stamps.cpu().double()
    mma = (st[:, 2] - st[:, 1])
    life = (st[:, 3] - st[:, 0])
    pre = (st[:, 1] - st[:, 0])
    alg = 2 * F * S * d * 2 + 2 * d * d * 2
    print(f"{'pair (cta_group::2)' if pair else '1-CTA'} kernel, F={F} S={S}: {tiles} tiles | fused+merge {us:.1f} us per layer ({alg / us / 1e3:.0f} GB/s of algorithmic bytes = "
          f"{alg / us / 1e3 / pk:.2f} of the measured HBM copy peak)")
    print(f"   MMA phase (first issue -> last MMA complete): median {mma.median():.0f} cycles, min {mma.min():.0f}, max {mma.max():.0f}"
          f"  => tensor pipe over the MMA phase = {100 * MMA_CYCLES / mma.median():.1f} % (ideal {MMA_CYCLES} cycles)")
    print(f"   CTA lifetime median {life.median():.0f} cycles (entry -> first MMA {pre.median():.0f});  MMA share of the lifetime "
          f"{100 * MMA_CYCLES / life.median():.1f} %")
lib().tdb_xattn_set_pair(0)
