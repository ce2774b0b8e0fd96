"""Self-attention core alone: tcgen05 kernels (tdb_attn_tc.cu) vs the CUDA-core kernels (tdb_attn.cu) on the step's shapes.
  python tools/attn_bench.py > gpurun_out/attn_bench.txt
20 calls captured in one CUDA graph (rotating over 4 buffer sets), CUDA events around 5 replays: device time without host enqueue cost."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tubedetr_b200 import kernels as K  # noqa: E402

H, d = 8, 256
scale = 1 / math.sqrt(32)


def timeit(fn, reps=20):
    """device time per call: `reps` calls captured in ONE CUDA graph (no host enqueue cost between kernels), replayed 5 times"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(4):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


for name, B, L, need_w in (("encoder 25 x 141", 25, 141, False), ("encoder 50 x 69", 50, 69, False), ("TSA 1 x 100", 1, 100, True),
                           ("TSA 2 x 200", 2, 200, True)):
    for drop in (0.0, 0.1):
        nb = 4
        sets = []
        for i in range(nb):
            qk = torch.randn(B * L, 512, device="cuda").to(torch.bfloat16)
            v = torch.randn(B * L, d, device="cuda").to(torch.bfloat16)
            do = torch.randn(B * L, d, device="cuda").to(torch.bfloat16)
            o = torch.empty(B * L, d, dtype=torch.bfloat16, device="cuda")
            p = torch.empty(B, H, L, L, device="cuda")
            pd = torch.empty(B, H, L, L, device="cuda") if (drop and need_w) else None
            pbar = torch.empty(B, L, L, device="cuda") if need_w else None
            dpbar = torch.randn(B, L, L, device="cuda") if need_w else None
            dqk, dv = torch.empty_like(qk), torch.empty_like(v)
            ds, pds = torch.empty_like(p), torch.empty_like(p)
            sets.append((qk, v, do, o, p, pd, pbar, dpbar, dqk, dv, ds, pds))
        kpm = torch.zeros(B, L, dtype=torch.uint8, device="cuda")
        seed = torch.tensor([5], dtype=torch.int64, device="cuda")
        keep = K.dropout_mask(torch.empty(B, H, L, L, dtype=torch.uint8, device="cuda"), seed, 3, drop) if drop else None
        dropt = (seed, 3, drop) if drop else None
        ks = 1 / (1 - drop) if drop else 1.0

        def f_tc(i):
            qk, v, do, o, p, pd, pbar, dpbar, dqk, dv, ds, pds = sets[i % nb]
            K.mha_tc_fwd(qk[:, :256], qk[:, 256:], v, kpm, o, p, pbar, B, H, L, L, scale, drop=dropt, pdrop=pd)

        def f_cc(i):
            qk, v, do, o, p, pd, pbar, dpbar, dqk, dv, ds, pds = sets[i % nb]
            if drop:
                K.dropout_mask(keep, seed, 3, drop)        # the mask kernel belongs to the CUDA-core path's cost
            K.mha_fwd_cuda_core(qk[:, :256], qk[:, 256:], v, kpm, o, p, pbar, B, H, L, L, scale, keep=keep, pdrop=pd, keep_scale=ks)

        def b_tc(i):
            qk, v, do, o, p, pd, pbar, dpbar, dqk, dv, ds, pds = sets[i % nb]
            K.mha_tc_bwd(qk[:, :256], qk[:, 256:], v, do, p, dpbar, dqk[:, :256], dqk[:, 256:], dv, B, H, L, L, scale, drop=dropt)

        def b_cc(i):
            qk, v, do, o, p, pd, pbar, dpbar, dqk, dv, ds, pds = sets[i % nb]
            K.mha_bwd_cuda_core(qk[:, :256], qk[:, 256:], v, do, p, dpbar, ds, dqk[:, :256], dqk[:, 256:], dv, B, H, L, L, scale,
                                keep=keep, keep_scale=ks, pd_scratch=pds if drop else None)

        f_cc(0)
        torch.cuda.synchronize()
        t = [timeit(f) for f in (f_tc, f_cc, b_tc, b_cc)]
        pbytes = B * H * L * L * 4
        print(f"{name:18s} drop {drop:.1f}: fwd tcgen05 {t[0]:7.1f} us | cuda-core {t[1]:7.1f} us || bwd tcgen05 {t[2]:7.1f} us | cuda-core {t[3]:7.1f} us"
              f"   (P = {pbytes / 1e6:.1f} MB: fwd tc {pbytes / t[0] / 1e3:.0f} GB/s of P writes)")
