"""Bring-up of the halo mode of the 2-CTA implicit 3x3 convolution: error vs torch for halo off / on / on without the
descriptor base offset, fprop and dgrad forms, three image widths (1 and 2 TMA boxes per A tile), then timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200.gemm import REMAP_P2C, gemm  # noqa: E402

TWO = 1 << 6
NOHALO = 1 << 7
NOBASE = 1 << 8


def pad_rows(x):
    N, H, W, C = x.shape
    xp = torch.zeros(N, H + 2, W + 2, C, dtype=x.dtype, device=x.device)
    xp[:, 1:-1, 1:-1] = x
    return xp.view(-1, C)


def err(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6)).item()


torch.manual_seed(0)
for (Nimg, H, W, C) in [(4, 11, 13, 256), (3, 22, 22, 256), (2, 44, 44, 256), (1, 88, 88, 256)]:
    x = torch.randn(Nimg, H, W, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, device="cuda") * 0.05).to(torch.bfloat16)
    g = torch.randn(Nimg, H, W, C, device="cuda").to(torch.bfloat16)
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    y = torch.nn.functional.conv2d(xf, w.float(), padding=1)
    y.backward(g.float().permute(0, 3, 1, 2))
    ref = y.detach().permute(0, 2, 3, 1)
    refdx = xf.grad.permute(0, 2, 3, 1)
    xp, gp = pad_rows(x), pad_rows(g)
    wk = w.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
    taps = [(kh - 1) * (W + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    for name, fl in [("halo off", TWO | NOHALO), ("halo on", TWO), ("halo on + desc base offset", TWO | NOBASE)]:
        o = torch.full((Nimg * H * W, C), float("nan"), dtype=torch.bfloat16, device="cuda")
        gemm(xp, wk, o, xp.shape[0], C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], remap=REMAP_P2C,
             img_hw=(H, W), debug_flags=fl)
        dx = torch.full((Nimg * H * W, C), float("nan"), dtype=torch.bfloat16, device="cuda")
        gemm(gp, wk, dx, gp.shape[0], C, C, b_major=1, ntaps=9, a_off1=[-t for t in taps], b_off0=[t * C for t in range(9)],
             remap=REMAP_P2C, img_hw=(H, W), debug_flags=fl)
        torch.cuda.synchronize()
        print(f"W={W:3d} {name:26s} fprop err {err(o.view(Nimg, H, W, C), ref):.3e}  dgrad err {err(dx.view(Nimg, H, W, C), refdx):.3e}",
              flush=True)

# timing on the layer3 shape (100 frames) and layer4 (C=512, 11x11)
for (Nimg, H, W, C) in [(100, 22, 22, 256), (25, 22, 22, 256), (100, 11, 11, 512)]:
    xp = torch.randn(Nimg * (H + 2) * (W + 2), C, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(C, 9 * C, device="cuda") * 0.02).to(torch.bfloat16)
    o = torch.empty(Nimg * H * W, C, dtype=torch.bfloat16, device="cuda")
    sc, sh = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    taps = [(kh - 1) * (W + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, fl in [("halo off", NOHALO), ("halo on", 0)]:
        def run():
            gemm(xp, wk, o, xp.shape[0], C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], scale=sc, bias=sh,
                 relu=True, remap=REMAP_P2C, img_hw=(H, W), debug_flags=fl)
        for _ in range(3):
            run()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        fl_ = 2.0 * Nimg * H * W * C * 9 * C
        print(f"N={Nimg} {H}x{W} C={C} {name:9s} median {ts[5]:.1f} us  min {ts[0]:.1f} us  {fl_ / ts[5] / 1e6:.0f} TF (algorithmic)",
              flush=True)

# ---- 1-CTA kernel (narrow convs of layer1 / layer2): ring and resident-B variants
print("1-CTA kernel", flush=True)
for (Nimg, H, W, C) in [(3, 11, 13, 64), (2, 44, 44, 128), (1, 88, 88, 64), (2, 88, 88, 64)]:
    x = torch.randn(Nimg, H, W, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, device="cuda") * 0.05).to(torch.bfloat16)
    g = torch.randn(Nimg, H, W, C, device="cuda").to(torch.bfloat16)
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    y = torch.nn.functional.conv2d(xf, w.float(), padding=1)
    y.backward(g.float().permute(0, 3, 1, 2))
    ref = y.detach().permute(0, 2, 3, 1)
    refdx = xf.grad.permute(0, 2, 3, 1)
    xp, gp = pad_rows(x), pad_rows(g)
    wk = w.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
    taps = [(kh - 1) * (W + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    for name, fl in [("halo off", NOHALO), ("halo on", 0)]:
        o = torch.full((Nimg * H * W, C), float("nan"), dtype=torch.bfloat16, device="cuda")
        gemm(xp, wk, o, xp.shape[0], C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], remap=REMAP_P2C,
             img_hw=(H, W), debug_flags=fl)
        dx = torch.full((Nimg * H * W, C), float("nan"), dtype=torch.bfloat16, device="cuda")
        gemm(gp, wk, dx, gp.shape[0], C, C, b_major=1, ntaps=9, a_off1=[-t for t in taps], b_off0=[t * C for t in range(9)],
             remap=REMAP_P2C, img_hw=(H, W), debug_flags=fl)
        torch.cuda.synchronize()
        print(f"C={C} W={W:3d} {name:10s} fprop err {err(o.view(Nimg, H, W, C), ref):.3e}  dgrad err {err(dx.view(Nimg, H, W, C), refdx):.3e}",
              flush=True)
for (Nimg, H, W, C) in [(100, 88, 88, 64), (25, 88, 88, 64), (100, 44, 44, 128), (25, 44, 44, 128)]:
    xp = torch.randn(Nimg * (H + 2) * (W + 2), C, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(C, 9 * C, device="cuda") * 0.02).to(torch.bfloat16)
    o = torch.empty(Nimg * H * W, C, dtype=torch.bfloat16, device="cuda")
    sc, sh = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    taps = [(kh - 1) * (W + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, fl in [("halo off", NOHALO), ("halo on", 0)]:
        def run():
            gemm(xp, wk, o, xp.shape[0], C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], scale=sc, bias=sh,
                 relu=True, remap=REMAP_P2C, img_hw=(H, W), debug_flags=fl)
        for _ in range(3):
            run()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        fl_ = 2.0 * Nimg * H * W * C * 9 * C
        byts = 2.0 * (xp.numel() + o.numel())
        print(f"N={Nimg} {H}x{W} C={C} {name:9s} median {ts[5]:.1f} us  {fl_ / ts[5] / 1e6:.0f} TF  {byts / ts[5] / 1e3:.0f} GB/s (in+out)",
              flush=True)
