"""Narrow (N = 64 / 128) backbone convolutions of layer1 / layer2 at cfg-2 (125 frames, res 352) on the 1-CTA kernel: row-per-thread
epilogue (mode 3, the default) against the per-warp transposed, row-coalesced epilogue (mode 1).  CUDA-graph replays over rotating
buffers.  Output: gpurun_out/narrow_bench.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200.gemm import REMAP_C2P, REMAP_NONE, REMAP_P2C, gemm  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/narrow_bench.txt", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


def timed(fn, reps):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(reps):
            fn(i)
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


def case(name, F, h, w, cin, cout, taps, remap):
    g = torch.Generator(device="cuda").manual_seed(0)
    Rc, Rp = F * h * w, F * (h + 2) * (w + 2)
    M = Rp if taps == 9 else Rc
    orows = Rp if remap == REMAP_C2P else Rc
    nb = 2
    sets = [((torch.randn(M, cin, device="cuda", generator=g) * 0.5).to(torch.bfloat16), torch.zeros(orows, cout, dtype=torch.bfloat16, device="cuda"))
            for _ in range(nb)]
    B = (torch.randn(cout, taps * cin, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    sc, bi = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
    kw = dict(scale=sc, bias=bi, relu=True)
    if remap != REMAP_NONE:
        kw.update(remap=remap, img_hw=(h, w))
    if taps == 9:
        kw.update(ntaps=9, a_off1=[(kh - 1) * (w + 2) + (kw_ - 1) for kh in range(3) for kw_ in range(3)], b_off0=[t * cin for t in range(9)])
    alg = (M * cin + Rc * cout + B.numel()) * 2
    line = f"{name:26s} M={M:8d} N={cout:4d} K={cin:4d}x{taps} alg {alg / 1e6:7.1f} MB |"
    outs = []
    for mode in (3, 1):
        def run(i, mode=mode):
            A, o = sets[i % nb]
            gemm(A, B, o, M, cout, cin, debug_flags=(mode + 1) << 1, **kw)
        us = timed(run, 8)
        outs.append(sets[0][1].clone())
        line += f" mode{mode} {us:7.1f} us {alg / us / 1e3:6.0f} GB/s |"
    P(line, "equal" if torch.equal(outs[0], outs[1]) else f"max diff {(outs[0].float() - outs[1].float()).abs().max().item():.3g}")


if __name__ == "__main__":
    P(torch.cuda.get_device_name(0))
    F = 125
    case("layer1 conv1 -> padded", F, 88, 88, 256, 64, 1, REMAP_C2P)
    case("layer1.0 conv1 -> padded", F, 88, 88, 64, 64, 1, REMAP_C2P)
    case("layer1 conv2 3x3", F, 88, 88, 64, 64, 9, REMAP_P2C)
    case("layer2 conv1 -> padded", F, 44, 44, 512, 128, 1, REMAP_C2P)
    case("layer2 conv2 3x3", F, 44, 44, 128, 128, 9, REMAP_P2C)
    case("layer2.0 conv1 (88x88)", F, 88, 88, 256, 128, 1, REMAP_NONE)
    case("layer3 conv2 3x3 (1-CTA n/a)", 25, 22, 22, 256, 256, 9, REMAP_P2C)
