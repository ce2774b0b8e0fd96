"""What a foreign kernel that holds some SMs does to a GEMM: static persistent schedule (tile += grid) against cluster launch control
(one cluster per tile, work stealing).  A spin kernel with a 200 KB shared-memory footprint (nothing can share its SMs) occupies
`hog` SMs on one stream while the layer3 3x3 convolution (125 frames) runs 6 times on another.  Output: gpurun_out/clc_hog.txt"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200 import _lib  # noqa: E402
from tubedetr_b200.gemm import REMAP_NONE, REMAP_P2C, gemm  # noqa: E402

DYNAMIC = 2048
os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/clc_hog.txt", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


def run_case(name, fn, reps=6):
    s_hog, s_g = torch.cuda.Stream(), torch.cuda.Stream()
    lib = _lib.lib()
    lib.tdb_debug_spin.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_void_p]
    for flags, tag in ((DYNAMIC, "clc"), (0, "static")):
        line = f"{name:30s} {tag:6s} |"
        for hog in (0, 8, 16, 32):
            for _ in range(2):
                fn(flags)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if hog:
                _lib.check(lib.tdb_debug_spin(hog, 200 * 1024, 6_000_000, C.c_void_p(s_hog.cuda_stream)), "spin")   # ~3 ms
            with torch.cuda.stream(s_g):
                torch.cuda._sleep(200_000)         # let the hog settle on its SMs first
                e0.record()
                for _ in range(reps):
                    fn(flags)
                e1.record()
            torch.cuda.synchronize()
            line += f" hog {hog:2d}: {e0.elapsed_time(e1) / reps * 1e3:7.1f} us |"
        P(line)


if __name__ == "__main__":
    P(torch.cuda.get_device_name(0))
    F, h, w, Cc = 125, 22, 22, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    Rp = F * (h + 2) * (w + 2)
    x = (torch.randn(Rp, Cc, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    wk = (torch.randn(Cc, 9 * Cc, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    y = torch.empty(F * h * w, Cc, dtype=torch.bfloat16, device="cuda")
    sc, sh = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    taps = [(kh - 1) * (w + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    run_case("layer3 conv2 3x3 (pair, halo)", lambda fl: gemm(x, wk, y, Rp, Cc, Cc, ntaps=9, a_off1=taps, b_off0=[t * Cc for t in range(9)], scale=sc,
                                                               bias=sh, relu=True, remap=REMAP_P2C, img_hw=(h, w), debug_flags=fl))
    M = F * h * w
    A = (torch.randn(M, 1024, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    B = (torch.randn(256, 1024, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    o = torch.empty(M, 256, dtype=torch.bfloat16, device="cuda")
    run_case("layer3 conv1 1024->256 (pair)", lambda fl: gemm(A, B, o, M, 256, 1024, scale=sc, bias=sh, relu=True, debug_flags=fl))
    A2 = (torch.randn(M, 256, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    B2 = (torch.randn(1024, 256, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    R2 = torch.randn(M, 1024, device="cuda", generator=g).to(torch.bfloat16)
    o2 = torch.empty(M, 1024, dtype=torch.bfloat16, device="cuda")
    sc2, sh2 = torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda")
    run_case("layer3 conv3 256->1024+res (1cta)", lambda fl: gemm(A2, B2, o2, M, 1024, 256, scale=sc2, bias=sh2, residual=R2, relu=True, debug_flags=fl))
