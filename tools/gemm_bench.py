"""Micro-benchmark of tdb_gemm on the shapes that dominate the TubeDETR step (from the TDB_GEMM_LOG of a profiled step),
per epilogue mode, with a correctness check against fp32 torch.  Output: gpurun_out/gemm_bench.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200.gemm import REMAP_C2P, REMAP_P2C, gemm  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/gemm_bench.txt", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    log.write(s + "\n")
    log.flush()


def bench(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def case(name, M, N, K, scale=False, residual=False, relu=False, mask=False, bmaj=0, f32=False, check=True, modes=(5, 6)):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = (torch.randn(K, N, device="cuda", generator=g) if bmaj else torch.randn(N, K, device="cuda", generator=g)).to(torch.bfloat16)
    sc = torch.rand(N, device="cuda") + 0.5 if scale else None
    bi = torch.randn(N, device="cuda") if scale else None
    R = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if residual else None
    Mk = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if mask else None
    out = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
    byts = (A.numel() + B.numel()) * 2 + out.numel() * out.element_size() + (R.numel() * 2 if residual else 0) + (Mk.numel() * 2 if mask else 0)
    res = []
    for mode in modes:
        def run():
            gemm(A, B, out, M, N, K, b_major=bmaj, scale=sc, bias=bi, residual=R, relu=relu, mask=Mk, debug_flags=mode << 1)
        us = bench(run)
        err = float("nan")
        if check and M * N <= 64 * 1024 * 1024:
            ref = A[:4096].float() @ (B.float() if bmaj else B.float().t())
            if scale:
                ref = ref * sc + bi
            if residual:
                ref = ref + R[:4096].float()
            if relu:
                ref = torch.relu(ref)
            if mask:
                ref = ref * (Mk[:4096].float() > 0)
            err = ((out[:4096].float() - ref).abs().max() / (ref.abs().max() + 1e-6)).item()
        res.append(f"mode{mode - 1}: {us:8.1f} us {2.0 * M * N * K / us / 1e6:7.1f} TF {byts / us / 1e3:7.1f} GB/s err {err:.2e}")
    P(f"{name:34s} M={M:7d} N={N:5d} K={K:5d} | " + " | ".join(res))


def experiment():
    """where does the epilogue time go?  mode 3 with stores / TMEM loads switched off (results are garbage on purpose)"""
    M, N, K = 48400, 1024, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    R = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    sc, bi = torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for name, kw, fl in (("full res+relu", dict(scale=sc, bias=bi, residual=R, relu=True), 0),
                         ("no residual", dict(scale=sc, bias=bi, relu=True), 0),
                         ("plain", dict(), 0), ("plain, no stores", dict(), 16), ("plain, no tmem ld", dict(), 32),
                         ("plain, no stores no ld", dict(), 48)):
        for bn in (128, 256):
            us = bench(lambda: gemm(A, B, out, M, N, K, block_n=bn, debug_flags=(4 << 1) | fl, **kw))
            P(f"experiment {name:24s} bn{bn}: {us:8.1f} us")


def two_cta():
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, N, K) in ((16384, 2048, 2048), (48400, 256, 1024), (48400, 1024, 256)):
        A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        for nm, fl in (("1cta", 0), ("2cta", 64)):
            us = bench(lambda: gemm(A, B, out, M, N, K, debug_flags=fl))
            P(f"{nm} plain {M}x{N}x{K}: {us:8.1f} us {2.0 * M * N * K / us / 1e6:7.1f} TF")
    N_, h, w, C = 100, 22, 22, 256
    Rp = N_ * (h + 2) * (w + 2)
    x = torch.randn(Rp, C, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(C, 9 * C, device="cuda") * 0.02).to(torch.bfloat16)
    y = torch.empty(N_ * h * w, C, dtype=torch.bfloat16, device="cuda")
    sc, sh = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    taps = [(kh - 1) * (w + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    for nm, fl in (("1cta", 0), ("2cta", 64)):
        us = bench(lambda: gemm(x, wk, y, Rp, C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], scale=sc, bias=sh,
                                relu=True, remap=REMAP_P2C, img_hw=(h, w), debug_flags=fl))
        P(f"{nm} l3 conv2 3x3 fast: {us:8.1f} us  {2.0 * N_ * h * w * C * 9 * C / us / 1e6:7.1f} TF (algorithmic)")


def main():
    P(torch.cuda.get_device_name(0))
    if "--2cta" in sys.argv:
        two_cta()
        return
    experiment()
    case("l3 conv3 fast (res+relu)", 48400, 1024, 256, scale=True, residual=True, relu=True)
    case("l3 conv3 slow (res+relu)", 12100, 1024, 256, scale=True, residual=True, relu=True)
    case("l3 dgrad1 slow (res+mask, Bmn)", 12100, 1024, 256, residual=True, mask=True, bmaj=1)
    case("l1 conv3 fast (res+relu)", 774400, 256, 64, scale=True, residual=True, relu=True, check=False)
    case("l2 conv3 fast (res+relu)", 193600, 512, 128, scale=True, residual=True, relu=True, check=False)
    case("l3 conv1 fast (relu)", 48400, 256, 1024, scale=True, relu=True)
    case("l4 conv3 fast (res+relu)", 12100, 2048, 512, scale=True, residual=True, relu=True)
    case("l1 conv1 fast (relu)", 774400, 64, 256, scale=True, relu=True, check=False)
    case("l3 dgrad3 slow (mask, Bmn)", 12100, 256, 1024, mask=True, bmaj=1)
    case("enc ffn1 (bias+relu)", 3525, 2048, 256, scale=True, relu=True)
    case("enc ffn2 (f32 out)", 3525, 256, 2048, scale=True, f32=True)
    case("xattn kv proj", 14100, 256, 256, scale=True)
    case("big square", 16384, 2048, 2048)
    # implicit 3x3 conv (layer3, 100 frames) both modes
    N_, h, w, C = 100, 22, 22, 256
    Rp = N_ * (h + 2) * (w + 2)
    x = torch.randn(Rp, C, device="cuda").to(torch.bfloat16)
    wk = (torch.randn(C, 9 * C, device="cuda") * 0.02).to(torch.bfloat16)
    y = torch.empty(N_ * h * w, C, dtype=torch.bfloat16, device="cuda")
    sc, sh = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    taps = [(kh - 1) * (w + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    for mode in (4,):
        for bn in (128, 256):
            us = bench(lambda: gemm(x, wk, y, Rp, C, C, ntaps=9, a_off1=taps, b_off0=[t * C for t in range(9)], scale=sc, bias=sh,
                                    relu=True, remap=REMAP_P2C, img_hw=(h, w), debug_flags=mode << 1, block_n=bn))
            P(f"l3 conv2 3x3 fast mode{mode - 1} bn{bn}: {us:8.1f} us  {2.0 * N_ * h * w * C * 9 * C / us / 1e6:7.1f} TF (algorithmic)")


if __name__ == "__main__":
    main()
