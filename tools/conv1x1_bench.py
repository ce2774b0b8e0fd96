"""Per-shape comparison of the three routes a short-reduction bf16 GEMM with an un-remapped output can take, on the cfg-2 backbone /
decoder shapes (125 frames at res 352): (a) the 1-CTA kernel (tdb_gemm_kernel<128,5>: TMA residual in / TMA store out, 128 x 128 tiles),
(b) the pair kernel with the row-per-thread epilogue (tdb_gemm2_kernel<0>), (c) the pair kernel with the TMA-box epilogue
(tdb_gemm2_kernel<1>).  Device time from CUDA-graph replays over rotating operand sets (> L2), so that neither the ctypes launch
path nor a warm L2 flatters a route.  Output: gpurun_out/conv1x1_bench.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200.gemm import REMAP_C2P, REMAP_NONE, REMAP_P2C, gemm  # noqa: E402

TWO_CTA, NO_TMA = 64, 512
os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/conv1x1_bench.txt", "w")


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


def case(name, M, N, K, res, relu, mask=False, bmaj=0, ntaps=1, remap=REMAP_NONE, img=None, frames=0):
    g = torch.Generator(device="cuda").manual_seed(0)
    per = (M * K + M * N * (1 + int(res) + int(mask))) * 2
    nb = max(2, min(6, int(400e6 // per) + 1))
    Kw = K * ntaps
    sets = []
    for _ in range(nb):
        A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        R = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if res else None
        Mk = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if mask else None
        orows = M
        if remap == REMAP_C2P:
            orows = frames * (img[0] + 2) * (img[1] + 2)
        elif remap == REMAP_P2C:
            orows = frames * img[0] * img[1]
        out = torch.zeros(orows, N, dtype=torch.bfloat16, device="cuda")
        sets.append((A, R, Mk, out))
    B = (torch.randn(K, ntaps * N, device="cuda", generator=g) if bmaj else torch.randn(N, Kw, device="cuda", generator=g)).to(torch.bfloat16) * 0.05
    sc = torch.rand(N, device="cuda") + 0.5
    bi = torch.randn(N, device="cuda")
    kw = dict(b_major=bmaj, scale=sc, bias=bi, relu=relu)
    if ntaps > 1:
        kw.update(ntaps=ntaps, a_off0=[0] * ntaps, b_off0=[t * (N if bmaj else K) for t in range(ntaps)])
    if remap != REMAP_NONE:
        kw.update(remap=remap, img_hw=img)
    if remap == REMAP_P2C and ntaps == 9:
        wp = img[1] + 2
        kw.update(a_off1=[(kh - 1) * wp + (kw_ - 1) for kh in range(3) for kw_ in range(3)])
    alg = (M * K + N * Kw + M * N * (1 + int(res) + int(mask))) * 2
    line = f"{name:34s} M={M:7d} N={N:5d} K={K:5d}x{ntaps} alg {alg / 1e6:7.1f} MB |"
    outs = {}
    for tag, fl in (("old", NO_TMA), ("pair/row", TWO_CTA | NO_TMA), ("pair/tma", TWO_CTA)):
        def run(i, fl=fl):
            A, R, Mk, out = sets[i % nb]
            gemm(A, B, out, M, N, K, residual=R, mask=Mk, debug_flags=fl, **kw)
        us = timed(run, 4 * nb)
        outs[tag] = sets[0][3].clone()
        line += f" {tag} {us:7.1f} us {alg / us / 1e3:6.0f} GB/s |"
    same = torch.equal(outs["pair/row"], outs["pair/tma"])
    d1 = (outs["old"].float() - outs["pair/tma"].float()).abs().max().item()
    P(line, "tma==row" if same else "TMA != ROW", f"max|old-tma| {d1:.3g}")


if __name__ == "__main__":
    P(torch.cuda.get_device_name(0))
    F, f25 = 125, 25
    case("layer1 conv3+res", F * 88 * 88, 256, 64, True, True)
    case("layer1 downsample", F * 88 * 88, 256, 64, False, False)
    case("layer2 conv3+res", F * 44 * 44, 512, 128, True, True)
    case("layer2 downsample", F * 44 * 44, 512, 256, False, False)
    case("layer3.0 conv1 (44x44)", F * 44 * 44, 256, 512, False, True)
    case("layer3 conv3+res", F * 22 * 22, 1024, 256, True, True)
    case("layer3 downsample", F * 22 * 22, 1024, 512, False, False)
    case("layer3 conv2 via im2col (s2)", F * 22 * 22, 256, 2304, False, True)
    case("layer4 conv3+res", F * 11 * 11, 2048, 512, True, True)
    case("layer4 downsample", F * 11 * 11, 2048, 1024, False, False)
    case("layer4.0 conv1 (22x22)", F * 22 * 22, 512, 1024, False, True)
    case("kv proj all layers (2 taps)", 100 * 141, 1536, 256, False, False, ntaps=2)
    case("enc FFN1 (2 taps)", 25 * 141, 2048, 256, False, True, ntaps=2)
    case("layer3 conv1 -> padded grid", F * 22 * 22, 256, 1024, False, True, remap=REMAP_C2P, img=(22, 22), frames=F)
    case("layer4 conv1 -> padded grid", F * 11 * 11, 512, 2048, False, True, remap=REMAP_C2P, img=(11, 11), frames=F)
    case("layer3 conv2 3x3 (halo)", F * 24 * 24, 256, 256, False, True, ntaps=9, remap=REMAP_P2C, img=(22, 22), frames=F)
    case("layer4 conv2 3x3 (halo)", F * 13 * 13, 512, 512, False, True, ntaps=9, remap=REMAP_P2C, img=(11, 11), frames=F)
    case("layer3 conv3 dgrad+mask -> padded (25f)", f25 * 22 * 22, 256, 1024, False, False, mask=True, bmaj=1, remap=REMAP_C2P, img=(22, 22), frames=f25)
    case("layer3 conv2 dgrad 3x3+mask (25f)", f25 * 24 * 24, 256, 256, False, False, mask=True, bmaj=1, ntaps=9, remap=REMAP_P2C, img=(22, 22), frames=f25)
    case("layer3 conv1 dgrad+res+mask (25f)", f25 * 22 * 22, 1024, 256, True, False, mask=True, bmaj=1)
    case("layer2 conv1 dgrad+res+mask (25f)", f25 * 44 * 44, 512, 128, True, False, mask=True, bmaj=1)
    case("layer4 conv1 dgrad+res+mask (25f)", f25 * 11 * 11, 2048, 512, True, False, mask=True, bmaj=1)
