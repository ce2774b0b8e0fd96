"""Bring-up diagnostic for the tcgen05 GEMM: runs structured patterns per operand layout and dumps what came out.
Usage (on the GPU box): python tools/gemm_diag.py  -> gpurun_out/gemm_diag.txt + .pt dumps"""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200.gemm import gemm  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/gemm_diag.txt", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    log.write(s + "\n")
    log.flush()


def run(name, A, B, M, N, K, ref, **kw):
    try:
        out = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
        gemm(A, B, out, M, N, K, **kw)
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        nan = torch.isnan(out).sum().item()
        P(f"[{name}] max_err={err:.4g} nans={nan} ref_absmax={ref.abs().max().item():.4g} kw={ {k: v for k, v in kw.items() if not torch.is_tensor(v)} }")
        if not (err < 1e-2 * (ref.abs().max().item() + 1e-6)):
            torch.save({"out": out.cpu(), "ref": ref.cpu()}, f"gpurun_out/diag_{name}.pt")
            bad = ((out - ref).abs() > 1e-2 * ref.abs().max()).nonzero()
            P("   first bad idx:", bad[:8].tolist(), "count", bad.shape[0])
            P("   out[0,:8]", out[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist())
        return err
    except Exception:
        P(f"[{name}] EXCEPTION\n" + traceback.format_exc())
        return float("inf")


def main():
    P(torch.cuda.get_device_name(0), torch.version.cuda)
    g = torch.Generator().manual_seed(0)
    for bn in (64, 128, 256):
        M, N, K = 128, bn, 64
        A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
        B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
        run(f"nt_1tile_bn{bn}", A, B, M, N, K, A.float() @ B.float().t(), block_n=bn)
        K = 256
        A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
        B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
        run(f"nt_k256_bn{bn}", A, B, M, N, K, A.float() @ B.float().t(), block_n=bn)
    # permutation probe: A = identity-ish so D[m,n] = B[n,m]
    M, N, K = 128, 64, 64
    A = torch.zeros(M, K)
    A[:64, :64] = torch.eye(64)
    B = (torch.arange(N)[:, None] * 64 + torch.arange(K)[None, :]).float() / 64.0
    run("perm_probe", A.to(torch.bfloat16).cuda(), B.to(torch.bfloat16).cuda(), M, N, K,
        A.to(torch.bfloat16).float().cuda() @ B.to(torch.bfloat16).float().cuda().t(), block_n=64)
    # multi-tile persistent
    M, N, K = 4096, 512, 512
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    run("nt_big", A, B, M, N, K, A.float() @ B.float().t())
    run("nt_big_8ctas", A, B, M, N, K, A.float() @ B.float().t(), max_ctas=8)
    # MN-major B (dgrad form) with both LBO/SBO conventions
    M, Kr, N = 256, 128, 128
    A = torch.randn(M, Kr, generator=g).to(torch.bfloat16).cuda()
    Bm = torch.randn(Kr, N, generator=g).to(torch.bfloat16).cuda()
    for bn in (64, 128):
        for dbg in (0, 1):
            run(f"b_mn_bn{bn}_dbg{dbg}", A, Bm, M, N, Kr, A.float() @ Bm.float(), b_major=1, block_n=bn, debug_flags=dbg)
    # MN-major A and B (wgrad form)
    R, Mo, No = 192, 128, 128
    Am = torch.randn(R, Mo, generator=g).to(torch.bfloat16).cuda()
    Bm = torch.randn(R, No, generator=g).to(torch.bfloat16).cuda()
    for dbg in (0, 1):
        run(f"ab_mn_dbg{dbg}", Am, Bm, Mo, No, R, Am.float().t() @ Bm.float(), a_major=1, b_major=1, block_n=128, debug_flags=dbg)
    # timing of a large GEMM
    M, N, K = 16384, 2048, 2048
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for bn in (128, 256):
        try:
            for _ in range(3):
                gemm(A, B, out, M, N, K, block_n=bn)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                gemm(A, B, out, M, N, K, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            P(f"[perf bn{bn}] {M}x{N}x{K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
        except Exception:
            P("[perf] EXCEPTION\n" + traceback.format_exc())
    ref = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        torch.matmul(A, B.t(), out=ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(A, B.t(), out=ref)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    P(f"[perf cublas] {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
