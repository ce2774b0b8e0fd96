"""Measurement probes shared by bench.py and tools/ (no product logic here)."""
import ctypes as C
import math

import torch

from . import kernels as K
from ._lib import lib

XATTN_MMA_CYCLES = 2 * 16 * 128      # K and V projection of a 128-row tile: 16 tcgen05.mma (M128 N256 K16) each, 128 cycles per MMA


def xattn_phase(F=100, S=141, pair=0, reps=24):
    """Decoder cross-attention kernel (fused KV projection + attention, reference models/transformer.py:724-745) at F frames x S
    memory tokens.  Returns (i) the tensor-pipe utilisation over the kernel's MMA phase from in-kernel SM
    clock stamps: 4096 MMA cycles / (last MMA complete - first MMA issue), median over tiles, and (ii) the streaming view:
    algorithmic bytes (memory + pos read once, weights) / device time of fused + merge kernels in a CUDA-graph replay over
    rotating HBM-resident inputs."""
    d = 256
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(1)
    q = torch.randn(F, d, generator=g).bfloat16().to(dev)
    nb = 8                       # rotate memory buffers so the timed launches read HBM, not L2
    mems = [(torch.randn(F * S, d, generator=g).bfloat16().to(dev), torch.randn(F * S, d, generator=g).bfloat16().to(dev))
            for _ in range(nb)]
    W = (torch.randn(2 * d, d, generator=g) / 16).bfloat16().to(dev)
    bv = torch.zeros(d, device=dev)
    kpm = torch.zeros(F, S, dtype=torch.uint8, device=dev)
    o = torch.empty(F, d, dtype=torch.bfloat16, device=dev)
    p = torch.empty(F, 8, 1, S, device=dev)
    pbar = torch.empty(F, 1, S, device=dev)
    tiles = (F * S + 127) // 128
    stamps = torch.zeros(tiles * 6 + 8, dtype=torch.int64, device=dev)     # [tiles][4] SM clocks + [tiles][2] globaltimer ns

    def run(i):
        K.xattn_fused_fwd(q, mems[i % nb][0], mems[i % nb][1], W, bv, kpm, o, p, pbar, F, S, 1 / math.sqrt(32))
    lib().tdb_xattn_set_pair(int(pair))
    try:
        for i in range(4):
            run(i)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()          # replay a graph: the eager ctypes launch path (~10 us per call) would hide the kernels
        with torch.cuda.graph(gr, capture_error_mode="thread_local"):   # an NCCL watchdog thread may be polling events (N > 1)
            for i in range(reps):
                run(i)
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (5 * reps) * 1e3
        lib().tdb_xattn_set_timing_buffer(C.c_void_p(stamps.data_ptr()))
        run(5)
        torch.cuda.synchronize()
    finally:
        lib().tdb_xattn_set_timing_buffer(None)
        lib().tdb_xattn_set_pair(0)
    raw = stamps.cpu()
    st = raw[:tiles * 4].view(tiles, 4).double()
    skew = None
    if not pair:                             # (the pair kernel does not write the global timer stamps)
        gt = raw[tiles * 4:tiles * 6].view(tiles, 2).double()
        skew = {"cta_entry_spread_us": float(gt[:, 0].max() - gt[:, 0].min()) / 1e3,
                "first_entry_to_last_exit_us": float(gt[:, 1].max() - gt[:, 0].min()) / 1e3,
                "cta_lifetime_us_median": float((gt[:, 1] - gt[:, 0]).median()) / 1e3}
    lead = st[::2] if pair else st           # the pair kernel stamps both tiles of a pair with the leader's clock
    mma = lead[:, 2] - lead[:, 1]
    life = lead[:, 3] - lead[:, 0]
    alg = 2 * F * S * d * 2 + 2 * d * d * 2
    return {"kernel": "xattn_fused2_kernel (CTA pair, cta_group::2)" if pair else "xattn_fused_kernel", "frames": F, "tokens": S, "tiles": tiles,
            "mma_phase_cycles_median": float(mma.median()), "mma_phase_cycles_min": float(mma.min()), "mma_phase_cycles_max": float(mma.max()),
            "tensor_pipe_pct_over_mma_phase": 100.0 * XATTN_MMA_CYCLES / float(mma.median()),
            "cta_lifetime_cycles_median": float(life.median()),
            "tensor_pipe_pct_of_cta_lifetime": 100.0 * XATTN_MMA_CYCLES / float(life.median()),
            "us_per_layer_fused_plus_merge": us, "globaltimer": skew, "algorithmic_bytes": alg, "algorithmic_gbs": alg / us / 1e3}


def decoder_attn_hoisted(F=100, S=141, nl=6, reps=12, peak_tflops=None, peak_gbs=None):
    """The decoder cross-attention as the model runs it by default (reference models/transformer.py:567-579, 724-745): (i) the K / V
    projections of ALL `nl` layers as two tcgen05 GEMMs [F*S, 256] x [nl*256, 256]^T with split-precision weights (two reduction
    taps: executed tensor-core work = 2 x algorithmic), (ii) per layer the streaming one-query attention core on a 256-column slice.
    Device time from CUDA-graph replays over rotating inputs (> L2 for the GEMM operands)."""
    from .gemm import gemm
    d = 256
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(2)
    R = F * S
    nb = 4
    mems = [(torch.randn(R, d, generator=g).bfloat16().to(dev), torch.randn(R, d, generator=g).bfloat16().to(dev)) for _ in range(nb)]
    outs = [(torch.empty(R, nl * d, dtype=torch.bfloat16, device=dev), torch.empty(R, nl * d, dtype=torch.bfloat16, device=dev)) for _ in range(nb)]
    Wk = (torch.randn(nl * d, 2 * d, generator=g) / 16).bfloat16().to(dev)
    Wv = (torch.randn(nl * d, 2 * d, generator=g) / 16).bfloat16().to(dev)
    bias = torch.zeros(nl * d, device=dev)
    q = torch.randn(F, d, generator=g).bfloat16().to(dev)
    kpm = torch.zeros(F, S, dtype=torch.uint8, device=dev)
    o = torch.empty(F, d, dtype=torch.bfloat16, device=dev)
    p = torch.empty(F, 8, 1, S, device=dev)
    pbar = torch.empty(F, 1, S, device=dev)

    def proj(i):
        m, out = mems[i % nb], outs[i % nb]
        gemm(m[0], Wk, out[0], R, nl * d, d, ntaps=2, a_off0=(0, 0), b_off0=(0, d), bias=bias)
        gemm(m[1], Wv, out[1], R, nl * d, d, ntaps=2, a_off0=(0, 0), b_off0=(0, d), bias=bias)

    def core(i):
        out = outs[i % nb]
        l = i % nl
        K.xattn_core_fwd(q, out[0][:, l * d:(l + 1) * d], out[1][:, l * d:(l + 1) * d], kpm, o, p, pbar, F, S, 1 / math.sqrt(32))

    def timed(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode="thread_local"):
            for i in range(reps):
                fn(i)
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (5 * reps) * 1e3

    us_proj = timed(proj)
    us_core = timed(core)
    flops = 2.0 * 2 * R * nl * d * d                      # algorithmic: K and V projections of all layers
    bytes_proj = 2 * (R * d + nl * d * 2 * d + R * nl * d) * 2
    bytes_core = 2 * R * d * 2 + F * d * 2 * 2 + F * 8 * S * 4
    res = {"path": "hoisted K/V projections (2 x tdb_gemm, all layers) + per-layer xattn_core_fwd_kernel", "frames": F, "tokens": S, "layers": nl,
           "kv_proj_us_all_layers": us_proj, "kv_proj_algorithmic_tflops": flops / us_proj / 1e6,
           "kv_proj_executed_tflops": 2 * flops / us_proj / 1e6, "kv_proj_algorithmic_bytes": bytes_proj,
           "kv_proj_gbs": bytes_proj / us_proj / 1e3, "core_us_per_layer": us_core, "core_algorithmic_bytes": bytes_core,
           "core_gbs": bytes_core / us_core / 1e3, "us_all_layers": us_proj + nl * us_core}
    if peak_tflops:
        res["kv_proj_executed_frac_of_tensor_peak"] = res["kv_proj_executed_tflops"] / peak_tflops
    if peak_gbs:
        res["kv_proj_frac_of_hbm_peak"] = res["kv_proj_gbs"] / peak_gbs
        res["core_frac_of_hbm_peak"] = res["core_gbs"] / peak_gbs
    return res
