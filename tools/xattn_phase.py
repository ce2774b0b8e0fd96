"""Decoder cross-attention kernel: tensor-pipe utilisation OVER ITS MMA PHASE from in-kernel SM clock
stamps, plus the streaming view (ii): algorithmic bytes / kernel time against the measured HBM peak.
  python tools/xattn_phase.py > gpurun_out/xattn_phase.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tubedetr_b200.probes import XATTN_MMA_CYCLES, xattn_phase  # noqa: E402

pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
pk = json.load(open(pj))["hbm_gbs"] if os.path.exists(pj) else 6555.2
for pair, F, S in ((0, 100, 141), (1, 100, 141), (0, 400, 141), (1, 400, 141), (0, 800, 141), (1, 800, 141)):
    r = xattn_phase(F, S, pair)
    print(f"{r['kernel']}, F={F} S={S}: {r['tiles']} tiles | fused+merge {r['us_per_layer_fused_plus_merge']:.1f} us per layer in a graph replay "
          f"({r['algorithmic_gbs']:.0f} GB/s of algorithmic bytes = {r['algorithmic_gbs'] / pk:.2f} of the measured HBM copy peak)")
    print(f"   MMA phase (first issue -> last MMA complete): median {r['mma_phase_cycles_median']:.0f} cycles, min {r['mma_phase_cycles_min']:.0f}, "
          f"max {r['mma_phase_cycles_max']:.0f}  => tensor pipe over the MMA phase = {r['tensor_pipe_pct_over_mma_phase']:.1f} % (ideal {XATTN_MMA_CYCLES} cycles)")
    print(f"   CTA lifetime median {r['cta_lifetime_cycles_median']:.0f} cycles;  MMA share of the lifetime {r['tensor_pipe_pct_of_cta_lifetime']:.1f} %")
    if r["globaltimer"]:
        g = r["globaltimer"]
        print(f"   %globaltimer: CTA entry spread across the grid {g['cta_entry_spread_us']:.2f} us, first entry -> last exit {g['first_entry_to_last_exit_us']:.2f} us, "
              f"CTA lifetime median {g['cta_lifetime_us_median']:.2f} us")
