"""Phase breakdown of one CTA of mha_tc_bwd_kernel (in-kernel SM clock stamps) at the encoder shape: python tools/attn_phase.py"""
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tubedetr_b200 import kernels as K  # noqa: E402
from tubedetr_b200._lib import lib  # noqa: E402

B, L, H, d = 25, 141, 8, 256
qk = torch.randn(B * L, 512, device="cuda").bfloat16()
v, do = torch.randn(B * L, d, device="cuda").bfloat16(), torch.randn(B * L, d, device="cuda").bfloat16()
p = torch.softmax(torch.randn(B, H, L, L, device="cuda"), -1)
dqk, dv = torch.empty_like(qk), torch.empty_like(v)
seed = torch.tensor([5], dtype=torch.int64, device="cuda")
st = torch.zeros(16, dtype=torch.int64, device="cuda")
for drop in (None, (seed, 3, 0.1)):
    for _ in range(3):
        K.mha_tc_bwd(qk[:, :256], qk[:, 256:], v, do, p, None, dqk[:, :256], dqk[:, 256:], dv, B, H, L, L, 1 / math.sqrt(32), drop=drop)
    lib().tdb_mha_tc_set_timing_buffer(C.c_void_p(st.data_ptr()))
    K.mha_tc_bwd(qk[:, :256], qk[:, 256:], v, do, p, None, dqk[:, :256], dqk[:, 256:], dv, B, H, L, L, 1 / math.sqrt(32), drop=drop)
    torch.cuda.synchronize()
    lib().tdb_mha_tc_set_timing_buffer(None)
    s = st.cpu().tolist()
    names = ["dPd ready", "pass A", "pass B", "MMA2 done", "epilogues"]
    print("dropout" if drop else "no dropout", "-- cycles from kernel entry of CTA 0 (thread = query row 0):")
    prev = s[0]
    for it in range(2):
        for j, n in enumerate(names):
            t = s[1 + 6 * it + j]
            print(f"  tile {it} {n:12s} +{t - prev:7d}  (at {t - s[0]:7d})")
            prev = t
