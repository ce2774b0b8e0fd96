"""Optimizer-side step of the reference training loop on flat buffers.

Reference, per iteration (engine.py:147-161):
    optimizer.zero_grad(); losses.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)          # args.clip_max_norm = 0.1
    optimizer.step()                                                      # torch.optim.AdamW, 3 LR groups (main.py:381-414)
    adjust_learning_rate(optimizer, ...)                                  # util/optim.py:28-87: writes param_group["lr"]
    update_ema(model, model_ema, args.ema_decay)                          # util/optim.py:8-25
`FusedAdamWEMA` keeps that interface (`param_groups` in the reference's order default / backbone / text_encoder so the
reference's `adjust_learning_rate` drives it unchanged, `zero_grad`, `step`, `state_dict`) and runs clip + AdamW + EMA as
three kernel launches of libtdb.so per LR group over flat buffers (tdb_optim.cu): parameters are re-homed as views of one flat
fp32 buffer laid out like `parallel.FlatGradBuffer` (text | transformer | backbone), so the all-reduced gradient buffer feeds
the optimizer directly.  For the transformer's linear weights the kernel also writes the bf16 copy the next forward's GEMMs
read (no per-layer cast kernels after a step).  No host synchronisation: the clip coefficient stays on the device.
"""
import ctypes as C

import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr
from .parallel import GROUP_BACKBONE, GROUP_REST, GROUP_TEXT, FlatGradBuffer, default_group_of


class _Group(C.Structure):
    _fields_ = [("begin", C.c_int64), ("end", C.c_int64), ("lr", C.c_float), ("weight_decay", C.c_float)]


class FusedAdamWEMA:
    def __init__(self, model, lr=5e-5, lr_backbone=1e-5, text_encoder_lr=5e-5, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=0.1, model_ema=None, ema_decay=0.9998, fgb=None, bf16_mirror=True):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        assert named and all(p.is_cuda and p.dtype == torch.float32 for _, p in named), "fp32 CUDA parameters expected"
        self.fgb = fgb if fgb is not None else FlatGradBuffer(named, named[0][1].device, groups=default_group_of)
        f = self.fgb
        names = {id(p): n for n, p in named}
        self.names = [names[id(p)] for p in f.params]
        self.flat_p = torch.zeros_like(f.flat)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(f.flat), torch.zeros_like(f.flat)
        self.p_views = f.views(self.flat_p)
        with torch.no_grad():
            torch._foreach_copy_(self.p_views, [p.data for p in f.params])
            for p, v in zip(f.params, self.p_views):
                p.data = v                                   # the module now reads/writes the flat buffer
        self.flat_ema = None
        if model_ema is not None:                            # deepcopy of the model (main.py:370): same names
            esd = dict(model_ema.named_parameters())
            self.flat_ema = torch.zeros_like(f.flat)
            self.ema_views = f.views(self.flat_ema)
            self.ema_params = [esd[n] for n in self.names]
            with torch.no_grad():
                torch._foreach_copy_(self.ema_views, [p.data for p in self.ema_params])
                for p, v in zip(self.ema_params, self.ema_views):
                    p.data = v
        # reference order (main.py:381-405); adjust_learning_rate (util/optim.py:83-87) zips base LRs over exactly these three
        self._gid = [GROUP_REST, GROUP_BACKBONE, GROUP_TEXT]
        self.param_groups = [{"params": f.group_params(g), "lr": l, "weight_decay": weight_decay, "betas": betas, "eps": eps}
                             for g, l in zip(self._gid, (lr, lr_backbone, text_encoder_lr))]
        self.max_norm, self.ema_decay, self.step_count = max_norm, ema_decay, 0
        self.norm = torch.zeros(1, dtype=torch.float32, device=f.flat.device)
        lib().tdb_optim_workspace_bytes.restype = C.c_int64
        self._ws_bytes = int(lib().tdb_optim_workspace_bytes())
        self._ws = torch.empty(self._ws_bytes // 8, dtype=torch.float64, device=f.flat.device)
        # bf16 operand mirror of the transformer group (the only parameters ops.bf16_weight serves: 2-D linear weights)
        self.mirror = None
        if bf16_mirror and GROUP_REST in f.bounds:
            lo, hi = f.bounds[GROUP_REST]
            self.mirror = torch.zeros(hi - lo, dtype=torch.bfloat16, device=f.flat.device)
            self._mirror_views = [(p, self.mirror[off - lo:off - lo + p.numel()].view_as(p))
                                  for p, off, g in zip(f.params, f.offsets, f.group_ids) if g == GROUP_REST and p.dim() == 2]

    # ---- torch.optim.Optimizer surface used by the reference loop
    def zero_grad(self, set_to_none=False):
        if set_to_none:
            for p in self.fgb.params:
                p.grad = None
        else:
            self.fgb.zero()
            for p, v in zip(self.fgb.params, self.fgb.views()):
                p.grad = v

    def _layout(self):
        """names / offsets / sizes of the flat moment buffers: the layout follows FlatGradBuffer's ordering and padding, so a
        checkpoint is only valid for the same layout"""
        return [(n, int(o), int(p.numel())) for n, o, p in zip(self.names, self.fgb.offsets, self.fgb.params)]

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "layout": self._layout(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        if "layout" in sd and [tuple(x) for x in sd["layout"]] != self._layout():
            raise ValueError("FusedAdamWEMA.load_state_dict: the checkpoint's flat-buffer layout (parameter names / offsets / sizes) "
                             "differs from this optimizer's; it was saved for another model configuration or buffer ordering")
        if sd["exp_avg"].numel() != self.exp_avg.numel():
            raise ValueError("FusedAdamWEMA.load_state_dict: moment buffers of %d elements, expected %d" % (sd["exp_avg"].numel(), self.exp_avg.numel()))
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)

    @torch.no_grad()
    def step(self, packed=False):
        """clip_grad_norm_ + AdamW.step + update_ema.  packed=True: the flat gradient buffer is already complete (after
        parallel.backward_overlapped / FlatGradBuffer.pack); otherwise gradients assigned by autograd are packed first.
        Returns the total gradient norm as a device scalar (what clip_grad_norm_ returns), without synchronising."""
        f = self.fgb
        if not packed:
            f.pack()
        self.step_count += 1
        st = stream_ptr()
        n = f.flat.numel()
        use_clip = self.max_norm is not None and self.max_norm > 0
        if use_clip:
            check(lib().tdb_grad_sqnorm(ptr(f.flat), C.c_int64(n), ptr(self._ws), C.c_int64(self._ws_bytes), ptr(self.norm), st),
                  "grad_sqnorm")
        esz = 4
        for gid, pg in zip(self._gid, self.param_groups):
            if gid not in f.bounds:
                continue
            lo, hi = f.bounds[gid]
            grp = _Group(0, hi - lo, pg["lr"], pg["weight_decay"])
            off = lambda t, e=esz: None if t is None else C.c_void_p(t.data_ptr() + lo * e)
            mir = C.c_void_p(self.mirror.data_ptr()) if (self.mirror is not None and gid == GROUP_REST) else None
            b1, b2 = pg["betas"]
            check(lib().tdb_adamw_ema_step(off(self.flat_p), off(f.flat), off(self.exp_avg), off(self.exp_avg_sq), off(self.flat_ema),
                                           mir, C.c_int64(hi - lo), C.byref(grp), 1, C.c_float(b1), C.c_float(b2), C.c_float(pg["eps"]),
                                           C.c_int64(self.step_count), ptr(self.norm) if use_clip else None,
                                           C.c_float(self.max_norm or 0.0), C.c_float(self.ema_decay), st), "adamw_ema_step")
        # the kernels wrote parameters behind autograd's back: bump the version counters (host-side only, no kernels) so
        # version-keyed caches (ops.bf16_weight, ResNet101Engine.prepare) see the new weights
        torch.autograd.graph.increment_version(f.params)
        if self.flat_ema is not None:
            torch.autograd.graph.increment_version(self.ema_params)
        if self.mirror is not None:
            for p, v in self._mirror_views:
                ops.adopt_bf16_weight(p, v)
        return self.norm[0]
