"""Data parallelism for the TubeDETR step: clips shard over ranks (one process per GPU), no activation exchange.

The only data-path collective is ONE all-reduce of a flat fp32 gradient buffer per step; it
replaces DistributedDataParallel's bucketed reducer (reference main.py:372-376).  Every parameter's .grad is a view into
the flat buffer, so backward writes gradients in place (static addresses => CUDA-graph replayable) and a single
NCCL call over NVLink moves them.  SetCriterion keeps the reference's own 4-byte num_boxes all-reduce
(models/tubedetr.py:411-413).
"""
import torch
import torch.distributed as dist


class FlatGradBuffer:
    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        device = device or self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def views(self):
        out, o = [], 0
        for p in self.params:
            out.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        return out

    def pack(self):
        """copy freshly assigned .grad tensors into the flat buffer with one multi-tensor copy (missing grads -> 0)."""
        if not hasattr(self, "_views"):
            self._views = self.views()
        have = [(v, p.grad) for v, p in zip(self._views, self.params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        for v, p in zip(self._views, self.params):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])

    def all_reduce(self, average=True):
        """sum over ranks (then / world): after this every rank holds the gradient of the mean loss over all clips."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat

    def nbytes(self):
        return self.flat.numel() * 4
