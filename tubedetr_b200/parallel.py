"""Data parallelism for the TubeDETR step: clips shard over ranks (one process per GPU), no activation exchange.

The only data-path collective is the all-reduce of a flat fp32 gradient buffer; it replaces
DistributedDataParallel's bucketed reducer (reference main.py:372-376).  The buffer is laid out in GROUPS
(text encoder | transformer + heads | backbone) so that a group is one contiguous slice:

  * `FlatGradBuffer.all_reduce()`            -- one NCCL call over the whole buffer after backward (serialised), or
  * `backward_overlapped(...)`               -- backward split at the two trunk outputs (RoBERTa hidden states, backbone
    features): the text-encoder slice (67 % of the bytes) is reduced on a side stream WHILE the
    backbone backward (the longest part of the step) still runs; the remaining slice follows it.  Both variants give
    bit-identical sums (same NCCL reduction per element); the split is pure scheduling.
SetCriterion keeps the reference's own 4-byte num_boxes all-reduce (models/tubedetr.py:411-413).
"""
import torch
import torch.distributed as dist

GROUP_TEXT, GROUP_REST, GROUP_BACKBONE = 0, 1, 2


def default_group_of(name):
    """the reference's own LR groups (main.py:381-405 select by these substrings)"""
    if "text_encoder" in name:
        return GROUP_TEXT
    if "backbone" in name:
        return GROUP_BACKBONE
    return GROUP_REST


class FlatGradBuffer:
    """Every trainable parameter owns a slice of ONE flat fp32 buffer, ordered group by group."""

    def __init__(self, params, device=None, groups=None, align=8):
        """params: iterable of tensors (one group) or, with groups=callable(name)->int, an iterable of (name, tensor).
        Every parameter's slice starts at a multiple of `align` elements (zero padding in between: 32 B for fp32, 16 B for a
        bf16 mirror of the same layout -> vector loads and TMA tensor maps can address each slice directly)."""
        if groups is None:
            items = [(0, p) for p in params if p.requires_grad]
        else:
            items = [(int(groups(n)), p) for n, p in params if p.requires_grad]
        order = sorted(range(len(items)), key=lambda i: (items[i][0], i))     # stable: group-major, declaration order inside
        self.params = [items[i][1] for i in order]
        self.group_ids = [items[i][0] for i in order]
        device = device or self.params[0].device
        pad = lambda n: (n + align - 1) // align * align
        self.offsets, o = [], 0
        for p in self.params:
            self.offsets.append(o)
            o += pad(p.numel())
        self.flat = torch.zeros(o, dtype=torch.float32, device=device)
        self.bounds = {}                      # group -> (lo, hi) element range (hi includes the last slice's padding)
        for g, p, off in zip(self.group_ids, self.params, self.offsets):
            lo, hi = self.bounds.get(g, (off, off))
            self.bounds[g] = (lo, off + pad(p.numel()))
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        self._views = None

    def zero(self):
        self.flat.zero_()

    def views(self, flat=None):
        """per-parameter views into `flat` (default: the gradient buffer; any tensor of the same length works: parameters,
        Adam moments, an EMA copy, a bf16 mirror)"""
        flat = self.flat if flat is None else flat
        return [flat[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]

    def group_params(self, *gs):
        return [p for g, p in zip(self.group_ids, self.params) if g in gs]

    def segment(self, *gs):
        """contiguous flat slice covering the given (adjacent) groups"""
        gs = [g for g in gs if g in self.bounds]
        if not gs:
            return self.flat[:0]
        lo, hi = min(self.bounds[g][0] for g in gs), max(self.bounds[g][1] for g in gs)
        assert hi - lo == sum(self.bounds[g][1] - self.bounds[g][0] for g in gs), "groups are not adjacent in the flat buffer"
        return self.flat[lo:hi]

    def pack(self, groups=None, grads=None):
        """copy freshly produced gradients into the flat buffer with one multi-tensor copy (missing grads -> 0).
        grads: optional list aligned with group_params(*groups) (e.g. the result of torch.autograd.grad); default p.grad."""
        if self._views is None:
            self._views = self.views()
        sel = [i for i, g in enumerate(self.group_ids) if groups is None or g in groups]
        if grads is None:
            grads = [self.params[i].grad for i in sel]
        dst, src = [], []
        for i, g in zip(sel, grads):
            v = self._views[i]
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(g.view_as(v) if g.shape != v.shape else g)
        if dst:
            torch._foreach_copy_(dst, src)

    def all_reduce(self, average=True, groups=None):
        """sum over ranks (then / world): after this every rank holds the gradient of the mean loss over all clips."""
        t = self.flat if groups is None else self.segment(*groups)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and t.numel():
            if average and dist.get_backend() == "nccl":
                dist.all_reduce(t, op=dist.ReduceOp.AVG)          # the division rides inside the NCCL kernel
            else:
                dist.all_reduce(t)
                if average:
                    t.div_(dist.get_world_size())
        return t

    def nbytes(self):
        return self.flat.numel() * 4


def backward_overlapped(total, fgb, text_out, backbone_out, side_stream=None):
    """Backward of `total` with the gradient all-reduce overlapped with the backbone backward.

    text_out / backbone_out: the two trunk outputs the graph is cut at (RoBERTa `last_hidden_state`, backbone features;
    `TubeDETR.trunk_outputs()`), either may be None / not require grad (frozen trunk).
      phase 1  d total / d (GROUP_REST parameters, trunk outputs)                      -- decoder, encoder, heads
      phase 2a text encoder backward -> pack -> all-reduce(text slice)                 -- on `side_stream`
      phase 2b backbone backward     -> pack -> all-reduce(rest + backbone slice)      -- on the current stream
    Everything is enqueued without host synchronisation (CUDA-graph capturable: the NCCL calls are captured with it).
    On return the current stream has joined the side stream; fgb.flat holds the averaged gradients.
    """
    rest = fgb.group_params(GROUP_REST)
    text = fgb.group_params(GROUP_TEXT)
    back = fgb.group_params(GROUP_BACKBONE)
    cuts = [t for t in (text_out, backbone_out) if t is not None and t.requires_grad]
    g1 = torch.autograd.grad(total, rest + cuts, allow_unused=True)
    g_rest, g_cuts = list(g1[:len(rest)]), dict(zip([id(t) for t in cuts], g1[len(rest):]))
    cuda = fgb.flat.is_cuda
    main = torch.cuda.current_stream(fgb.flat.device) if cuda else None
    if cuda and side_stream is None:
        side_stream = torch.cuda.Stream(device=fgb.flat.device)

    # ---- 2a: text trunk on the side stream
    def text_part():
        g_text = [None] * len(text)
        gt = g_cuts.get(id(text_out)) if text_out is not None else None
        if gt is not None and text:
            g_text = torch.autograd.grad(text_out, text, gt, allow_unused=True)
        fgb.pack(groups=(GROUP_TEXT,), grads=g_text)
        fgb.all_reduce(groups=(GROUP_TEXT,))
        return g_text

    if cuda:
        side_stream.wait_stream(main)
        with torch.cuda.stream(side_stream):
            g_text = text_part()
    else:
        g_text = text_part()

    # ---- 2b: backbone trunk on the main stream, then the remaining slice
    g_back = [None] * len(back)
    gb = g_cuts.get(id(backbone_out)) if backbone_out is not None else None
    if gb is not None and back:
        g_back = torch.autograd.grad(backbone_out, back, gb, allow_unused=True)
    fgb.pack(groups=(GROUP_REST, GROUP_BACKBONE), grads=g_rest + list(g_back))
    fgb.all_reduce(groups=(GROUP_REST, GROUP_BACKBONE))
    if cuda:
        main.wait_stream(side_stream)
    # hand the results to the optimizer the usual way: .grad = view into the flat buffer
    for p, v in zip(fgb.params, fgb._views):
        p.grad = v
    return fgb.flat
