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


def default_subgroup_of(name):
    """order INSIDE the backbone group = the order in which the backbone backward completes its stages (layer4, layer3, layer2),
    so that a finished stage is one contiguous slice that can be all-reduced while the earlier stages still run"""
    if "backbone" in name:
        for li in (4, 3, 2):
            if f"layer{li}." in name:
                return 4 - li
    return 0


class FlatGradBuffer:
    """Every trainable parameter owns a slice of ONE flat fp32 buffer, ordered group by group."""

    def __init__(self, params, device=None, groups=None, align=8, subgroups=None):
        """params: iterable of tensors (one group) or, with groups=callable(name)->int, an iterable of (name, tensor).
        Every parameter's slice starts at a multiple of `align` elements (zero padding in between: 32 B for fp32, 16 B for a
        bf16 mirror of the same layout -> vector loads and TMA tensor maps can address each slice directly)."""
        if groups is None:
            items = [(0, 0, p) for p in params if p.requires_grad]
        else:
            items = [(int(groups(n)), int(subgroups(n)) if subgroups else 0, p) for n, p in params if p.requires_grad]
        order = sorted(range(len(items)), key=lambda i: (items[i][0], items[i][1], i))   # stable: group-major, declaration order inside
        self.params = [items[i][2] for i in order]
        self.group_ids = [items[i][0] for i in order]
        self.sub_ids = [items[i][1] for i in order]
        device = device or self.params[0].device
        pad = lambda n: (n + align - 1) // align * align
        self.offsets, o = [], 0
        for p in self.params:
            self.offsets.append(o)
            o += pad(p.numel())
        self.flat = torch.zeros(o, dtype=torch.float32, device=device)
        self.bounds = {}                      # group -> (lo, hi) element range (hi includes the last slice's padding)
        self.sub_bounds = {}                  # (group, subgroup) -> (lo, hi)
        for g, sgp, p, off in zip(self.group_ids, self.sub_ids, self.params, self.offsets):
            lo, hi = self.bounds.get(g, (off, off))
            self.bounds[g] = (lo, off + pad(p.numel()))
            lo, hi = self.sub_bounds.get((g, sgp), (off, off))
            self.sub_bounds[(g, sgp)] = (lo, off + pad(p.numel()))
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        self._views = None
        self.bound = False

    def bind_destinations(self):
        """let the backward kernels write gradients straight into this buffer (ops.set_grad_dest): no pack copy for them, and a
        slice is complete -- ready for its all-reduce -- as soon as its producers have run"""
        from . import ops
        if self._views is None:
            self._views = self.views()
        ops.set_grad_dest(self.params, self._views)
        self.bound = True

    def sub_segment(self, g, sub):
        lo, hi = self.sub_bounds[(g, sub)]
        return self.flat[lo:hi]

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

    def all_reduce(self, average=True, groups=None, tensor=None, group=None):
        """sum over ranks (then / world): after this every rank holds the gradient of the mean loss over all clips.
        tensor: an explicit slice of the flat buffer (default: `groups`, default: everything); group: process group (communicator)."""
        t = tensor if tensor is not None else (self.flat if groups is None else self.segment(*groups))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and t.numel():
            if average and dist.get_backend(group) == "nccl":
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)          # the division rides inside the NCCL kernel
            else:
                dist.all_reduce(t, group=group)
                if average:
                    t.div_(dist.get_world_size())
        return t

    def nbytes(self):
        return self.flat.numel() * 4


def backward_overlapped(total, fgb, text_out, backbone_out, side_stream=None, engine=None, comm_stream=None, comm_group=None):
    """Backward of `total` with the gradient all-reduce overlapped with the backbone backward.

    text_out / backbone_out: the two trunk outputs the graph is cut at (RoBERTa `last_hidden_state`, backbone features;
    `TubeDETR.trunk_outputs()`), either may be None / not require grad (frozen trunk).
      phase 1  d total / d (GROUP_REST parameters, trunk outputs)                      -- decoder, encoder, heads
               -> all-reduce(rest slice) on `comm_stream`                              -- overlaps everything below
      phase 2a text encoder backward -> pack -> all-reduce(text slice)                 -- on `side_stream`
      phase 2b backbone backward on the current stream; with `engine` (ResNet101Engine) and a buffer whose views are bound as
               gradient destinations (fgb.bind_destinations()) every finished stage (layer4, layer3, layer2) is all-reduced on
               `comm_stream` while the earlier stages still run, so only the last, small slice (layer2: 5 MB) is exposed;
               otherwise the backbone slice is packed and reduced after the backward.
    The collectives of `comm_stream` use `comm_group` (a second communicator: two NCCL streams of ONE communicator may not
    overlap safely); the text slice uses the default group.  Everything is enqueued without host synchronisation (CUDA-graph
    capturable).  On return the current stream has joined both helper streams; fgb.flat holds the averaged gradients.
    """
    rest = fgb.group_params(GROUP_REST)
    text = fgb.group_params(GROUP_TEXT)
    back = fgb.group_params(GROUP_BACKBONE)
    cuts = [t for t in (text_out, backbone_out) if t is not None and t.requires_grad]
    if fgb.bound:
        from . import ops
        ops.begin_backward()
    g1 = torch.autograd.grad(total, rest + cuts, allow_unused=True)
    g_rest, g_cuts = list(g1[:len(rest)]), dict(zip([id(t) for t in cuts], g1[len(rest):]))
    cuda = fgb.flat.is_cuda
    main = torch.cuda.current_stream(fgb.flat.device) if cuda else None
    if cuda and side_stream is None:
        side_stream = torch.cuda.Stream(device=fgb.flat.device)
    split = cuda and comm_stream is not None           # rest / backbone slices on their own stream + communicator

    def on_comm(tensor):
        """all-reduce a finished slice on the communication stream, ordered after everything enqueued on the main stream so far"""
        comm_stream.wait_stream(main)
        with torch.cuda.stream(comm_stream):
            fgb.all_reduce(tensor=tensor, group=comm_group)

    fgb.pack(groups=(GROUP_REST,), grads=g_rest)
    if split:
        on_comm(fgb.segment(GROUP_REST))

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

    # ---- 2b: backbone trunk on the main stream
    g_back = [None] * len(back)
    gb = g_cuts.get(id(backbone_out)) if backbone_out is not None else None
    staged = split and engine is not None and fgb.bound and gb is not None and bool(back)
    if gb is not None and back:
        if staged:       # stage li of the backbone has written all its gradients (in place, into fgb.flat): reduce that slice now
            engine.stage_callback = lambda li: on_comm(fgb.sub_segment(GROUP_BACKBONE, 4 - li))
        try:
            g_back = torch.autograd.grad(backbone_out, back, gb, allow_unused=True)
        finally:
            if staged:
                engine.stage_callback = None
    if staged:
        assert all(g is not None and g.data_ptr() == v.data_ptr() for g, v in
                   zip(g_back, [fgb._views[i] for i, gid in enumerate(fgb.group_ids) if gid == GROUP_BACKBONE])), \
            "staged all-reduce needs every backbone gradient written in place"
    else:
        fgb.pack(groups=(GROUP_BACKBONE,), grads=list(g_back))
        if split:
            on_comm(fgb.segment(GROUP_BACKBONE))
        else:
            fgb.all_reduce(groups=(GROUP_REST, GROUP_BACKBONE))
    if cuda:
        main.wait_stream(side_stream)
        if split:
            main.wait_stream(comm_stream)
    # hand the results to the optimizer the usual way: .grad = view into the flat buffer
    for p, v in zip(fgb.params, fgb._views):
        p.grad = v
    return fgb.flat
