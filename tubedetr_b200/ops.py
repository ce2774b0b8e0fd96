"""torch.autograd glue around the C-ABI kernels.  Each Function only marshals buffers: forward and backward math
runs in libtdb.so (tcgen05 GEMM, attention, LayerNorm kernels).  Activations that feed a GEMM are bf16, the residual
stream and normalisation statistics are fp32, parameters/gradients stay fp32 (master weights)."""
import weakref

import torch

from . import kernels as K
from .gemm import effective_splits, gemm, splitk_reduce

_WCACHE = {}

# ---- weight-gradient work on a second stream ------------------------------------------------------------------------------
# In backward only the data gradients (dgrad) feed the next node; weight gradients (wgrad GEMM + split-K reduce + bias column
# sums) are leaves.  Most backward GEMMs of this model fill less than one wave of the 148 SMs (25 slow frames, 3525 encoder
# tokens, 100 time queries), so every backward node forks its wgrad work onto a side stream that runs concurrently with its
# dgrad, and joins before it returns (the caching allocator and the reused scratch buffers never see cross-stream lifetimes
# beyond one node).  Captured by CUDA graphs as parallel branches.  TDB_WGRAD_STREAM=0 disables it.
import os as _os

WGRAD_STREAM = _os.environ.get("TDB_WGRAD_STREAM", "1") != "0"
_SIDE = {}


class wgrad_scope:
    """with sc: ...   -> the enclosed launches go to the side stream, ordered after everything enqueued on the main stream so
    far; sc.join() -> the main stream waits for the side stream."""

    def __init__(self, device):
        self.on = WGRAD_STREAM and torch.device(device).type == "cuda"
        if self.on:
            self.main = torch.cuda.current_stream(device)
            side = _SIDE.get(device)
            if side is None:
                side = _SIDE[device] = torch.cuda.Stream(device=device)
            self.side = side
            self.used = False

    def __enter__(self):
        if self.on:
            self.side.wait_stream(self.main)
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
            self.used = True
        return self

    def __exit__(self, *a):
        if self.on:
            self._ctx.__exit__(*a)
        return False

    def join(self):
        if self.on and self.used:
            self.main.wait_stream(self.side)
            self.used = False

    def mark(self):
        """event after everything enqueued on the side stream so far (None when the side stream is off / idle)"""
        if not (self.on and self.used):
            return None
        ev = torch.cuda.Event()
        ev.record(self.side)
        return ev

    def wait(self, ev):
        if ev is not None:
            self.main.wait_event(ev)


def _root(w):
    return w._base if w._base is not None else w


# ---- gradient destinations ----------------------------------------------------------------------------------------------------
# Data-parallel steps keep every gradient in ONE flat buffer (parallel.FlatGradBuffer).  When the buffer's per-parameter views are
# registered here, the backward nodes of this module write weight / bias gradients STRAIGHT into them (the split-K reduce and the
# column sums take the view as their output), so no pack copy of these tensors is needed before the all-reduce and a slice can be
# reduced as soon as its last producer has run.  Only used with torch.autograd.grad (parallel.backward_overlapped): with
# .backward() AccumulateGrad would clone a tensor that is referenced elsewhere.
_GRAD_DEST = {}
_GRAD_CLAIMED = set()


def set_grad_dest(params, views):
    _GRAD_DEST.clear()
    _GRAD_CLAIMED.clear()
    for p, v in zip(params, views):
        _GRAD_DEST[id(p)] = (weakref.ref(p), v)


def begin_backward():
    """start of a backward pass: every destination may be claimed ONCE per pass.  A parameter that feeds two backward nodes (e.g.
    input_proj.bias: slow and fast rows) gets its in-place view from the first and a fresh tensor from the second; autograd then sums
    the two into a new tensor, which the driver's pack() copies over the view."""
    _GRAD_CLAIMED.clear()


def grad_dest(p):
    ent = _GRAD_DEST.get(id(p))
    if ent is not None and ent[0]() is p and id(p) not in _GRAD_CLAIMED:
        _GRAD_CLAIMED.add(id(p))
        return ent[1]
    return None


def _grad_out(p, dtype=torch.float32):
    """output tensor for the gradient of parameter p: its flat-buffer view when registered, else a fresh tensor"""
    v = grad_dest(p)
    return v if v is not None else torch.empty(p.shape, dtype=dtype, device=p.device)


def bf16_weight(w):
    """bf16 copy of an fp32 parameter, refreshed when the parameter changes (optimizer steps bump _version).
    The entry holds a weak reference to the parameter (its base tensor for views): an address can be recycled by the caching
    allocator for a different tensor of the same shape and version, which must never hit a stale copy."""
    key = (w.data_ptr(), tuple(w.shape))
    ck = (w._version,)
    root = _root(w)
    ent = _WCACHE.get(key)
    if ent is None or ent[0] != ck or ent[2]() is not root:
        wb = ent[1] if (ent is not None and ent[1].shape == w.shape and ent[2]() is root) else torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
        K.cast_add_bf16(w.detach().contiguous(), None, wb)
        ent = (ck, wb, weakref.ref(root))
        _WCACHE[key] = ent
        if len(_WCACHE) > 4096:              # drop entries whose parameter is gone
            for k_ in [k_ for k_, e in _WCACHE.items() if e[2]() is None]:
                del _WCACHE[k_]
    return ent[1]


# ---- split-precision weights for the forward GEMMs -------------------------------------------------------------------------
# Weight rounding is the dominant error of a bf16 transformer forward against the fp32 reference: it is a fixed perturbation of
# the network shared by every token and frame, whereas activation rounding is independent noise that averages out (measured with
# the CPU oracle at cfg-5: bf16 weights alone -> pred_sted error 0.017, bf16 activations alone -> 0.006).  The forward GEMMs
# therefore read W as bf16 hi + bf16 lo ([N][2K] = [hi | lo], ~16 mantissa bits) through two reduction taps over the SAME
# activation tile: y = x W_hi^T + x W_lo^T.  Twice the tensor-core work on GEMMs that are launch- or HBM-bound anyway; backward
# (dgrad) keeps the plain bf16 copy.  TDB_WSPLIT=0 restores one-tap bf16 weights.
WSPLIT = _os.environ.get("TDB_WSPLIT", "1") != "0"
_WSCACHE = {}


def bf16_weight_split(w):
    """[N][2K] bf16 = [hi | lo] copy of an fp32 parameter [N][K], refreshed when the parameter changes"""
    key = (w.data_ptr(), tuple(w.shape))
    ck = (w._version,)
    root = _root(w)
    ent = _WSCACHE.get(key)
    if ent is None or ent[0] != ck or ent[2]() is not root:
        N, Kd = w.shape
        wb = ent[1] if (ent is not None and ent[2]() is root) else torch.empty(N, 2 * Kd, dtype=torch.bfloat16, device=w.device)
        K.split_bf16(w.detach().contiguous(), wb)
        ent = (ck, wb, weakref.ref(root))
        _WSCACHE[key] = ent
        if len(_WSCACHE) > 4096:
            for k_ in [k_ for k_, e in _WSCACHE.items() if e[2]() is None]:
                del _WSCACHE[k_]
    return ent[1]


def gemm_fwd_w(x, W, y, R, N, Kd, rows=None, **kw):
    """y = x @ W[rows]^T (+ epilogue): forward GEMM against an fp32 parameter, split-precision weights by default"""
    lo, hi = rows if rows is not None else (0, W.shape[0])
    if WSPLIT and Kd % 64 == 0:
        Ws = bf16_weight_split(W)
        return gemm(x, Ws[lo:hi], y, R, N, Kd, ntaps=2, a_off0=(0, 0), b_off0=(0, Kd), **kw)
    return gemm(x, bf16_weight(W)[lo:hi], y, R, N, Kd, **kw)


def adopt_bf16_weight(w, wb):
    """register an up-to-date bf16 copy of `w` produced elsewhere (optim.FusedAdamWEMA writes it in the optimizer kernel),
    so the next forward does not launch a cast kernel for it"""
    _WCACHE[(w.data_ptr(), tuple(w.shape))] = ((w._version,), wb, weakref.ref(_root(w)))


_DROP = {}          # device -> [seed tensor (int64, device), host site counter, generation (bumped with the seed)]


def _drop_state(device):
    st = _DROP.get(device)
    if st is None:
        seed = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFF], dtype=torch.int64, device=device)
        st = _DROP[device] = [seed, 0, 0]
    return st


def advance_dropout_seed(device):
    """once per training step (TubeDETR._encode): a captured in-place add, so every CUDA-graph replay draws new masks"""
    st = _drop_state(device)
    st[0].add_(1)
    st[2] += 1


def dropout_keep(shape, p, device):
    """uint8 keep mask of `shape` (1 = keep with probability 1 - p) from ONE kernel launch"""
    st = _drop_state(device)
    st[1] += 1
    keep = torch.empty(shape, dtype=torch.uint8, device=device)
    return K.dropout_mask(keep, st[0], st[1], p)


def _as_bf16(t):
    """bf16 view of a gradient.  Kernels that produce an fp32 gradient (LayerNorm backward) also write its bf16 copy and
    attach it as `_tdb_bf16`, so the consumer's GEMM operand needs no separate cast kernel."""
    if t.dtype == torch.bfloat16:
        return t
    side = getattr(t, "_tdb_bf16", None)
    if side is not None and side.shape == t.shape:
        return side
    return t.to(torch.bfloat16)


def wgrad_into(dy_b, x_b, out):
    """out[N,K] (fp32, contiguous) = dy_b[R,N]^T @ x_b[R,K]   (both operands read MN-major by the GEMM)."""
    R, N = dy_b.shape
    Kd = x_b.shape[1]
    tiles = ((N + 127) // 128) * max(1, Kd // 256 if Kd % 256 == 0 else Kd // 64)
    want = max(1, min((148 + tiles - 1) // tiles, 32))
    s = effective_splits(R, want)
    part = torch.empty(s, N, Kd, dtype=torch.float32, device=dy_b.device)
    gemm(dy_b, x_b, part, N, Kd, R, a_major=1, b_major=1, splits=want)
    splitk_reduce(part, s, N, Kd, out)
    return out


class LinearFn(torch.autograd.Function):
    """y = relu?(x @ W^T + b): x bf16 [R,K], W fp32 [N,K] (bf16 copy cached), y bf16 or fp32 [R,N]."""

    @staticmethod
    def forward(ctx, x, W, b, relu, out_fp32, mask_dx=False, masked_by_consumer=False, dx_scale=1.0):
        # mask_dx: x is a post-ReLU activation -> dgrad applies (x > 0) in its epilogue (ReLU backward of the producer);
        # masked_by_consumer: this op's own ReLU backward is done by the consumer that way;
        # dx_scale: constant factor on dx in the same epilogue (1 / (1 - p) when x went through HiddenDropoutFn).
        ctx.mask_dx, ctx.masked_by_consumer, ctx.dx_scale = mask_dx, masked_by_consumer, float(dx_scale)
        assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
        R, Kd = x.shape
        N = W.shape[0]
        y = torch.empty(R, N, dtype=torch.float32 if out_fp32 else torch.bfloat16, device=x.device)
        gemm_fwd_w(x, W, y, R, N, Kd, bias=b, relu=relu)
        ctx.relu = relu
        ctx.bias_ref = b if (b is not None and b.requires_grad) else None
        ctx.save_for_backward(x, W, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        Wb = bf16_weight(W)
        dyb = _as_bf16(dy)
        if ctx.relu and not ctx.masked_by_consumer:
            dyb = dyb * (y > 0)
        dyb = dyb.contiguous()
        R, Kd = x.shape
        N = W.shape[0]
        dx = dW = db = None
        sc = wgrad_scope(x.device)
        with sc:                                          # weight / bias gradients: side stream, concurrent with the dgrad below
            if ctx.needs_input_grad[1]:
                dW = wgrad_into(dyb, x, _grad_out(W))
            if ctx.needs_input_grad[2]:
                pre = getattr(dy, "_tdb_colsum", None)       # LayerNorm backward already summed this gradient over rows
                if pre is not None and pre.numel() == N and not (ctx.relu and not ctx.masked_by_consumer):
                    db = pre
                else:
                    db = K.colsum_bf16(dyb, _grad_out(ctx.bias_ref) if ctx.bias_ref is not None else torch.empty(N, dtype=torch.float32, device=x.device))
        if ctx.needs_input_grad[0]:
            dx = torch.empty(R, Kd, dtype=torch.bfloat16, device=x.device)
            gemm(dyb, Wb, dx, R, Kd, N, b_major=1, mask=x if ctx.mask_dx else None,
                 scale=_const_vec(x.device, Kd, ctx.dx_scale) if ctx.dx_scale != 1.0 else None)
        sc.join()
        return dx, dW, db, None, None, None, None, None


def linear(x, W, b, relu=False, out_fp32=False, mask_dx=False, masked_by_consumer=False, dx_scale=1.0):
    return LinearFn.apply(x, W, b, relu, out_fp32, mask_dx, masked_by_consumer, dx_scale)


_CONST = {}


def _const_vec(device, n, value):
    key = (device, n, value)
    if key not in _CONST:
        _CONST[key] = torch.full((n,), value, dtype=torch.float32, device=device)
    return _CONST[key]


class HiddenDropoutFn(torch.autograd.Function):
    """y = keep * x / (1 - p) on a post-ReLU bf16 activation (FFN hidden dropout).  ONE kernel, no mask tensor; the backward
    is the identity because the consuming linear does it in its dgrad epilogue (mask_dx=True: y > 0 <=> kept and ReLU-active;
    dx_scale = 1 / (1 - p)) and the producing linear is built with masked_by_consumer=True."""

    @staticmethod
    def forward(ctx, x, p):
        st = _drop_state(x.device)
        st[1] += 1
        x = x.contiguous()
        return K.dropout_bf16(x, torch.empty_like(x), st[0], st[1], p)

    @staticmethod
    def backward(ctx, dy):
        return dy, None


def hidden_dropout(x, p):
    return HiddenDropoutFn.apply(x, float(p))


class InProjFn(torch.autograd.Function):
    """Packed attention in-projection (torch MultiheadAttention in_proj_weight [3d,d], in_proj_bias [3d]).
    segs = ((lo, hi), ...) row ranges of W applied to xs[i]; returns one bf16 tensor per segment.
    Reference call sites: models/transformer.py:637-640, 698-719, 734-740 (q/k share an input, v differs)."""

    @staticmethod
    def forward(ctx, W, b, segs, *xs):
        outs = []
        for (lo, hi), x in zip(segs, xs):
            y = torch.empty(x.shape[0], hi - lo, dtype=torch.bfloat16, device=x.device)
            gemm_fwd_w(x, W, y, x.shape[0], hi - lo, x.shape[1], rows=(lo, hi), bias=b[lo:hi])
            outs.append(y)
        ctx.segs = segs
        ctx.bias_ref = b if b.requires_grad else None
        ctx.save_for_backward(W, *xs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dys):
        W, *xs = ctx.saved_tensors
        Wb = bf16_weight(W)
        # gradient rows that no segment (or an unused output) covers must read zero
        full = sum(hi - lo for (lo, hi), dy in zip(ctx.segs, dys) if dy is not None) == W.shape[0]
        dW = db = None
        if ctx.needs_input_grad[0]:
            dW = _grad_out(W)
            if not full:
                dW.zero_()
        if ctx.needs_input_grad[1]:
            bref = ctx.bias_ref
            db = _grad_out(bref) if bref is not None else torch.empty(W.shape[0], dtype=torch.float32, device=W.device)
            if not full:
                db.zero_()
        dxs = []
        sc = wgrad_scope(W.device)
        for i, ((lo, hi), x, dy) in enumerate(zip(ctx.segs, xs, dys)):
            if dy is None:
                dxs.append(None)
                continue
            dyb = _as_bf16(dy).contiguous()
            with sc:
                if dW is not None:
                    wgrad_into(dyb, x, dW[lo:hi])
                if db is not None:
                    K.colsum_bf16(dyb, db[lo:hi])
            if ctx.needs_input_grad[3 + i]:
                dx = torch.empty_like(x)
                gemm(dyb, Wb[lo:hi], dx, x.shape[0], x.shape[1], hi - lo, b_major=1)
                dxs.append(dx)
            else:
                dxs.append(None)
        sc.join()
        return (dW, db, None) + tuple(dxs)


def in_proj(W, b, segs, *xs):
    return InProjFn.apply(W, b, tuple(segs), *xs)


def _check_drop_gen(device, gen):
    """fused dropouts re-hash their keep bits from the DEVICE seed, which the next training forward bumps in place"""
    if _drop_state(device)[2] != gen:
        raise RuntimeError("tubedetr_b200: the dropout seed advanced between this forward and its backward (a second training "
                           "forward ran first); run backward before the next forward (one forward per backward, INTEGRATION.md)")


class MHAFn(torch.autograd.Function):
    """Attention core on projected q/k/v (bf16 [B*L, 256]); returns (o bf16 [B*Lq,256], pbar fp32 [B,Lq,Lk]).
    packed: q and k are the two halves of one [R,512] tensor (self-attention).  drop_p > 0 applies torch's
    attention-probability dropout inside the kernel (pbar averages the dropped P): on the tcgen05 path the keep bits come from the
    counter-based hash stream (no mask tensor, the backward regenerates them); the CUDA-core path takes an explicit mask drawn
    from the same stream."""

    @staticmethod
    def forward(ctx, q, k, v, kpm, B, H, Lq, Lk, scale, packed, drop_p, need_weights=True):
        if packed:
            qv, kv = q[:, :256], q[:, 256:]
        else:
            qv, kv = q, k
        o = torch.empty(B * Lq, H * 32, dtype=torch.bfloat16, device=q.device)
        p = torch.empty(B, H, Lq, Lk, dtype=torch.float32, device=q.device)
        # the encoder never reads its attention weights (reference transformer.py:638-640 discards them): no head-mean launch and,
        # under dropout, no second probability tensor
        pbar = torch.empty(B, Lq, Lk, dtype=torch.float32, device=q.device) if need_weights else None
        keep = drop = None
        tc = K.mha_uses_tc(H, Lq, Lk, 1)
        if drop_p > 0:
            st = _drop_state(q.device)
            st[1] += 1
            drop = (st[0], st[1], float(drop_p))
            ctx.drop_gen = st[2]
        pdrop = torch.empty_like(p) if (drop is not None and need_weights) else None
        if tc:
            K.mha_tc_fwd(qv, kv, v, kpm, o, p, pbar, B, H, Lq, Lk, scale, drop=drop, pdrop=pdrop)
        else:
            if drop is not None:
                keep = K.dropout_mask(torch.empty((B, H, Lq, Lk), dtype=torch.uint8, device=q.device), *drop)
            K.mha_fwd_cuda_core(qv, kv, v, kpm, o, p, pbar, B, H, Lq, Lk, scale, keep=keep, pdrop=pdrop, keep_scale=1.0 / (1.0 - drop_p))
        ctx.cfg = (B, H, Lq, Lk, scale, packed, drop_p)
        ctx.drop = drop
        ctx.save_for_backward(q, k if not packed else None, v, p, keep)
        if pbar is None:
            pbar = p.new_empty(0)
            ctx.mark_non_differentiable(pbar)
        return o, pbar

    @staticmethod
    def backward(ctx, do, dpbar):
        q, k, v, p, keep = ctx.saved_tensors
        B, H, Lq, Lk, scale, packed, drop_p = ctx.cfg
        if do is None:
            do = torch.zeros(B * Lq, H * 32, dtype=torch.bfloat16, device=q.device)
        do = _as_bf16(do).contiguous()
        if dpbar is not None and dpbar.numel() == 0:      # need_weights=False: placeholder output
            dpbar = None
        if dpbar is not None:
            dpbar = dpbar.contiguous().float()
        if ctx.drop is not None:
            _check_drop_gen(q.device, ctx.drop_gen)
        dv = torch.empty_like(v)
        if packed:
            dqk = torch.empty_like(q)
            qv, kv, dq, dk = q[:, :256], q[:, 256:], dqk[:, :256], dqk[:, 256:]
        else:
            qv, kv, dq, dk = q, k, torch.empty_like(q), torch.empty_like(k)
        if K.mha_uses_tc(H, Lq, Lk, 2):     # tcgen05 backward keeps dS / dropped P in shared memory: no fp32 scratch tensors
            K.mha_tc_bwd(qv, kv, v, do, p, dpbar, dq, dk, dv, B, H, Lq, Lk, scale, drop=ctx.drop)
        else:
            if ctx.drop is not None and keep is None:       # forward ran on tcgen05 with hash dropout: materialise the same bits
                keep = K.dropout_mask(torch.empty((B, H, Lq, Lk), dtype=torch.uint8, device=q.device), *ctx.drop)
            K.mha_bwd_cuda_core(qv, kv, v, do, p, dpbar, torch.empty_like(p), dq, dk, dv, B, H, Lq, Lk, scale, keep=keep,
                                keep_scale=1.0 / (1.0 - drop_p), pd_scratch=torch.empty_like(p) if keep is not None else None)
        if packed:
            return (dqk, None, dv) + (None,) * 9
        return (dq, dk, dv) + (None,) * 9


def mha(q, k, v, kpm, B, H, Lq, Lk, scale, packed=False, drop_p=0.0, need_weights=True):
    return MHAFn.apply(q, k, v, kpm, B, H, Lq, Lk, scale, packed, float(drop_p), bool(need_weights))


class XAttnFusedFn(torch.autograd.Function):
    """Time-aligned cross-attention with the K/V projections fused into the attention kernel (tcgen05 + TMEM):
    q [F,256] bf16 (already projected), mempb/memb [F*S,256] bf16, W/b = packed in_proj (rows 256:768 used here).
    Returns (o bf16 [F,256], pbar fp32 [F,1,S]).  Backward re-projects K/V with tdb_gemm (they were never stored)."""

    @staticmethod
    def forward(ctx, q, mempb, memb, W, b, kpm, F, S, scale, drop_p):
        Wb = bf16_weight(W)
        o = torch.empty(F, 256, dtype=torch.bfloat16, device=q.device)
        p = torch.empty(F, 8, 1, S, dtype=torch.float32, device=q.device)
        pbar = torch.empty(F, 1, S, dtype=torch.float32, device=q.device)
        keep = dropout_keep((F, 8, 1, S), drop_p, q.device) if drop_p > 0 else None
        K.xattn_fused_fwd(q, mempb, memb, Wb[256:768], b[512:768], kpm, o, p, pbar, F, S, scale, keep=keep,
                          keep_scale=1.0 / (1.0 - drop_p))
        ctx.cfg = (F, S, scale, drop_p)
        ctx.save_for_backward(q, mempb, memb, W, b, p, keep)
        return o, pbar

    @staticmethod
    def backward(ctx, do, dpbar):
        q, mempb, memb, W, b, p, keep = ctx.saved_tensors
        F, S, scale, drop_p = ctx.cfg
        Wb = bf16_weight(W)
        R = F * S
        kp = torch.empty(R, 256, dtype=torch.bfloat16, device=q.device)
        vp = torch.empty(R, 256, dtype=torch.bfloat16, device=q.device)
        gemm(mempb, Wb[256:512], kp, R, 256, 256, bias=b[256:512])
        gemm(memb, Wb[512:768], vp, R, 256, 256, bias=b[512:768])
        if do is None:
            do = torch.zeros(F, 256, dtype=torch.bfloat16, device=q.device)
        do = _as_bf16(do).contiguous()
        dpbar = dpbar.contiguous().float() if dpbar is not None else None
        dq, dk, dv = torch.empty_like(q), torch.empty_like(kp), torch.empty_like(vp)
        K.xattn_bwd(q.contiguous(), kp, vp, do, p, dpbar, dq, dk, dv, F, S, scale, keep=keep, keep_scale=1.0 / (1.0 - drop_p))
        dW = torch.zeros_like(W) if ctx.needs_input_grad[3] else None
        db = torch.zeros(W.shape[0], dtype=torch.float32, device=W.device) if ctx.needs_input_grad[4] else None
        sc = wgrad_scope(q.device)
        with sc:
            if dW is not None:
                wgrad_into(dk, mempb, dW[256:512])
                wgrad_into(dv, memb, dW[512:768])
            if db is not None:
                K.colsum_bf16(dk, db[256:512])
                K.colsum_bf16(dv, db[512:768])
        dmp = dmb = None
        if ctx.needs_input_grad[1]:
            dmp = torch.empty_like(mempb)
            gemm(dk, Wb[256:512], dmp, R, 256, 256, b_major=1)
        if ctx.needs_input_grad[2]:
            dmb = torch.empty_like(memb)
            gemm(dv, Wb[512:768], dmb, R, 256, 256, b_major=1)
        sc.join()
        return dq, dmp, dmb, dW, db, None, None, None, None, None


def xattn_fused(q, mempb, memb, W, b, kpm, F, S, scale, drop_p=0.0):
    return XAttnFusedFn.apply(q, mempb, memb, W, b, kpm, F, S, scale, float(drop_p))


# ---- decoder cross-attention with the K / V projections of ALL layers hoisted out of the layer loop -------------------------------
# Every decoder layer projects the SAME memory (reference models/transformer.py:567-579 passes memory / pos unchanged to each layer;
# :734-740 key = memory + pos, value = memory), so K and V of all layers are two GEMMs [F*S, 256] x [layers*256, 256]^T instead of
# 2 x layers launches on 111-tile grids, and they are SAVED: the backward needs no re-projection, the data gradients of the six
# layers reach `memory` through two GEMMs with a layers*256-deep reduction (instead of 12 GEMMs + 10 accumulation passes over
# 7 MB tensors), and each layer's attention core reads / writes its 256-column slice of the [F*S, layers*256] buffers in place.
class DecKVShared:
    """per-forward state shared by DecoderKVFn and the per-layer XAttnCoreFn nodes (gradient buffers written slice by slice)"""

    def __init__(self):
        self.dK = self.dV = self.dW = self.db = self.db_kv = None


_DECKV_CACHE = {}


def _deckv_weights(Ws, bs):
    """split-precision [layers*256][512] copies of the key / value rows of the layers' in_proj_weight + concatenated biases"""
    key = tuple(w.data_ptr() for w in Ws)
    ck = tuple(w._version for w in Ws) + tuple(b._version for b in bs)
    ent = _DECKV_CACHE.get(key)
    if ent is None or ent[0] != ck or any(r() is not _root(w) for r, w in zip(ent[2], Ws)):
        nl, d = len(Ws), Ws[0].shape[1]
        if ent is not None and ent[1][0].shape[0] == nl * d:
            Wk, Wv = ent[1][0], ent[1][1]
        else:
            Wk = torch.empty(nl * d, 2 * d, dtype=torch.bfloat16, device=Ws[0].device)
            Wv = torch.empty_like(Wk)
        for l, w in enumerate(Ws):
            wd = w.detach()
            K.split_bf16(wd[d:2 * d], Wk[l * d:(l + 1) * d])
            K.split_bf16(wd[2 * d:], Wv[l * d:(l + 1) * d])
        bk = torch.cat([b.detach()[d:2 * d] for b in bs]).contiguous()
        bv = torch.cat([b.detach()[2 * d:] for b in bs]).contiguous()
        ent = (ck, (Wk, Wv, bk, bv), [weakref.ref(_root(w)) for w in Ws])
        _DECKV_CACHE.clear()
        _DECKV_CACHE[key] = ent
    return ent[1]


class DecoderKVFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mempb, memb, shared, nl, *wb):
        Ws, bs = wb[:nl], wb[nl:]
        Wk, Wv, bk, bv = _deckv_weights(Ws, bs)
        R, d = mempb.shape
        K_all = torch.empty(R, nl * d, dtype=torch.bfloat16, device=mempb.device)
        V_all = torch.empty_like(K_all)
        gemm(mempb, Wk, K_all, R, nl * d, d, ntaps=2, a_off0=(0, 0), b_off0=(0, d), bias=bk)
        gemm(memb, Wv, V_all, R, nl * d, d, ntaps=2, a_off0=(0, 0), b_off0=(0, d), bias=bv)
        tok = torch.empty(1, dtype=torch.float32, device=mempb.device)     # gradient token: orders this node's backward after the layers'
        ctx.shared, ctx.nl = shared, nl
        ctx.wk, ctx.wv = Wk, Wv              # cache-owned bf16 copies: valid until the weights change, i.e. beyond this step's backward
        ctx.save_for_backward(mempb, memb)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(K_all, V_all)
        return K_all, V_all, tok

    @staticmethod
    def backward(ctx, _gk, _gv, _gtok):
        mempb, memb = ctx.saved_tensors
        sh, nl = ctx.shared, ctx.nl
        if sh.dK is None:
            return (None,) * (4 + 2 * nl)
        Wk, Wv = ctx.wk, ctx.wv
        R, d = mempb.shape
        for l in range(nl):            # a layer whose cross-attention received no gradient (never in the reference path): zero q rows
            if sh.dW[l] is None:
                sh.dW[l] = torch.zeros(3 * d, d, dtype=torch.float32, device=mempb.device)
                sh.db[l] = torch.zeros(3 * d, dtype=torch.float32, device=mempb.device)
        sc = wgrad_scope(mempb.device)
        with sc:
            for l in range(nl):
                wgrad_into(sh.dK[:, l * d:(l + 1) * d], mempb, sh.dW[l][d:2 * d])
                wgrad_into(sh.dV[:, l * d:(l + 1) * d], memb, sh.dW[l][2 * d:])
            K.colsum_bf16(sh.dK, sh.db_kv[0])
            K.colsum_bf16(sh.dV, sh.db_kv[1])
        dmp = dmb = None
        if ctx.needs_input_grad[0]:
            dmp = torch.empty_like(mempb)
            gemm(sh.dK, Wk[:, :d], dmp, R, d, nl * d, b_major=1)
        if ctx.needs_input_grad[1]:
            dmb = torch.empty_like(memb)
            gemm(sh.dV, Wv[:, :d], dmb, R, d, nl * d, b_major=1)
        sc.join()
        for l in range(nl):                                           # tiny strided copies: bias gradients of the k / v rows
            sh.db[l][d:2 * d].copy_(sh.db_kv[0][l * d:(l + 1) * d])
            sh.db[l][2 * d:].copy_(sh.db_kv[1][l * d:(l + 1) * d])
        return (dmp, dmb, None, None) + tuple(sh.dW[l] for l in range(nl)) + tuple(sh.db[l] for l in range(nl))


def decoder_kv(mempb, memb, Ws, bs):
    """-> (K_all, V_all [F*S, layers*256] bf16, token, shared)"""
    shared = DecKVShared()
    K_all, V_all, tok = DecoderKVFn.apply(mempb, memb, shared, len(Ws), *Ws, *bs)
    return K_all, V_all, tok, shared


class XAttnCoreFn(torch.autograd.Function):
    """One decoder layer's time-aligned cross-attention on the hoisted K / V: q-projection (rows 0:256 of in_proj) + the one-query
    attention core.  W / b receive their gradients from DecoderKVFn (one gradient tensor per parameter, filled by both nodes)."""

    @staticmethod
    def forward(ctx, xq, K_all, V_all, tok, W, b, kpm, shared, layer, nl, F, S, scale, drop_p):
        d = xq.shape[1]
        q = torch.empty(F, d, dtype=torch.bfloat16, device=xq.device)
        gemm_fwd_w(xq, W, q, F, d, d, rows=(0, d), bias=b[:d])
        o = torch.empty(F, d, dtype=torch.bfloat16, device=xq.device)
        p = torch.empty(F, 8, 1, S, dtype=torch.float32, device=xq.device)
        pbar = torch.empty(F, 1, S, dtype=torch.float32, device=xq.device)
        keep = dropout_keep((F, 8, 1, S), drop_p, xq.device) if drop_p > 0 else None
        Kl, Vl = K_all[:, layer * d:(layer + 1) * d], V_all[:, layer * d:(layer + 1) * d]
        K.xattn_core_fwd(q, Kl, Vl, kpm, o, p, pbar, F, S, scale, keep=keep, keep_scale=1.0 / (1.0 - drop_p))
        ctx.cfg = (shared, layer, nl, F, S, scale, drop_p, tok is not None)
        ctx.bias_ref = b
        ctx.save_for_backward(xq, q, K_all, V_all, W, p, keep)
        return o, pbar

    @staticmethod
    def backward(ctx, do, dpbar):
        xq, q, K_all, V_all, W, p, keep = ctx.saved_tensors
        sh, layer, nl, F, S, scale, drop_p, has_tok = ctx.cfg
        d = xq.shape[1]
        dev = xq.device
        if sh.dK is None:            # first layer to run its backward (the last decoder layer) allocates the shared gradient buffers
            sh.dK, sh.dV = torch.empty_like(K_all), torch.empty_like(V_all)
            sh.dW, sh.db = [None] * nl, [None] * nl
            sh.db_kv = torch.empty(2, nl * d, dtype=torch.float32, device=dev)
        if sh.dW[layer] is None:
            sh.dW[layer], sh.db[layer] = _grad_out(W), _grad_out(ctx.bias_ref)
        if do is None:
            do = torch.zeros(F, d, dtype=torch.bfloat16, device=dev)
        do = _as_bf16(do).contiguous()
        dpbar = dpbar.contiguous().float() if dpbar is not None else None
        dq = torch.empty_like(q)
        sl = slice(layer * d, (layer + 1) * d)
        K.xattn_core_bwd(q, K_all[:, sl], V_all[:, sl], do, p, dpbar, dq, sh.dK[:, sl], sh.dV[:, sl], F, S, scale, keep=keep,
                         keep_scale=1.0 / (1.0 - drop_p))
        sc = wgrad_scope(dev)
        with sc:
            wgrad_into(dq, xq, sh.dW[layer][:d])
            K.colsum_bf16(dq, sh.db[layer][:d])
        dxq = torch.empty_like(xq)
        gemm(dq, bf16_weight(W)[:d], dxq, F, d, d, b_major=1)
        sc.join()
        gtok = dq.new_empty(1, dtype=torch.float32) if has_tok else None
        return (dxq, None, None, gtok) + (None,) * 10


def xattn_core(xq, K_all, V_all, tok, W, b, kpm, shared, layer, nl, F, S, scale, drop_p=0.0):
    return XAttnCoreFn.apply(xq, K_all, V_all, tok, W, b, kpm, shared, layer, nl, F, S, scale, float(drop_p))


class AddLayerNormFn(torch.autograd.Function):
    """y = LayerNorm(x + dropout_p(r)) over d=256 (fp32 statistics).  Returns (y fp32, y bf16, (y + pos) bf16 or None).
    drop_p > 0: the residual dropout of the reference (`src + self.dropoutN(src2)`, transformer.py:641-645, 721-750) runs inside
    the LayerNorm kernels from a counter-based hash (forward and backward regenerate the same keep bits; no mask tensor)."""

    @staticmethod
    def forward(ctx, x, r, gamma, beta, pos, eps, drop_p=0.0):
        assert x.dtype == torch.float32 and x.is_contiguous()
        rows, D = x.shape
        if r is not None:
            r = r.contiguous()
            assert r.dtype == torch.float32
        y = torch.empty_like(x)
        yb = torch.empty(rows, D, dtype=torch.bfloat16, device=x.device)
        ypb = torch.empty_like(yb) if pos is not None else None
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        drop = None
        if drop_p > 0 and r is not None:
            st = _drop_state(x.device)
            st[1] += 1
            drop = (st[0], st[1], float(drop_p))
            ctx.drop_gen = st[2]
        K.layernorm_fwd(x, r, gamma, beta, pos, y, yb, ypb, mean, rstd, rows, D, eps, drop=drop)
        ctx.save_for_backward(x, r, gamma, mean, rstd)
        ctx.has_r = r is not None
        ctx.drop = drop
        ctx.pos_dtype = pos.dtype if pos is not None else None
        if ypb is None:
            return y, yb
        return y, yb, ypb

    @staticmethod
    def backward(ctx, dy, dyb, dypb=None):
        x, r, gamma, mean, rstd = ctx.saved_tensors
        rows, D = x.shape
        if dy is None and dyb is None and dypb is None:
            return None, None, None, None, None, None, None
        if ctx.drop is not None:
            _check_drop_gen(x.device, ctx.drop_gen)
        dy = dy.contiguous() if dy is not None else None
        dyb = dyb.contiguous() if dyb is not None else None
        dypb = dypb.contiguous() if dypb is not None else None
        dz = torch.empty_like(x)
        dzb = torch.empty(rows, D, dtype=torch.bfloat16, device=x.device) if (ctx.has_r and ctx.drop is None) else None
        # contiguous [dgamma | dbeta | column sums of r's gradient]: one reduction launch; the third block IS the bias gradient of
        # the linear layer that produced r (out_proj / linear2), handed over through the `_tdb_colsum` side channel
        dgb = torch.empty(3 * D if ctx.has_r else 2 * D, dtype=torch.float32, device=x.device)
        dg, db = dgb[:D], dgb[D:2 * D]
        dbias = dgb[2 * D:] if ctx.has_r else None
        dr = drb = None
        if ctx.drop is not None:
            dr, drb = torch.empty_like(x), torch.empty(rows, D, dtype=torch.bfloat16, device=x.device)
        K.layernorm_bwd(dy, x, r, gamma, mean, rstd, dz, dg, db, rows, D, dy2=dyb, dy3=dypb, dz_bf=dzb,   # sums the three grads in-kernel
                        drop=ctx.drop, dr=dr, dr_bf=drb, dbias=dbias)
        if dzb is not None:
            dz._tdb_bf16 = dzb
        if dr is not None:
            dr._tdb_bf16 = drb
            dr._tdb_colsum = dbias
        elif ctx.has_r:
            dz._tdb_colsum = dbias
        dpos = None
        if ctx.needs_input_grad[4] and dypb is not None:   # (y + pos): pos is the time-query embedding in the decoder
            dpos = dypb.to(ctx.pos_dtype)
        gr = (dr if dr is not None else dz) if ctx.has_r else None
        return dz, gr, dg, db, dpos, None, None


def add_layernorm(x, r, gamma, beta, pos=None, eps=1e-5, drop_p=0.0):
    return AddLayerNormFn.apply(x, r, gamma, beta, pos, eps, float(drop_p))


class BackboneJointFn(torch.autograd.Function):
    """Slow (with grad) and fast (no grad) frames through the backbone as ONE batch (reference models/tubedetr.py:120-131
    runs them as two calls; the fast call is under no_grad).  Returns (feat_slow, feat_fast); only feat_slow is
    differentiable, and backward runs on the slow frames' row prefix of the saved activations."""

    @staticmethod
    def forward(ctx, frames, frames_fast, engine, W, names, tag, *params):
        ns = frames.shape[0]
        feat, h, w, bctx = engine.forward([frames, frames_fast], W, save=True, tag=tag, n_keep=ns)
        fs, ff = feat[:ns * h * w], feat[ns * h * w:]
        ctx.engine, ctx.W, ctx.names, ctx.bctx = engine, W, names, bctx
        ctx.save_for_backward(fs, *params)
        ctx.mark_non_differentiable(ff)
        return fs, ff

    @staticmethod
    def backward(ctx, g, _g_fast):
        feat, *params = ctx.saved_tensors
        g = (_as_bf16(g) * (feat > 0)).contiguous()
        grads = {n: _grad_out(p) for n, p in zip(ctx.names, params)}
        ctx.engine.backward(ctx.bctx, ctx.W, g, grads)
        return (None, None, None, None, None, None) + tuple(grads[n] for n in ctx.names)


class BackboneFn(torch.autograd.Function):
    """ResNet-101 layer4 features WITH grad for layer2-4 conv weights (reference models/backbone.py:82-89)."""

    @staticmethod
    def forward(ctx, frames, engine, W, names, tag, *params):
        feat, h, w, bctx = engine.forward(frames, W, save=True, tag=tag)
        feat = feat[:]          # a fresh tensor object per call: the engine hands out the SAME cached buffer view every time
        ctx.engine, ctx.W, ctx.names, ctx.bctx = engine, W, names, bctx
        ctx.save_for_backward(feat, *params)
        ctx.hw = (h, w)
        return feat

    @staticmethod
    def backward(ctx, g):
        feat, *params = ctx.saved_tensors
        g = (_as_bf16(g) * (feat > 0)).contiguous()
        grads = {n: _grad_out(p) for n, p in zip(ctx.names, params)}
        ctx.engine.backward(ctx.bctx, ctx.W, g, grads)
        return (None, None, None, None, None) + tuple(grads[n] for n in ctx.names)


# ---- fused data movement around the encoder (tdb_glue.cu) ----------------------------------------------------------------------
def _cf32(t):
    return t if (t is None or (t.dtype == torch.float32 and t.is_contiguous())) else t.float().contiguous()


def _cbf(t):
    return t if (t is None or (t.dtype == torch.bfloat16 and t.is_contiguous())) else t.to(torch.bfloat16).contiguous()


class EncAssembleFn(torch.autograd.Function):
    """[image tokens | per-clip repeated text tokens] -> x32 fp32, bf16(x32), bf16(x32 + pe), pe  (one kernel; reference
    models/transformer.py:269-331 builds these with repeat / stack / cat loops)"""

    @staticmethod
    def forward(ctx, src, txt, pos, n_clips):
        n, HW, D = src.shape
        L = txt.shape[1]
        S = HW + L
        dev = src.device
        x32 = torch.empty(n * S, D, dtype=torch.float32, device=dev)
        pe = torch.empty_like(x32)
        xb = torch.empty(n * S, D, dtype=torch.bfloat16, device=dev)
        xpb = torch.empty_like(xb)
        K.enc_assemble_fwd(_cf32(src), _cf32(txt), _cf32(pos), x32, xb, xpb, pe, n, HW, L, n_clips)
        ctx.dims = (n, HW, L, n_clips, txt.shape[0])
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(pe)
        return x32, xb, xpb, pe

    @staticmethod
    def backward(ctx, g32, gb, gpb, _gpe):
        n, HW, L, n_clips, B = ctx.dims
        ref = g32 if g32 is not None else (gb if gb is not None else gpb)
        if ref is None:
            return None, None, None, None
        dsrc = torch.empty(n, HW, 256, dtype=torch.float32, device=ref.device)
        dtxt = torch.empty(B, L, 256, dtype=torch.float32, device=ref.device)
        K.enc_assemble_bwd(_cf32(g32), _cbf(gb), _cbf(gpb), dsrc, dtxt, n, HW, L, n_clips)
        return dsrc, dtxt, None, None


def enc_assemble(src, txt, pos, n_clips):
    return EncAssembleFn.apply(src, txt, pos, int(n_clips))


class FastMixFn(torch.autograd.Function):
    """z = bf16(enc[clip(t)] + fm) on the image rows: the input of fast_residual (reference models/transformer.py:373-391)"""

    @staticmethod
    def forward(ctx, enc, fm, B, T, k, HW, S):
        z = torch.empty(B * T * HW, 256, dtype=torch.bfloat16, device=enc.device)
        K.fast_mix_fwd(_cf32(enc), _cbf(fm), z, B, T, k, HW, S)
        ctx.dims = (B, T, k, HW, S, enc.shape)
        return z

    @staticmethod
    def backward(ctx, dz):
        B, T, k, HW, S, eshape = ctx.dims
        dz = _cbf(dz)
        denc = torch.empty(eshape, dtype=torch.float32, device=dz.device)
        K.fast_mix_bwd(dz, denc, B, T, k, HW, S)
        return denc, dz.view(B * T * HW, 256), None, None, None, None, None


def fast_mix(enc, fm, B, T, k, HW, S):
    return FastMixFn.apply(enc, fm, B, T, k, HW, S)


class AggregateFn(torch.autograd.Function):
    """temporal replication + fast-branch aggregation + the decoder's bf16 memory operands in one pass (reference
    models/transformer.py:393-446): -> mem fp32 [B*T*S, 256], mem_pos, bf16(mem), bf16(mem + mem_pos)"""

    @staticmethod
    def forward(ctx, enc, pe, upd, B, T, k, HW, S):
        dev = enc.device
        mem = torch.empty(B * T * S, 256, dtype=torch.float32, device=dev)
        mem_pos = torch.empty_like(mem)
        memb = torch.empty(B * T * S, 256, dtype=torch.bfloat16, device=dev)
        mempb = torch.empty_like(memb)
        K.aggregate_fwd(_cf32(enc), _cf32(pe), _cf32(upd), mem, mem_pos, memb, mempb, B, T, k, HW, S)
        ctx.dims = (B, T, k, HW, S, enc.shape, upd is not None)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(mem_pos)
        return mem, mem_pos, memb, mempb

    @staticmethod
    def backward(ctx, gmem, _gpos, gmemb, gmempb):
        B, T, k, HW, S, eshape, has_upd = ctx.dims
        ref = gmem if gmem is not None else (gmemb if gmemb is not None else gmempb)
        if ref is None:
            return (None,) * 8
        dev = ref.device
        denc = torch.empty(eshape, dtype=torch.float32, device=dev)
        dupd = dupd_b = None
        if has_upd and ctx.needs_input_grad[2]:
            dupd = torch.empty(B * T * HW, 256, dtype=torch.float32, device=dev)
            dupd_b = torch.empty(B * T * HW, 256, dtype=torch.bfloat16, device=dev)
        K.aggregate_bwd(_cf32(gmem), _cbf(gmemb), _cbf(gmempb), denc, dupd, dupd_b, B, T, k, HW, S)
        if dupd is not None:
            dupd._tdb_bf16 = dupd_b
        return denc, None, dupd, None, None, None, None, None


def aggregate(enc, pe, upd, B, T, k, HW, S):
    return AggregateFn.apply(enc, pe, upd, B, T, k, HW, S)


class HeadOutFn(torch.autograd.Function):
    """last layer of a prediction head fused with its activation / logit dropout (tdb_glue.cu); x is the bf16 output of the previous
    layer's GEMM (post-ReLU, possibly after hidden dropout): with mask_dx its ReLU / dropout backward happens here"""

    @staticmethod
    def forward(ctx, x, W, b, act, drop_p, mask_dx, dx_scale):
        x = _cbf(x)
        R = x.shape[0]
        J = W.shape[0]
        y = torch.empty(R, J, dtype=torch.float32, device=x.device)
        drop = None
        if drop_p > 0:
            st = _drop_state(x.device)
            st[1] += 1
            drop = (st[0], st[1], float(drop_p))
            ctx.drop_gen = st[2]
        K.head_out_fwd(x, W.detach().float().contiguous(), b.detach().float().contiguous(), y, act, drop=drop)
        ctx.cfg = (act, drop, mask_dx, float(dx_scale))
        ctx.save_for_backward(x, W, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        act, drop, mask_dx, dx_scale = ctx.cfg
        if drop is not None:
            _check_drop_gen(x.device, ctx.drop_gen)
        R, J = y.shape
        dpre = torch.empty(R, J, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        dW = torch.empty(J, 256, dtype=torch.float32, device=x.device)
        db = torch.empty(J, dtype=torch.float32, device=x.device)
        K.head_out_bwd(dy.contiguous().float(), y, x, W.detach().float().contiguous(), dpre, dx, dW, db, act, mask_dx, dx_scale, drop=drop)
        return dx, dW, db, None, None, None, None


def head_out(x, W, b, act=0, drop_p=0.0, mask_dx=False, dx_scale=1.0):
    return HeadOutFn.apply(x, W, b, int(act), float(drop_p), bool(mask_dx), float(dx_scale))
