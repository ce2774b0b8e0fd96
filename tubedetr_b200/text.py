"""The text encoder (RoBERTa-base) on this library's kernels.

Reference: models/transformer.py:130-135, 250-263 -- `RobertaModel.from_pretrained(...)` called as a library; SURVEY.md 8(f).1.  The HF
module stays the PARAMETER CONTAINER (state_dict names `transformer.text_encoder.*` unchanged, checkpoints load as before); its forward
is restated here on libtdb.so: at 20 tokens every linear layer is a one-tile tcgen05 GEMM over split-precision weights, the weight
gradients are rank-L outer-product sums (tdb_skinny_wgrad: one launch for dW and db), attention (12 heads of 64) and erf-GELU are small
dedicated kernels, LayerNorm + residual + dropout is the d = 768 instantiation of the encoder's kernel.  ~125 launches forward and
~220 backward instead of ~700 library launches, all bf16 operands with fp32 accumulation / statistics.
Embedding lookups (three gathers + two adds) stay torch ops: they are index operations on a 154 MB table, not arithmetic.
"""
import torch
import torch.nn.functional as F

from . import kernels as K
from . import ops
from .gemm import gemm

MAX_TOKENS = 96        # tdb_text_attn_bwd keeps Q, K, V, dO and two L x L tiles in shared memory


class SkinnyLinearFn(torch.autograd.Function):
    """y = x W^T + b for a handful of rows (tokens): forward / dgrad on tdb_gemm, dW + db in ONE CUDA-core launch."""

    @staticmethod
    def forward(ctx, x, W, b, out_fp32):
        R, Kd = x.shape
        N = W.shape[0]
        y = torch.empty(R, N, dtype=torch.float32 if out_fp32 else torch.bfloat16, device=x.device)
        ops.gemm_fwd_w(x, W, y, R, N, Kd, bias=b)
        ctx.save_for_backward(x, W)
        ctx.bias_ref = b
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        R, Kd = x.shape
        N = W.shape[0]
        dyb = ops._as_bf16(dy).contiguous()
        dW = db = dx = None
        sc = ops.wgrad_scope(x.device)
        with sc:
            if ctx.needs_input_grad[1]:
                dW = ops._grad_out(W)
                db = ops._grad_out(ctx.bias_ref) if ctx.needs_input_grad[2] else None
                K.skinny_wgrad(dyb, x, dW, db)
        if ctx.needs_input_grad[0]:
            dx = torch.empty(R, Kd, dtype=torch.bfloat16, device=x.device)
            Wb = ops.bf16_weight_split(W)[:, :Kd] if ops.WSPLIT else ops.bf16_weight(W)      # hi half of the split copy: no second copy
            gemm(dyb, Wb, dx, R, Kd, N, b_major=1)
        sc.join()
        return dx, dW, db, None


def slinear(x, W, b, out_fp32=False):
    return SkinnyLinearFn.apply(x, W, b, bool(out_fp32))


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return K.gelu_fwd(x, torch.empty_like(x))

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return K.gelu_bwd(ops._as_bf16(dy).contiguous(), x, torch.empty_like(x))


class TextAttnFn(torch.autograd.Function):
    """softmax(q k^T / 8 + key mask) v for 12 heads of 64 over L <= 96 tokens; attention dropout from the hash stream"""

    @staticmethod
    def forward(ctx, q, k, v, kpm, B, H, L, drop_p):
        o = torch.empty(B * L, H * 64, dtype=torch.bfloat16, device=q.device)
        p = torch.empty(B, H, L, L, dtype=torch.float32, device=q.device)
        drop = None
        if drop_p > 0:
            st = ops._drop_state(q.device)
            st[1] += 1
            drop = (st[0], st[1], float(drop_p))
            ctx.drop_gen = st[2]
        K.text_attn_fwd(q, k, v, kpm, o, p, B, H, L, 0.125, drop=drop)
        ctx.cfg = (B, H, L, drop)
        ctx.save_for_backward(q, k, v, p)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, p = ctx.saved_tensors
        B, H, L, drop = ctx.cfg
        if drop is not None:
            ops._check_drop_gen(q.device, ctx.drop_gen)
        do = ops._as_bf16(do).contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        K.text_attn_bwd(q, k, v, do, p, dq, dk, dv, B, H, L, 0.125, drop=drop)
        return dq, dk, dv, None, None, None, None, None


_POS_CACHE = {}


def _position_ids(ids, pad):
    """HF create_position_ids_from_input_ids: cumulative count of non-pad tokens + pad index (cached per id tensor version)"""
    key = (ids.data_ptr(), tuple(ids.shape))
    ent = _POS_CACHE.get(key)
    if ent is None or ent[0] != ids._version:
        m = ids.ne(pad).int()
        pos = (torch.cumsum(m, 1).type_as(m) * m).long() + pad
        if len(_POS_CACHE) > 16:
            _POS_CACHE.clear()
        ent = _POS_CACHE[key] = (ids._version, pos)
    return ent[1]


def supported(mod, ids):
    cfg = mod.config
    return (ids.is_cuda and ids.shape[1] <= MAX_TOKENS and cfg.hidden_size == 768 and cfg.num_attention_heads == 12
            and cfg.hidden_act == "gelu" and getattr(cfg, "position_embedding_type", "absolute") == "absolute")


def roberta_forward(mod, ids, am, training):
    """last_hidden_state of HF `RobertaModel` `mod` (B, L, 768): -> (fp32 rows [B*L, 768], their bf16 copy)"""
    B, L = ids.shape
    emb = mod.embeddings
    cfg = mod.config
    dp = float(cfg.hidden_dropout_prob) if training else 0.0
    adp = float(cfg.attention_probs_dropout_prob) if training else 0.0
    pos = _position_ids(ids, emb.padding_idx)
    x = emb.word_embeddings(ids) + emb.token_type_embeddings.weight[0] + emb.position_embeddings(pos)
    ln = emb.LayerNorm
    x32, xb = ops.add_layernorm(x.reshape(B * L, 768).float().contiguous(), None, ln.weight, ln.bias, eps=ln.eps)
    if dp > 0:
        x32 = F.dropout(x32, dp, True)
        xb = x32.to(torch.bfloat16)
    kpm = am.eq(0).to(torch.uint8).contiguous()
    for lyr in mod.encoder.layer:
        at, sa = lyr.attention, lyr.attention.self
        q = slinear(xb, sa.query.weight, sa.query.bias)
        k = slinear(xb, sa.key.weight, sa.key.bias)
        v = slinear(xb, sa.value.weight, sa.value.bias)
        ctx = TextAttnFn.apply(q, k, v, kpm, B, 12, L, adp)
        ao = slinear(ctx, at.output.dense.weight, at.output.dense.bias, out_fp32=True)
        ln = at.output.LayerNorm
        x32, xb = ops.add_layernorm(x32, ao, ln.weight, ln.bias, eps=ln.eps, drop_p=dp)
        h = slinear(xb, lyr.intermediate.dense.weight, lyr.intermediate.dense.bias)
        g = GeluFn.apply(h)
        o2 = slinear(g, lyr.output.dense.weight, lyr.output.dense.bias, out_fp32=True)
        ln = lyr.output.LayerNorm
        x32, xb = ops.add_layernorm(x32, o2, ln.weight, ln.bias, eps=ln.eps, drop_p=dp)
    return x32, xb
