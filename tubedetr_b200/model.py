"""Drop-in TubeDETR module, SetCriterion and build() on the B200 kernels.

Boundary kept from the reference:
  * TubeDETR.forward(samples, durations, captions, encode_and_save=True, memory_cache=None, samples_fast=None)
    -- reference models/tubedetr.py:93-101; two-phase protocol of engine.py:67-80; memory_cache keys of
    models/transformer.py:448-458; output keys of models/tubedetr.py:223-254.
  * SetCriterion(losses, sigma).forward(outputs, targets, inter_idx, time_mask) -- models/tubedetr.py:257-460.
  * build(args) -> (model, criterion, weight_dict) -- models/tubedetr.py:463-506.
  * state_dict names/shapes identical to the reference (923 entries) so its checkpoints load with strict=True.
Internals are NOT the reference's: activations are batch-major pixel/token rows ([frames*tokens, C]), all convolutions
and linear layers run on the tcgen05 GEMM, attention / LayerNorm on the kernels of libtdb.so, and the per-clip Python
loops of the reference (tubedetr.py:167-179, transformer.py:275-308, 400-417) are single gathers.
Only the reference default path is implemented (temporal stride > 0, fast_mode "", sine embeddings); other flag values
are rejected loudly in build().  In train() mode every dropout of the reference is applied (attention probabilities inside the
attention kernel, residual / FFN / resizer / sted-head dropouts as torch ops); eval() is deterministic.
"""
import math
import os
import zlib

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from . import ops
from . import text
from .resnet import STAGES, ResNet101Engine

D_MODEL, NHEAD, DFF = 256, 8, 2048


class NestedTensor(object):
    """Same container as reference util/misc.py:106-120 (tensors + bool pad mask)."""

    def __init__(self, tensors, mask):
        self.tensors = tensors
        self.mask = mask

    def to(self, *args, **kwargs):
        return type(self)(self.tensors.to(*args, **kwargs), self.mask.to(*args, **kwargs) if self.mask is not None else None)

    def decompose(self):
        return self.tensors, self.mask

    @classmethod
    def from_tensor_list(cls, clips):
        from .synthetic import pack_clips
        return cls(*pack_clips(clips))


# ----------------------------------------------------------------------------- parameter containers (names only)
class _Conv(nn.Module):
    def __init__(self, cin, cout, k, bias=False, trainable=True):
        super().__init__()
        w = torch.empty(cout, cin, k, k)
        nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        self.weight = nn.Parameter(w, requires_grad=trainable)
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))


class _FrozenBN(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *a, **k)


class _Bottleneck(nn.Module):
    def __init__(self, cin, width, first, trainable):
        super().__init__()
        self.conv1, self.bn1 = _Conv(cin, width, 1, trainable=trainable), _FrozenBN(width)
        self.conv2, self.bn2 = _Conv(width, width, 3, trainable=trainable), _FrozenBN(width)
        self.conv3, self.bn3 = _Conv(width, width * 4, 1, trainable=trainable), _FrozenBN(width * 4)
        if first:
            self.downsample = nn.Sequential(_Conv(cin, width * 4, 1, trainable=trainable), _FrozenBN(width * 4))


class _Body(nn.Module):
    def __init__(self, train_backbone=True):
        super().__init__()
        self.conv1, self.bn1 = _Conv(3, 64, 7, trainable=False), _FrozenBN(64)
        cin = 64
        for li, (width, nb, _) in enumerate(STAGES, start=1):
            tr = train_backbone and li >= 2  # reference models/backbone.py:82-89
            setattr(self, f"layer{li}", nn.Sequential(*[_Bottleneck(cin if b == 0 else width * 4, width, b == 0, tr)
                                                        for b in range(nb)]))
            cin = width * 4


class _BackboneBase(nn.Module):
    def __init__(self, train_backbone):
        super().__init__()
        self.body = _Body(train_backbone)
        self.num_channels = 2048


class _PosSine(nn.Module):
    """reference models/position_encoding.py:52-94 with N_steps=128, normalize=True."""

    def forward(self, mask):
        nm = (~mask).float()
        y, x = nm.cumsum(1), nm.cumsum(2)
        y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
        x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
        i = torch.arange(128, dtype=torch.float32, device=mask.device)
        dim_t = 10000.0 ** (2 * torch.div(i, 2, rounding_mode="floor") / 128)

        def enc(e):
            pe = e[..., None] / dim_t
            return torch.stack((pe[..., 0::2].sin(), pe[..., 1::2].cos()), dim=-1).flatten(-2)

        return torch.cat((enc(y), enc(x)), dim=-1)  # (N,h,w,256) channel-last

    def rows(self, mask):
        """(N, h*w, 256) embedding rows; on CUDA one kernel (tdb_pos_sine) instead of ~15 elementwise launches"""
        N, h, w = mask.shape
        if not mask.is_cuda:
            return self.forward(mask).view(N, h * w, 256)
        from . import kernels as K
        out = torch.empty(N, h * w, 256, dtype=torch.float32, device=mask.device)
        return K.pos_sine(mask.contiguous().view(torch.uint8), out, N, h, w)


class _Lin(nn.Module):
    def __init__(self, cin, cout, xavier=False, zero=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.zeros(cout))
        if zero:
            nn.init.zeros_(self.weight)
        elif xavier:
            nn.init.xavier_uniform_(self.weight)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            nn.init.uniform_(self.bias, -1 / math.sqrt(cin), 1 / math.sqrt(cin))


class _LN(nn.Module):
    def __init__(self, d, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))
        self.eps = eps


class _MHA(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        nn.init.xavier_uniform_(self.in_proj_weight)
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = _Lin(d, d, xavier=True)
        nn.init.zeros_(self.out_proj.bias)


class _EncLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.self_attn = _MHA(D_MODEL)
        self.linear1, self.linear2 = _Lin(D_MODEL, DFF, xavier=True), _Lin(DFF, D_MODEL, xavier=True)
        self.norm1, self.norm2 = _LN(D_MODEL), _LN(D_MODEL)


class _DecLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.self_attn, self.cross_attn_image = _MHA(D_MODEL), _MHA(D_MODEL)
        self.linear1, self.linear2 = _Lin(D_MODEL, DFF, xavier=True), _Lin(DFF, D_MODEL, xavier=True)
        self.norm1, self.norm3, self.norm4 = _LN(D_MODEL), _LN(D_MODEL), _LN(D_MODEL)


class _Stack(nn.Module):
    def __init__(self, cls, n, final_norm):
        super().__init__()
        self.layers = nn.ModuleList([cls() for _ in range(n)])
        if final_norm:
            self.norm = _LN(D_MODEL)


class _TimeSine(nn.Module):
    def __init__(self, max_len, d):
        super().__init__()
        from .weights import _time_sine
        self.register_buffer("te", _time_sine(max_len, d))


class _Resizer(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = _Lin(768, D_MODEL)
        self.layer_norm = _LN(D_MODEL, eps=1e-12)  # reference models/transformer.py:765


class _MLP(nn.Module):
    """reference models/tubedetr.py:23-42 (dropout, when set, follows EVERY layer including the last one)"""

    def __init__(self, din, dh, dout, n, dropout=0.0):
        super().__init__()
        h = [dh] * (n - 1)
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip([din] + h, h + [dout]))
        self.dropout = dropout

    def forward(self, x):
        for i, l in enumerate(self.layers):
            x = l(x)
            if i < len(self.layers) - 1:
                x = F.relu(x)
            if self.dropout and self.training:
                x = F.dropout(x, self.dropout, True)
        return x

    def rows(self, xb, act=0):
        """CUDA path on bf16 rows [R, 256]: hidden layers on the tcgen05 GEMM (bias + ReLU epilogue, ReLU / dropout backward in the
        consumer's dgrad), last layer + activation (+ logit dropout) in one small kernel (tdb_head_out).  -> fp32 [R, dout]"""
        dp = self.dropout if (self.dropout and self.training) else 0.0
        nl = len(self.layers)
        first = True
        for i, l in enumerate(self.layers[:-1]):
            xb = ops.linear(xb, l.weight, l.bias, relu=True, masked_by_consumer=True, mask_dx=not first,
                            dx_scale=(1.0 / (1.0 - dp)) if (dp > 0 and not first) else 1.0)
            if dp > 0:
                xb = ops.hidden_dropout(xb, dp)
            first = False
        last = self.layers[-1]
        return ops.head_out(xb, last.weight, last.bias, act=act, drop_p=dp, mask_dx=nl > 1, dx_scale=(1.0 / (1.0 - dp)) if dp > 0 else 1.0)


class HashTokenizer:
    """Offline stand-in when the roberta-base vocabulary is not on disk: deterministic word hashing into RoBERTa's id
    range with BOS=0 / EOS=2 / PAD=1.  (The real RobertaTokenizerFast is used when available.)"""

    def __call__(self, texts, device):
        toks = [[0] + [3 + zlib.crc32(w.encode()) % 50000 for w in t.lower().split()] + [2] for t in texts]
        L = max(len(t) for t in toks)
        ids = torch.full((len(toks), L), 1, dtype=torch.long)
        am = torch.zeros(len(toks), L, dtype=torch.long)
        for i, t in enumerate(toks):
            ids[i, :len(t)] = torch.tensor(t)
            am[i, :len(t)] = 1
        return ids.to(device), am.to(device)


def _make_text_encoder(name="roberta-base", offline_stand_in=None):
    """roberta-base tokenizer + encoder exactly as the reference loads them (models/transformer.py:130-135); any failure to
    load them RAISES.  The offline stand-in (randomly initialised RobertaModel of the same architecture + HashTokenizer) exists
    for boxes without the checkpoint (tests, bench.py, smoke) and must be asked for: `offline_stand_in=True`
    (build(args) passes args.offline_text_encoder) or env TDB_OFFLINE_TEXT_ENCODER=1.  It warns when active: token ids then do
    NOT match a pretrained embedding table."""
    import os
    import warnings
    from transformers import RobertaConfig, RobertaModel, RobertaTokenizerFast
    if offline_stand_in is None:
        offline_stand_in = os.environ.get("TDB_OFFLINE_TEXT_ENCODER", "0") not in ("0", "")
    if offline_stand_in:
        warnings.warn("tubedetr_b200: offline text-encoder stand-in active (random-init RoBERTa architecture + hash tokenizer); "
                      "not suitable for training or for loading pretrained checkpoints", stacklevel=2)
        cfg = RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1,
                            bos_token_id=0, eos_token_id=2, layer_norm_eps=1e-5)
        return RobertaModel(cfg), HashTokenizer()
    tok = RobertaTokenizerFast.from_pretrained(name)
    enc = RobertaModel.from_pretrained(name)

    def tokenize(texts, device):
        be = tok.batch_encode_plus(texts, padding="longest", return_tensors="pt").to(device)
        return be["input_ids"], be["attention_mask"]
    return enc, tokenize


class Transformer(nn.Module):
    """Parameter container + the video-text encoder / space-time decoder drivers."""

    def __init__(self, num_encoder_layers=6, num_decoder_layers=6, video_max_len=200, stride=5, no_tsa=False, fast=True,
                 dropout=0.1, offline_text_encoder=None):
        super().__init__()
        self.dropout = dropout          # reference models/transformer.py:608-676 (attention, residual and FFN dropouts)
        self.encoder = _Stack(_EncLayer, num_encoder_layers, final_norm=False)
        self.decoder = _Stack(_DecLayer, num_decoder_layers, final_norm=True)
        self.time_embed = _TimeSine(video_max_len, D_MODEL)
        self.fast = fast
        if fast:
            self.fast_encoder = _Lin(D_MODEL, D_MODEL)
            self.fast_residual = _Lin(D_MODEL, D_MODEL, zero=True)  # reference zero-inits it (transformer.py:173-174)
        self.text_encoder, self._tokenize = _make_text_encoder(offline_stand_in=offline_text_encoder)
        self.resizer = _Resizer()
        self.d_model, self.nhead, self.stride, self.no_tsa = D_MODEL, NHEAD, stride, no_tsa
        self.video_max_len = video_max_len
        # decoder cross-attention: "hoist" (default) = K / V of ALL layers projected by two GEMMs over the layer-invariant memory,
        # per-layer streaming attention core on their column slices; "fused" = per-layer tcgen05 kernel with the K/V projection fused
        # in (K, V never reach HBM; S >= 43); "unfused" = per-layer projections + generic attention kernel
        self.xattn_mode = os.environ.get("TDB_XATTN_MODE", "hoist")
        self.fused_xattn = True

    def _reset_temporal_parameters(self):  # called by reference main.py:545 after loading MDETR weights
        if self.fast:
            nn.init.zeros_(self.fast_residual.weight)
            nn.init.zeros_(self.fast_residual.bias)

    # ---- one post-norm encoder layer on token rows (reference transformer.py:629-646)
    def _drop(self):
        return self.dropout if self.training else 0.0

    def _ffn(self, l, xb, dp):
        """linear2(dropout(relu(linear1(x)))) -> fp32.  Without dropout the ReLU backward rides in linear2's dgrad epilogue."""
        if dp > 0:      # the dropout on the FFN output is applied by the add_layernorm that consumes it (drop_p=dp)
            # hidden dropout: one kernel forward; its backward and the ReLU backward ride in linear2's dgrad epilogue
            hdn = ops.linear(xb, l.linear1.weight, l.linear1.bias, relu=True, masked_by_consumer=True)
            hdn = ops.hidden_dropout(hdn, dp)
            return ops.linear(hdn, l.linear2.weight, l.linear2.bias, out_fp32=True, mask_dx=True, dx_scale=1.0 / (1.0 - dp))
        hdn = ops.linear(xb, l.linear1.weight, l.linear1.bias, relu=True, masked_by_consumer=True)
        return ops.linear(hdn, l.linear2.weight, l.linear2.bias, out_fp32=True, mask_dx=True)

    def _enc_layer(self, l, x32, xb, xpb, pos, kpm, n, S, need_pos):
        a, dp = l.self_attn, self._drop()
        qk, v = ops.in_proj(a.in_proj_weight, a.in_proj_bias, ((0, 512), (512, 768)), xpb, xb)
        o, _ = ops.mha(qk, None, v, kpm, n, NHEAD, S, S, 32 ** -0.5, packed=True, drop_p=dp, need_weights=False)
        att = ops.linear(o, a.out_proj.weight, a.out_proj.bias, out_fp32=True)
        x32, xb = ops.add_layernorm(x32, att, l.norm1.weight, l.norm1.bias, drop_p=dp)      # residual dropouts ride in the LN kernels
        f = self._ffn(l, xb, dp)
        if need_pos:
            return ops.add_layernorm(x32, f, l.norm2.weight, l.norm2.bias, pos, drop_p=dp)
        return ops.add_layernorm(x32, f, l.norm2.weight, l.norm2.bias, drop_p=dp) + (None,)

    # ---- one decoder layer (reference transformer.py:684-751); rows are (b,t) batch-major
    def _dec_layer(self, l, x32, xb, xqb, qp, memb, mempb, kpm_mem, kpm_q, B, T, S, kv=None, li=0):
        a, dp = l.self_attn, self._drop()
        if self.no_tsa:  # each time query attends to itself only (transformer.py:701-711)
            (v,) = ops.in_proj(a.in_proj_weight, a.in_proj_bias, ((512, 768),), xb)
            if dp > 0:   # softmax over one key = 1, then attention dropout on that single probability
                v = v * ((torch.rand(B * T, NHEAD, 1, device=v.device) >= dp).to(v.dtype) / (1 - dp)).expand(B * T, NHEAD, 32).reshape(B * T, 256)
            att = ops.linear(v, a.out_proj.weight, a.out_proj.bias, out_fp32=True)
            w = torch.ones(B * T, 1, 1, dtype=torch.float32, device=x32.device)
        else:
            qk, v = ops.in_proj(a.in_proj_weight, a.in_proj_bias, ((0, 512), (512, 768)), xqb, xb)
            o, w = ops.mha(qk, None, v, kpm_q, B, NHEAD, T, T, 32 ** -0.5, packed=True, drop_p=dp)
            att = ops.linear(o, a.out_proj.weight, a.out_proj.bias, out_fp32=True)
        x32, xb, xqb = ops.add_layernorm(x32, att, l.norm1.weight, l.norm1.bias, qp, drop_p=dp)
        c = l.cross_attn_image
        if kv is not None:                              # K / V of all layers were projected once, outside the layer loop (default)
            K_all, V_all, tok, shared, nl = kv
            o, cw = ops.xattn_core(xqb, K_all, V_all, tok if li == 0 else None, c.in_proj_weight, c.in_proj_bias, kpm_mem, shared, li,
                                   nl, B * T, S, 32 ** -0.5, drop_p=dp)
        elif self.fused_xattn and S >= 43:              # K/V projection fused into the attention kernel (K, V never reach HBM)
            (q,) = ops.in_proj(c.in_proj_weight, c.in_proj_bias, ((0, 256),), xqb)
            o, cw = ops.xattn_fused(q, mempb, memb, c.in_proj_weight, c.in_proj_bias, kpm_mem, B * T, S, 32 ** -0.5, drop_p=dp)
        else:
            q, k, v = ops.in_proj(c.in_proj_weight, c.in_proj_bias, ((0, 256), (256, 512), (512, 768)), xqb, mempb, memb)
            o, cw = ops.mha(q, k, v, kpm_mem, B * T, NHEAD, 1, S, 32 ** -0.5, drop_p=dp)
        att = ops.linear(o, c.out_proj.weight, c.out_proj.bias, out_fp32=True)
        x32, xb = ops.add_layernorm(x32, att, l.norm3.weight, l.norm3.bias, drop_p=dp)
        f = self._ffn(l, xb, dp)
        x32, xb, xqb = ops.add_layernorm(x32, f, l.norm4.weight, l.norm4.bias, qp, drop_p=dp)
        return x32, xb, xqb, w, cw


_TEXT_STREAMS = {}   # device -> side stream of the text encoder (module level: streams do not deepcopy with the model)


class TubeDETR(nn.Module):
    def __init__(self, num_queries=1, aux_loss=True, video_max_len=200, stride=5, guided_attn=True, fast=True,
                 fast_mode="", sted=True, no_tsa=False, enc_layers=6, dec_layers=6, train_backbone=True, dropout=0.1,
                 offline_text_encoder=None):
        super().__init__()
        assert num_queries == 1 and fast_mode == "" and stride > 0, "only the reference default path is implemented"
        self.num_queries = num_queries
        self.transformer = Transformer(enc_layers, dec_layers, video_max_len, stride, no_tsa, fast, dropout, offline_text_encoder)
        self.bbox_embed = _MLP(D_MODEL, D_MODEL, 4, 3)
        self.query_embed = nn.Embedding(num_queries, D_MODEL)
        self.input_proj = _Conv(2048, D_MODEL, 1, bias=True)
        self.backbone = nn.Sequential(_BackboneBase(train_backbone), _PosSine())
        self.backbone.num_channels = 2048
        self.aux_loss, self.video_max_len, self.stride = aux_loss, video_max_len, stride
        self.guided_attn, self.fast, self.fast_mode, self.sted = guided_attn, fast, fast_mode, sted
        if sted:
            self.sted_embed = _MLP(D_MODEL, D_MODEL, 2, 2, dropout=0.5)   # reference models/tubedetr.py:91
        self._engine = ResNet101Engine()
        self.text_side_stream = True   # RoBERTa on its own CUDA stream, concurrent with the backbone (forward and backward)
        self.joint_backbone = True  # slow + fast frames share one backbone batch when their spatial sizes agree
        self.fast_l2_chunk = None   # frames per chunk for the L2-resident schedule of stem+layer1+layer2 in the no-grad pass
        # Opt-in contract: samples.tensors[b, i] IS samples_fast.tensors[b, i*k] (what the reference's datasets produce:
        # datasets/vidstg.py:250-251 returns images[:, ::stride] next to images).  The reference then runs the backbone on those
        # frames twice (models/tubedetr.py:121 with grad, :128 under no_grad) with identical results; with this flag the
        # backbone runs once per distinct frame (T instead of T + ceil(T/k) frames) and the fast branch reuses the slow
        # frames' features (detached, as the fast branch never backpropagates into the backbone).
        self.slow_frames_alias_fast = False
        self.text_autocast = False  # library path only: run HF RoBERTa's GEMMs under bf16 autocast
        # text encoder as the HF library call it is in the reference (default), or on this library's kernels (text.py, TDB_OWN_TEXT=1):
        # parity-tested, but measured SLOWER inside the step (20.1 vs 18.5 ms): its one-tile tcgen05 GEMMs need a whole SM's shared
        # memory each and queue behind the backbone's persistent GEMM waves, where the library's small-footprint kernels slip in
        self.own_text_encoder = os.environ.get("TDB_OWN_TEXT", "0") != "0"

    # ------------------------------------------------------------------ helpers
    _TRANSIENT = ("_trunk", "_idx_cache")     # per-step state that must not follow a copy of the model

    def __deepcopy__(self, memo):
        """EMA copies (reference main.py:370 `deepcopy(model)`) may be taken at any time, also after a training forward: the
        non-leaf trunk tensors of the last step and the device index cache stay behind."""
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k_, v in self.__dict__.items():
            if k_ not in self._TRANSIENT:
                new.__dict__[k_] = copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        return {k_: v for k_, v in self.__dict__.items() if k_ not in self._TRANSIENT}

    def _backbone_tensors(self):
        return {"backbone.0.body." + k: v for k, v in list(self.backbone[0].body.named_parameters()) +
                list(self.backbone[0].body.named_buffers())}

    @staticmethod
    def _resize_mask(mask, h, w):
        H, W = mask.shape[-2:]
        ii = (torch.arange(h, device=mask.device) * (H / h)).floor().long().clamp(max=H - 1)
        jj = (torch.arange(w, device=mask.device) * (W / w)).floor().long().clamp(max=W - 1)
        return mask[:, ii][:, :, jj]

    def _index_tensors(self, durations, k, dev):
        """duration-dependent index tensors, built on the host once per (durations, stride) and cached on the device so a
        steady-state step issues no host->device copies (CUDA-graph capturable)."""
        key = (durations, k, str(dev))
        c = self.__dict__.setdefault("_idx_cache", {})
        if key not in c:
            if len(c) >= 64:          # durations vary per batch in real training: keep the cache bounded
                c.clear()
            B, T = len(durations), max(durations)
            n_clips = math.ceil(T / k)
            dur = torch.tensor(durations)
            tt = torch.arange(T)
            valid = tt[None] < dur[:, None]
            clip_of_t = (torch.arange(B)[:, None] * n_clips + tt[None] // k).flatten()
            c[key] = tuple(t.to(dev) for t in (dur, tt, valid, clip_of_t))
        return c[key]

    def _dedup_index_tensors(self, B, T, k, dev):
        """slow frames = fast frames [::k] of every video (uniform durations): indices of the fast frames that are NOT slow
        frames, and for every fast frame its row in the joint batch [slow frames | remaining fast frames]."""
        key = ("dedup", B, T, k, str(dev))
        c = self.__dict__.setdefault("_idx_cache", {})
        if key not in c:
            n_clips = math.ceil(T / k)
            t = torch.arange(B * T)
            tv, bv = t % T, t // T
            is_slow = tv % k == 0
            rest_idx = t[~is_slow]
            order = torch.empty(B * T, dtype=torch.long)
            order[is_slow] = (bv * n_clips + tv // k)[is_slow]
            order[~is_slow] = B * n_clips + torch.arange(rest_idx.numel())
            c[key] = (rest_idx.to(dev), order.to(dev))
        return c[key]

    def trunk_outputs(self):
        """(RoBERTa last_hidden_state, backbone features of the slow frames) of the latest encode call: the tensors
        `parallel.backward_overlapped` splits the backward at so the text-encoder gradients can be all-reduced while the
        backbone backward runs."""
        return self.__dict__.pop("_trunk", (None, None))     # handed over once: nothing of the step's graph stays on the module

    def text_stream(self, device):
        return _TEXT_STREAMS.get(device) if self.text_side_stream else None

    def _tokenize(self, captions, device):
        if isinstance(captions, (tuple, list)) and len(captions) == 2 and torch.is_tensor(captions[0]):
            return captions[0].to(device), captions[1].to(device)  # pre-tokenised (input_ids, attention_mask)
        return self.transformer._tokenize(list(captions), device)

    # ------------------------------------------------------------------ forward
    def forward(self, samples, durations, captions, encode_and_save=True, memory_cache=None, samples_fast=None):
        _lib.lib()  # fail loudly if libtdb.so is missing: there is no PyTorch fallback
        if encode_and_save:
            assert memory_cache is None
            return self._encode(samples, durations, captions, samples_fast)
        assert memory_cache is not None
        return self._decode(memory_cache)

    def _encode(self, samples, durations, captions, samples_fast):
        tr = self.transformer
        frames, fmask = samples.decompose()
        if not frames.is_cuda:
            raise RuntimeError("tubedetr_b200 runs on CUDA (sm_100a) only; move the inputs to the GPU")
        dev = frames.device
        B, T, k = len(durations), max(durations), self.stride
        n_clips = math.ceil(T / k)
        # Text encoder (library call, reference transformer.py:250-263) on a side stream: its ~10^3 tiny kernels (L = 20 tokens)
        # are independent of the backbone in forward AND backward (autograd replays a node on the stream it ran on), so they
        # fill the launch gaps of the convolution GEMMs instead of extending the critical path.  Joined right before first use.
        ids, am = self._tokenize(captions, dev)
        if self.training:
            ops.advance_dropout_seed(dev)
        side = self.text_side_stream
        if side:
            tstream = _TEXT_STREAMS.get(dev)
            if tstream is None:
                # high priority (default): its tiny kernels slip in at GEMM boundaries; TDB_TEXT_PRIO=0 = same priority as the main stream
                tstream = _TEXT_STREAMS[dev] = torch.cuda.Stream(device=dev, priority=int(os.environ.get("TDB_TEXT_PRIO", "-1")))
            main = torch.cuda.current_stream(dev)
            tstream.wait_stream(main)
        with torch.cuda.stream(tstream if side else torch.cuda.current_stream(dev)):
            if self.own_text_encoder and text.supported(tr.text_encoder, ids):
                # RoBERTa on this library's kernels (text.py): the HF module only holds the parameters
                _, hid = text.roberta_forward(tr.text_encoder, ids, am, self.training)      # bf16 rows (B*L, 768)
            else:
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.text_autocast):
                    hid = tr.text_encoder(input_ids=ids, attention_mask=am).last_hidden_state    # (B,L,768)

        sd = self._backbone_tensors()
        W = self._engine.prepare(sd)
        names = [n for n, p in sd.items() if isinstance(p, nn.Parameter) and p.requires_grad]
        feat_f = None
        joint = (self.fast and self.joint_backbone and samples_fast is not None
                 and samples_fast.tensors.shape[2:] == frames.shape[2:])
        dedup = (joint and self.slow_frames_alias_fast and k > 1 and all(d == T for d in durations)
                 and samples_fast.tensors.shape[0] == B * T and frames.shape[0] == B * n_clips)
        order = None
        if joint:      # slow + fast frames as one backbone batch (same launches, better SM fill); fast rows carry no grad
            ffr = samples_fast.tensors.float()
            if dedup:  # only the fast frames that are not slow frames; `order` maps fast frame -> row block of the joint batch
                rest_idx, order = self._dedup_index_tensors(B, T, k, dev)
                ffr = ffr.index_select(0, rest_idx)
            if names and torch.is_grad_enabled():
                feat, feat_f = ops.BackboneJointFn.apply(frames.float(), ffr, self._engine, W, names, "joint",
                                                         *[sd[n] for n in names])
                h, w = self._engine.last_hw
            else:
                fa, h, w, _ = self._engine.forward([frames.float(), ffr], W, save=False, tag="joint")
                feat, feat_f = fa[:frames.shape[0] * h * w], fa[frames.shape[0] * h * w:]
        elif names and torch.is_grad_enabled():
            feat = ops.BackboneFn.apply(frames.float(), self._engine, W, names, "slow", *[sd[n] for n in names])
            h, w = self._engine.last_hw
        else:
            feat, h, w, _ = self._engine.forward(frames.float(), W, save=False, tag="slow")
        # the two trunk outputs a data-parallel step may cut the backward at (parallel.backward_overlapped)
        self.__dict__["_trunk"] = (hid, feat) if torch.is_grad_enabled() else (None, None)
        n, HW = frames.shape[0], h * w
        assert n == B * n_clips, "with temporal stride every video of the batch needs the same number of clips"
        with torch.no_grad():
            m_s = self._resize_mask(fmask, h, w).clone()
            m_s[:, 0, 0] = False
            pos = self.backbone[1].rows(m_s)                     # (n, HW, 256) sine embedding, one kernel
            dur, tt, valid, clip_of_t = self._index_tensors(tuple(durations), k, dev)
        Win, bin_ = self.input_proj.weight.view(D_MODEL, 2048), self.input_proj.bias
        src = ops.linear(feat, Win, bin_, out_fp32=True).view(n, HW, D_MODEL)

        fsrc = None
        if self.fast:
            ff, fm_ = samples_fast.decompose()
            with torch.no_grad():
                if feat_f is None:
                    feat_f, hf, wf, _ = self._engine.forward(ff.float(), W, save=False, tag="fast",
                                                             l2_chunk=self.fast_l2_chunk)
                else:
                    hf, wf = h, w
                m_f = self._resize_mask(fm_, hf, wf)
            if order is not None:    # joint batch rows [slow | rest] -> fast frame order; slow features enter detached
                fsrc_all = ops.linear(torch.cat([feat.detach(), feat_f], 0), Win, bin_).view(-1, HW, D_MODEL).index_select(0, order)
            else:
                fsrc_all = ops.linear(feat_f, Win, bin_).view(-1, HW, D_MODEL)   # input_proj gets weight-grad here too
            ragged = any(d != T for d in durations)
            if ragged:
                idx = ((torch.cumsum(dur, 0) - dur)[:, None] + tt[None]).clamp(max=ff.shape[0] - 1).flatten()
                fsrc = fsrc_all[idx] * valid.flatten()[:, None, None].to(fsrc_all.dtype)
                m_t = torch.where(valid.flatten()[:, None], m_f.flatten(1)[idx], torch.ones((), dtype=torch.bool, device=dev))
            else:
                fsrc, m_t = fsrc_all, m_f.flatten(1)
        else:
            m_t = torch.where(valid.flatten()[:, None], m_s.flatten(1)[clip_of_t], torch.ones((), dtype=torch.bool, device=dev))
        m_t = m_t.clone()
        m_t[:, 0] = False

        # time queries + masks (reference transformer.py:211-238)
        qp = (self.query_embed.weight[0][None, None, :] + tr.time_embed.te[:T, 0][None]).expand(B, T, D_MODEL)
        q_kpm = ~valid
        q_kpm[:, 0] = False

        # text features -> resizer on our kernels
        if side:
            main.wait_stream(tstream)
            hid.record_stream(main)
        L = ids.shape[1]
        hidb = hid if hid.dtype == torch.bfloat16 else hid.float().reshape(B * L, 768).to(torch.bfloat16)
        r = ops.linear(hidb.reshape(B * L, 768), tr.resizer.fc.weight, tr.resizer.fc.bias, out_fp32=True)
        txt, _ = ops.add_layernorm(r, None, tr.resizer.layer_norm.weight, tr.resizer.layer_norm.bias, eps=1e-12)
        if self.training:
            txt = F.dropout(txt, 0.1, True)                     # FeatureResizer dropout (reference transformer.py:141,772)
        txt = txt.view(B, L, D_MODEL)
        txt_kpm = am.ne(1)
        txt_rep = txt.repeat_interleave(n_clips, 0)                                      # (n,L,d)

        S = HW + L
        # [image tokens | the video's text tokens] per clip, position rows, and the bf16 operands of the first in-projection: one kernel
        x32, xb, xpb, pe = ops.enc_assemble(src, txt, pos, n_clips)
        kpm_enc_b = torch.cat([m_s.flatten(1), txt_kpm.repeat_interleave(n_clips, 0)], 1)
        kpm_enc = kpm_enc_b.to(torch.uint8).contiguous()
        nl = len(tr.encoder.layers)
        for i, l in enumerate(tr.encoder.layers):
            x32, xb, xpb = tr._enc_layer(l, x32, xb, xpb, pe, kpm_enc, n, S, need_pos=i < nl - 1)

        # temporal replication: frame (b,t) <- clip (b, t//k)  (reference transformer.py:393-427)
        kpm_dec = torch.cat([m_t, txt_kpm.repeat_interleave(T, 0)], 1)
        kpm_dec[:, 0] = False
        BT = B * T
        upd = None
        if self.fast:      # fast branch: z = enc[clip(t)] + fast_encoder(fast features) -> fast_residual (reference transformer.py:373-391)
            fm = ops.linear(fsrc.reshape(BT * HW, D_MODEL), tr.fast_encoder.weight, tr.fast_encoder.bias)
            z = ops.fast_mix(x32, fm, B, T, k, HW, S)
            upd = ops.linear(z, tr.fast_residual.weight, tr.fast_residual.bias, out_fp32=True)
        # replication over time + aggregation + the decoder's bf16 memory operands in one pass (reference transformer.py:393-446)
        mem, mem_pos, memb, mempb = ops.aggregate(x32, pe, upd, B, T, k, HW, S)
        mem, mem_pos = mem.view(BT, S, D_MODEL), mem_pos.view(BT, S, D_MODEL)
        return {
            "text_memory_resized": txt_rep.transpose(0, 1), "text_memory": mem[:, HW:].transpose(0, 1),
            "text_attention_mask": txt_kpm.repeat_interleave(n_clips, 0), "tokenized": {"input_ids": ids, "attention_mask": am},
            "img_memory": mem.transpose(0, 1), "mask": kpm_dec, "pos_embed": mem_pos.transpose(0, 1),
            "query_embed": qp.transpose(0, 1), "query_mask": q_kpm,
            # private: bf16 operands of the decoder's key / value projections, produced by the same kernel as img_memory (decode
            # recomputes them from img_memory / pos_embed when a caller passes a cache without them)
            "_memb": memb, "_mempb": mempb,
        }

    def _decode(self, mc):
        tr = self.transformer
        mem = mc["img_memory"].transpose(0, 1).contiguous()          # (B*T,S,d) fp32
        mem_pos = mc["pos_embed"].transpose(0, 1)
        qp = mc["query_embed"].transpose(0, 1).contiguous()          # (B,T,d)
        B, T, _ = qp.shape
        BT, S, _ = mem.shape
        kpm_mem = mc["mask"].to(torch.uint8).contiguous()
        kpm_q = mc["query_mask"].to(torch.uint8).contiguous()
        memb, mempb = mc.get("_memb"), mc.get("_mempb")
        if memb is None or mempb is None or memb.shape[0] != BT * S:
            memb = mem.reshape(BT * S, D_MODEL).to(torch.bfloat16)
            mempb = (mem + mem_pos).reshape(BT * S, D_MODEL).to(torch.bfloat16)
        qp = qp.reshape(B * T, D_MODEL).float().contiguous()
        x32 = torch.zeros(B * T, D_MODEL, device=mem.device)          # tgt = 0 (reference transformer.py:463-464)
        xb, xqb = x32.to(torch.bfloat16), qp.to(torch.bfloat16)
        dn = tr.decoder.norm
        hs, ws, cws = [], [], []
        kv = None
        if tr.xattn_mode == "hoist":
            ca = [l.cross_attn_image for l in tr.decoder.layers]
            kv = ops.decoder_kv(mempb, memb, [c.in_proj_weight for c in ca], [c.in_proj_bias for c in ca]) + (len(ca),)
        tr.fused_xattn = tr.xattn_mode != "unfused"
        for li, l in enumerate(tr.decoder.layers):
            x32, xb, xqb, w, cw = tr._dec_layer(l, x32, xb, xqb, qp, memb, mempb, kpm_mem, kpm_q, B, T, S, kv=kv, li=li)
            hs.append(ops.add_layernorm(x32, None, dn.weight, dn.bias)[1])      # bf16 copy of decoder.norm(output): the heads' GEMM operand
            ws.append(w)
            cws.append(cw)
        nlay = len(hs)
        hsb = torch.stack(hs).view(nlay * B * T, D_MODEL)
        out = {}
        boxes = self.bbox_embed.rows(hsb, act=1).view(nlay, B * T, 4)           # sigmoid inside the head kernel
        out["pred_boxes"] = boxes[-1]
        if self.sted:
            sted = self.sted_embed.rows(hsb).view(nlay, B, T, 2)
            out["pred_sted"] = sted[-1]
        if self.guided_attn:
            out["weights"], out["ca_weights"] = ws[-1], cws[-1]
        if self.aux_loss:
            out["aux_outputs"] = []
            for i in range(len(ws) - 1):
                a = {"pred_boxes": boxes[i]}
                if self.sted:
                    a["pred_sted"] = sted[i]
                if self.guided_attn:
                    a["weights"], a["ca_weights"] = ws[i], cws[i]
                out["aux_outputs"].append(a)
        return out


# ----------------------------------------------------------------------------- criterion
def _box_cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def _giou_pairs(a, b):
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, 2:], b[:, 2:]) - torch.max(a[:, :2], b[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = area_a + area_b - inter
    whc = (torch.max(a[:, 2:], b[:, 2:]) - torch.min(a[:, :2], b[:, :2])).clamp(min=0)
    areac = whc[:, 0] * whc[:, 1]
    return inter / union - (areac - union) / areac


class SetCriterion(nn.Module):
    """Same API and loss values as reference models/tubedetr.py:257-460 (L1 + GIoU on kept boxes, KL on start/end,
    guided attention), computed pairwise instead of through an NxN GIoU matrix and without the .item() host sync."""

    def __init__(self, losses, sigma=1):
        super().__init__()
        self.losses, self.sigma = losses, sigma
        self.fused = os.environ.get("TDB_FUSED_CRITERION", "1") != "0"    # CUDA inputs: two kernels instead of ~150 (tdb_loss.cu)

    def prepare(self, targets, inter_idx=None, time_mask=None, device=None):
        """Host-side part of the loss (target tensors, Gaussian start/end distributions, negative-frame map, the
        num_boxes all-reduce of reference models/tubedetr.py:411-413).  Hoistable: set `criterion.static = prepare(...)`
        to keep it out of a captured CUDA graph when targets do not change."""
        dev = device or targets[0]["boxes"].device
        num_boxes = torch.as_tensor([float(sum(len(t["boxes"]) for t in targets))], device=dev)
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(num_boxes)
            world = torch.distributed.get_world_size()
        prep = {"num_boxes": torch.clamp(num_boxes / world, min=1)[0],
                "tgt_boxes": torch.cat([t["boxes"] for t in targets], 0)}
        if inter_idx is not None and time_mask is not None:
            pm = torch.zeros(time_mask.shape, dtype=torch.bool)
            for k_, idx in enumerate(inter_idx):
                if idx[0] >= 0:
                    pm[k_, idx[0]:idx[1] + 1] = True
            neg = pm.to(dev) | ~time_mask
            prep["neg"], prep["nneg"] = neg, (~neg).sum(1) + 1e-6
            tt = torch.arange(time_mask.shape[1])
            gs = []
            for c in range(2):
                tgt = torch.tensor([x[c] for x in inter_idx])
                g = (-((tt[None] - tgt[:, None]) ** 2) / (2 * self.sigma ** 2)).exp()
                gs.append(F.normalize(g + 1e-6, p=1, dim=1).to(dev))
            prep["gauss"] = gs
            prep["gauss_bt2"] = torch.stack(gs, -1).contiguous()
            prep["nneg"] = prep["nneg"].float().contiguous()
        return prep

    def forward(self, outputs, targets, inter_idx=None, time_mask=None):
        prep = getattr(self, "static", None) or self.prepare(targets, inter_idx, time_mask, outputs["pred_boxes"].device)
        # all decoder layers at once: the main output and the 5 aux outputs are stacked on a leading axis so every loss
        # term is ONE batched expression (6x fewer kernels than the reference's per-layer Python loop, same values)
        layers = [outputs] + list(outputs.get("aux_outputs", []))
        names = [""] + [f"_{i}" for i in range(len(layers) - 1)]
        if self.fused and outputs["pred_boxes"].is_cuda and len(layers) <= 8:
            return self._fused(layers, names, prep, time_mask)
        vals = self._all(layers, prep, time_mask)
        return {k + sfx: v[i] for k, v in vals.items() for i, sfx in enumerate(names)}

    def _fused(self, layers, names, prep, time_mask):
        """all loss terms of all decoder layers in ONE kernel (tdb_loss.cu), their gradients in one more"""
        fams = [f for f in ("boxes", "sted", "guided_attn") if f in self.losses]
        f32 = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
        pbs = [f32(o["pred_boxes"]) for o in layers] if "boxes" in fams else None
        sts = [f32(o["pred_sted"]) for o in layers] if "sted" in fams else None
        ws = [f32(o["weights"]) for o in layers] if "guided_attn" in fams else None
        tm = time_mask.contiguous().view(torch.uint8) if time_mask is not None else None
        neg = prep["neg"].contiguous().view(torch.uint8) if "neg" in prep else None
        flat = CriterionFn.apply(self, prep, tm, neg, len(layers), *(pbs or []), *(sts or []), *(ws or []))
        nl = len(layers)
        out = {}
        for fi, key in enumerate(("loss_bbox", "loss_giou", "loss_sted", "loss_guided_attn")):
            fam = ("boxes", "boxes", "sted", "guided_attn")[fi]
            if fam in fams:
                for i, sfx in enumerate(names):
                    out[key + sfx] = flat[fi * nl + i]
        return out

    def _all(self, layers, prep, time_mask):
        l, eps = {}, 1e-6
        if "boxes" in self.losses:
            pb = torch.stack([o["pred_boxes"] for o in layers])                     # (Lyr,K,4)
            tb = prep["tgt_boxes"][None]
            l["loss_bbox"] = (pb - tb).abs().sum((1, 2)) / prep["num_boxes"]
            a, b = _box_cxcywh_to_xyxy(pb), _box_cxcywh_to_xyxy(tb.expand_as(pb))
            giou = _giou_pairs(a.reshape(-1, 4), b.reshape(-1, 4)).view(pb.shape[0], -1)
            l["loss_giou"] = (1 - giou).sum(1) / prep["num_boxes"]
        if "sted" in self.losses:
            sted = torch.stack([o["pred_sted"] for o in layers])                    # (Lyr,B,T,2)
            sted = sted.masked_fill(~time_mask[None, :, :, None], -1e32)
            p = sted.softmax(2)                                                     # over time
            g = torch.stack(prep["gauss"], -1)[None]                                # (1,B,T,2)
            kl = p * ((p + eps) / g).log() * time_mask[None, :, :, None]
            l["loss_sted"] = kl.sum(3).mean((1, 2))
        if "guided_attn" in self.losses:
            w = torch.stack([o["weights"] for o in layers])                         # (Lyr,B,T,T)
            ga = -(1 - w + eps).log()
            ga = ga.masked_fill(prep["neg"][None, :, :, None], 0)
            l["loss_guided_attn"] = (ga.sum(3) / prep["nneg"][None, :, None]).sum(2).mean(1)
        return l


class CriterionFn(torch.autograd.Function):
    """SetCriterion on the fused kernels: forward -> 4 * nlayers scalars (bbox, giou, sted, guided x layer) as separate 0-dim
    outputs, backward -> gradients of every pred_boxes / pred_sted / weights tensor in one launch."""

    @staticmethod
    def forward(ctx, crit, prep, tm, neg, nl, *tensors):
        from . import kernels as K
        fams = [f for f in ("boxes", "sted", "guided_attn") if f in crit.losses]
        it = iter(range(0, len(tensors), nl))
        pbs = list(tensors[next(it):][:nl]) if "boxes" in fams else None
        sts = list(tensors[next(it):][:nl]) if "sted" in fams else None
        ws = list(tensors[next(it):][:nl]) if "guided_attn" in fams else None
        B, T = (sts[0].shape[0], sts[0].shape[1]) if sts else ((ws[0].shape[0], ws[0].shape[1]) if ws else (1, 1))
        Kb = pbs[0].shape[0] if pbs else 0
        desc = K.loss_desc(pbs, sts, ws, prep["tgt_boxes"].float().contiguous() if pbs else None, prep["num_boxes"].reshape(1).float().contiguous(),
                           prep.get("gauss_bt2"), tm, neg, prep.get("nneg"), Kb, B, T)
        dev = tensors[0].device
        losses = torch.zeros(4 * nl, dtype=torch.float32, device=dev) if len(fams) < 3 else torch.empty(4 * nl, dtype=torch.float32, device=dev)
        K.criterion_fwd(desc, losses)
        ctx.desc, ctx.keep = desc, (pbs, sts, ws, prep, tm, neg)      # keep = owners of the pointers in desc
        ctx.dims = (nl, Kb, B, T)
        return tuple(losses.unbind(0))

    @staticmethod
    def backward(ctx, *gs):
        from . import kernels as K
        pbs, sts, ws, prep, tm, neg = ctx.keep
        nl, Kb, B, T = ctx.dims
        dev = (pbs or sts or ws)[0].device
        zero = None
        parts = []
        for g in gs:
            if g is None:
                zero = torch.zeros((), dtype=torch.float32, device=dev) if zero is None else zero
                g = zero
            parts.append(g.reshape(()).float())
        gl = torch.stack(parts)
        d_boxes = torch.empty(nl, Kb, 4, dtype=torch.float32, device=dev) if pbs else None
        d_sted = torch.empty(nl, B, T, 2, dtype=torch.float32, device=dev) if sts else None
        d_w = torch.empty(nl, B, T, T, dtype=torch.float32, device=dev) if ws else None
        K.criterion_bwd(ctx.desc, gl, d_boxes, d_sted, d_w)
        grads = []
        for d_, lst in ((d_boxes, pbs), (d_sted, sts), (d_w, ws)):
            if lst:
                grads += [d_[i] for i in range(nl)]
        return (None, None, None, None, None) + tuple(grads)


# ----------------------------------------------------------------------------- factory
_UNSUPPORTED = {"fast_mode": "", "no_time_embed": False, "learn_time_embed": False, "dilation": False,
                "position_embedding": "sine", "hidden_dim": 256, "nheads": 8, "dim_feedforward": 2048,
                "num_queries": 1, "backbone": "resnet101", "pass_pos_and_query": True}


def build(args):
    """reference models/tubedetr.py:463-506; supports the default value of every architecture flag plus
    --stride/--resolution/--no_fast/--no_tsa/--no_guided_attn/--no_sted/--no_aux_loss; anything else raises."""
    for k_, v in _UNSUPPORTED.items():
        if hasattr(args, k_) and getattr(args, k_) != v:
            raise NotImplementedError(f"tubedetr_b200: --{k_}={getattr(args, k_)} is not implemented (only {v!r})")
    if not getattr(args, "stride", 5):
        raise NotImplementedError("tubedetr_b200: stride=0 (no temporal sampling) is not implemented")
    model = TubeDETR(num_queries=args.num_queries, aux_loss=args.aux_loss, video_max_len=args.video_max_len_train,
                     stride=args.stride, guided_attn=args.guided_attn, fast=args.fast, fast_mode=args.fast_mode,
                     sted=args.sted, no_tsa=args.no_tsa, enc_layers=args.enc_layers, dec_layers=args.dec_layers,
                     train_backbone=args.lr_backbone > 0, dropout=getattr(args, "dropout", 0.1),
                     offline_text_encoder=getattr(args, "offline_text_encoder", None))
    if getattr(args, "freeze_backbone", False):
        for p in model.backbone.parameters():
            p.requires_grad_(False)
    if getattr(args, "freeze_text_encoder", False):
        for p in model.transformer.text_encoder.parameters():
            p.requires_grad_(False)
    weight_dict = {"loss_bbox": args.bbox_loss_coef, "loss_giou": args.giou_loss_coef, "loss_sted": args.sted_loss_coef}
    if args.guided_attn:
        weight_dict["loss_guided_attn"] = args.guided_attn_loss_coef
    if args.aux_loss:
        weight_dict.update({f"{k_}_{i}": v for i in range(args.dec_layers - 1) for k_, v in list(weight_dict.items())})
    losses = (["boxes", "sted"] if args.sted else ["boxes"]) + (["guided_attn"] if args.guided_attn else [])
    criterion = SetCriterion(losses=losses, sigma=args.sigma).to(torch.device(args.device))
    return model, criterion, weight_dict
