"""CPU ORACLE for the TubeDETR forward/backward hot path -- TEST INFRASTRUCTURE ONLY.

A functional, loop-free fp32 PyTorch restatement of the reference algorithm
(antoyang/TubeDETR @ d70c0eed).  It is imported only by tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke(); the product path
(tubedetr_b200/) never imports it and fails loudly without its CUDA library.

Parity pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference itself, executed in the build
container by tests/golden/make_golden.py (committed) and stored under tests/golden/.
tests/test_oracle_golden.py checks every fixture.  Third-party arithmetic on the path
(torchvision ResNet-101 topology, torch MultiheadAttention math, HF RoBERTa) is restated
here from its published definition, except RoBERTa which stays the library call it is in
the reference (reference models/transformer.py:130-135,255).

Everything takes a flat ``sd`` (reference state_dict names) so no nn.Module mirrors the
reference's class structure.  Layouts are batch-major ([frames, tokens, d]) internally.
"""
import math

import torch
import torch.nn.functional as F

RESNET101_BLOCKS = (3, 4, 23, 3)


# ----------------------------------------------------------------------------- backbone
def frozen_bn(x, sd, p):
    """reference models/backbone.py:60-70 (eps 1e-5, buffers only)."""
    scale = sd[p + ".weight"] * torch.rsqrt(sd[p + ".running_var"] + 1e-5)
    shift = sd[p + ".bias"] - sd[p + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def _q(t, on):
    """straight-through bf16 rounding (forward value rounded, gradient passes unchanged)"""
    return t + (t.to(torch.bfloat16).float() - t).detach() if on else t


def resnet101_layer4(x, sd, p="backbone.0.body.", emulate_bf16=False):
    """torchvision ResNet-101 v1.5 truncated at layer4 (reference models/backbone.py:93-122).

    emulate_bf16=True rounds weights and every stored activation to bf16 exactly where the CUDA path does (bf16
    operands, fp32 accumulate + FrozenBN/residual/ReLU epilogue, bf16 store).  Used only to validate the backward pass:
    ReLU masks then agree with the CUDA forward, which a comparison against the pure-fp32 network cannot guarantee.
    In the first block of every stage the CUDA path computes conv3 and the downsample branch as ONE GEMM over [y2 | x] with the
    FrozenBN scales folded into the bf16 weights and no bf16 store of the identity branch; the emulation rounds there too.
    """
    e = emulate_bf16

    def bn_affine(q):
        scale = sd[q + ".weight"] * torch.rsqrt(sd[q + ".running_var"] + 1e-5)
        return scale, sd[q + ".bias"] - sd[q + ".running_mean"] * scale

    def conv(t, w, **kw):
        return F.conv2d(t, _q(w, e), **kw)

    x = conv(_q(x, e), sd[p + "conv1.weight"], stride=2, padding=3)
    x = _q(F.relu(frozen_bn(x, sd, p + "bn1")), e)
    x = F.max_pool2d(x, 3, stride=2, padding=1)
    for li, nblocks in enumerate(RESNET101_BLOCKS, start=1):
        for bi in range(nblocks):
            q = f"{p}layer{li}.{bi}."
            stride = 2 if (bi == 0 and li > 1) else 1
            idt = x
            y = _q(F.relu(frozen_bn(conv(x, sd[q + "conv1.weight"]), sd, q + "bn1")), e)
            y = _q(F.relu(frozen_bn(conv(y, sd[q + "conv2.weight"], stride=stride, padding=1), sd, q + "bn2")), e)
            if bi == 0 and e:
                s3, b3 = bn_affine(q + "bn3")
                sdn, bdn = bn_affine(q + "downsample.1")
                y = (F.conv2d(y, _q(sd[q + "conv3.weight"] * s3.view(-1, 1, 1, 1), e))
                     + F.conv2d(x, _q(sd[q + "downsample.0.weight"] * sdn.view(-1, 1, 1, 1), e), stride=stride)
                     + (b3 + bdn).view(1, -1, 1, 1))
                x = _q(F.relu(y), e)
                continue
            y = frozen_bn(conv(y, sd[q + "conv3.weight"]), sd, q + "bn3")
            if bi == 0:
                idt = _q(frozen_bn(conv(x, sd[q + "downsample.0.weight"], stride=stride), sd, q + "downsample.1"), e)
            x = _q(F.relu(y + idt), e)
    return x


def resize_mask_nearest(mask, h, w):
    """reference models/backbone.py:101-103: F.interpolate(nearest) on the pad mask."""
    H, W = mask.shape[-2:]
    ii = (torch.arange(h) * (H / h)).floor().long().clamp(max=H - 1)
    jj = (torch.arange(w) * (W / w)).floor().long().clamp(max=W - 1)
    return mask[:, ii][:, :, jj]


def pos_sine(mask, num_pos_feats=128, temperature=10000.0):
    """reference models/position_encoding.py:71-94 (normalize=True, scale 2pi). -> (N,256,h,w)"""
    nm = (~mask).float()
    y = nm.cumsum(1)
    x = nm.cumsum(2)
    y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)

    def enc(e):
        pe = e[..., None] / dim_t
        out = torch.empty_like(pe)
        out[..., 0::2] = pe[..., 0::2].sin()
        out[..., 1::2] = pe[..., 1::2].cos()
        return out

    return torch.cat((enc(y), enc(x)), dim=3).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------- attention
def mha(q_in, k_in, v_in, sd, p, kpm, nhead=8):
    """torch.nn.MultiheadAttention, need_weights path, batch-major restatement.

    q_in (B,Lq,d), k_in/v_in (B,Lk,d), kpm (B,Lk) bool True=masked or None.
    Returns out (B,Lq,d) and head-averaged probabilities (B,Lq,Lk).
    Call sites: reference models/transformer.py:638-640, 712-719, 734-740.
    """
    d = q_in.shape[-1]
    hd = d // nhead
    W, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, W[:d], b[:d]) * (hd ** -0.5)
    k = F.linear(k_in, W[d:2 * d], b[d:2 * d])
    v = F.linear(v_in, W[2 * d:], b[2 * d:])
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    q = q.view(B, Lq, nhead, hd).transpose(1, 2)
    k = k.view(B, Lk, nhead, hd).transpose(1, 2)
    v = v.view(B, Lk, nhead, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    pr = s.softmax(-1)
    o = (pr @ v).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"]), pr.mean(1)


def layer_norm(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def ffn(x, sd, p):
    h = F.relu(F.linear(x, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"]))
    return F.linear(h, sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])


def encoder_layer(x, pos, kpm, sd, p):
    """reference models/transformer.py:629-646 (post-norm, eval mode)."""
    qk = x + pos
    a, _ = mha(qk, qk, x, sd, p + ".self_attn", kpm)
    x = layer_norm(x + a, sd, p + ".norm1")
    return layer_norm(x + ffn(x, sd, p), sd, p + ".norm2")


def decoder_layer(x, qp, mem, mem_pos, mem_kpm, q_kpm, sd, p, no_tsa=False):
    """reference models/transformer.py:684-751.  x,qp (B,T,d); mem,mem_pos (B*T,S,d)."""
    B, T, d = x.shape
    if no_tsa:  # :701-711, sequence length 1 => softmax over a single key
        a, w = mha((x + qp).reshape(B * T, 1, d), (x + qp).reshape(B * T, 1, d), x.reshape(B * T, 1, d),
                   sd, p + ".self_attn", None)
        a = a.view(B, T, d)
    else:
        a, w = mha(x + qp, x + qp, x, sd, p + ".self_attn", q_kpm)
    x = layer_norm(x + a, sd, p + ".norm1")
    a2, cw = mha((x + qp).reshape(B * T, 1, d), mem + mem_pos, mem, sd, p + ".cross_attn_image", mem_kpm)
    x = layer_norm(x + a2.view(B, T, d), sd, p + ".norm3")
    x = layer_norm(x + ffn(x, sd, p), sd, p + ".norm4")
    return x, w, cw


# ----------------------------------------------------------------------------- full model
def text_encode(sd, input_ids, attention_mask, roberta=None):
    """HF RobertaModel stays a library call (reference models/transformer.py:250-263)."""
    if roberta is None:
        from transformers import RobertaConfig, RobertaModel
        roberta = RobertaModel(RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1,
                                             pad_token_id=1, bos_token_id=0, eos_token_id=2, layer_norm_eps=1e-5))
        pre = "transformer.text_encoder."
        roberta.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
        roberta.eval()
    h = roberta(input_ids=input_ids, attention_mask=attention_mask).last_hidden_state  # (B,L,768)
    r = F.linear(h, sd["transformer.resizer.fc.weight"], sd["transformer.resizer.fc.bias"])
    r = layer_norm(r, sd, "transformer.resizer.layer_norm", eps=1e-12)  # reference :765
    return r, attention_mask.ne(1)


def forward(sd, frames_slow, mask_slow, frames_fast, mask_fast, durations, input_ids, attention_mask,
            stride, fast=True, no_tsa=False, nlayers=6, roberta=None):
    """Whole two-phase forward (reference models/tubedetr.py:93-254, models/transformer.py:195-491), eval mode.

    frames_* (N,3,H,W) fp32, mask_* (N,H,W) bool True=pad.  Returns dict of outputs plus the
    memory_cache tensors in the reference's (seq-first) shapes.
    """
    B, T = len(durations), max(durations)
    assert stride > 0, "only the temporal-stride path (reference default) is restated"
    n_clips = math.ceil(T / stride)
    W_in = sd["input_proj.weight"]
    d = W_in.shape[0]

    def backbone(fr, mk, grad=True):
        with torch.set_grad_enabled(grad and torch.is_grad_enabled()):
            f = resnet101_layer4(fr, sd)
        m = resize_mask_nearest(mk, f.shape[-2], f.shape[-1])
        # input_proj runs WITH grad on the fast features too (tubedetr.py:128-131): it gets weight-grad from both
        return F.conv2d(f, W_in, sd["input_proj.bias"]), m, pos_sine(m), f

    src, m_s, pos, feat_slow = backbone(frames_slow, mask_slow)
    n, _, h, w = src.shape
    HW = h * w
    assert n == B * n_clips
    m_s = m_s.clone()
    m_s[:, 0, 0] = False                                            # tubedetr.py:186
    x = src.flatten(2).transpose(1, 2)                              # (n,HW,d)
    pos = pos.flatten(2).transpose(1, 2)

    dur = torch.tensor(durations)
    tt = torch.arange(T)
    valid = tt[None, :] < dur[:, None]                              # (B,T)
    if fast:
        fsrc_all, m_f_all, _, feat_fast = backbone(frames_fast, mask_fast, grad=False)
        idx = (torch.cumsum(dur, 0) - dur)[:, None] + tt[None, :]   # index into packed fast frames
        idx = idx.clamp(max=frames_fast.shape[0] - 1)
        fsrc = fsrc_all[idx.flatten()].view(B, T, d, HW) * valid[:, :, None, None]
        fsrc = fsrc.view(B * T, d, HW).transpose(1, 2)              # (B*T,HW,d), zeros on pad frames
        m_t = torch.where(valid[:, :, None], m_f_all[idx.flatten()].view(B, T, HW), torch.ones(1, dtype=torch.bool))
    else:
        # tubedetr.py:171-178: slow mask replicated over the frames of its clip, ones on pad frames
        clip_of_t = (torch.arange(B)[:, None] * n_clips + tt[None, :] // stride)
        m_t = torch.where(valid[:, :, None], m_s.flatten(1)[clip_of_t.flatten()].view(B, T, HW),
                          torch.ones(1, dtype=torch.bool))
    m_t = m_t.reshape(B * T, HW).clone()
    m_t[:, 0] = False                                               # tubedetr.py:187

    # time queries (transformer.py:211-238)
    qp = sd["query_embed.weight"][0][None, None, :] + sd["transformer.time_embed.te"][:T, 0][None]
    qp = qp.expand(B, T, d)
    q_kpm = ~valid
    q_kpm[:, 0] = False

    txt, txt_kpm = text_encode(sd, input_ids, attention_mask, roberta)        # (B,L,d),(B,L)
    L = txt.shape[1]
    txt_rep = txt.repeat_interleave(n_clips, 0)                     # (n,L,d) transformer.py:286-293
    kpm_enc = torch.cat([m_s.flatten(1), txt_kpm.repeat_interleave(n_clips, 0)], 1)
    xe = torch.cat([x, txt_rep], 1)
    pe = torch.cat([pos, torch.zeros_like(txt_rep)], 1)
    enc_layers = [xe]
    for l in range(nlayers):
        xe = encoder_layer(xe, pe, kpm_enc, sd, f"transformer.encoder.layers.{l}")
        enc_layers.append(xe)

    # temporal replication (transformer.py:393-427): frame (b,t) <- clip (b, t//k)
    clip_of_t = (torch.arange(B)[:, None] * n_clips + tt[None, :] // stride).flatten()
    mem = xe[clip_of_t].clone()                                     # (B*T,S,d)
    mem_pos = pe[clip_of_t]
    kpm_dec = torch.cat([m_t, txt_kpm.repeat_interleave(T, 0)], 1)
    kpm_dec[:, 0] = False                                           # transformer.py:424
    if fast:
        fm = F.linear(fsrc, sd["transformer.fast_encoder.weight"], sd["transformer.fast_encoder.bias"])
        upd = F.linear(mem[:, :HW] + fm, sd["transformer.fast_residual.weight"], sd["transformer.fast_residual.bias"])
        mem = torch.cat([mem[:, :HW] + upd, mem[:, HW:]], 1)        # transformer.py:440-445

    # decoder (transformer.py:462-491, 548-605); tgt = 0
    y = torch.zeros(B, T, d)
    hs, ws, cws = [], [], []
    for l in range(nlayers):
        y, wl, cwl = decoder_layer(y, qp, mem, mem_pos, kpm_dec, q_kpm, sd,
                                   f"transformer.decoder.layers.{l}", no_tsa=no_tsa)
        hs.append(layer_norm(y, sd, "transformer.decoder.norm"))
        ws.append(wl)
        cws.append(cwl)
    hs = torch.stack(hs)                                            # (6,B,T,d)

    # heads (tubedetr.py:226-252), eval => no dropout
    def mlp(z, p, nl):
        for i in range(nl):
            z = F.linear(z, sd[f"{p}.layers.{i}.weight"], sd[f"{p}.layers.{i}.bias"])
            if i < nl - 1:
                z = F.relu(z)
        return z

    sted = mlp(hs, "sted_embed", 2)                                 # (6,B,T,2)
    boxes = mlp(hs.flatten(1, 2), "bbox_embed", 3).sigmoid()        # (6,B*T,4)
    out = {"pred_boxes": boxes[-1], "pred_sted": sted[-1], "weights": ws[-1], "ca_weights": cws[-1],
           "aux_outputs": [{"pred_boxes": boxes[i], "pred_sted": sted[i], "weights": ws[i], "ca_weights": cws[i]}
                           for i in range(nlayers - 1)]}
    cache = {"img_memory": mem.transpose(0, 1), "pos_embed": mem_pos.transpose(0, 1), "mask": kpm_dec,
             "query_embed": qp.transpose(0, 1), "query_mask": q_kpm, "text_memory_resized": txt_rep.transpose(0, 1),
             "text_attention_mask": txt_kpm.repeat_interleave(n_clips, 0), "feat_slow": feat_slow, "hs": hs,
             "src": src, "enc": xe, "mem": mem, "enc_layers": enc_layers, "txt": txt}
    return out, cache


# ----------------------------------------------------------------------------- criterion
def box_cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def giou_diag(a, b):
    """diag of reference util/box_ops.py:94-115 (paired boxes, xyxy)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, 2:], b[:, 2:]) - torch.max(a[:, :2], b[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = area_a + area_b - inter
    whc = (torch.max(a[:, 2:], b[:, 2:]) - torch.min(a[:, :2], b[:, :2])).clamp(min=0)
    areac = whc[:, 0] * whc[:, 1]
    return inter / union - (areac - union) / areac


def criterion(out, target_boxes, inter_idx, time_mask, keep, sigma=1.0, world_size=1):
    """reference models/tubedetr.py:270-460 + the keep-index of engine.py:83-102.

    target_boxes (K,4) cxcywh for the kept frames (concatenated), keep LongTensor (K,).
    Returns the 24-entry loss dict (main + _0.._4).
    """
    num_boxes = max(float(target_boxes.shape[0]) / world_size, 1.0)
    B, T = time_mask.shape
    tt = torch.arange(T)
    pos_map = torch.zeros(B, T, dtype=torch.bool)
    for k, (s, e) in enumerate(inter_idx):
        if s >= 0:
            pos_map[k, s:e + 1] = True
    eps = 1e-6

    def one(o):
        l = {}
        pb = o["pred_boxes"][keep]
        l["loss_bbox"] = (pb - target_boxes).abs().sum() / num_boxes
        l["loss_giou"] = (1 - giou_diag(box_cxcywh_to_xyxy(pb), box_cxcywh_to_xyxy(target_boxes))).sum() / num_boxes
        sted = o["pred_sted"].masked_fill(~time_mask[:, :, None], -1e32)
        tot = 0
        for c in range(2):
            tgt = torch.tensor([x[c] for x in inter_idx])
            g = (-((tt[None] - tgt[:, None]) ** 2) / (2 * sigma ** 2)).exp()
            g = F.normalize(g + eps, p=1, dim=1)
            p = sted[:, :, c].softmax(1)
            tot = tot + p * ((p + eps) / g).log() * time_mask
        l["loss_sted"] = tot.mean()
        neg = pos_map | ~time_mask
        ga = -(1 - o["weights"] + eps).log()
        ga = ga.masked_fill(neg[:, :, None], 0)
        l["loss_guided_attn"] = (ga.sum(2) / ((~neg).sum(1) + eps)[:, None]).sum(1).mean()
        return l

    losses = one(out)
    for i, aux in enumerate(out["aux_outputs"]):
        losses.update({f"{k}_{i}": v for k, v in one(aux).items()})
    return losses


WEIGHT_DICT_BASE = {"loss_bbox": 5.0, "loss_giou": 2.0, "loss_sted": 10.0, "loss_guided_attn": 1.0}


def weight_dict(nlayers=6):
    """reference models/tubedetr.py:482-494 with main.py defaults."""
    wd = dict(WEIGHT_DICT_BASE)
    for i in range(nlayers - 1):
        wd.update({f"{k}_{i}": v for k, v in WEIGHT_DICT_BASE.items()})
    return wd
