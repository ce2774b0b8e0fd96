"""Synthetic VidSTG-shaped batches: there is no dataset offline.

Used by the golden-fixture generator, the parity tests and bench.py so all three see the
same inputs for a given (shape, seed).
"""
import torch


def make_batch(durations, res, stride, ntok, seed=0):
    """durations: frames per video; res: (H,W) or list of per-video (H,W); ntok: caption length per video.

    Returns dict with
      clips          list of (3,T_i,H_i,W_i) fp32 N(0,1)   (real data is ImageNet-normalised)
      input_ids      (B,Lmax) long, BOS=0 ... EOS=2, pad=1;  attention_mask (B,Lmax) long
      inter_idx      [[s,e]] annotated moment = [T_i//4, 3*T_i//4]
      target_boxes   (K,4) cxcywh, one per kept frame;  time_mask (B,Tmax) bool
    """
    g = torch.Generator().manual_seed(1000 + seed)
    B = len(durations)
    if isinstance(res[0], int):
        res = [tuple(res)] * B
    clips = [torch.randn(3, t, h, w, generator=g) for t, (h, w) in zip(durations, res)]
    L = max(ntok)
    ids = torch.full((B, L), 1, dtype=torch.long)
    am = torch.zeros(B, L, dtype=torch.long)
    for b, n in enumerate(ntok):
        ids[b, :n] = torch.randint(3, 50000, (n,), generator=g)
        ids[b, 0], ids[b, n - 1] = 0, 2
        am[b, :n] = 1
    T = max(durations)
    inter_idx = [[t // 4, (3 * t) // 4] for t in durations]
    K = sum(e - s + 1 for s, e in inter_idx)
    cxcy = torch.rand(K, 2, generator=g) * 0.5 + 0.25
    wh = torch.rand(K, 2, generator=g) * 0.3 + 0.1
    time_mask = torch.zeros(B, T, dtype=torch.bool)
    for b, t in enumerate(durations):
        time_mask[b, :t] = True
    return {"clips": clips, "input_ids": ids, "attention_mask": am, "inter_idx": inter_idx,
            "target_boxes": torch.cat([cxcy, wh], 1), "time_mask": time_mask, "durations": list(durations),
            "stride": stride}


def pack_clips(clips):
    """Pad-and-pack a list of (3,T,H,W) clips into frames (sum T,3,Hmax,Wmax) + bool pad mask (sum T,Hmax,Wmax).

    Same result as the reference's NestedTensor.from_tensor_list on videos (util/misc.py:142-172).
    """
    H = max(c.shape[2] for c in clips)
    W = max(c.shape[3] for c in clips)
    n = sum(c.shape[1] for c in clips)
    frames = torch.zeros(n, 3, H, W, dtype=clips[0].dtype, device=clips[0].device)
    mask = torch.ones(n, H, W, dtype=torch.bool, device=clips[0].device)
    o = 0
    for c in clips:
        t, h, w = c.shape[1:]
        frames[o:o + t, :, :h, :w] = c.transpose(0, 1)
        mask[o:o + t, :h, :w] = False
        o += t
    return frames, mask
