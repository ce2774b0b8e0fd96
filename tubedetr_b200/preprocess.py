"""Input pipeline on the GPU.

The reference decodes a clip with ffmpeg to packed rgb24 frames (datasets/vidstg.py:104-116), then per frame on the CPU: resize
(datasets/video_transforms.py:128-215, shorter side -> `resolution`, capped by max_size), ToTensor (/255), Normalize(ImageNet mean / std),
and finally pads the clips of a batch into a NestedTensor (util/misc.py:142-172).  `clips_to_nested` does all of that after the decode in
one kernel launch per clip: upload the uint8 frames (3 bytes per pixel instead of 12), and tdb_frames_preprocess writes the resized,
normalised fp32 frames straight into their slot of the padded batch tensor together with the pad mask.
"""
import torch

from . import kernels as K
from .model import NestedTensor

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # datasets/video_transforms.py:41


def target_size(h0, w0, size, max_size=None):
    """(h, w) after the reference's aspect-preserving resize (video_transforms.py:131-150)"""
    if max_size is not None:
        mn, mx = float(min(w0, h0)), float(max(w0, h0))
        if mx / mn * size > max_size:
            size = int(round(max_size * mn / mx))
    if (w0 <= h0 and w0 == size) or (h0 <= w0 and h0 == size):
        return h0, w0
    if w0 < h0:
        return int(size * h0 / w0), size
    return size, int(size * w0 / h0)


def clips_to_nested(clips_u8, size, max_size=None, stride=1, device="cuda"):
    """clips_u8: list of uint8 tensors (T_i, H_i, W_i, 3) (the decoder's rgb24 output).  Returns (samples_fast, samples_slow): the padded
    NestedTensors of all frames and of every `stride`-th frame (datasets/vidstg.py:250-252), frames (sum T, 3, Hmax, Wmax) fp32."""
    sizes = [target_size(c.shape[1], c.shape[2], size, max_size) for c in clips_u8]
    Hp, Wp = max(s[0] for s in sizes), max(s[1] for s in sizes)
    n = sum(c.shape[0] for c in clips_u8)
    frames = torch.empty(n, 3, Hp, Wp, dtype=torch.float32, device=device)
    mask = torch.empty(n, Hp, Wp, dtype=torch.uint8, device=device)
    o, slow_idx = 0, []
    for c, (h, w) in zip(clips_u8, sizes):
        cu = c.to(device, non_blocking=True).contiguous()
        T, H0, W0, _ = cu.shape
        K.frames_preprocess(cu, frames[o:o + T], mask[o:o + T], T, H0, W0, h, w, Hp, Wp, MEAN, STD)
        slow_idx += list(range(o, o + T, stride))
        o += T
    fast = NestedTensor(frames, mask.bool())
    if stride == 1:
        return fast, fast
    idx = torch.tensor(slow_idx, device=device)
    return fast, NestedTensor(frames.index_select(0, idx), mask.bool().index_select(0, idx))
