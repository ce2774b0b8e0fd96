"""ResNet-101 + FrozenBatchNorm2d feature extractor on the tcgen05 GEMM (forward, dgrad, wgrad).

Replaces reference models/backbone.py:60-70 (FrozenBatchNorm2d), :73-124 (BackboneBase/Backbone over torchvision
resnet101).  Topology restated from torchvision ResNet v1.5: stem 7x7/2 + maxpool, bottlenecks (3,4,23,3), stride on the
3x3 conv of the first block of layer2-4.  Trainable: conv weights of layer2-4 only (reference :82-89).

Data layout: NHWC bf16 "pixel rows" [N*H*W][C].  The input of every stride-1 3x3 conv lives in a zero-haloed
[N*(H+2)*(W+2)][C] buffer so the convolution is 9 row-shifted taps of one GEMM.  FrozenBN is an epilogue
(scale,shift per channel) in forward and folded into the weights (dgrad) / the split-K reduce (wgrad) in backward; ReLU
backward is the `mask` epilogue of the producing dgrad GEMM, so no elementwise pass ever touches HBM on its own.
This module only sequences kernel launches; every FLOP runs in libtdb.so.
"""
import os

import torch

from . import kernels as K
from .gemm import DYNAMIC_TILES, REMAP_C2P, REMAP_C2P1, REMAP_C2S, REMAP_P2C, REMAP_S2C, effective_splits, gemm, splitk_reduce

STAGES = ((64, 3, 1), (128, 4, 2), (256, 23, 2), (512, 3, 2))  # width, blocks, stride of first block
BN_EPS = 1e-5


WGRAD_LAG = max(1, min(2, int(os.environ.get("TDB_WGRAD_LAG", "1"))))
# experiment switch: cap the persistent grids of the backbone BACKWARD GEMMs so that the text encoder's tiny backward kernels (side
# stream, concurrent with the backbone backward) find idle SMs instead of delaying a GEMM CTA that needs a whole SM (0 = no cap)
BWD_MAX_CTAS = int(os.environ.get("TDB_BB_BWD_MAX_CTAS", "0"))
STEM_FUSED = os.environ.get("TDB_STEM_FUSED", "1") != "0"    # one kernel for conv1 + FrozenBN + ReLU + maxpool (tdb_stem.cu)
S2_IMPLICIT = os.environ.get("TDB_S2_IMPLICIT", "1") != "0"  # stride-2 3x3 convs as implicit GEMMs over a space-to-depth layout (no im2col matrix)
# first block of a stage: out = relu(bn3(conv3(y2)) + bnd(convd(xs))) as ONE GEMM over K = width + cin: y2 and the (subsampled) block
# input sit side by side in one matrix [rows, width + cin], the FrozenBN scales are folded into the concatenated weights (the
# `ws` copies the backward already uses), the shifts are summed.  Removes the downsample GEMM and the write + re-read of its
# [rows, 4 * width] output as a residual (layer1.0 at 125 frames: 0.99 GB of HBM traffic)
DS_JOINT = os.environ.get("TDB_DS_JOINT", "1") != "0"
# zero-haloed grids of the stride-1 3x3 convs with ONE shared halo row / column per image ((H + 1) x (W + 1) positions instead of
# (H + 2) x (W + 2): the cell right of a row's last pixel is the next row's halo cell): 8 % (22 x 22) .. 15 % (11 x 11) fewer GEMM rows
HALO1 = os.environ.get("TDB_HALO1", "1") != "0"
PADH = 1 if HALO1 else 2
C2P, P2C = (REMAP_C2P1, REMAP_S2C) if HALO1 else (REMAP_C2P, REMAP_P2C)


def _dyn(env, auto):
    """scheduling hint for the GEMMs of one pass: TDB_GEMM_FLAG_DYNAMIC_TILES (tiles handed out by cluster launch control, i.e. work
    stealing) when that pass shares the GPU with other streams' kernels.  Measured on B200 (tools/clc_hog.py): +5 % on an idle GPU,
    -25..45 % when 8-32 SMs are held by a foreign kernel, because a static persistent schedule makes the CTAs that start late run
    their full share after everybody else has finished.  env: 0 / 1 / unset = `auto`."""
    v = os.environ.get(env)
    on = auto if v is None else v != "0"
    return DYNAMIC_TILES if on else 0


def _multi_rank():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def conv_out(h, k, s, p):
    return (h + 2 * p - k) // s + 1


class ResNet101Engine:
    def __init__(self):
        self._bufs = {}
        self._gen = {}        # tag -> generation, bumped by every forward that (re)writes the tag's buffers
        self._wcache = {}
        self._bn = None
        self._bn_key = None
        self.last_hw = None
        self.stage_callback = None   # callable(stage index) fired by backward() when a stage's weight gradients are all enqueued
        self._bf = 0                 # scheduling hint of the current backward pass (see _dyn)

    def __deepcopy__(self, memo):  # EMA copies of the model (reference main.py:370) get a fresh, empty engine
        return ResNet101Engine()

    # ------------------------------------------------------------------ buffers
    def buf(self, tag, shape, dtype=torch.bfloat16, zero=False, device="cuda"):
        """Scratch / activation buffer of the engine, ONE allocation per tag sized to the largest request seen: the reference's
        data pipeline changes the frame count and the resolution nearly every iteration (durations up to video_max_len,
        RandomResize), so keying by shape would grow without bound.  A view of the tag's storage with the requested shape is
        returned; `zero` buffers (zero-haloed conv inputs: kernels only ever write their interior rows) are re-zeroed whenever
        the shape changes."""
        shape = tuple(int(x) for x in shape)
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        ent = self._bufs.get(tag)
        if ent is None or ent[0].numel() < nbytes:
            store = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
            ent = [store, None, None]
            self._bufs[tag] = ent
        if ent[1] != (shape, dtype):
            view = ent[0][:nbytes].view(dtype).view(shape)
            if zero:
                view.zero_()
            ent[1], ent[2] = (shape, dtype), view
        return ent[2]

    def allocated_bytes(self):
        return sum(e[0].numel() for e in self._bufs.values())

    def check_generation(self, ctx):
        """Saved activations are engine-owned buffers that the next forward with the same tag overwrites in place (no autograd
        version counter sees raw kernel writes).  One forward per backward is the contract; anything else
        (gradient accumulation over a combined loss, checkpoint recompute, an interleaved forward) raises instead of returning
        silently wrong gradients."""
        if self._gen.get(ctx["tag"]) != ctx["gen"]:
            raise RuntimeError(f"tubedetr_b200: backbone activations of tag '{ctx['tag']}' were overwritten by a later forward "
                               "(generation %s, now %s): run backward before the next forward of the same model" % (ctx["gen"], self._gen.get(ctx["tag"])))

    # ------------------------------------------------------------------ weights
    def prepare(self, sd, prefix="backbone.0.body."):
        """sd: name -> tensor (module parameters/buffers, fp32 cuda).  Builds bf16 GEMM-layout weights
        (cached on tensor version) and folded FrozenBN scale/shift."""
        names = [("conv1", "bn1", 3, 64, 49)]
        for li, (width, nb, _) in enumerate(STAGES, start=1):
            cin = 64 if li == 1 else STAGES[li - 2][0] * 4
            for bi in range(nb):
                p = f"layer{li}.{bi}."
                bc = cin if bi == 0 else width * 4
                names += [(p + "conv1", p + "bn1", bc, width, 1), (p + "conv2", p + "bn2", width, width, 9),
                          (p + "conv3", p + "bn3", width, width * 4, 1)]
                if bi == 0:
                    names.append((p + "downsample.0", p + "downsample.1", bc, width * 4, 1))
        bn_key = tuple(sd[prefix + bn + ".running_var"]._version + sd[prefix + bn + ".weight"]._version
                       + sd[prefix + bn + ".bias"]._version + sd[prefix + bn + ".running_mean"]._version
                       for _, bn, _, _, _ in names) + (sd[prefix + "bn1.weight"].data_ptr(),)
        if self._bn is None or self._bn_key != bn_key:
            self._bn = {}
            for conv, bn, _, _, _ in names:
                q = prefix + bn
                scale = (sd[q + ".weight"] * torch.rsqrt(sd[q + ".running_var"] + BN_EPS)).float().contiguous()
                shift = (sd[q + ".bias"] - sd[q + ".running_mean"] * scale).float().contiguous()
                self._bn[conv] = (scale, shift)
            self._bn_key = bn_key
            self._wcache = {}
        W = {}
        for conv, bn, cin, cout, taps in names:
            w = sd[prefix + conv + ".weight"]
            ck = (w.data_ptr(), w._version)
            ent = self._wcache.get(conv)
            if ent is None or ent[0] != ck:
                kpad = 192 if taps == 49 else taps * cin
                wb = self.buf("w:" + conv, (cout, kpad))
                ws = self.buf("ws:" + conv, (cout, kpad))
                K.prep_weight(w.detach().float().contiguous(), wb, ws, self._bn[conv][0], cout, cin, taps, kpad)
                ent = (ck, wb, ws)
                self._wcache[conv] = ent
            W[conv] = (ent[1], ent[2], self._bn[conv][0], self._bn[conv][1])
        for li, (width, _, _) in enumerate(STAGES, start=1):
            p = f"layer{li}.0."
            e3, ed = self._wcache[p + "conv3"], self._wcache[p + "downsample.0"]
            ck = (e3[0], ed[0])
            ent = self._wcache.get("j:" + p)
            if ent is None or ent[0] != ck:
                wj = torch.cat([e3[2], ed[2]], 1).contiguous()                      # [cout, width + cin], FrozenBN scales folded in
                ent = (ck, wj, (self._bn[p + "conv3"][1] + self._bn[p + "downsample.0"][1]).contiguous())
                self._wcache["j:" + p] = ent
            W["j:" + p] = (ent[1], ent[2])
        # fused stem: conv1 weight in K order (c, kh, kw padded to 8) -- frozen, so a handful of torch ops once per weight version
        w1 = sd[prefix + "conv1.weight"]
        ck = (w1.data_ptr(), w1._version)
        ent = self._wcache.get("conv1:fused")
        if ent is None or ent[0] != ck:
            wk = torch.zeros(64, 192, dtype=torch.bfloat16, device=w1.device)
            wk[:, :168] = torch.nn.functional.pad(w1.detach().float(), (0, 1)).reshape(64, 168).to(torch.bfloat16)
            ent = (ck, wk)
            self._wcache["conv1:fused"] = ent
        W["conv1:fused"] = ent[1]
        return W

    # ------------------------------------------------------------------ forward
    def forward(self, frames, W, save, tag, l2_chunk=None, n_keep=None):
        """frames fp32 [N,3,H,W] cuda, or a LIST of such tensors that is processed as ONE batch (concatenated along N without
        a copy: the stem gathers each source into one row buffer).  Returns (feat rows bf16 [N*h*w, 2048], h, w, ctx or None).

        n_keep (with save): only the first n_keep frames are differentiated -- ctx records row-prefix views of the saved
        activations, so backward() runs on n_keep frames.  This is how the slow (with grad) and fast (no grad) passes of
        reference models/tubedetr.py:120-131 share one launch sequence: 125 frames per GEMM instead of 25 + 100 fills the
        148 SMs better (wave quantisation) and halves the launch count; per-frame results are bit-identical.

        l2_chunk (no-grad pass only): run stem + layer1 + layer2 -- the HBM-bound, high-resolution part -- over chunks of that
        many frames with per-chunk buffers that are reused by every chunk, so the intermediate activations of a chunk
        (~8 MB/frame at res 352) live in the 126 MB L2 instead of streaming through HBM; layer3/4 then run on the whole batch."""
        srcs = [f.contiguous() for f in (frames if isinstance(frames, (list, tuple)) else [frames])]
        self._gen[tag] = self._gen.get(tag, 0) + 1
        _, _, H, Wd = srcs[0].shape
        N = sum(f.shape[0] for f in srcs)
        if l2_chunk and not save and N > l2_chunk and len(srcs) == 1:
            frames = srcs[0]
            big, h2, w2 = None, 0, 0
            for n0 in range(0, N, l2_chunk):
                n = min(l2_chunk, N - n0)
                x, h2, w2, _ = self._stages(self._stem([frames[n0:n0 + n]], W, f"{tag}c{n}"), n, W, False, f"{tag}c{n}", 1, 2,
                                            H, Wd, out_into=(big, n0, N))
                big = x if big is None else big
            x, h, w, ctx = self._stages((big, h2, w2), N, W, False, tag, 3, 4, H, Wd)
        else:
            x, h, w, ctx = self._stages(self._stem(srcs, W, tag), N, W, save, tag, 1, 4, H, Wd, n_keep=n_keep)
        self.last_hw = (h, w)
        return x, h, w, ctx

    def _stem(self, srcs, W, tag):
        _, _, H, Wd = srcs[0].shape
        N = sum(f.shape[0] for f in srcs)
        H1, W1 = conv_out(H, 7, 2, 3), conv_out(Wd, 7, 2, 3)
        H2, W2 = conv_out(H1, 3, 2, 1), conv_out(W1, 3, 2, 1)
        wb, _, sc, sh = W["conv1"]
        if STEM_FUSED:
            if DS_JOINT:      # the pooled rows land in the right half of layer1.0's [rows, conv2 output | block input] matrix
                self._l1j = self.buf(tag + ":L1J", (N * H2 * W2, 128))
                x = self._l1j[:, 64:]
            else:
                self._l1j = None
                x = self.buf(tag + ":pool", (N * H2 * W2, 64))
            base = 0
            for frames in srcs:
                assert frames.shape[2:] == (H, Wd), "frames of one backbone batch must share their spatial size"
                n = frames.shape[0]
                K.stem_fused(frames, W["conv1:fused"], sc, sh, x[base * H2 * W2:(base + n) * H2 * W2], n, H, Wd)
                base += n
            return x, H2, W2
        self._l1j = None
        stem = self.buf(tag + ":stem", (N * H1 * W1, 64))
        chunk = max(1, min(N, (1 << 28) // (H1 * W1 * 192 * 2)))
        col = self.buf(tag + ":stemcol", (chunk * H1 * W1, 192))
        base = 0
        for frames in srcs:
            assert frames.shape[2:] == (H, Wd), "frames of one backbone batch must share their spatial size"
            for n0 in range(0, frames.shape[0], chunk):
                n = min(chunk, frames.shape[0] - n0)
                K.stem_im2col(frames[n0:n0 + n], col, n, H, Wd)
                gemm(col, wb, stem[(base + n0) * H1 * W1:(base + n0 + n) * H1 * W1], n * H1 * W1, 64, 192, scale=sc, bias=sh,
                     relu=True)
            base += frames.shape[0]
        x = self.buf(tag + ":pool", (N * H2 * W2, 64))
        K.maxpool3x3s2(stem, x, N, H1, W1, 64)
        return x, H2, W2

    def _stages(self, xhw, N, W, save, tag, li_lo, li_hi, H, Wd, out_into=None, n_keep=None):
        """bottleneck stages li_lo..li_hi on pixel rows x [N*h*w, C].  out_into=(big, n0, Ntot): the last block writes its
        rows into rows [n0*ho*wo, ...) of a full-batch buffer (allocated on first use) instead of a private one."""
        x, h, w = xhw
        ff = _dyn("TDB_CLC_FWD", False)        # the text encoder's forward runs beside this pass (side stream)
        nk = N if n_keep is None else n_keep
        ctx = {"N": nk, "blocks": [], "tag": tag, "gen": self._gen.get(tag)} if save else None
        big = None
        for li, (width, nb, stride0) in enumerate(STAGES, start=1):
            if li < li_lo or li > li_hi:
                continue
            for bi in range(nb):
                name = f"layer{li}.{bi}."
                stride = stride0 if bi == 0 else 1
                keep = save and li >= 2                     # frozen stem/layer1 are never differentiated
                btag = f"{tag}:{name}" if keep else f"{tag}:L{li}"
                cin = x.shape[1]
                cout = width * 4
                ho, wo = (conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)) if stride == 2 else (h, w)
                R, Ro = N * h * w, N * ho * wo
                w1, _, s1, b1 = W[name + "conv1"]
                w2, _, s2, b2 = W[name + "conv2"]
                w3, _, s3, b3 = W[name + "conv3"]
                rec = {"name": name, "x": x, "h": h, "w": w, "ho": ho, "wo": wo, "stride": stride, "width": width,
                       "cin": cin, "cout": cout, "first": bi == 0}
                J = None
                if DS_JOINT and bi == 0:
                    if stride == 2:
                        J = self.buf(btag + "J", (Ro, width + cin))
                    elif getattr(self, "_l1j", None) is not None and x.data_ptr() == self._l1j.data_ptr() + 2 * width:
                        J = self._l1j                                            # layer1.0: the stem wrote x into J[:, width:]
                y2buf = (lambda: J[:, :width]) if J is not None else (lambda: self.buf(btag + "y2", (Ro, width)))
                if stride == 1:
                    Rp = N * (h + PADH) * (w + PADH)
                    y1 = self.buf(btag + "y1p", (Rp, width), zero=True)
                    gemm(x, w1, y1, R, width, cin, scale=s1, bias=b1, relu=True, remap=C2P, img_hw=(h, w), debug_flags=ff)
                    y2 = y2buf()
                    wp = w + PADH
                    taps = [(kh - 1) * wp + (kw - 1) for kh in range(3) for kw in range(3)]
                    gemm(y1, w2, y2, Rp, width, width, ntaps=9, a_off1=taps, b_off0=[t * width for t in range(9)],
                         scale=s2, bias=b2, relu=True, remap=P2C, img_hw=(h, w), debug_flags=ff)
                elif S2_IMPLICIT and not (keep and nk == N):
                    # stride-2 3x3 conv WITHOUT an im2col matrix: conv1 writes its rows space-to-depth (4 parity planes side by side,
                    # one-pixel zero halo on the top / left of every plane); tap (kh, kw) of the convolution then is the plane
                    # ((kh - 1) & 1, (kw - 1) & 1) shifted by (-1 if k == 0 else 0) rows / columns: 9 constant (row, column) offsets
                    # of ONE implicit GEMM (same tap-major K order as the im2col matrix, so the result is unchanged).
                    ohp, owp = ho + 1, wo + 1
                    y1s = self.buf(f"{tag}:L{li}y1s", (N * ohp * owp, 4 * width), zero=True)
                    gemm(x, w1, y1s, R, width, cin, scale=s1, bias=b1, relu=True, remap=REMAP_C2S, img_hw=(h, w), debug_flags=ff)
                    a0, a1 = [], []
                    for kh in range(3):
                        for kw in range(3):
                            a0.append((((kh - 1) & 1) * 2 + ((kw - 1) & 1)) * width)
                            a1.append((-1 if kh == 0 else 0) * owp + (-1 if kw == 0 else 0))
                    y2 = y2buf()
                    gemm(y1s, w2, y2, N * ohp * owp, width, width, ntaps=9, a_off0=a0, a_off1=a1, b_off0=[t * width for t in range(9)],
                         scale=s2, bias=b2, relu=True, remap=REMAP_S2C, img_hw=(ho, wo), debug_flags=ff)
                    y1 = None
                    if keep:
                        # the differentiated frames (a row prefix) also get the compact conv1 output and the im2col matrix their
                        # backward reads (ReLU mask, weight gradient): 1/5 of the old cost at 25 of 125 frames
                        y1 = self.buf(btag + "y1", (nk * h * w, width))
                        gemm(x[:nk * h * w], w1, y1, nk * h * w, width, cin, scale=s1, bias=b1, relu=True, debug_flags=ff)
                        colb = self.buf(btag + "col", (nk * ho * wo, 9 * width))
                        K.im2col3x3s2(y1, colb, nk, h, w, width)
                        rec["col"] = colb
                else:
                    y1 = self.buf(btag + "y1", (R, width))
                    gemm(x, w1, y1, R, width, cin, scale=s1, bias=b1, relu=True, debug_flags=ff)
                    colb = self.buf(btag + "col", (Ro, 9 * width))
                    K.im2col3x3s2(y1, colb, N, h, w, width)
                    y2 = y2buf()
                    gemm(colb, w2, y2, Ro, width, 9 * width, scale=s2, bias=b2, relu=True, debug_flags=ff)
                    rec["col"] = colb
                if bi == 0 and J is not None:
                    xs = x
                    if stride == 2:
                        xs = J[:, width:]
                        K.subsample2(x, xs, N, h, w, cin)
                    rec["xs"] = xs
                    idt = None
                elif bi == 0:
                    wd, _, sd_, bd = W[name + "downsample.0"]
                    xs = x
                    if stride == 2:
                        xs = self.buf(btag + "xs", (Ro, cin))
                        K.subsample2(x, xs, N, h, w, cin)
                    idt = self.buf(btag + "idt", (Ro, cout))
                    gemm(xs, wd, idt, Ro, cout, cin, scale=sd_, bias=bd, debug_flags=ff)
                    rec["xs"] = xs
                else:
                    idt = x
                last = out_into is not None and li == li_hi and bi == nb - 1
                if last:
                    big, n0, ntot = out_into
                    if big is None:
                        big = self.buf(f"{tag}:big{li}", (ntot * ho * wo, cout))
                    out = big[n0 * ho * wo:(n0 + N) * ho * wo]
                else:
                    # ping-pong the block output so the no-grad pass needs two buffers per stage
                    out = self.buf(f"{btag}out{0 if keep else bi % 2}", (Ro, cout))
                if J is not None:
                    wj, bj = W["j:" + name]
                    gemm(J, wj, out, Ro, cout, width + cin, bias=bj, relu=True, debug_flags=ff)
                else:
                    gemm(y2, w3, out, Ro, cout, width, scale=s3, bias=b3, residual=idt, relu=True, debug_flags=ff)
                rec.update(y1=y1, y2=y2, out=out)
                if keep:
                    if nk != N:      # differentiate the first nk frames only: row-prefix views of every saved activation
                        pre = {"x": nk * h * w, "y1": nk * (h + PADH) * (w + PADH) if stride == 1 else nk * h * w, "y2": nk * ho * wo,
                               "out": nk * ho * wo, "col": nk * ho * wo, "xs": nk * ho * wo}
                        for key, rows in pre.items():
                            if key in rec:
                                rec[key] = rec[key][:rows]
                    ctx["blocks"].append(rec)
                x, h, w = out, ho, wo
        if big is not None:
            x = big
        return x, h, w, ctx

    # ------------------------------------------------------------------ backward
    def _wgrad(self, g, xin, M, Ncols, Kred, rowscale, out, taps=1, z_b_off1=None):
        """out (fp32, torch weight layout) = rowscale[:,None] * g^T @ xin  (both operands read MN-major)."""
        nz = 9 if z_b_off1 is not None else 0
        ntot = Ncols * (9 if nz else 1)
        work = ((M + 127) // 128) * max(1, Ncols // 256 if Ncols % 256 == 0 else Ncols // 64) * max(nz, 1)
        want = max(1, min((148 + work - 1) // work, 32))
        s = effective_splits(Kred, want)
        part = torch.empty(s, M, ntot, dtype=torch.float32, device=g.device)
        gemm(g, xin, part, M, Ncols, Kred, a_major=1, b_major=1, nz=nz, z_b_off1=z_b_off1,
             z_out_col=[t * Ncols for t in range(9)] if nz else None, splits=want, max_ctas=BWD_MAX_CTAS, debug_flags=self._bf)
        splitk_reduce(part, s, M, ntot, out, rowscale=rowscale, taps=9 if (nz or taps == 9) else 1)

    def backward(self, ctx, W, g_out, grads, prefix="backbone.0.body."):
        """g_out: bf16 [N*h*w, 2048] = dL/d(pre-ReLU output of layer4.2), already masked by (feat > 0).
        grads: dict name -> fp32 tensor (torch layout) written in place for every layer2-4 conv weight."""
        from .ops import wgrad_scope
        self.check_generation(ctx)
        bf = self._bf = _dyn("TDB_CLC_BWD", _multi_rank())   # gradient all-reduces (and the text encoder's backward) run beside this pass
        N, tag = ctx["N"], ctx["tag"]
        blocks = ctx["blocks"]
        # Weight gradients run on a side stream, concurrent with the dgrad chain.  The scratch gradients are double buffered by
        # block parity (the block-output gradient triple buffered), so the main stream could run up to TWO blocks ahead of the
        # side stream (TDB_WGRAD_LAG=2).  Measured on B200 (cfg-2 step, same box): lag 1 (wait for the previous block's weight
        # gradients before starting a block) 20.68 ms, lag 2 21.25 ms -- the extra run-ahead only adds contention between the
        # dgrad chain and a longer queue of split-K wgrad CTAs -- so lag 1 is the default.
        sc = wgrad_scope(g_out.device)
        pending = []
        for i in range(len(blocks) - 1, -1, -1):
            while len(pending) >= WGRAD_LAG:
                sc.wait(pending.pop(0))
            par = i % 2
            r = blocks[i]
            name, h, w, ho, wo, width, cin, cout = (r[k] for k in ("name", "h", "w", "ho", "wo", "width", "cin", "cout"))
            R, Ro = N * h * w, N * ho * wo
            _, w1s, s1, _ = W[name + "conv1"]
            _, w2s, s2, _ = W[name + "conv2"]
            _, w3s, s3, _ = W[name + "conv3"]
            x, y1, y2 = r["x"], r["y1"], r["y2"]
            last = i == 0                                   # layer2.0: its input is the frozen layer1 output
            # ---- conv3
            with sc:
                self._wgrad(g_out, y2, cout, width, Ro, s3, grads[prefix + name + "conv3.weight"])
            if r["stride"] == 1:
                Rp = N * (h + PADH) * (w + PADH)
                wp = w + PADH
                taps = [(kh - 1) * wp + (kw - 1) for kh in range(3) for kw in range(3)]
                g2 = self.buf(f"{tag}:g2p{par}", (Rp, width), zero=True)
                gemm(g_out, w3s, g2, Ro, width, cout, b_major=1, mask=y2, remap=C2P, img_hw=(h, w), max_ctas=BWD_MAX_CTAS, debug_flags=bf)
                # ---- conv2 (implicit 3x3 over the haloed grid)
                with sc:
                    self._wgrad(g2, y1, width, width, Rp, s2, grads[prefix + name + "conv2.weight"], z_b_off1=taps)
                g1 = self.buf(f"{tag}:g1{par}", (R, width))
                gemm(g2, w2s, g1, Rp, width, width, b_major=1, ntaps=9, a_off1=[-t for t in taps],
                     b_off0=[t * width for t in range(9)], mask=y1, remap=P2C, img_hw=(h, w), max_ctas=BWD_MAX_CTAS, debug_flags=bf)
            else:
                g2 = self.buf(f"{tag}:g2c{par}", (Ro, width))
                gemm(g_out, w3s, g2, Ro, width, cout, b_major=1, mask=y2, max_ctas=BWD_MAX_CTAS, debug_flags=bf)
                with sc:
                    self._wgrad(g2, r["col"], width, 9 * width, Ro, s2, grads[prefix + name + "conv2.weight"], taps=9)
                dcol = self.buf(f"{tag}:dcol{par}", (Ro, 9 * width))
                gemm(g2, w2s, dcol, Ro, 9 * width, width, b_major=1, max_ctas=BWD_MAX_CTAS, debug_flags=bf)
                g1 = self.buf(f"{tag}:g1{par}", (R, width))
                K.col2im3x3s2_mask(dcol, y1, g1, N, h, w, width)
            # ---- conv1
            with sc:
                self._wgrad(g1, x, width, cin, R, s1, grads[prefix + name + "conv1.weight"])
            resid = g_out
            if r["first"]:
                _, wds, sdn, _ = W[name + "downsample.0"]
                with sc:
                    self._wgrad(g_out, r["xs"], cout, cin, Ro, sdn, grads[prefix + name + "downsample.0.weight"])
                if not last:
                    dxs = self.buf(f"{tag}:dxs{par}", (Ro, cin))
                    gemm(g_out, wds, dxs, Ro, cin, cout, b_major=1, max_ctas=BWD_MAX_CTAS, debug_flags=bf)
                    resid = dxs
                    if r["stride"] == 2:
                        resid = self.buf(f"{tag}:dxsu{par}", (R, cin))
                        K.upsample2_zero(dxs, resid, N, h, w, cin)
            if last:
                break
            if r["first"] and self.stage_callback is not None:
                # every weight gradient of this stage has been enqueued (main + wgrad side stream): join, then let the data-parallel
                # driver all-reduce the stage's slice of the flat gradient buffer while the earlier stages still run
                while pending:
                    sc.wait(pending.pop(0))
                sc.join()
                self.stage_callback(int(name[5]))
            gprev = self.buf(f"{tag}:gout{i % 3}", (R, cin))
            gemm(g1, w1s, gprev, R, cin, width, b_major=1, residual=resid, mask=x, max_ctas=BWD_MAX_CTAS, debug_flags=bf)
            g_out = gprev
            pending.append(sc.mark())
        sc.join()
        if self.stage_callback is not None:
            self.stage_callback(int(blocks[0]["name"][5]))
        return grads
