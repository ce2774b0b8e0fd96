"""Deterministic synthetic weights for the TubeDETR state_dict.

There are no pretrained checkpoints offline, so parity tests, the bench and the golden
fixtures all use the same seeded random state_dict.  Values depend only on
(seed, parameter name, shape) -- not on iteration order or on any model class -- so the
reference model (in the fixture generator), the CPU oracle and the CUDA path all see
bit-identical fp32 weights.

The init laws are chosen so a random-init ResNet-101 with FrozenBatchNorm keeps O(1)
activations (He conv init, small bn3 gain), which keeps the bf16 path well conditioned.
`transformer.fast_residual.*` is deliberately non-zero (the reference zero-inits it,
reference models/transformer.py:173-174, which would leave the fast branch untested).
"""
import math
import zlib

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _time_sine(max_len: int, d_model: int) -> torch.Tensor:
    # reference models/position_encoding.py:35-45 (buffer, not random)
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    te = torch.zeros(max_len, 1, d_model)
    te[:, 0, 0::2] = torch.sin(position * div_term)
    te[:, 0, 1::2] = torch.cos(position * div_term)
    return te


def seeded_tensor(seed: int, name: str, shape, dtype=torch.float32) -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(seed, name)
    if dtype not in (torch.float32, torch.float64, torch.float16, torch.bfloat16):
        if name.endswith("position_ids"):
            return torch.arange(shape[-1]).expand(shape).clone().to(dtype)
        return torch.zeros(shape, dtype=dtype)

    def randn(std=1.0, mean=0.0):
        return torch.randn(shape, generator=g) * std + mean

    def rand(lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    leaf = name.rsplit(".", 1)[-1]
    if name.endswith("time_embed.te"):
        return _time_sine(shape[0], shape[2])
    in_backbone = name.startswith("backbone.")
    is_norm = any(s in name for s in ("norm", "LayerNorm", "layer_norm", ".bn", "downsample.1"))
    if in_backbone:
        if leaf == "running_var":
            return rand(0.5, 1.5)
        if leaf == "running_mean":
            return randn(0.1)
        if is_norm and leaf == "weight":
            if ".bn3." in name:
                return rand(0.04, 0.10)
            if "downsample.1" in name:
                return rand(0.5, 0.7)
            return rand(0.8, 1.2)
        if is_norm and leaf == "bias":
            return randn(0.1)
        if len(shape) == 4:  # conv weight (Cout, Cin, kh, kw)
            fan_in = shape[1] * shape[2] * shape[3]
            return randn(math.sqrt(2.0 / fan_in))
    if is_norm and len(shape) == 1:
        return randn(0.05, 1.0) if leaf == "weight" else randn(0.05)
    if "text_encoder" in name:
        if len(shape) >= 2:
            return randn(0.02)
        return randn(0.02)
    if name == "query_embed.weight":
        return randn(1.0)
    if len(shape) == 4:  # input_proj conv
        fan_in = shape[1] * shape[2] * shape[3]
        return randn(1.0 / math.sqrt(fan_in))
    if len(shape) == 2:
        return randn(1.0 / math.sqrt(shape[1]))
    if len(shape) == 1:
        return randn(0.02)
    return randn(0.02)


def seeded_state_dict(manifest, seed: int = 0):
    """manifest: iterable of (name, shape, dtype_str). Returns {name: tensor} (CPU, fp32)."""
    out = {}
    for name, shape, dt in manifest:
        dtype = getattr(torch, dt.replace("torch.", ""))
        out[name] = seeded_tensor(seed, name, shape, dtype)
    return out
