"""Deterministic synthetic weights and inputs.

Shared by the tests, ``bench.py`` and ``__graft_entry__.smoke()`` so that every leg of a
comparison sees bit-identical data.  All generation happens on the CPU generator, which is
reproducible across machines for a fixed torch version.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def randomise_state(sd: Dict[str, torch.Tensor], seed: int = 1) -> Dict[str, torch.Tensor]:
    """Perturb BN running stats / affine and LN affine so that folding bugs are visible.

    Default init makes every BatchNorm the identity (mean 0, var 1, weight 1, bias 0) and
    every LayerNorm affine-free, which would hide mistakes in the load-time weight algebra.
    running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5), weight ~ U(0.5, 1.5), bias ~ N(0, 0.1).
    Keys are visited in state_dict order, so the result is a pure function of (sd, seed).
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        v = v.clone()
        is_bn = k.endswith("running_mean") or k.endswith("running_var")
        base = k.rsplit(".", 1)[0]
        is_norm_affine = (k.endswith(".weight") or k.endswith(".bias")) and v.dim() == 1 and (
            (base + ".running_mean") in sd or "norm" in base.rsplit(".", 1)[-1]
            or base.endswith("projection.1"))
        if k.endswith("running_mean"):
            v = torch.randn(v.shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            v = torch.rand(v.shape, generator=g) + 0.5
        elif is_norm_affine and k.endswith(".weight"):
            v = torch.rand(v.shape, generator=g) + 0.5
        elif is_norm_affine and k.endswith(".bias"):
            v = torch.randn(v.shape, generator=g) * 0.1
        elif is_bn:
            pass
        out[k] = v
    return out


def synth_images_u8(batch: int, seed: int = 1234, size: int = 224) -> torch.Tensor:
    """uint8 HWC images [B, size, size, 3]."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, size, size, 3), generator=g, dtype=torch.uint8)


def normalise_u8(images_u8: torch.Tensor) -> torch.Tensor:
    """(u8/255 - mean)/std, HWC -> NCHW fp32: what the reference's transform yields at 224x224."""
    x = images_u8.to(torch.float32).div(255.0).permute(0, 3, 1, 2)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return ((x - mean) / std).contiguous()


def synth_questions(batch: int, seed: int = 1234, max_len: int = 20, vocab: int = 10000,
                    full_length: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """ids int64 [B, L] in [4, vocab), ragged lengths in [3, L], pad id 0 / mask 0 beyond."""
    g = torch.Generator().manual_seed(seed + 7919)
    ids = torch.randint(4, vocab, (batch, max_len), generator=g, dtype=torch.int64)
    if full_length:
        return ids, torch.ones(batch, max_len, dtype=torch.int64)
    lens = torch.randint(min(3, max_len), max_len + 1, (batch,), generator=g)
    mask = (torch.arange(max_len).unsqueeze(0) < lens.unsqueeze(1)).to(torch.int64)
    return ids * mask, mask


def synth_batch(batch: int, seed: int = 1234, max_len: int = 20, vocab: int = 10000,
                full_length: bool = False):
    """(images_u8 [B,224,224,3], images_f32 NCHW, ids, mask)."""
    u8 = synth_images_u8(batch, seed)
    ids, mask = synth_questions(batch, seed, max_len, vocab, full_length)
    return u8, normalise_u8(u8), ids, mask


def state_fingerprint(sd: Dict[str, torch.Tensor]) -> Dict[str, float]:
    """Order-sensitive float64 fingerprint of a state_dict (sum, abs-sum, weighted sum)."""
    s = a = w = 0.0
    n = 0
    for i, (k, v) in enumerate(sd.items()):
        d = v.detach().to(torch.float64).flatten()
        s += float(d.sum())
        a += float(d.abs().sum())
        if d.numel():
            w += float((d * torch.arange(1, d.numel() + 1, dtype=torch.float64)).sum()) * (i + 1) * 1e-6
        n += d.numel()
    return {"sum": s, "abs_sum": a, "weighted": w, "numel": float(n), "keys": float(len(sd))}
