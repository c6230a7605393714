"""Seeded synthetic weights and images.

There is no network for checkpoints, so every parity case and the bench use a ``state_dict`` generated
here.  The same dict is loaded into the reference (when generating goldens), the oracle and the CUDA
encoder.  The scales mirror PyTorch's default inits the reference ends up with
(``build_sam.py:48-106`` builds with ``checkpoint=None``): Linear/Conv ~ U(+-1/sqrt(fan_in)).
``pos_embed`` and ``rel_pos_*`` are zero-initialised by the reference
(``image_encoder.py:68-70,236-237``); they are overwritten with noise here, otherwise those code
paths would be numerically invisible.  Norm weights/biases are perturbed away from 1/0 so the
affine terms are exercised.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .config import EncoderConfig, state_dict_spec


def make_state_dict(cfg: EncoderConfig, seed: int = 1234, rel_std: float = 0.02,
                    pos_std: float = 0.02) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in state_dict_spec(cfg):
        if key == "pos_embed":
            t = torch.randn(shape, generator=g) * pos_std
        elif "rel_pos" in key:
            t = torch.randn(shape, generator=g) * rel_std
        elif len(shape) == 1:
            is_norm = (".norm" in key) or key.startswith("orig_neck.1") or key.startswith("orig_neck.3") \
                or _is_groupnorm(key)
            if is_norm and key.endswith("weight"):
                t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            elif is_norm:
                t = 0.1 * torch.randn(shape, generator=g)
            else:  # a bias of a Linear/Conv: bound 1/sqrt(fan_in) of the matching weight
                wshape = dict(state_dict_spec(cfg))[key[:-4] + "weight"]
                t = _uniform(shape, _fan_in(key[:-4] + "weight", wshape), g)
        else:
            t = _uniform(shape, _fan_in(key, shape), g)
        sd[key] = t.to(torch.float32).contiguous()
    return sd


def _is_groupnorm(key: str) -> bool:
    # SimpleFPN norm slots (image_encoder.py:417-447)
    gn = ("down_4.1.", "down_4.4.", "down_4.6.", "down_8.1.", "down_8.3.", "down_16.1.",
          "down_32.1.", "down_32.3.")
    return any(s in key for s in gn)


def _fan_in(key: str, shape) -> int:
    if len(shape) == 2:
        return shape[1]
    if len(shape) == 4:
        # ConvTranspose2d weights are (in, out, kh, kw): torch computes fan_in from dim 1
        return shape[1] * shape[2] * shape[3]
    return shape[0]


def _uniform(shape, fan_in: int, g: torch.Generator) -> torch.Tensor:
    bound = 1.0 / math.sqrt(max(1, fan_in))
    return (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound


def make_images(batch: int, cfg: EncoderConfig, seed: int = 0, hw=None) -> torch.Tensor:
    """``randn(B,3,S,S)``: the post-normalisation range of real inputs (xdecoder_model.py:333).  ``hw`` overrides the canvas
    (scope row N3: e.g. the 1024 x 2048 pads of the reference's COCO evaluation)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    h, w = hw if hw else (cfg.img_size, cfg.img_size)
    return torch.randn(batch, cfg.in_chans, h, w, generator=g)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in float64."""
    a64, b64 = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float((a64 - b64).norm() / b64.norm().clamp_min(1e-30))
