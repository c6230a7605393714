"""CPU ORACLE for the SAM ViT image-encoder forward path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file.  The product (``interactable-unified-vision-language_b200``) never does
and has no CPU fallback.

What this is: a functional restatement (no ``nn.Module``; explicit GEMMs, explicit window loops,
explicit relative-position gathers) of the reference algorithm in
``/root/reference/sam/modeling/image_encoder.py`` and ``sam/modeling/common.py``, operating on a plain
``state_dict``.  Every function cites the reference lines it follows.

Parity pin: the reference publishes no tests/golden vectors for this path, so this
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the
unmodified reference from ``/root/reference`` in the build container, runs it on the seeded synthetic
weights/images of ``synthetic.py`` and commits sampled outputs + per-stage taps as
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this file against those fixtures.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def stage_images(images, pixel_mean, pixel_std, size: int = 1024) -> Tensor:
    """What the encoder's callers do before calling it (scope row N2; modeling/architectures/xdecoder_model.py:481-484):
    ``images = [(x - pixel_mean) / pixel_std for x in images]`` on uint8 (C,h,w) tensors, then
    ``ImageList.from_tensors(images, 1024)`` (detectron2): a (B,C,H,W) canvas whose sides are the batch maxima rounded up to a
    multiple of ``size``, each image in the top-left corner, zeros elsewhere.  The detectron2 class is absent from the reference
    tree; its documented semantics (pad_value = 0.0, bottom/right padding) are restated here.  Only canvases of exactly
    ``size`` x ``size`` reach the encoder in the BASELINE configurations."""
    mean = torch.as_tensor(pixel_mean, dtype=torch.float32).view(-1, 1, 1)
    std = torch.as_tensor(pixel_std, dtype=torch.float32).view(-1, 1, 1)
    norm = [(x.to(torch.float32) - mean) / std for x in images]
    H = -(-max(t.shape[-2] for t in norm) // size) * size
    W = -(-max(t.shape[-1] for t in norm) // size) * size
    out = torch.zeros(len(norm), norm[0].shape[0], H, W, dtype=torch.float32)
    for i, t in enumerate(norm):
        out[i, :, : t.shape[-2], : t.shape[-1]] = t
    return out


def patch_embed(x: Tensor, w: Tensor, b: Tensor, patch: int) -> Tensor:
    """Conv2d(k=stride=patch) + NCHW->NHWC (image_encoder.py:402-410) as an im2col GEMM.
    x (B,C,H,W) -> (B,H/p,W/p,D)."""
    B, C, H, W = x.shape
    gh, gw = H // patch, W // patch
    cols = x.reshape(B, C, gh, patch, gw, patch).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, C * patch * patch)
    out = cols @ w.reshape(w.shape[0], -1).t() + b
    return out.reshape(B, gh, gw, -1)


def bicubic_pos_embed(pos: Tensor, h: int, w: int) -> Tensor:
    """Fallback of image_encoder.py:124-132 (used when the token grid differs from pos_embed's)."""
    H0, W0 = pos.shape[1:3]
    return F.interpolate(pos.permute(0, 3, 1, 2), scale_factor=(h / H0, w / W0), mode="bicubic").permute(0, 2, 3, 1)


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance (image_encoder.py:166,176; build_sam.py:65)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (common.py:23; image_encoder.py:420)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def rel_pos_rows(q_size: int, k_size: int, rel_pos: Tensor) -> Tensor:
    """get_rel_pos (image_encoder.py:307-337): (q_size,k_size,hd) table rows R[q,k] = rel_pos[q-k+(k_size-1)]
    (after the q/k coordinate scaling, which is the identity when q_size == k_size); the table is
    linearly resized first when its length is not 2*max(q,k)-1 (:319-330)."""
    L = 2 * max(q_size, k_size) - 1
    if rel_pos.shape[0] != L:
        rel_pos = F.interpolate(rel_pos.t()[None], size=L, mode="linear")[0].t()
    qs = torch.arange(q_size, dtype=torch.float64)[:, None] * max(k_size / q_size, 1.0)
    ks = torch.arange(k_size, dtype=torch.float64)[None, :] * max(q_size / k_size, 1.0)
    idx = (qs - ks + (k_size - 1) * max(q_size / k_size, 1.0)).long()
    return rel_pos[idx]


def attention(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int) -> Tensor:
    """Attention.forward (image_encoder.py:239-255) with add_decomposed_rel_pos (:340-376).
    x (B',H,W,D) -> (B',H,W,D).  The bias uses the UNSCALED q (:249)."""
    Bp, H, W, D = x.shape
    hd = D // num_heads
    S = H * W
    qkv = x.reshape(Bp * S, D) @ p[prefix + "qkv.weight"].t() + p[prefix + "qkv.bias"]
    qkv = qkv.reshape(Bp, S, 3, num_heads, hd)
    q = qkv[:, :, 0].permute(0, 2, 1, 3)   # (B',h,S,hd)
    k = qkv[:, :, 1].permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)
    scores = (q * hd ** -0.5) @ k.transpose(-1, -2)          # (B',h,S,S)
    Rh = rel_pos_rows(H, H, p[prefix + "rel_pos_h"])          # (H,H,hd)
    Rw = rel_pos_rows(W, W, p[prefix + "rel_pos_w"])
    q5 = q.reshape(Bp, num_heads, H, W, hd)
    bias_h = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)         # (B',h,H,W,Kh)
    bias_w = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)         # (B',h,H,W,Kw)
    scores = (scores.reshape(Bp, num_heads, H, W, H, W)
              + bias_h[..., :, None] + bias_w[..., None, :]).reshape(Bp, num_heads, S, S)
    probs = torch.softmax(scores, dim=-1)
    out = (probs @ v).permute(0, 2, 1, 3).reshape(Bp, H, W, D)
    return out.reshape(Bp * S, D).matmul(p[prefix + "proj.weight"].t()).add(p[prefix + "proj.bias"]).reshape(Bp, H, W, D)


def window_partition(x: Tensor, ws: int):
    """image_encoder.py:258-279: zero-pad bottom/right to a multiple of ws, cut into ws x ws windows."""
    B, H, W, C = x.shape
    ph, pw = (-H) % ws, (-W) % ws
    Hp, Wp = H + ph, W + pw
    xp = x.new_zeros(B, Hp, Wp, C)
    xp[:, :H, :W] = x
    nh, nw = Hp // ws, Wp // ws
    wins = torch.stack([xp[:, i * ws:(i + 1) * ws, j * ws:(j + 1) * ws] for i in range(nh) for j in range(nw)], dim=1)
    return wins.reshape(B * nh * nw, ws, ws, C), (Hp, Wp)


def window_unpartition(wins: Tensor, ws: int, pad_hw, hw) -> Tensor:
    """image_encoder.py:282-304: inverse of window_partition, cropping the pad."""
    Hp, Wp = pad_hw
    H, W = hw
    nh, nw = Hp // ws, Wp // ws
    B = wins.shape[0] // (nh * nw)
    wins = wins.reshape(B, nh, nw, ws, ws, -1)
    out = wins.new_empty(B, Hp, Wp, wins.shape[-1])
    for i in range(nh):
        for j in range(nw):
            out[:, i * ws:(i + 1) * ws, j * ws:(j + 1) * ws] = wins[:, i, j]
    return out[:, :H, :W].contiguous()


def block(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int, window: int, ln_eps: float) -> Tensor:
    """Block.forward (image_encoder.py:181-197).  Note the pad is applied AFTER norm1, so pad tokens
    enter qkv as zeros and become real keys/values equal to the qkv bias (:183-187,274)."""
    shortcut = x
    y = layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], ln_eps)
    if window > 0:
        H, W = y.shape[1:3]
        y, pad_hw = window_partition(y, window)
    y = attention(y, p, prefix + "attn.", num_heads)
    if window > 0:
        y = window_unpartition(y, window, pad_hw, (H, W))
    x = shortcut + y
    z = layer_norm(x, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], ln_eps)
    Bq, Hq, Wq, D = z.shape
    z = z.reshape(-1, D) @ p[prefix + "mlp.lin1.weight"].t() + p[prefix + "mlp.lin1.bias"]     # common.py:25-26
    z = gelu_erf(z) @ p[prefix + "mlp.lin2.weight"].t() + p[prefix + "mlp.lin2.bias"]
    return x + z.reshape(Bq, Hq, Wq, D)


# ----------------------------------------------------------------------------------------------
# SimpleFPN neck (image_encoder.py:413-466), all NCHW
# ----------------------------------------------------------------------------------------------
def group_norm1(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.GroupNorm(1, C): one group = whole sample (C,H,W), biased variance, per-channel affine."""
    B = x.shape[0]
    flat = x.reshape(B, -1)
    mu = flat.mean(1).reshape(B, 1, 1, 1)
    var = ((flat - flat.mean(1, keepdim=True)) ** 2).mean(1).reshape(B, 1, 1, 1)
    return (x - mu) / torch.sqrt(var + eps) * w.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1)


def conv_transpose_2x2(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """ConvTranspose2d(k=2,s=2): non-overlapping, out[b,o,2y+i,2x+j] = sum_c x[b,c,y,x] w[c,o,i,j] + b[o]."""
    B, C, H, W = x.shape
    O = w.shape[1]
    y = torch.einsum("bchw,coij->bohiwj", x, w).reshape(B, O, 2 * H, 2 * W)
    return y + b.reshape(1, -1, 1, 1)


def conv_2x2_s2(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """Conv2d(k=2,s=2): out[b,o,y,x] = sum_{c,i,j} x[b,c,2y+i,2x+j] w[o,c,i,j] + b[o]."""
    B, C, H, W = x.shape
    xr = x.reshape(B, C, H // 2, 2, W // 2, 2)
    return torch.einsum("bchiwj,ocij->bohw", xr, w) + b.reshape(1, -1, 1, 1)


def conv_1x1(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return torch.einsum("bchw,oc->bohw", x, w[:, :, 0, 0]) + b.reshape(1, -1, 1, 1)


def simple_fpn(x: Tensor, p: Dict[str, Tensor], gn_eps: float, prefix: str = "neck.") -> Dict[str, Tensor]:
    """SimpleFPN.forward (image_encoder.py:455-466); GELU placement differs per branch (:417-447)."""
    def W(name):
        return p[prefix + name + ".weight"], p[prefix + name + ".bias"]

    d4 = conv_transpose_2x2(x, *W("down_4.0"))
    d4 = gelu_erf(group_norm1(d4, *W("down_4.1"), gn_eps))
    d4 = conv_transpose_2x2(d4, *W("down_4.3"))
    d4 = group_norm1(d4, *W("down_4.4"), gn_eps)
    d4 = conv_1x1(d4, *W("down_4.5"))
    d4 = gelu_erf(group_norm1(d4, *W("down_4.6"), gn_eps))

    d8 = conv_transpose_2x2(x, *W("down_8.0"))
    d8 = group_norm1(d8, *W("down_8.1"), gn_eps)
    d8 = conv_1x1(d8, *W("down_8.2"))
    d8 = gelu_erf(group_norm1(d8, *W("down_8.3"), gn_eps))

    d16 = conv_1x1(x, *W("down_16.0"))
    d16 = gelu_erf(group_norm1(d16, *W("down_16.1"), gn_eps))

    d32 = conv_2x2_s2(x, *W("down_32.0"))
    d32 = group_norm1(d32, *W("down_32.1"), gn_eps)
    d32 = conv_1x1(d32, *W("down_32.2"))
    d32 = gelu_erf(group_norm1(d32, *W("down_32.3"), gn_eps))
    return {"res2": d4, "res3": d8, "res4": d16, "res5": d32}


# ----------------------------------------------------------------------------------------------
# whole encoder
# ----------------------------------------------------------------------------------------------
def encoder_forward(sd: Dict[str, Tensor], x: Tensor, *, depth: int, num_heads: int, window_size: int,
                    global_attn_indexes, patch_size: int = 16, ln_eps: float = 1e-6, gn_eps: float = 1e-5,
                    tap: Optional[Callable[[str, Tensor], None]] = None) -> Dict[str, Tensor]:
    """ImageEncoderViT.forward (image_encoder.py:107-120).  ``tap(name, tensor)`` receives the token
    stream after the patch-embed(+pos) and after every block (for bisecting a mismatch)."""
    t = patch_embed(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], patch_size)
    pos = sd.get("pos_embed")
    if pos is not None:
        if pos.shape[1:3] != t.shape[1:3]:
            pos = bicubic_pos_embed(pos, t.shape[1], t.shape[2])
        t = t + pos
    if tap:
        tap("embed", t)
    for i in range(depth):
        ws = 0 if i in global_attn_indexes else window_size
        t = block(t, sd, f"blocks.{i}.", num_heads, ws, ln_eps)
        if tap:
            tap(f"block{i}", t)
    return simple_fpn(t.permute(0, 3, 1, 2), sd, gn_eps)


def encoder_forward_cfg(sd, x, cfg, tap=None, dtype=torch.float32):
    """Convenience wrapper taking an ``EncoderConfig``; runs image by image (no cross-sample coupling:
    LayerNorm is per token, GroupNorm(1,C) per sample) to bound the (h,S,S) score memory."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    outs = []
    with torch.no_grad():
        for b in range(x.shape[0]):
            outs.append(encoder_forward(
                sd, x[b:b + 1].to(dtype), depth=cfg.depth, num_heads=cfg.num_heads, window_size=cfg.window_size,
                global_attn_indexes=cfg.global_attn_indexes, patch_size=cfg.patch_size, ln_eps=cfg.ln_eps,
                gn_eps=cfg.gn_eps, tap=tap if b == 0 else None))
    return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}
