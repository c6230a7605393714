"""Encoder hyper-parameters for the SAM ViT image-encoder hot path.

The three presets restate the constants the reference fixes in
``sam/build_sam.py:14-44`` (embed dim / depth / heads / global-attention block
indexes) and ``sam/build_sam.py:60-73`` (img 1024, patch 16, window 14, mlp ratio 4,
LayerNorm eps 1e-6, qkv bias, decomposed rel-pos on).  The neck widths follow
``sam/modeling/image_encoder.py:413-447`` (``SimpleFPN``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class EncoderConfig:
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    global_attn_indexes: Tuple[int, ...] = (2, 5, 8, 11)
    img_size: int = 1024
    patch_size: int = 16
    in_chans: int = 3
    mlp_ratio: float = 4.0
    out_chans: int = 256          # only sizes the never-executed ``orig_neck`` (image_encoder.py:88-104)
    window_size: int = 14
    ln_eps: float = 1e-6          # build_sam.py:65
    gn_eps: float = 1e-5          # nn.GroupNorm default (image_encoder.py:419)
    fpn_dims: Tuple[int, int, int, int] = (128, 256, 512, 1024)   # image_encoder.py:105

    # ---- derived ----
    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def tokens(self) -> int:
        return self.grid * self.grid

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def mlp_dim(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)

    @property
    def down4_chan(self) -> int:   # image_encoder.py:416
        return max(self.fpn_dims[0] * 2, self.embed_dim // 2)

    @property
    def down8_chan(self) -> int:   # image_encoder.py:427
        return max(self.fpn_dims[1], self.embed_dim // 2)

    @property
    def down32_chan(self) -> int:  # image_encoder.py:440
        return max(self.fpn_dims[3], self.embed_dim * 2)

    def is_global(self, i: int) -> bool:
        return i in self.global_attn_indexes

    def rel_table_len(self, i: int) -> int:
        s = self.grid if self.is_global(i) else self.window_size
        return 2 * s - 1

    # ---- algorithmic work (2*MAC), minimal token count: SURVEY.md section 8(a) ----
    def flops_per_image(self) -> float:
        D, T, hd, h = self.embed_dim, self.tokens, self.head_dim, self.num_heads
        g, w = self.grid, self.window_size
        k_pe = self.in_chans * self.patch_size ** 2
        fl = 2.0 * T * k_pe * D                                   # patch embed
        n_glob = len([i for i in range(self.depth) if self.is_global(i)])
        n_win = self.depth - n_glob
        fl += self.depth * (2.0 * T * D * (3 * D + D + 2 * self.mlp_dim))   # qkv, proj, lin1, lin2
        nwin = ((g + w - 1) // w) ** 2
        fl += n_win * 4.0 * nwin * (w * w) ** 2 * D               # QK^T + PV, windowed
        fl += n_glob * 4.0 * float(T) ** 2 * D                    # QK^T + PV, global
        # rel-pos einsums (image_encoder.py:369-370): per query, (k_h + k_w) dots of length hd per head
        fl += n_win * 2.0 * nwin * (w * w) * (2 * w) * hd * h
        fl += n_glob * 2.0 * T * (2 * g) * hd * h
        # neck (image_encoder.py:417-447)
        d4, d8, d32 = self.down4_chan, self.down8_chan, self.down32_chan
        o = self.fpn_dims
        fl += 2.0 * T * D * 4 * d4 + 2.0 * 4 * T * d4 * 4 * (d4 // 2) + 2.0 * 16 * T * (d4 // 2) * o[0]
        fl += 2.0 * T * D * 4 * d8 + 2.0 * 4 * T * d8 * o[1]
        fl += 2.0 * T * D * o[2]
        fl += 2.0 * (T // 4) * 4 * D * d32 + 2.0 * (T // 4) * d32 * o[3]
        return fl


PRESETS: Dict[str, EncoderConfig] = {
    "vit_b": EncoderConfig(768, 12, 12, (2, 5, 8, 11)),
    "vit_l": EncoderConfig(1024, 24, 16, (5, 11, 17, 23)),
    "vit_h": EncoderConfig(1280, 32, 16, (7, 15, 23, 31)),
    # small shapes for fast parity cases (not reference presets): head_dim 64 and head_dim 80
    "tiny64": EncoderConfig(128, 2, 2, (1,)),
    "tiny80": EncoderConfig(160, 3, 2, (2,)),
}


def state_dict_spec(cfg: EncoderConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) for every entry of ``ImageEncoderViT.state_dict()`` in the reference's order
    (probed list in SURVEY.md section 8(b); ``image_encoder.py:58-105``)."""
    D, g, p = cfg.embed_dim, cfg.grid, cfg.patch_size
    hd, oc = cfg.head_dim, cfg.out_chans
    spec: List[Tuple[str, Tuple[int, ...]]] = [
        ("pos_embed", (1, g, g, D)),
        ("patch_embed.proj.weight", (D, cfg.in_chans, p, p)),
        ("patch_embed.proj.bias", (D,)),
    ]
    for i in range(cfg.depth):
        L = cfg.rel_table_len(i)
        b = f"blocks.{i}."
        spec += [
            (b + "norm1.weight", (D,)), (b + "norm1.bias", (D,)),
            (b + "attn.rel_pos_h", (L, hd)), (b + "attn.rel_pos_w", (L, hd)),
            (b + "attn.qkv.weight", (3 * D, D)), (b + "attn.qkv.bias", (3 * D,)),
            (b + "attn.proj.weight", (D, D)), (b + "attn.proj.bias", (D,)),
            (b + "norm2.weight", (D,)), (b + "norm2.bias", (D,)),
            (b + "mlp.lin1.weight", (cfg.mlp_dim, D)), (b + "mlp.lin1.bias", (cfg.mlp_dim,)),
            (b + "mlp.lin2.weight", (D, cfg.mlp_dim)), (b + "mlp.lin2.bias", (D,)),
        ]
    spec += [
        ("orig_neck.0.weight", (oc, D, 1, 1)),
        ("orig_neck.1.weight", (oc,)), ("orig_neck.1.bias", (oc,)),
        ("orig_neck.2.weight", (oc, oc, 3, 3)),
        ("orig_neck.3.weight", (oc,)), ("orig_neck.3.bias", (oc,)),
    ]
    d4, d8, d32 = cfg.down4_chan, cfg.down8_chan, cfg.down32_chan
    o = cfg.fpn_dims
    n = "neck."
    spec += [
        (n + "down_4.0.weight", (D, d4, 2, 2)), (n + "down_4.0.bias", (d4,)),
        (n + "down_4.1.weight", (d4,)), (n + "down_4.1.bias", (d4,)),
        (n + "down_4.3.weight", (d4, d4 // 2, 2, 2)), (n + "down_4.3.bias", (d4 // 2,)),
        (n + "down_4.4.weight", (d4 // 2,)), (n + "down_4.4.bias", (d4 // 2,)),
        (n + "down_4.5.weight", (o[0], d4 // 2, 1, 1)), (n + "down_4.5.bias", (o[0],)),
        (n + "down_4.6.weight", (o[0],)), (n + "down_4.6.bias", (o[0],)),
        (n + "down_8.0.weight", (D, d8, 2, 2)), (n + "down_8.0.bias", (d8,)),
        (n + "down_8.1.weight", (d8,)), (n + "down_8.1.bias", (d8,)),
        (n + "down_8.2.weight", (o[1], d8, 1, 1)), (n + "down_8.2.bias", (o[1],)),
        (n + "down_8.3.weight", (o[1],)), (n + "down_8.3.bias", (o[1],)),
        (n + "down_16.0.weight", (o[2], D, 1, 1)), (n + "down_16.0.bias", (o[2],)),
        (n + "down_16.1.weight", (o[2],)), (n + "down_16.1.bias", (o[2],)),
        (n + "down_32.0.weight", (d32, D, 2, 2)), (n + "down_32.0.bias", (d32,)),
        (n + "down_32.1.weight", (d32,)), (n + "down_32.1.bias", (d32,)),
        (n + "down_32.2.weight", (o[3], d32, 1, 1)), (n + "down_32.2.bias", (o[3],)),
        (n + "down_32.3.weight", (o[3],)), (n + "down_32.3.bias", (o[3],)),
    ]
    return spec
