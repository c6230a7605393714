"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def sampled_rel_l2(t: torch.Tensor, g: dict, key: str) -> float:
    """rel-L2 between tensor ``t`` and the golden SAMPLE stored under ``key`` (e.g. 'out.res2')."""
    idx = torch.from_numpy(g[key + ".idx"])
    ref = torch.from_numpy(g[key + ".val"]).double()
    assert tuple(t.shape) == tuple(int(s) for s in g[key + ".shape"]), (t.shape, g[key + ".shape"])
    got = t.detach().reshape(-1).cpu().double()[idx]
    return float((got - ref).norm() / ref.norm())


def norm_ratio(t: torch.Tensor, g: dict, key: str) -> float:
    return float(t.detach().double().norm().item() / float(g[key + ".norm"]))


# ---------------------------------------------------------------------------------------------
# reference for the attention core operating directly on a qkv tensor (uses the oracle's pieces)
def ref_attention_core(qkv: torch.Tensor, rel_h: torch.Tensor, rel_w: torch.Tensor, qkv_bias: torch.Tensor, B: int, g: int,
                       ws: int, heads: int) -> torch.Tensor:
    """qkv (B*g*g, 3D) token order -> (B*g*g, D).  Windows of ws x ws with zero-padded-then-biased pad tokens
    (image_encoder.py:183-187,258-279: the pad is applied after norm1 so pad tokens' qkv equals the bias)."""
    from oracle import sam_vit_oracle as orc
    D = qkv.shape[1] // 3
    hd = D // heads
    x = qkv.reshape(B, g, g, 3 * D).double()
    if ws < g:
        Hp = -(-g // ws) * ws
        xp = qkv_bias.double().reshape(1, 1, 1, -1).expand(B, Hp, Hp, 3 * D).clone()
        xp[:, :g, :g] = x
        wins, pad_hw = orc.window_partition(xp, ws)
    else:
        wins, pad_hw = x, (g, g)
    Bp = wins.shape[0]
    S = ws * ws
    q, k, v = wins.reshape(Bp, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    scores = (q * hd ** -0.5) @ k.transpose(-1, -2)
    Rh = orc.rel_pos_rows(ws, ws, rel_h.double())
    Rw = orc.rel_pos_rows(ws, ws, rel_w.double())
    q5 = q.reshape(Bp, heads, ws, ws, hd)
    bh = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)
    bw = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)
    scores = (scores.reshape(Bp, heads, ws, ws, ws, ws) + bh[..., :, None] + bw[..., None, :]).reshape(Bp, heads, S, S)
    out = (torch.softmax(scores, -1) @ v).permute(0, 2, 1, 3).reshape(Bp, ws, ws, D)
    if ws < g:
        out = orc.window_unpartition(out, ws, pad_hw, (g, g))
    return out.reshape(B * g * g, D)


def seeded_state_dict(items, seed):
    """Deterministic weights for the fixtures at the reference's REAL widths (step1.yaml: 512 channels, 8 heads, 6 / 9 layers), whose
    state_dicts (tens of millions of values) are not stored: every entry is drawn from a generator seeded by (seed, crc32 of its key), by
    a rule of its name — the golden scripts load the result into the unmodified reference classes, the tests into the drop-ins."""
    import zlib
    out = {}
    for k, v in items:
        shape = tuple(v.shape)
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        r = torch.randn(shape, generator=g)
        norm = ("norm" in k) or (k.startswith("input_proj.") and k.split(".")[2] == "1")
        if k.endswith("sampling_offsets.weight") or k.endswith("attention_weights.weight"):
            t = r * 0.05
        elif k.endswith("sampling_offsets.bias"):
            t = r * 0.5
        elif norm and k.endswith("weight"):
            t = 1.0 + 0.2 * r
        elif norm or k.endswith("bias"):
            t = 0.1 * r
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if k in ("query_feat.weight", "query_embed.weight", "level_embed.weight"):
                t = r                                             # nn.Embedding default: N(0, 1)
            else:
                t = r / fan_in ** 0.5
        else:
            t = r
        out[k] = t.to(v.dtype) if v.dtype.is_floating_point else v.clone()
    return out


def sampled(t, n, seed):
    """(flat indices, values) of n pseudo-random entries of t"""
    idx = torch.from_numpy(np.random.RandomState(seed).randint(0, t.numel(), size=min(n, t.numel())).astype(np.int64))
    return idx, t.detach().reshape(-1)[idx].float()
