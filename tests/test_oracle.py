"""The oracle (oracle/sam_vit_oracle.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

import iuvl_b200 as ib
from oracle import sam_vit_oracle as orc
from tests.util import load_golden, sampled_rel_l2, norm_ratio

TOL = 2e-5   # fp32 CPU vs fp32 CPU, different summation orders


def _run(case):
    g = load_golden(case)
    cfg = ib.PRESETS[str(g["meta_preset"])]
    sd = ib.make_state_dict(cfg, int(g["meta_weight_seed"]), rel_std=float(g["meta_rel_std"]))
    x = ib.make_images(int(g["meta_batch"]), cfg, int(g["meta_image_seed"]))
    taps = {}
    out = orc.encoder_forward_cfg(sd, x, cfg, tap=lambda n, t: taps.__setitem__(n, t))
    return g, out, taps


@pytest.mark.parametrize("case", ["tiny64_std", "tiny64_stress", "tiny80_std", "tiny80_stress"])
def test_oracle_matches_reference_tiny(case):
    g, out, taps = _run(case)
    for k in ("res2", "res3", "res4", "res5"):
        assert sampled_rel_l2(out[k], g, "out." + k) < TOL, k
        assert abs(norm_ratio(out[k], g, "out." + k) - 1) < 1e-4
    for name, t in taps.items():
        assert sampled_rel_l2(t, g, "tap." + name) < TOL, name


@pytest.mark.slow
def test_oracle_matches_reference_vit_b():
    g, out, taps = _run("vit_b_std")
    for k in ("res2", "res3", "res4", "res5"):
        assert sampled_rel_l2(out[k], g, "out." + k) < TOL, k
    for name, t in taps.items():
        assert sampled_rel_l2(t, g, "tap." + name) < TOL, name


def test_rel_pos_rows_index_rule():
    # R[q,k] = table[q - k + (S-1)]  (image_encoder.py:333-337, probed in SURVEY.md section 7)
    tab = torch.arange(27 * 4, dtype=torch.float32).reshape(27, 4)
    R = orc.rel_pos_rows(14, 14, tab)
    for q in (0, 5, 13):
        for k in (0, 7, 13):
            assert torch.equal(R[q, k], tab[q - k + 13])


def test_rel_pos_rows_interpolates_short_table():
    tab = torch.randn(27, 8)
    R = orc.rel_pos_rows(64, 64, tab)          # table length 27 != 127 -> linear resize (:321-330)
    assert R.shape == (64, 64, 8)
    ref = torch.nn.functional.interpolate(tab.t()[None], size=127, mode="linear")[0].t()
    assert torch.allclose(R[10, 3], ref[10 - 3 + 63])


def test_window_partition_roundtrip_and_padding():
    x = torch.randn(2, 64, 64, 8)
    w, pad_hw = orc.window_partition(x, 14)
    assert w.shape == (2 * 25, 14, 14, 8) and pad_hw == (70, 70)
    # last window of image 0: rows 56..69 -> 8 real rows + 6 zero rows
    assert torch.equal(w[24, :8, :8], x[0, 56:64, 56:64]) and w[24, 8:].abs().sum() == 0
    assert torch.equal(orc.window_unpartition(w, 14, pad_hw, (64, 64)), x)


def test_pad_tokens_are_real_keys():
    """Pad tokens are zeros after norm1, so k = v = qkv bias and they take part in the softmax
    (hard part 3): masking them out changes the edge windows."""
    cfg = ib.PRESETS["tiny64"]
    sd = ib.make_state_dict(cfg, 7, rel_std=0.02)
    D, h, hd = cfg.embed_dim, cfg.num_heads, cfg.head_dim
    win = torch.zeros(1, 14, 14, D)
    win[:, :8, :8] = torch.randn(1, 8, 8, D)              # the bottom-right window: 8x8 real tokens
    got = orc.attention(win, sd, "blocks.0.attn.", h)
    # manual: explicit softmax over ALL 196 keys (pad keys have k = b_k, v = b_v)
    Wq, bq = sd["blocks.0.attn.qkv.weight"], sd["blocks.0.attn.qkv.bias"]
    qkv = (win.reshape(196, D) @ Wq.t() + bq).reshape(196, 3, h, hd)
    assert torch.allclose(qkv[195, 1].reshape(-1), bq[D:2 * D])       # pad key == k bias
    Rh = orc.rel_pos_rows(14, 14, sd["blocks.0.attn.rel_pos_h"])
    Rw = orc.rel_pos_rows(14, 14, sd["blocks.0.attn.rel_pos_w"])
    real = torch.zeros(14, 14, dtype=torch.bool)
    real[:8, :8] = True
    real = real.reshape(-1)
    outs, outs_masked = [], []
    for n in range(h):
        q, k, v = qkv[:, 0, n], qkv[:, 1, n], qkv[:, 2, n]
        s = (q * hd ** -0.5) @ k.t()
        qi = torch.arange(196)
        bias = torch.einsum("qc,qkc->qk", q, Rh[qi // 14][:, qi // 14]) + torch.einsum("qc,qkc->qk", q, Rw[qi % 14][:, qi % 14])
        s = s + bias
        outs.append(torch.softmax(s, -1) @ v)
        outs_masked.append(torch.softmax(s.masked_fill(~real[None, :], float("-inf")), -1) @ v)
    man = torch.cat(outs, 1) @ sd["blocks.0.attn.proj.weight"].t() + sd["blocks.0.attn.proj.bias"]
    msk = torch.cat(outs_masked, 1) @ sd["blocks.0.attn.proj.weight"].t() + sd["blocks.0.attn.proj.bias"]
    assert ib.rel_l2(got.reshape(196, D)[real], man[real]) < 1e-5
    assert ib.rel_l2(msk[real], man[real]) > 1e-2         # masking pads is a different function
