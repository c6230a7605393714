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
    hw = tuple(int(v) for v in g["meta_hw"]) if "meta_hw" in g else None
    x = ib.make_images(int(g["meta_batch"]), cfg, int(g["meta_image_seed"]), hw=hw)
    taps = {}
    out = orc.encoder_forward_cfg(sd, x, cfg, tap=lambda n, t: taps.__setitem__(n, t))
    return g, out, taps


# *_wide / *_tall: canvases other than 1024 x 1024 -> the reference's bicubic pos_embed / linear rel_pos fallbacks (scope row N3)
@pytest.mark.parametrize("case", ["tiny64_std", "tiny64_stress", "tiny80_std", "tiny80_stress", "tiny64_wide", "tiny80_tall"])
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


@pytest.mark.slow
@pytest.mark.parametrize("case,image", [("vit_b_std3", 1), ("vit_h_std3", 2)])
def test_oracle_matches_reference_batch3_goldens(case, image):
    """The three-image goldens behind the GPU parity cases at the benchmarked batch sizes: the oracle on ONE of their images
    (the samples whose flat index falls into that image) — pins the fixture's image order / seeding on the CPU."""
    g = load_golden(case)
    cfg = ib.PRESETS[str(g["meta_preset"])]
    sd = ib.make_state_dict(cfg, int(g["meta_weight_seed"]), rel_std=float(g["meta_rel_std"]))
    x = ib.make_images(int(g["meta_batch"]), cfg, int(g["meta_image_seed"]))
    out = orc.encoder_forward_cfg(sd, x[image:image + 1], cfg)
    for k in ("res2", "res3", "res4", "res5"):
        shape = [int(v) for v in g[f"out.{k}.shape"]]
        per = shape[1] * shape[2] * shape[3]
        idx = torch.from_numpy(g[f"out.{k}.idx"])
        sel = (idx // per) == image
        assert int(sel.sum()) > 1000
        ref = torch.from_numpy(g[f"out.{k}.val"]).double()[sel]
        got = out[k].reshape(-1).double()[idx[sel] % per]
        assert float((got - ref).norm() / ref.norm()) < TOL, k


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


def test_stage_images_restates_the_callers_preprocessing():
    """(x - mean) / std on uint8 CHW tensors + detectron2 ImageList.from_tensors(images, 1024) zero padding
    (modeling/architectures/xdecoder_model.py:481-484)."""
    import torch
    from oracle import sam_vit_oracle as orc
    g = torch.Generator().manual_seed(0)
    imgs = [torch.randint(0, 256, (3, 5, 7), generator=g, dtype=torch.uint8), torch.randint(0, 256, (3, 8, 3), generator=g, dtype=torch.uint8)]
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    out = orc.stage_images(imgs, mean, std, size=8)
    assert out.shape == (2, 3, 8, 8)
    m, s = torch.tensor(mean).view(3, 1, 1), torch.tensor(std).view(3, 1, 1)
    assert torch.equal(out[0, :, :5, :7], (imgs[0].float() - m) / s)
    assert torch.equal(out[1, :, :8, :3], (imgs[1].float() - m) / s)
    assert out[0, :, 5:, :].abs().sum() == 0 and out[0, :, :, 7:].abs().sum() == 0 and out[1, :, :, 3:].abs().sum() == 0
    big = orc.stage_images([torch.zeros(3, 9, 2, dtype=torch.uint8)], mean, std, size=8)
    assert big.shape == (1, 3, 16, 8)          # sides round UP to a multiple of the divisibility


@pytest.mark.parametrize("case", ["toy", "small", "heads8"])
def test_msda_oracle_against_reference_goldens(case):
    """Both restatements of the multi-scale deformable attention forward against outputs of the reference's own
    ms_deform_attn_core_pytorch (tests/golden/make_golden_msda.py)."""
    import numpy as np
    import torch
    from oracle import msda_oracle as mo
    from tests.util import GOLDEN
    import os
    z = np.load(os.path.join(GOLDEN, f"msda_{case}.npz"))
    value, loc, aw = (torch.from_numpy(z[k]).double() for k in ("value", "loc", "aw"))
    shapes = [tuple(int(v) for v in hw) for hw in z["shapes"]]
    ref = torch.from_numpy(z["out"])
    out = mo.ms_deform_attn_core(value, shapes, loc, aw)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    if case != "heads8":
        starts = [0]
        for h, w in shapes[:-1]:
            starts.append(starts[-1] + h * w)
        out2 = mo.ms_deform_attn_loops(value, shapes, starts, loc, aw)
        assert torch.allclose(out2, ref, rtol=1e-9, atol=1e-10), float((out2 - ref).abs().max())


@pytest.mark.parametrize("case", ["points", "boxes_masked", "heads64"])
def test_msda_module_oracle_against_reference_goldens(case):
    """oracle.ms_deform_attn_module against outputs of the UNMODIFIED reference MSDeformAttn module
    (tests/golden/make_golden_msda_module.py; ops/modules/ms_deform_attn.py:82-125)."""
    import os
    import numpy as np
    from oracle import msda_oracle as mo
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"msda_module_{case}.npz"))
    C, M, P, L = (int(v) for v in z["meta"])
    sd = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("sd.")}
    mask = torch.from_numpy(z["mask"]) if z["mask"].size else None
    out = mo.ms_deform_attn_module(sd, torch.from_numpy(z["query"]).double(), torch.from_numpy(z["ref"]).double(),
                                   torch.from_numpy(z["inp"]).double(), [tuple(int(v) for v in hw) for hw in z["shapes"]], mask, M, L, P)
    assert torch.allclose(out, torch.from_numpy(z["out"]), rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("case", ["small", "heads64"])
def test_deform_encoder_oracle_against_reference_goldens(case):
    """oracle.deform_encoder_only against outputs of the UNMODIFIED reference encoder classes
    (tests/golden/make_golden_deform_encoder.py; transformer_encoder_deform.py:23-161)."""
    import os
    import numpy as np
    from oracle import msda_oracle as mo
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"deform_encoder_{case}.npz"))
    C, M, NL, F_, P, L = (int(v) for v in z["meta"])
    sd = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("sd.")}
    out = mo.deform_encoder_only(sd, [torch.from_numpy(z[f"src{i}"]).double() for i in range(L)],
                                 [torch.from_numpy(z[f"pos{i}"]).double() for i in range(L)], M, P, NL)
    ref = torch.from_numpy(z["memory"]).double()
    assert float((out - ref).abs().max()) < 1e-5      # the golden is stored in fp32


def test_deform_encoder_module_keys_match_reference():
    """The drop-in encoder's state_dict keys / shapes are the reference's (so its checkpoints load with strict=True)."""
    import os
    import numpy as np
    from iuvl_b200.msda import MSDeformAttnTransformerEncoderOnly
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "deform_encoder_small.npz"))
    C, M, NL, F_, P, L = (int(v) for v in z["meta"])
    mod = MSDeformAttnTransformerEncoderOnly(C, M, NL, F_, 0.1, "relu", L, P)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    mod.load_state_dict(sd, strict=True)
    assert {k: tuple(v.shape) for k, v in mod.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    with pytest.raises(RuntimeError):        # no CPU path
        with torch.no_grad():
            mod([torch.from_numpy(z[f"src{i}"]) for i in range(L)], [torch.from_numpy(z[f"pos{i}"]) for i in range(L)])


def _pixel_decoder_case(case):
    import os
    import numpy as np
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"pixel_decoder_{case}.npz"))
    C, MD, M, NL, F_ = (int(v) for v in z["meta"])
    seed, N, side = (int(v) for v in z["feat_seed"])
    gf = torch.Generator().manual_seed(seed)
    feats = {f"res{2 + i}": torch.randn(N, c, side >> i, side >> i, generator=gf) for i, c in enumerate((128, 256, 512, 1024))}
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, (C, MD, M, NL, F_), feats, sd


@pytest.mark.parametrize("case", ["small", "wide"])
def test_pixel_decoder_oracle_against_reference_goldens(case):
    """oracle.pixel_decoder against outputs of the UNMODIFIED reference MSDeformAttnPixelDecoder
    (tests/golden/make_golden_pixel_decoder.py; transformer_encoder_deform.py:315-359; the reference ran in fp32)."""
    from oracle import pixel_decoder_oracle as po
    import iuvl_b200 as ib
    z, (C, MD, M, NL, F_), feats, sd = _pixel_decoder_case(case)
    mask, multi = po.pixel_decoder({k: v.double() for k, v in sd.items()}, {k: v.double() for k, v in feats.items()}, M, NL)
    assert ib.rel_l2(mask, torch.from_numpy(z["mask_features"])) < 2e-5
    for i, m in enumerate(multi):
        assert ib.rel_l2(m, torch.from_numpy(z[f"multi{i}"])) < 2e-5


def _sampled_err(t, z, name):
    idx = torch.from_numpy(z[name + ".idx"])
    ref = torch.from_numpy(z[name + ".val"]).double()
    assert tuple(t.shape) == tuple(int(v) for v in z[name + ".shape"]), (name, t.shape)
    got = t.detach().reshape(-1).double().cpu()[idx]
    return float((got - ref).norm() / ref.norm())


def _step1_pixel_decoder():
    """The step1.yaml geometry (conv_dim = mask_dim = 512, 8 heads, d_ffn 1024, 6 layers): weights regenerated by seeded_state_dict."""
    import os
    import numpy as np
    from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder
    from tests.util import GOLDEN, seeded_state_dict
    z = np.load(os.path.join(GOLDEN, "pixel_decoder_step1.npz"))
    C, MD, M, NL, F_, N, side, seed = (int(v) for v in z["meta"])
    mod = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=M, transformer_dim_feedforward=F_, transformer_enc_layers=NL,
                                   conv_dim=C, mask_dim=MD, norm="GN", transformer_in_features=["res3", "res4", "res5"], common_stride=4)
    sd = seeded_state_dict(list(mod.state_dict().items()), seed)
    mod.load_state_dict(sd, strict=True)
    gf = torch.Generator().manual_seed(seed + 1)
    feats = {f"res{2 + i}": torch.randn(N, c, side >> i, side >> i, generator=gf) for i, c in enumerate((128, 256, 512, 1024))}
    return z, mod, sd, feats, (M, NL)


@pytest.mark.slow
def test_pixel_decoder_oracle_at_the_step1_geometry():
    """oracle.pixel_decoder against the UNMODIFIED reference class at conv_dim 512 / 8 heads / 6 layers (samples of its outputs)."""
    from oracle import pixel_decoder_oracle as po
    z, mod, sd, feats, (M, NL) = _step1_pixel_decoder()
    mask, multi = po.pixel_decoder({k: v.double() for k, v in sd.items()}, {k: v.double() for k, v in feats.items()}, M, NL)
    assert _sampled_err(mask, z, "mask_features") < 5e-5
    for i, m in enumerate(multi):
        assert _sampled_err(m, z, f"multi{i}") < 5e-5


def _step1_mask_path():
    import os
    import numpy as np
    from iuvl_b200.mask_head import XDecoderMaskPath
    from tests.util import GOLDEN, seeded_state_dict
    z = np.load(os.path.join(GOLDEN, "xdecoder_mask_path_step1.npz"))
    C, MD, Q, NH, FF, NL, B, ms, seed = (int(v) for v in z["meta"][:9])
    sides = [int(v) for v in z["meta"][9:]]
    path = XDecoderMaskPath(C, MD, Q, NH, FF, 3, [0, 1, 2, 0, 1, 2, 0, 1, 2][:NL])
    sd = seeded_state_dict(list(path.state_dict().items()), seed)
    path.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    x = [torch.randn(B, C, s_, s_, generator=g) for s_ in sides]
    mf = torch.randn(B, MD, ms, ms, generator=g)
    return z, path, sd, x, mf, (Q, NH, NL)


@pytest.mark.slow
def test_xdecoder_mask_path_oracle_at_the_step1_geometry():
    """oracle.xdecoder_mask_path against the UNMODIFIED reference forward at hidden 512 / 101 queries / 8 heads / 9 layers."""
    from oracle import mask_head_oracle as mo
    z, path, sd, x, mf, (Q, NH, NL) = _step1_mask_path()
    masks = mo.xdecoder_mask_path({k: v.double() for k, v in sd.items()}, [t.double() for t in x], mf.double(), Q, NH, [0, 1, 2, 0, 1, 2, 0, 1, 2][:NL])
    assert _sampled_err(masks[0], z, "aux0") < 1e-5
    assert _sampled_err(masks[-1], z, "pred_masks") < 5e-3


def test_pixel_decoder_module_keys_match_reference():
    from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder
    z, (C, MD, M, NL, F_), feats, sd = _pixel_decoder_case("small")
    mod = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=M, transformer_dim_feedforward=F_, transformer_enc_layers=NL,
                                   conv_dim=C, mask_dim=MD, norm="GN", transformer_in_features=["res3", "res4", "res5"], common_stride=4)
    mod.load_state_dict(sd, strict=True)
    assert {k: tuple(v.shape) for k, v in mod.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    with pytest.raises(RuntimeError):        # no CPU path
        with torch.no_grad():
            mod(feats)


@pytest.mark.parametrize("shape,size", [((3, 32, 32), (8, 8)), ((2, 48, 40), (12, 10)), ((1, 24, 24), (12, 12)), ((2, 17, 23), (5, 9))])
def test_resize_bicubic_aa_oracle_is_f_interpolate(shape, size):
    """The written-out antialiased bicubic filter against the op the reference calls (xdecoder.py:463)."""
    import torch.nn.functional as F
    from oracle import mask_head_oracle as mo
    x = torch.randn(1, *shape, generator=torch.Generator().manual_seed(7), dtype=torch.float64)
    want = F.interpolate(x, size=size, mode="bicubic", align_corners=False, antialias=True)
    got = mo.resize_bicubic_aa(x, size)
    assert float((got - want).abs().max()) < 1e-12


@pytest.mark.parametrize("case", ["small", "odd", "half"])
def test_mask_head_oracle_against_reference_goldens(case):
    """oracle.mask_branch against outputs of the UNMODIFIED reference method XDecoder.forward_prediction_heads
    (tests/golden/make_golden_mask_head.py; xdecoder.py:429-470; the reference ran in fp32)."""
    import os
    import numpy as np
    import iuvl_b200 as ib
    from oracle import mask_head_oracle as mo
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"mask_head_{case}.npz"))
    C, MD, Q, NH, th, tw = (int(v) for v in z["meta"])
    sd = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("sd.")}
    masks, logits, attn = mo.mask_branch(sd, torch.from_numpy(z["output"]).double(), torch.from_numpy(z["mask_features"]).double(), (th, tw), Q, NH)
    assert ib.rel_l2(masks, torch.from_numpy(z["outputs_mask"])) < 2e-6
    ref_attn = torch.from_numpy(z["attn_mask"])
    differ = attn != ref_attn
    # a bool can only differ where the fp32 reference and the fp64 oracle straddle zero
    rep = logits.flatten(2).unsqueeze(1).repeat(1, NH, 1, 1).flatten(0, 1)
    assert float(differ.float().mean()) < 1e-3 and (not differ.any() or float(rep[differ].abs().max()) < 1e-4)


@pytest.mark.parametrize("case", ["full_small", "full_q101"])
def test_class_box_oracle_against_reference_goldens(case):
    """oracle.class_box_branch against the class logits / boxes / caption embeddings of the UNMODIFIED reference method
    (xdecoder.py:452-484) with the reference's own compute_similarity (vlpencoder.py:239-245)."""
    import os
    import numpy as np
    import iuvl_b200 as ib
    from oracle import mask_head_oracle as mo
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"mask_head_{case}.npz"))
    Q = int(z["meta"][2])
    sd = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("sd.")}
    cls, box, cap = mo.class_box_branch(sd, torch.from_numpy(z["output"]).double(), Q, torch.from_numpy(z["text_embeddings"]).double(),
                                        float(z["logit_scale"]))
    assert ib.rel_l2(cls, torch.from_numpy(z["outputs_class"])) < 2e-6
    assert ib.rel_l2(box, torch.from_numpy(z["outputs_bbox"])) < 2e-6
    assert ib.rel_l2(cap, torch.from_numpy(z["outputs_caption"])) < 2e-6


@pytest.mark.parametrize("case", ["small", "q101", "nomask"])
def test_cross_attention_oracle_against_reference_goldens(case):
    """oracle.cross_attention_layer against outputs of the UNMODIFIED reference CrossAttentionLayer
    (tests/golden/make_golden_cross_attn.py; interface/modules.py:72-131)."""
    import os
    import numpy as np
    import iuvl_b200 as ib
    from oracle import mask_head_oracle as mo
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"cross_attn_{case}.npz"))
    C, NH = (int(v) for v in z["meta"])
    sd = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("sd.")}
    mask = torch.from_numpy(z["mask"]) if z["mask"].size else None
    out = mo.cross_attention_layer(sd, torch.from_numpy(z["tgt"]).double(), torch.from_numpy(z["memory"]).double(), mask,
                                   torch.from_numpy(z["pos"]).double(), torch.from_numpy(z["query_pos"]).double(), NH)
    assert ib.rel_l2(out, torch.from_numpy(z["out"])) < 2e-6


@pytest.mark.parametrize("case", ["small", "q101"])
def test_decoder_layer_oracles_against_reference_goldens(case):
    """oracle.self_attention_layer / ffn_layer against outputs of the UNMODIFIED reference SelfAttentionLayer / FFNLayer
    (tests/golden/make_golden_decoder_layers.py; interface/modules.py:14-69,134-174)."""
    import os
    import numpy as np
    import iuvl_b200 as ib
    from oracle import mask_head_oracle as mo
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"decoder_layers_{case}.npz"))
    C, NH, FF = (int(v) for v in z["meta"])
    sa = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("sa.")}
    ffn = {k[4:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("ffn.")}
    y = mo.self_attention_layer(sa, torch.from_numpy(z["tgt"]).double(), torch.from_numpy(z["mask"]), torch.from_numpy(z["query_pos"]).double(), NH)
    assert ib.rel_l2(y, torch.from_numpy(z["self_out"])) < 2e-6
    assert ib.rel_l2(mo.ffn_layer(ffn, torch.from_numpy(z["self_out"]).double()), torch.from_numpy(z["ffn_out"])) < 2e-6


def _mask_path_case(case):
    import os
    import numpy as np
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"xdecoder_mask_path_{case}.npz"))
    C, MD, Q, NH, FF, NL = (int(v) for v in z["meta"])
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    x = [torch.from_numpy(z[f"x{i}"]) for i in range(3)]
    return z, (C, MD, Q, NH, FF, NL), sd, x, torch.from_numpy(z["mask_features"])


@pytest.mark.parametrize("case", ["small", "q101"])
def test_xdecoder_mask_path_oracle_against_reference_goldens(case):
    """oracle.xdecoder_mask_path against outputs of the UNMODIFIED reference XDecoder.forward (task='seg') executed on a stand-in object
    (tests/golden/make_golden_xdecoder_mask_path.py; interface/xdecoder.py:191-329).  The attention masks are thresholds of fp32 logits:
    the fp64 oracle may flip single mask bits, so the comparison allows the small drift that causes in the later layers."""
    import iuvl_b200 as ib
    from oracle import mask_head_oracle as mo
    z, (C, MD, Q, NH, FF, NL), sd, x, mf = _mask_path_case(case)
    masks = mo.xdecoder_mask_path({k: v.double() for k, v in sd.items()}, [t.double() for t in x], mf.double(), Q, NH, [0, 1, 2, 0, 1, 2, 0, 1, 2][:NL])
    assert ib.rel_l2(masks[0], torch.from_numpy(z["aux0"]) if "aux0" in z.files else masks[0]) < 1e-5
    assert ib.rel_l2(masks[-1], torch.from_numpy(z["pred_masks"])) < 2e-3


def test_mask_head_oracle_properties():
    """Size-independent properties of the row-N4 oracles: the antialiased filter preserves constants and is the identity at scale 1;
    an all-False attention mask equals no mask; a fully masked key range leaves the output independent of those keys."""
    from oracle import mask_head_oracle as mo
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 19, 23, generator=g, dtype=torch.float64)
    assert float((mo.resize_bicubic_aa(torch.full((1, 17, 31), 2.5, dtype=torch.float64), (5, 7)) - 2.5).abs().max()) < 1e-12
    assert float((mo.resize_bicubic_aa(x, (19, 23)) - x).abs().max()) < 1e-12
    for n_in, n_out in ((256, 32), (40, 10), (24, 12), (9, 16)):
        w = mo.aa_weights(n_in, n_out)
        assert float((w.sum(1) - 1).abs().max()) < 1e-12
    C, NH, Q, B, HW = 128, 2, 5, 2, 12
    sd = {"multihead_attn.in_proj_weight": torch.randn(3 * C, C, generator=g, dtype=torch.float64) * 0.1,
          "multihead_attn.in_proj_bias": torch.randn(3 * C, generator=g, dtype=torch.float64) * 0.1,
          "multihead_attn.out_proj.weight": torch.randn(C, C, generator=g, dtype=torch.float64) * 0.1,
          "multihead_attn.out_proj.bias": torch.randn(C, generator=g, dtype=torch.float64) * 0.1,
          "norm.weight": torch.ones(C, dtype=torch.float64), "norm.bias": torch.zeros(C, dtype=torch.float64)}
    tgt, mem = torch.randn(Q, B, C, generator=g, dtype=torch.float64), torch.randn(HW, B, C, generator=g, dtype=torch.float64)
    none = mo.cross_attention_layer(sd, tgt, mem, None, None, None, NH)
    allf = mo.cross_attention_layer(sd, tgt, mem, torch.zeros(B * NH, Q, HW, dtype=torch.bool), None, None, NH)
    assert torch.equal(none, allf)
    mask = torch.zeros(B * NH, Q, HW, dtype=torch.bool)
    mask[:, :, 6:] = True
    mem2 = mem.clone()
    mem2[6:] = torch.randn(HW - 6, B, C, generator=g, dtype=torch.float64)
    a = mo.cross_attention_layer(sd, tgt, mem, mask, None, None, NH)
    b = mo.cross_attention_layer(sd, tgt, mem2, mask, None, None, NH)
    assert float((a - b).abs().max()) < 1e-12


@pytest.mark.parametrize("g,ws", [(20, 14), (16, 16)])
def test_key_bias_is_invisible_and_value_bias_folds_out(g, ws):
    """The identity behind DESIGN.md's next step 4 (image_encoder.py:239-255 with the pad tokens of :183-187, 271-275): a constant added
    to every key row shifts all scores of a query by the same amount (invisible to the softmax) and a constant added to every value
    row comes out of the attention unchanged.  So with qkv = x W^T + b, the attention of the biased rows — pad tokens' rows equal to b —
    equals the attention of the rows with the K and V parts of the bias REMOVED (pad rows: b_q | 0 | 0), plus b_v."""
    from tests.util import ref_attention_core
    gen = torch.Generator().manual_seed(g * 100 + ws)
    B, heads, hd = 2, 3, 8
    D = heads * hd
    raw = torch.randn(B * g * g, 3 * D, generator=gen, dtype=torch.float64)
    bias = torch.randn(3 * D, generator=gen, dtype=torch.float64)
    rel_h = torch.randn(2 * ws - 1, hd, generator=gen, dtype=torch.float64) * 0.5
    rel_w = torch.randn(2 * ws - 1, hd, generator=gen, dtype=torch.float64) * 0.5
    full = ref_attention_core(raw + bias, rel_h, rel_w, bias, B, g, ws, heads)
    bias_q_only = bias.clone()
    bias_q_only[D:] = 0
    reduced = ref_attention_core(raw + bias_q_only, rel_h, rel_w, bias_q_only, B, g, ws, heads)
    assert torch.allclose(full, reduced + bias[2 * D:], rtol=1e-10, atol=1e-10)
