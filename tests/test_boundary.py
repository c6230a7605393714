"""Drop-in boundary (CPU): constructor signature, state_dict keys / shapes / module types identical to the reference
(fixture tests/golden/reference_state_dict_keys.json was dumped from the unmodified reference via sam.build_sam),
C-ABI library loads and exports every symbol include/samvit_b200.h declares, and the product fails loudly
without a GPU instead of falling back."""
import inspect
import json
import os
import re

import pytest
import torch

import iuvl_b200 as ib
from iuvl_b200 import cabi
from iuvl_b200.encoder import ImageEncoderViT, build_encoder
from tests.util import GOLDEN, ROOT


@pytest.fixture(scope="module")
def ref_keys():
    with open(os.path.join(GOLDEN, "reference_state_dict_keys.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("preset", ["vit_b", "vit_l", "vit_h"])
def test_state_dict_matches_reference(preset, ref_keys):
    with torch.device("meta"):
        enc = build_encoder(ib.PRESETS[preset])
    mine = [[k, list(v.shape)] for k, v in enc.state_dict().items()]
    assert mine == ref_keys[preset]                      # same keys, same order, same shapes
    assert [[k, list(s)] for k, s in ib.state_dict_spec(ib.PRESETS[preset])] == ref_keys[preset]
    assert all(v.dtype == torch.float32 for v in enc.state_dict().values())


@pytest.mark.parametrize("preset", ["vit_b", "vit_h"])
def test_module_tree_matches_reference(preset, ref_keys):
    """named_modules() names and leaf types: the optimizer grouping tests isinstance(LayerNorm/GroupNorm/...)
    (trainer/xdecoder_trainer.py:61-73,103-131)."""
    with torch.device("meta"):
        enc = build_encoder(ib.PRESETS[preset])
    mine = [[n, type(m).__name__] for n, m in enc.named_modules()]
    assert mine == ref_keys[preset + "_modules"]


def test_constructor_signature_is_the_reference_one():
    # image_encoder.py:18-36
    expected = ["self", "img_size", "patch_size", "in_chans", "embed_dim", "depth", "num_heads", "mlp_ratio", "out_chans",
                "qkv_bias", "norm_layer", "act_layer", "use_abs_pos", "use_rel_pos", "rel_pos_zero_init", "window_size",
                "global_attn_indexes"]
    sig = inspect.signature(ImageEncoderViT.__init__)
    assert list(sig.parameters) == expected
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["img_size"], d["patch_size"], d["in_chans"], d["embed_dim"], d["depth"], d["num_heads"]) == (1024, 16, 3, 768, 12, 12)
    assert d["mlp_ratio"] == 4.0 and d["out_chans"] == 256 and d["window_size"] == 0 and d["global_attn_indexes"] == ()


def test_load_state_dict_strict_and_eval_train_to():
    cfg = ib.PRESETS["tiny64"]
    enc = build_encoder(cfg)
    sd = ib.make_state_dict(cfg, 3)
    res = enc.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(enc.blocks[1].attn.rel_pos_h, sd["blocks.1.attn.rel_pos_h"])
    assert enc.blocks[1].attn.rel_pos_h.shape[0] == 127 and enc.blocks[0].attn.rel_pos_h.shape[0] == 27
    enc.train(); enc.eval(); enc.to("cpu")
    assert enc.img_size == 1024
    # SAM-style prefixed, non-strict load as in build_sam.py:96-99
    holder = torch.nn.Module()
    holder.image_encoder = enc
    msg = holder.load_state_dict({"image_encoder." + k: v for k, v in sd.items()}, strict=False)
    assert not msg.unexpected_keys


def test_no_cpu_fallback():
    enc = build_encoder(ib.PRESETS["tiny64"])
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        enc(torch.zeros(1, 3, 1024, 1024))


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        ImageEncoderViT(use_rel_pos=False, window_size=14)
    with pytest.raises(NotImplementedError):
        ImageEncoderViT(use_rel_pos=True, window_size=14, act_layer=torch.nn.ReLU)


def test_cabi_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "samvit_b200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(cabi.SYMBOLS), declared ^ set(cabi.SYMBOLS)
    lib = cabi.lib()                                   # loads the .so and resolves every symbol
    for name in declared:
        assert hasattr(lib, name)
    assert lib.svb_version() >= 100
    # the separate probe library (test / measurement infrastructure) and its header
    with open(os.path.join(ROOT, "include", "samvit_b200_probe.h")) as f:
        phdr = f.read()
    pdecl = set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", phdr))
    assert pdecl == set(cabi.PROBE_SYMBOLS), pdecl ^ set(cabi.PROBE_SYMBOLS)
    plib = cabi.probe_lib()
    for name in pdecl:
        assert hasattr(plib, name)
    assert not (pdecl - {"svb_probe_last_error"}) & set(cabi.SYMBOLS)        # no probe entry point in the product library


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "interactable-unified-vision-language_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "sam_vit_oracle" not in src and "import oracle" not in src and "from oracle" not in src, fn


def test_flops_per_image_match_survey():
    # 0.98682 / 2.92971 / 5.78735 TFLOP
    for k, v in (("vit_b", 0.98682e12), ("vit_l", 2.92971e12), ("vit_h", 5.78735e12)):
        assert abs(ib.PRESETS[k].flops_per_image() / v - 1) < 1e-5


def test_decoder_side_drop_ins_load_reference_state_dicts_and_refuse_cpu():
    """Host logic of the rows N1 / N4 drop-ins without a GPU: the reference's state_dict keys load with strict=True (fixtures written by
    the unmodified reference classes) and a CPU call raises instead of falling back."""
    import os
    import numpy as np
    import pytest
    import torch
    from iuvl_b200.mask_head import CrossAttentionLayer, FFNLayer, MaskPredictionHead, SelfAttentionLayer, XDecoderMaskPath
    from tests.util import GOLDEN

    def sd_of(z, prefix):
        return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}

    z = np.load(os.path.join(GOLDEN, "cross_attn_small.npz"))
    C, NH = (int(v) for v in z["meta"])
    layer = CrossAttentionLayer(C, NH)
    layer.load_state_dict(sd_of(z, "sd."), strict=True)
    with pytest.raises(RuntimeError), torch.no_grad():
        layer(torch.from_numpy(z["tgt"]), torch.from_numpy(z["memory"]))
    z = np.load(os.path.join(GOLDEN, "decoder_layers_small.npz"))
    C, NH, FF = (int(v) for v in z["meta"])
    sa, ffn = SelfAttentionLayer(C, NH), FFNLayer(C, FF)
    sa.load_state_dict(sd_of(z, "sa."), strict=True)
    ffn.load_state_dict(sd_of(z, "ffn."), strict=True)
    with pytest.raises(RuntimeError), torch.no_grad():
        ffn(torch.from_numpy(z["tgt"]))
    z = np.load(os.path.join(GOLDEN, "mask_head_small.npz"))
    C, MD, Q, NH, th, tw = (int(v) for v in z["meta"])
    head = MaskPredictionHead(C, MD, Q, NH)
    head.load_state_dict(sd_of(z, "sd."), strict=True)
    with pytest.raises(RuntimeError), torch.no_grad():
        head(torch.from_numpy(z["output"]), torch.from_numpy(z["mask_features"]), (th, tw))
    z = np.load(os.path.join(GOLDEN, "xdecoder_mask_path_small.npz"))
    C, MD, Q, NH, FF, NL = (int(v) for v in z["meta"])
    path = XDecoderMaskPath(C, MD, Q, NH, FF, 3, [0, 1, 2, 0, 1, 2, 0, 1, 2][:NL])
    path.load_state_dict(sd_of(z, "sd."), strict=True)
    assert set(path.state_dict().keys()) == set(sd_of(z, "sd.").keys())
    with pytest.raises(RuntimeError), torch.no_grad():
        path([torch.from_numpy(z[f"x{i}"]) for i in range(3)], torch.from_numpy(z["mask_features"]))
    # autograd-enabled calls are refused as well (forward only)
    with pytest.raises(RuntimeError):
        head(torch.from_numpy(z["x0"]).requires_grad_(), torch.from_numpy(z["mask_features"]), (4, 4))


def test_install_into_reference_builds_the_drop_in():
    """The drop-in mechanism itself: after install_into_reference() the reference's OWN factory (sam/build_sam.py:108-112,
    `sam_model_registry['vit_b'](checkpoint=None)`) builds this repo's encoder class with the reference's 209 state_dict keys, and a
    reference state_dict loads strictly.  Needs the reference package: /root/reference in the build container, else the unmodified
    copy under oracle/_ref that build() makes."""
    import importlib
    import sys
    roots = [p for p in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")) if os.path.exists(os.path.join(p, "sam", "build_sam.py"))]
    if not roots:
        pytest.skip("no copy of the reference's sam package here")
    sys.path.insert(0, roots[0])
    try:
        for m in [m for m in sys.modules if m == "sam" or m.startswith("sam.")]:
            del sys.modules[m]
        ref_cls = importlib.import_module("sam.modeling.image_encoder").ImageEncoderViT
        import torch
        from iuvl_b200 import encoder as enc_mod
        torch.manual_seed(0)
        ref_sd = importlib.import_module("sam.build_sam").sam_model_registry["vit_b"](checkpoint=None).image_encoder.state_dict()
        enc_mod.install_into_reference()
        sam = importlib.import_module("sam.build_sam").sam_model_registry["vit_b"](checkpoint=None)
        e = sam.image_encoder
        assert type(e) is enc_mod.ImageEncoderViT and type(e) is not ref_cls
        sd = e.state_dict()
        assert len(sd) == 209 and list(sd) == list(ref_sd)
        assert all(tuple(sd[k].shape) == tuple(ref_sd[k].shape) and sd[k].dtype == ref_sd[k].dtype for k in sd)
        res = e.load_state_dict(ref_sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        # _build_sam marks the encoder parameters trainable through named_parameters() (build_sam.py:101-105)
        assert all(p.requires_grad for n, p in sam.named_parameters() if n.startswith("image_encoder"))
        with pytest.raises(RuntimeError):
            e(torch.zeros(1, 3, 1024, 1024))                 # no CPU path
    finally:
        sys.path.remove(roots[0])
        for m in [m for m in sys.modules if m == "sam" or m.startswith("sam.")]:
            del sys.modules[m]


def _header_prototypes(path):
    """name -> list of C parameter declarations, from the (comment-stripped) header text."""
    with open(path) as f:
        txt = re.sub(r"/\*.*?\*/", " ", f.read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(svb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        params = " ".join(m.group(2).split())
        protos[m.group(1)] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
    return protos


def _kind(cdecl):
    if "*" in cdecl or "svb_stream_t" in cdecl:
        return "ptr"
    if re.search(r"\b(float|double)\b", cdecl):
        return "real"
    return "int"


def test_ctypes_prototypes_agree_with_the_header():
    """Every binding in cabi.SYMBOLS has as many arguments as the declaration in include/samvit_b200.h, of the same kind (pointer /
    integer / floating point) in the same order: a slip here corrupts arguments silently at call time."""
    import ctypes as C
    protos = _header_prototypes(os.path.join(ROOT, "include", "samvit_b200.h"))
    assert set(protos) == set(cabi.SYMBOLS)
    ptr_types = (C.c_void_p, C.c_char_p)
    for name, (_, argtypes) in cabi.SYMBOLS.items():
        params = protos[name]
        assert len(params) == len(argtypes), (name, params, argtypes)
        for p, a in zip(params, argtypes):
            if a in ptr_types or (isinstance(a, type) and issubclass(a, C._Pointer)):
                want = "ptr"
            elif a in (C.c_float, C.c_double):
                want = "real"
            else:
                want = "int"
            assert _kind(p) == want, (name, p, a)
        # float vs double must match exactly (a float passed where a double is read is garbage)
        for p, a in zip(params, argtypes):
            if a is C.c_double:
                assert re.search(r"\bdouble\b", p) and "*" not in p, (name, p)
            if a is C.c_float:
                assert re.search(r"\bfloat\b", p) and "*" not in p, (name, p)
