"""GPU parity of the pixel decoder (scope row N1): the streaming kernels of csrc/pixdec.cu against the torch ops they replace, and
the drop-in MSDeformAttnPixelDecoder against outputs of the UNMODIFIED reference class (tests/golden/pixel_decoder_*.npz)."""
import pytest
import torch
import torch.nn.functional as F

import iuvl_b200 as ib
from iuvl_b200 import cabi
from tests.test_oracle import _pixel_decoder_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dt(t):
    return cabi.DTYPE_BF16 if t == torch.bfloat16 else cabi.DTYPE_F32


@pytest.mark.parametrize("src_dtype,dst_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)])
def test_layout_changes_are_exact(src_dtype, dst_dtype):
    g = torch.Generator().manual_seed(1)
    B, C, H, W = 2, 72, 9, 13                                   # ragged against the 32 x 32 tiles
    x = torch.randn(B, C, H, W, generator=g).to(src_dtype).to(DEV)
    rows = torch.full((B, H * W + 5, C), 7.0, dtype=dst_dtype, device=DEV)       # sample stride larger than the level
    cabi.check(cabi.lib().svb_nchw_to_rows(x.data_ptr(), _dt(src_dtype), rows.data_ptr(), _dt(dst_dtype), B, C, H * W, (H * W + 5) * C,
                                           cabi.stream_ptr()), "nchw_to_rows")
    want = x.flatten(2).transpose(1, 2).to(dst_dtype)
    assert torch.equal(rows[:, :H * W], want) and bool((rows[:, H * W:] == 7.0).all())
    if dst_dtype == torch.float32:
        back = torch.empty(B, C, H, W, device=DEV)
        cabi.check(cabi.lib().svb_rows_to_nchw(rows.data_ptr(), (H * W + 5) * C, back.data_ptr(), B, C, H * W, cabi.stream_ptr()), "rows_to_nchw")
        assert torch.equal(back, x.float())


@pytest.mark.parametrize("C,groups,relu,out_dtype", [(64, 32, False, torch.float32), (512, 32, True, torch.float32), (128, 32, True, torch.bfloat16)])
def test_groupnorm_rows(C, groups, relu, out_dtype):
    g = torch.Generator().manual_seed(2)
    B, H, W = 3, 11, 7
    x = (torch.randn(B, H * W, C, generator=g) * 2 + 0.5).to(DEV)
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    out = torch.empty(B, H * W, C, dtype=out_dtype, device=DEV)
    ws = torch.empty(B * groups * 2, dtype=torch.float64, device=DEV)
    cabi.check(cabi.lib().svb_groupnorm_rows(x.data_ptr(), 0, gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), _dt(out_dtype), 0, B, H * W, C,
                                             groups, 1e-5, 1 if relu else 0, ws.data_ptr(), cabi.stream_ptr()), "groupnorm_rows")
    want = F.group_norm(x.transpose(1, 2).reshape(B, C, H, W).double(), groups, gamma.double(), beta.double(), 1e-5)
    if relu:
        want = torch.relu(want)
    want = want.flatten(2).transpose(1, 2)
    assert ib.rel_l2(out, want) < (1e-6 if out_dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("h,w,oh,ow", [(6, 5, 12, 10), (12, 12, 24, 24), (5, 7, 9, 16)])
def test_upsample_add_rows(h, w, oh, ow):
    g = torch.Generator().manual_seed(3)
    B, C = 2, 64
    src = torch.randn(B, h * w + 3, C, generator=g).to(DEV)                       # strided samples (a level inside src_flatten)
    dst = torch.randn(B, oh * ow, C, generator=g).to(DEV)
    want = dst.double() + F.interpolate(src[:, :h * w].transpose(1, 2).reshape(B, C, h, w).double(), size=(oh, ow), mode="bilinear",
                                        align_corners=False).flatten(2).transpose(1, 2)
    cabi.check(cabi.lib().svb_upsample_add_rows(src.data_ptr(), (h * w + 3) * C, dst.data_ptr(), B, h, w, oh, ow, C, cabi.stream_ptr()), "upsample")
    assert ib.rel_l2(dst, want) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_im2col3x3_matches_conv(dtype):
    g = torch.Generator().manual_seed(4)
    B, C, H, W, CO = 2, 64, 7, 9, 48
    x = torch.randn(B, H * W, C, generator=g).to(DEV)
    wt = torch.randn(CO, C, 3, 3, generator=g).to(DEV)
    col = torch.empty(B * H * W, 9 * C, dtype=dtype, device=DEV)
    cabi.check(cabi.lib().svb_im2col3x3_rows(x.data_ptr(), col.data_ptr(), _dt(dtype), B, H, W, C, cabi.stream_ptr()), "im2col3x3")
    got = col.double() @ wt.permute(0, 2, 3, 1).reshape(CO, -1).double().t()
    xin = x.to(dtype).double() if dtype == torch.bfloat16 else x.double()
    want = F.conv2d(xin.transpose(1, 2).reshape(B, C, H, W), wt.double(), padding=1).flatten(2).transpose(1, 2).reshape(B * H * W, CO)
    assert ib.rel_l2(got, want) < 1e-9


def test_add_cast_bcast():
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(3, 40, 64, generator=g).to(DEV), torch.randn(40, 64, generator=g).to(DEV)
    for dt in (torch.float32, torch.bfloat16):
        out = torch.empty(3, 40, 64, dtype=dt, device=DEV)
        cabi.check(cabi.lib().svb_add_cast_bcast(a.data_ptr(), b.data_ptr(), b.numel(), out.data_ptr(), _dt(dt), a.numel(), cabi.stream_ptr()), "bcast")
        assert torch.equal(out, (a + b[None]).to(dt))


@pytest.mark.parametrize("case", ["small", "wide"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_pixel_decoder_against_reference_goldens(case, precision, tol):
    """Drop-in MSDeformAttnPixelDecoder against the UNMODIFIED reference class (transformer_encoder_deform.py:164-359): fp32 validation
    mode <= 1e-4 relative L2 per output map, bf16 (tensor cores, fp32 accumulate, 2-6 GEMMs + GroupNorms deep per map) <= 2e-2."""
    from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder
    z, (C, MD, M, NL, F_), feats, sd = _pixel_decoder_case(case)
    mod = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=M, transformer_dim_feedforward=F_, transformer_enc_layers=NL,
                                   conv_dim=C, mask_dim=MD, norm="GN", transformer_in_features=["res3", "res4", "res5"], common_stride=4)
    mod.load_state_dict(sd, strict=True)
    mod.to(DEV).eval()
    mod.precision = precision
    with torch.no_grad():
        mask, multi = mod({k: v.to(DEV) for k, v in feats.items()})
    errs = {"mask_features": ib.rel_l2(mask, torch.from_numpy(z["mask_features"]))}
    assert tuple(mask.shape) == tuple(z["mask_features"].shape) and len(multi) == 3
    for i, m in enumerate(multi):
        assert tuple(m.shape) == tuple(z[f"multi{i}"].shape)
        errs[f"multi{i}"] = ib.rel_l2(m, torch.from_numpy(z[f"multi{i}"]))
    print(case, precision, errs)
    assert max(errs.values()) < tol, errs


def test_conv3x3_implicit_gemm_matches_conv2d():
    """svb_conv3x3_rows (zero-padded bf16 map + row-shifted TMA boxes per filter tap inside the tcgen05 GEMM) against F.conv2d of the
    bf16-rounded operands in fp64: the `output_conv` of the pixel decoder (transformer_encoder_deform.py:259-268) at 256-wide maps."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(21)
    for (B, H, W, Cin, Cout, relu) in ((2, 16, 128, 64, 64, 0), (1, 24, 256, 128, 96, 1), (1, 256, 256, 512, 512, 1)):
        x = torch.randn(B, H, W, Cin, generator=g)
        wt = torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5
        bias = torch.randn(Cout, generator=g)
        ref = F.conv2d(x.bfloat16().double().permute(0, 3, 1, 2), wt.bfloat16().double(), bias.double(), padding=1)
        if relu:
            ref = torch.relu(ref)
        ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout)
        xd, bd = x.to(DEV).contiguous(), bias.to(DEV)
        w3 = wt.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(DEV).bfloat16().contiguous()          # (Cout, ky, kx, Cin)
        out = torch.full((B * H * W, Cout), float("nan"), device=DEV)
        pad = torch.empty(B * (H + 2) * (W + 2), Cin, dtype=torch.bfloat16, device=DEV)
        cabi.check(cabi.lib().svb_conv3x3_rows(xd.data_ptr(), w3.data_ptr(), bd.data_ptr(), out.data_ptr(), pad.data_ptr(), B, H, W, Cin, Cout, relu,
                                               cabi.stream_ptr()), "svb_conv3x3_rows")
        torch.cuda.synchronize()
        err = ib.rel_l2(out, ref)
        assert err < 1e-5, ((B, H, W, Cin, Cout), err)
        # edges: the first / last row and column of the map see the zero border
        o4, r4 = out.cpu().reshape(B, H, W, Cout), ref.reshape(B, H, W, Cout)
        for sl in (o4[:, 0], o4[:, -1], o4[:, :, 0], o4[:, :, -1]), (r4[:, 0], r4[:, -1], r4[:, :, 0], r4[:, :, -1]):
            pass
        assert ib.rel_l2(torch.cat([o4[:, 0].flatten(), o4[:, -1].flatten(), o4[:, :, 0].flatten(), o4[:, :, -1].flatten()]),
                         torch.cat([r4[:, 0].flatten(), r4[:, -1].flatten(), r4[:, :, 0].flatten(), r4[:, :, -1].flatten()])) < 1e-5


def test_pixel_decoder_implicit_conv_equals_im2col_path():
    """The whole pixel decoder at the step1.yaml geometry on the maps of one 1024 x 1024 image (res2 = 256 x 256: the implicit-GEMM
    3x3 convolution is taken) against the same module with the im2col operand: same bf16 operands, same fp32 accumulation order per
    K block -> the outputs agree to fp32 rounding."""
    from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder
    torch.manual_seed(5)
    dec = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=1,
                                   conv_dim=512, mask_dim=512, norm="GN").to(DEV).eval()
    feats = {f"res{2 + i}": torch.randn(1, c, 256 >> i, 256 >> i, device=DEV).bfloat16() for i, c in enumerate((128, 256, 512, 1024))}
    outs = []
    with torch.no_grad():
        for flag in (True, False):
            dec.implicit_conv = flag
            mask, multi = dec(feats)
            outs.append(mask)
    assert torch.isfinite(outs[0]).all()
    assert ib.rel_l2(outs[0], outs[1]) < 1e-5, ib.rel_l2(outs[0], outs[1])


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 2e-2)])
def test_pixel_decoder_at_the_step1_geometry(precision, tol):
    """The composed module at the widths the reference runs (configs/step1.yaml: conv_dim = mask_dim = 512, 8 heads, d_ffn 1024, 6
    deformable encoder layers) against samples of the UNMODIFIED reference class's outputs; weights regenerated from a seed."""
    from tests.test_oracle import _sampled_err, _step1_pixel_decoder
    z, mod, sd, feats, _ = _step1_pixel_decoder()
    mod.to(DEV).eval()
    mod.precision = precision
    with torch.no_grad():
        mask, multi = mod({k: v.to(DEV) for k, v in feats.items()})
    errs = {"mask_features": _sampled_err(mask, z, "mask_features")}
    for i, m in enumerate(multi):
        errs[f"multi{i}"] = _sampled_err(m, z, f"multi{i}")
    print("step1", precision, errs)
    assert max(errs.values()) < tol, errs


def test_encoder_feeds_pixel_decoder():
    """BASELINE config 5 in miniature: the ViT encoder's four maps go straight into the pixel decoder (bf16 hand-over, no `.float()`
    up-cast of transformer_encoder_deform.py:320,345), against the two CPU oracles chained in fp64."""
    from oracle import pixel_decoder_oracle as po
    from oracle import sam_vit_oracle as orc
    from iuvl_b200.encoder import build_encoder
    from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder
    cfg = ib.PRESETS["tiny80"]
    sd = ib.make_state_dict(cfg, 99, rel_std=0.1)
    x = ib.make_images(2, cfg, 5)
    enc = build_encoder(cfg)
    enc.load_state_dict(sd)
    enc.to(DEV)
    torch.manual_seed(11)
    dec = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=4, transformer_dim_feedforward=128, transformer_enc_layers=2,
                                   conv_dim=64, mask_dim=32, norm="GN")
    with torch.no_grad():
        for layer in dec.transformer.encoder.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.2)
            layer.self_attn.attention_weights.weight.normal_(0, 0.2)
    dsd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    feats_ref = orc.encoder_forward_cfg(sd, x, cfg)
    mask_ref, multi_ref = po.pixel_decoder({k: v.double() for k, v in dsd.items()}, {k: v.double() for k, v in feats_ref.items()}, 4, 2)
    dec.to(DEV).eval()
    for precision, tol in (("fp32", 2e-4), ("bf16", 3e-2)):
        enc.precision = precision
        enc.out_dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        dec.precision = precision
        with torch.no_grad():
            mask, multi = dec(enc(x.to(DEV)))
        errs = [ib.rel_l2(mask, mask_ref)] + [ib.rel_l2(a, b) for a, b in zip(multi, multi_ref)]
        assert max(errs) < tol, (precision, errs)
