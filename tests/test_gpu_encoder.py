"""GPU parity of the whole encoder (through ImageEncoderViT -> C ABI) against
 (a) golden samples produced by the UNMODIFIED reference (tests/golden/*.npz), at full ViT-B/L/H sizes, and
 (b) the CPU oracle run on the box for the small configurations,
in both modes: fp32 validation (<= 1e-4 rel-L2) and bf16 tensor-core (<= 1e-2 rel-L2 per embedding) — the tolerances
BASELINE.json's north_star states."""
import pytest
import torch

import iuvl_b200 as ib
from iuvl_b200.encoder import build_encoder
from tests.util import load_golden, sampled_rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_FP32 = 1e-4
TOL_BF16 = 1e-2
KEYS = ("res2", "res3", "res4", "res5")


def _setup(case):
    g = load_golden(case)
    cfg = ib.PRESETS[str(g["meta_preset"])]
    sd = ib.make_state_dict(cfg, int(g["meta_weight_seed"]), rel_std=float(g["meta_rel_std"]))
    hw = tuple(int(v) for v in g["meta_hw"]) if "meta_hw" in g else None
    x = ib.make_images(int(g["meta_batch"]), cfg, int(g["meta_image_seed"]), hw=hw)
    enc = build_encoder(cfg)
    enc.load_state_dict(sd, strict=True)
    enc.to(DEV)
    return g, cfg, sd, x, enc


def _check(enc, x, g, precision, tol, taps=True, chunk=8):
    enc.precision = precision
    enc.max_chunk = chunk
    if taps:
        enc.enable_taps(True)
    with torch.no_grad():
        out = enc(x.to(DEV))
    torch.cuda.synchronize()
    errs = {}
    if taps:
        names = ["embed"] + [f"block{i}" for i in range(enc.cfg.depth)]
        for i, n in enumerate(names):
            errs["tap." + n] = sampled_rel_l2(enc.read_tap(i - 1), g, "tap." + n)
    for k in KEYS:
        assert out[k].dtype == enc.out_dtype
        errs["out." + k] = sampled_rel_l2(out[k].float(), g, "out." + k)
    bad = {k: v for k, v in errs.items() if not (v < tol)}
    assert not bad, f"{precision}: over tolerance {tol}: {bad}\nall: {errs}"
    return out, errs


@pytest.mark.parametrize("case", ["tiny64_std", "tiny64_stress", "tiny80_std", "tiny80_stress"])
def test_tiny_fp32_validation_mode(case):
    g, cfg, sd, x, enc = _setup(case)
    _check(enc, x, g, "fp32", TOL_FP32)


@pytest.mark.parametrize("case", ["tiny64_std", "tiny64_stress", "tiny80_std", "tiny80_stress"])
def test_tiny_bf16(case):
    g, cfg, sd, x, enc = _setup(case)
    # with N(0, 0.5^2) rel-pos tables the bias dominates the logits; bf16 q/k rounding is amplified -> looser bar
    _check(enc, x, g, "bf16", TOL_BF16 if case.endswith("std") else 3e-2)


@pytest.mark.parametrize("case,precision", [("tiny64_wide", "fp32"), ("tiny64_wide", "bf16"), ("tiny80_tall", "fp32"), ("tiny80_tall", "bf16"),
                                            ("vit_b_wide", "bf16")])
def test_other_canvases_against_reference_goldens(case, precision):
    """Scope row N3: 1024 x 2048 and 1536 x 512 inputs against goldens of the unmodified reference, which takes its
    bicubic pos_embed (image_encoder.py:124-132) and linear rel_pos (:319-330) fallbacks on them."""
    g, cfg, sd, x, enc = _setup(case)
    out, errs = _check(enc, x, g, precision, TOL_FP32 if precision == "fp32" else TOL_BF16, taps=False)
    H, W = x.shape[2:]
    assert tuple(out["res2"].shape) == (1, 128, H // 4, W // 4) and tuple(out["res5"].shape) == (1, 1024, H // 32, W // 32)


def test_resize_kernels_match_torch_interpolate():
    """The two table fallbacks alone against the calls the reference makes: F.interpolate(..., mode='bicubic') on pos_embed
    (image_encoder.py:124-132) and F.interpolate(..., mode='linear') on a rel_pos table (:321-330)."""
    import torch.nn.functional as F
    from iuvl_b200 import cabi
    g = torch.Generator().manual_seed(3)
    for (h1, w1) in ((64, 128), (96, 32), (32, 32), (128, 128)):
        pos = torch.randn(1, 64, 64, 48, generator=g)
        ref = F.interpolate(pos.permute(0, 3, 1, 2), scale_factor=(h1 / 64, w1 / 64), mode="bicubic").permute(0, 2, 3, 1)
        src, dst = pos.to(DEV), torch.empty(1, h1, w1, 48, device=DEV)
        cabi.check(cabi.lib().svb_resize_pos_embed(src.data_ptr(), dst.data_ptr(), 64, 64, h1, w1, 48, cabi.stream_ptr()))
        assert torch.allclose(dst.cpu(), ref, rtol=1e-5, atol=2e-6), float((dst.cpu() - ref).abs().max())
    for L1 in (255, 63, 191, 127, 27):
        tab = torch.randn(127, 80, generator=g)
        ref = F.interpolate(tab.reshape(1, 127, -1).permute(0, 2, 1), size=L1, mode="linear").reshape(-1, L1).permute(1, 0)
        src, dst = tab.to(DEV), torch.empty(L1, 80, device=DEV)
        cabi.check(cabi.lib().svb_resize_rel_pos(src.data_ptr(), dst.data_ptr(), 127, L1, 80, cabi.stream_ptr()))
        assert torch.allclose(dst.cpu(), ref, rtol=1e-5, atol=2e-6), (L1, float((dst.cpu() - ref).abs().max()))


def test_tiny_against_cpu_oracle_full_tensors():
    """Full-tensor comparison (not samples) with the oracle run here on the host CPU."""
    from oracle import sam_vit_oracle as orc
    cfg = ib.PRESETS["tiny80"]
    sd = ib.make_state_dict(cfg, 99, rel_std=0.1)
    x = ib.make_images(2, cfg, 5)
    ref = orc.encoder_forward_cfg(sd, x, cfg)
    enc = build_encoder(cfg)
    enc.load_state_dict(sd)
    enc.to(DEV)
    for precision, tol in (("fp32", TOL_FP32), ("bf16", TOL_BF16)):
        enc.precision = precision
        with torch.no_grad():
            out = enc(x.to(DEV))
        for k in KEYS:
            assert out[k].shape == ref[k].shape
            assert ib.rel_l2(out[k], ref[k]) < tol, (precision, k, ib.rel_l2(out[k], ref[k]))


@pytest.mark.parametrize("case,precision", [
    ("vit_b_std", "fp32"), ("vit_b_std", "bf16"), ("vit_b_stress", "fp32"), ("vit_b_stress", "bf16"),
    ("vit_l_std", "fp32"), ("vit_l_std", "bf16"), ("vit_h_std", "fp32"), ("vit_h_std", "bf16"), ("vit_h_stress", "fp32"),
    ("vit_h_stress", "bf16"),
])
def test_full_size_against_reference_goldens(case, precision):
    g, cfg, sd, x, enc = _setup(case)
    tol = TOL_FP32 if precision == "fp32" else (TOL_BF16 if case.endswith("std") else 3e-2)
    _check(enc, x, g, precision, tol)


def _positions_check(out, g, perm, tol):
    """Every position p of a batch built as x[perm[p]] against the reference samples of image perm[p] (the golden holds samples of a
    3-image reference batch: flat indices into (3, C, H, W)); returns the worst rel-L2 over positions and outputs."""
    worst = 0.0
    for k in KEYS:
        shape = [int(v) for v in g[f"out.{k}.shape"]]
        per = shape[1] * shape[2] * shape[3]
        idx = torch.from_numpy(g[f"out.{k}.idx"])
        val = torch.from_numpy(g[f"out.{k}.val"]).double()
        img, rest = idx // per, idx % per
        assert tuple(out[k].shape[1:]) == tuple(shape[1:])
        flat = out[k].reshape(out[k].shape[0], -1)
        for p, b in enumerate(perm):
            sel = img == b
            got = flat[p][rest[sel].to(flat.device)].double().cpu()
            err = float((got - val[sel]).norm() / val[sel].norm())
            assert err < tol, f"{k}: position {p} (image {b}): rel-L2 {err:.3e} >= {tol}"
            worst = max(worst, err)
    return worst


@pytest.mark.parametrize("case,batch,out_dtype", [
    ("vit_b_std3", 16, torch.float32),       # BASELINE.json configs[1]: ViT-B, batch 16
    ("vit_l_std3", 32, torch.float32),       # configs[2]: ViT-L, batch 32
    ("vit_h_std3", 13, torch.float32),       # a pass size with a ragged last GEMM tile (13 * 4096 rows)
    ("vit_h_std3", 64, torch.bfloat16),      # configs[3] = the benchmarked step: 64 images, passes of 12+12+12+12+16, bf16 outputs
])
def test_benchmarked_batch_sizes_against_reference_goldens(case, batch, out_dtype):
    """The configurations bench.py measures, checked against the unmodified reference: the batch is built from the three
    golden images in a shuffled order (max_chunk = 16 as in the bench), EVERY position is compared with the reference's
    samples of its image, and positions that hold the same image must be bit-identical (different rows of the GEMM tiles,
    different persistent-grid items of the attention kernels, different passes)."""
    g, cfg, sd, x3, enc = _setup(case)
    assert x3.shape[0] == 3
    gen = torch.Generator().manual_seed(batch)
    perm = [0, 1, 2] + torch.randint(0, 3, (batch - 3,), generator=gen).tolist()
    perm = [perm[i] for i in torch.randperm(batch, generator=gen).tolist()]
    x = x3.to(DEV)[torch.tensor(perm, device=DEV)]
    enc.precision, enc.max_chunk, enc.out_dtype = "bf16", 16, out_dtype
    with torch.no_grad():
        out = enc(x)
    torch.cuda.synchronize()
    worst = _positions_check(out, g, perm, TOL_BF16)
    first = {b: perm.index(b) for b in (0, 1, 2)}
    for k in KEYS:
        assert out[k].dtype == out_dtype
        for p, b in enumerate(perm):
            assert torch.equal(out[k][p], out[k][first[b]]), f"{k}: positions {p} and {first[b]} hold image {b} but differ"
    print(f"{case} batch {batch}: worst rel-L2 over positions / outputs = {worst:.3e}")


@pytest.mark.parametrize("preset,precision", [("vit_b", "fp32"), ("vit_b", "bf16"), ("vit_l", "bf16"), ("vit_h", "fp32"), ("vit_h", "bf16")])
def test_full_tensor_against_oracle_on_host(preset, precision):
    """ALL elements of the four embeddings and of the last block's token stream (not samples) against the oracle run on this
    box's host cores for one full-size image (the oracle itself is pinned to the reference by tests/test_oracle.py).  Besides the
    whole-tensor rel-L2: the rel-L2 of every (14 x 14 window, head-width channel slice) cell of the token stream and of every
    window-aligned spatial tile of the embeddings, and the largest absolute error in units of the tensor's RMS — a wrong
    edge window, last tile or head slice moves a whole-tensor norm by little and these by a lot."""
    from oracle import sam_vit_oracle as orc
    cfg = ib.PRESETS[preset]
    sd = ib.make_state_dict(cfg, 4242, rel_std=0.05)
    x = ib.make_images(1, cfg, 17)
    taps = {}
    ref = orc.encoder_forward_cfg(sd, x, cfg, tap=lambda n, t: taps.__setitem__(n, t.clone()))
    enc = build_encoder(cfg)
    enc.load_state_dict(sd)
    enc.to(DEV)
    enc.precision = precision
    enc.enable_taps(True)
    with torch.no_grad():
        out = enc(x.to(DEV))
    torch.cuda.synchronize()
    tol = TOL_FP32 if precision == "fp32" else TOL_BF16
    cell_tol, peak_tol = (3 * tol, 0.15) if precision == "bf16" else (3 * tol, 2e-3)

    def cells(a, b, ty, tx, tc):
        """rel-L2 per (ty x tx spatial tile, tc channels) cell of two (H, W, C) tensors (ragged last tiles included)"""
        H, W, Cc = a.shape
        worst = 0.0
        for y0 in range(0, H, ty):
            for x0 in range(0, W, tx):
                da = (a[y0:y0 + ty, x0:x0 + tx] - b[y0:y0 + ty, x0:x0 + tx]).double()
                rb = b[y0:y0 + ty, x0:x0 + tx].double()
                n = -(-Cc // tc)
                pad = n * tc - Cc
                if pad:
                    da = torch.nn.functional.pad(da, (0, pad)); rb = torch.nn.functional.pad(rb, (0, pad))
                e = da.reshape(-1, n, tc).pow(2).sum((0, 2)).sqrt() / rb.reshape(-1, n, tc).pow(2).sum((0, 2)).sqrt()
                worst = max(worst, float(e.max()))
        return worst

    last = enc.read_tap(cfg.depth - 1)[0].cpu()
    want = taps[f"block{cfg.depth - 1}"][0]
    assert ib.rel_l2(last, want) < tol
    w = cells(last, want, cfg.window_size, cfg.window_size, cfg.head_dim)
    assert w < cell_tol, f"token stream: worst (window, head slice) cell rel-L2 {w:.3e}"
    for k, stride in zip(KEYS, (4, 8, 16, 32)):
        a, b = out[k][0].float().cpu().permute(1, 2, 0), ref[k][0].permute(1, 2, 0)
        assert a.shape == b.shape
        err = ib.rel_l2(a, b)
        assert err < tol, (k, err)
        t = max(1, cfg.window_size * 16 // stride)
        w = cells(a, b, t, t, 64)
        assert w < cell_tol, f"{k}: worst window-aligned tile rel-L2 {w:.3e}"
        peak = float((a - b).abs().max() / b.double().pow(2).mean().sqrt())
        assert peak < peak_tol, f"{k}: largest absolute error = {peak:.3e} x RMS"


def test_batch_chunking_and_output_dtype_and_host_path():
    """Images are independent (no cross-sample coupling): any chunking gives the same embeddings; bf16 outputs are the
    rounded fp32 ones; the host-buffer entry point returns what the device one does."""
    g, cfg, sd, x, enc = _setup("tiny64_std")
    x3 = torch.cat([x, x[:1]], 0)               # 3 images, the third equals the first
    enc.precision = "bf16"
    outs = {}
    for chunk in (1, 2, 8):
        enc.max_chunk = chunk
        with torch.no_grad():
            outs[chunk] = enc(x3.to(DEV))
    for k in KEYS:
        assert ib.rel_l2(outs[1][k], outs[8][k]) < 1e-6 and ib.rel_l2(outs[2][k], outs[8][k]) < 1e-6
        assert ib.rel_l2(outs[8][k][2], outs[8][k][0]) < 1e-6
    enc.out_dtype = torch.bfloat16
    with torch.no_grad():
        ob = enc(x3.to(DEV))
    for k in KEYS:
        assert ob[k].dtype == torch.bfloat16
        assert ib.rel_l2(ob[k].float(), outs[8][k]) < 4e-3
    enc.out_dtype = torch.float32
    enc.max_chunk = 2
    with torch.no_grad():
        oh = enc.forward_host(x3.pin_memory())
    for k in KEYS:
        assert not oh[k].is_cuda
        assert ib.rel_l2(oh[k], outs[8][k]) < 1e-6


PIXEL_MEAN, PIXEL_STD = [123.675, 116.280, 103.530], [58.395, 57.120, 57.375]      # configs/step1.yaml:320-321


def _u8_images(sizes, seed=11):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8) for h, w in sizes]


def test_uint8_staging_is_exact():
    """(x - mean) / std + zero padding to the canvas + patch im2col, fp32: bit-exact against the oracle's restatement."""
    import ctypes as C
    from iuvl_b200 import cabi
    from oracle import sam_vit_oracle as orc
    imgs = _u8_images([(1024, 1024), (700, 1000), (37, 53), (1, 1), (1024, 5), (1021, 1023)])
    canvas = orc.stage_images(imgs, PIXEL_MEAN, PIXEL_STD, 1024)
    B = len(imgs)
    ref = canvas.reshape(B, 3, 64, 16, 64, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * 4096, 768)
    dev = [t.to(DEV) for t in imgs]
    ptrs = (C.c_void_p * B)(*[t.data_ptr() for t in dev])
    hs = (C.c_int * B)(*[t.shape[1] for t in dev])
    ws = (C.c_int * B)(*[t.shape[2] for t in dev])
    mean, std = (C.c_float * 3)(*PIXEL_MEAN), (C.c_float * 3)(*PIXEL_STD)
    out = torch.full((B * 4096, 768), float("nan"), device=DEV)
    cabi.check(cabi.lib().svb_stage_images_u8(ptrs, hs, ws, B, 3, 1024, 16, mean, std, out.data_ptr(), cabi.DTYPE_F32, cabi.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)
    outb = torch.empty(B * 4096, 768, device=DEV, dtype=torch.bfloat16)
    cabi.check(cabi.lib().svb_stage_images_u8(ptrs, hs, ws, B, 3, 1024, 16, mean, std, outb.data_ptr(), cabi.DTYPE_BF16, cabi.stream_ptr()))
    assert torch.equal(outb.cpu(), ref.bfloat16())


def test_uint8_forward_matches_the_oracle_pipeline():
    """forward_uint8 == oracle(stage_images(...)) (xdecoder_model.py:481-484 + image_encoder.py:107-120), ragged image sizes."""
    from oracle import sam_vit_oracle as orc
    cfg = ib.PRESETS["tiny80"]
    sd = ib.make_state_dict(cfg, 77, rel_std=0.1)
    imgs = _u8_images([(1024, 1024), (600, 911), (333, 1024)], seed=3)
    ref = orc.encoder_forward_cfg(sd, orc.stage_images(imgs, PIXEL_MEAN, PIXEL_STD, 1024), cfg)
    enc = build_encoder(cfg)
    enc.load_state_dict(sd)
    enc.to(DEV)
    for precision, tol in (("fp32", TOL_FP32), ("bf16", TOL_BF16)):
        enc.precision = precision
        with torch.no_grad():
            out = enc.forward_uint8([t.to(DEV) for t in imgs], PIXEL_MEAN, PIXEL_STD)
            same = enc(orc.stage_images(imgs, PIXEL_MEAN, PIXEL_STD, 1024).to(DEV))
        for k in KEYS:
            assert ib.rel_l2(out[k], ref[k]) < tol, (precision, k, ib.rel_l2(out[k], ref[k]))
            assert torch.equal(out[k], same[k])            # same arithmetic as the fp32-canvas entry point
    with torch.no_grad(), pytest.raises(NotImplementedError):
        enc.forward_uint8([torch.zeros(3, 1025, 10, dtype=torch.uint8, device=DEV)], PIXEL_MEAN, PIXEL_STD)


def test_half_precision_inputs_are_read_directly():
    """The reference's pipeline hands the encoder fp16 images (cast_batch_to_half, pipeline/XDecoderPipeline.py:93-95): fp16 / bf16
    inputs go through the same patch-embedding loader (svb_encoder_forward_x) and give bit-identical embeddings to the fp32 tensor
    holding the same values."""
    g, cfg, sd, x, enc = _setup("tiny64_std")
    for dt in (torch.float16, torch.bfloat16):
        xh = x.to(DEV).to(dt)
        for precision in ("bf16", "fp32"):
            enc.precision = precision
            with torch.no_grad():
                a = enc(xh)
                b = enc(xh.float())
            for k in KEYS:
                assert torch.equal(a[k], b[k]), (dt, precision, k)
    with torch.no_grad(), pytest.raises(TypeError):
        enc(x.to(DEV).double())


def _offset_case():
    """tiny80 weights whose residual stream carries a common offset of +40 on every channel of every token from the position
    embedding to the last block (removed again by the last lin2 bias): |row mean| / row std ~ 40."""
    cfg = ib.PRESETS["tiny80"]
    sd = ib.make_state_dict(cfg, 31, rel_std=0.05)
    sd["pos_embed"] = sd["pos_embed"] + 40.0
    last = f"blocks.{cfg.depth - 1}.mlp.lin2.bias"
    sd[last] = sd[last] - 40.0
    return cfg, sd, ib.make_images(1, cfg, 9)


def test_layernorm_fold_survives_a_large_common_offset():
    """The folded LayerNorm reads bf16(x) instead of bf16(LayerNorm(x)): without care its rounding error grows with |row mean| / row
    std and the single-pass variance cancels.  The producers therefore hand over the CENTRED row (x minus its mean before the update,
    Epilogue::shift_out) — LayerNorm is invariant under that shift.  With an offset of 40 standard deviations the bf16 bar must hold;
    the same run with the centring switched off (SVB_LN_CENTRE=0, in a subprocess) must miss it by a wide margin, which shows that
    this case has the power to detect the problem."""
    import json
    import os
    import subprocess
    import sys
    from oracle import sam_vit_oracle as orc
    cfg, sd, x = _offset_case()
    ref = orc.encoder_forward_cfg(sd, x, cfg)
    enc = build_encoder(cfg)
    enc.load_state_dict(sd)
    enc.to(DEV)
    enc.precision = "bf16"
    with torch.no_grad():
        out = enc(x.to(DEV))
    errs = {k: ib.rel_l2(out[k], ref[k]) for k in KEYS}
    assert max(errs.values()) < TOL_BF16, errs
    code = ("import json, torch, iuvl_b200 as ib\n"
            "from iuvl_b200.encoder import build_encoder\n"
            "from tests.test_gpu_encoder import _offset_case, KEYS\n"
            "from oracle import sam_vit_oracle as orc\n"
            "cfg, sd, x = _offset_case()\n"
            "ref = orc.encoder_forward_cfg(sd, x, cfg)\n"
            "enc = build_encoder(cfg); enc.load_state_dict(sd); enc.to('cuda'); enc.precision = 'bf16'\n"
            "with torch.no_grad():\n"
            "    out = enc(x.to('cuda'))\n"
            "print('ERRS', json.dumps({k: ib.rel_l2(out[k], ref[k]) for k in KEYS}))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, SVB_LN_CENTRE="0", PYTHONPATH=root),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    off = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("ERRS")][0][5:])
    assert max(off.values()) > 3 * max(errs.values()), (off, errs)


def test_weights_resync_after_load_state_dict():
    g, cfg, sd, x, enc = _setup("tiny64_std")
    enc.precision = "fp32"
    with torch.no_grad():
        a = enc(x[:1].to(DEV))
        sd2 = ib.make_state_dict(cfg, 4321, rel_std=0.02)
        enc.load_state_dict(sd2)
        b = enc(x[:1].to(DEV))
        enc.load_state_dict(sd)
        c = enc(x[:1].to(DEV))
    assert ib.rel_l2(b["res4"], a["res4"]) > 1e-2
    assert ib.rel_l2(c["res4"], a["res4"]) < 1e-6


def test_errors_are_loud():
    g, cfg, sd, x, enc = _setup("tiny64_std")
    with pytest.raises(RuntimeError, match="forward pass only"):
        enc(x[:1].to(DEV))                       # grad enabled + requires_grad params
    with torch.no_grad(), pytest.raises(NotImplementedError):
        enc(torch.zeros(1, 3, 1024, 1000, device=DEV))        # sides must be multiples of 32 patches


def test_pass_schedules_cover_the_batch():
    """The pass schedules of the device and the host entry points: every image once, no pass above max_chunk; the host schedule keeps
    its first and last pass (the exposed upload / download) no larger than the device schedule's largest pass."""
    cfg = ib.PRESETS["vit_h"]
    enc = build_encoder(cfg).to(DEV)                      # (no weights needed: the schedule depends on the geometry only)
    enc.out_dtype = torch.bfloat16
    for chunk in (8, 16):
        enc.max_chunk = chunk
        for batch in (1, 3, 8, 13, 16, 33, 64):
            dev_s, host_s = enc.pass_schedule(batch), enc.pass_schedule(batch, host_path=True)
            for sch in (dev_s, host_s):
                assert sum(sch) == batch and all(1 <= c <= chunk for c in sch), (batch, chunk, sch)
            assert host_s[0] <= max(dev_s) and host_s[-1] <= max(dev_s)
    enc.max_chunk = 16
    assert enc.pass_schedule(64) == [16, 12, 12, 12, 12]
    print("host schedule of 64 ViT-H images:", enc.pass_schedule(64, host_path=True))
