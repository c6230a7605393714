"""GPU parity of the multi-scale deformable attention forward (scope row N1), through the C ABI, against
 (a) the committed outputs of the reference's own ms_deform_attn_core_pytorch (tests/golden/msda_*.npz) and
 (b) the CPU oracle on seeded inputs at larger sizes, incl. locations outside [0,1] (zero padding) and exact edge hits."""
import os

import numpy as np
import pytest
import torch

import iuvl_b200 as ib
from iuvl_b200.msda import ms_deform_attn_forward
from tests.util import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _starts(shapes):
    s = [0]
    for h, w in shapes[:-1]:
        s.append(s[-1] + h * w)
    return s


@pytest.mark.parametrize("case", ["toy", "small", "heads8"])
def test_msda_against_reference_goldens(case):
    z = np.load(os.path.join(GOLDEN, f"msda_{case}.npz"))
    shapes = [tuple(int(v) for v in hw) for hw in z["shapes"]]
    value, loc, aw = (torch.from_numpy(z[k]) for k in ("value", "loc", "aw"))
    ref = torch.from_numpy(z["out"])
    N, S, M, D = value.shape
    if D % 4:                                    # the toy case of ops/test.py has 2 channels per head: pad to the 16-byte vector
        value = torch.nn.functional.pad(value, (0, 4 - D % 4))
    with torch.no_grad():
        out = ms_deform_attn_forward(value.to(DEV), torch.tensor(shapes), torch.tensor(_starts(shapes)), loc.to(DEV), aw.to(DEV))
    Dp = value.shape[-1]
    out = out.reshape(N, -1, M, Dp)[..., :D].reshape(N, -1, M * D)
    assert torch.allclose(out.cpu().double(), ref, rtol=1e-5, atol=1e-6), float((out.cpu().double() - ref).abs().max())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 4e-3)])
def test_msda_against_oracle_step1_geometry(dtype, tol):
    """8 heads x 64 channels, 3 levels, 4 points (transformer_encoder_deform.py / configs/step1.yaml) on 32/16/8 maps."""
    from oracle import msda_oracle as mo
    g = torch.Generator().manual_seed(5)
    shapes = [(32, 32), (16, 16), (8, 8)]
    N, M, D, P, L = 2, 8, 64, 4, 3
    S = sum(h * w for h, w in shapes)
    Lq = S
    value = torch.randn(N, S, M, D, generator=g)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.3 - 0.15
    loc[0, :5] = 0.0                             # exact corners / edges
    loc[0, 5:10] = 1.0
    loc[0, 10:15, :, :, :, 0] = 0.5 / 32         # centre of the first pixel column: integer sample position
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).reshape(N, Lq, M, L, P)
    v = value.to(dtype)
    ref = mo.ms_deform_attn_core(v.double(), shapes, loc.double(), aw.double())
    with torch.no_grad():
        out = ms_deform_attn_forward(v.to(DEV), torch.tensor(shapes), torch.tensor(_starts(shapes)), loc.to(DEV), aw.to(DEV))
    assert out.dtype == dtype and tuple(out.shape) == (N, Lq, M * D)
    assert ib.rel_l2(out, ref) < tol, ib.rel_l2(out, ref)


def test_msda_properties_and_errors():
    """Linearity in value and in the weights; zero weights -> zero; loud errors."""
    g = torch.Generator().manual_seed(9)
    shapes = [(20, 12), (7, 9)]
    N, M, D, P, L, Lq = 1, 2, 16, 3, 2, 101
    S = sum(h * w for h, w in shapes)
    va, vb = torch.randn(N, S, M, D, generator=g).to(DEV), torch.randn(N, S, M, D, generator=g).to(DEV)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g).to(DEV)
    aw = torch.rand(N, Lq, M, L, P, generator=g).to(DEV)
    sh, st = torch.tensor(shapes), torch.tensor(_starts(shapes))
    with torch.no_grad():
        fa, fb = ms_deform_attn_forward(va, sh, st, loc, aw), ms_deform_attn_forward(vb, sh, st, loc, aw)
        fab = ms_deform_attn_forward(va + 2 * vb, sh, st, loc, aw)
        assert torch.allclose(fab, fa + 2 * fb, rtol=1e-4, atol=1e-5)
        assert torch.allclose(ms_deform_attn_forward(va, sh, st, loc, 3 * aw), 3 * fa, rtol=1e-5, atol=1e-6)
        assert ms_deform_attn_forward(va, sh, st, loc, torch.zeros_like(aw)).abs().max() == 0
        assert ms_deform_attn_forward(va, sh, st, loc + 5.0, aw).abs().max() == 0          # every sample outside: zero padding
        with pytest.raises(RuntimeError):
            ms_deform_attn_forward(va.cpu(), sh, st, loc.cpu(), aw.cpu())
        with pytest.raises(Exception, match="levels cover"):
            ms_deform_attn_forward(va, torch.tensor([(20, 12), (7, 8)]), st, loc, aw)
    with pytest.raises(RuntimeError, match="forward pass only"):
        ms_deform_attn_forward(va.requires_grad_(), sh, st, loc, aw)


@pytest.mark.parametrize("case", ["points", "boxes_masked", "heads64"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_msda_module_against_reference_goldens(case, precision, tol):
    """The drop-in MSDeformAttn module (four tcgen05 / fp32 GEMMs + the fused softmax / location / gather kernel) against outputs
    of the UNMODIFIED reference module (ops/modules/ms_deform_attn.py:82-125), state_dict loaded by the reference's keys."""
    from iuvl_b200.msda import MSDeformAttn
    z = np.load(os.path.join(GOLDEN, f"msda_module_{case}.npz"))
    C, M, P, L = (int(v) for v in z["meta"])
    mod = MSDeformAttn(d_model=C, n_levels=L, n_heads=M, n_points=P)
    mod.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}, strict=True)
    mod.to(DEV)
    mod.precision = precision
    shapes = [tuple(int(v) for v in hw) for hw in z["shapes"]]
    mask = torch.from_numpy(z["mask"]).to(DEV) if z["mask"].size else None
    with torch.no_grad():
        out = mod(torch.from_numpy(z["query"]).to(DEV), torch.from_numpy(z["ref"]).to(DEV), torch.from_numpy(z["inp"]).to(DEV),
                  torch.tensor(shapes), torch.tensor(_starts(shapes)), mask)
    ref = torch.from_numpy(z["out"])
    assert tuple(out.shape) == tuple(ref.shape)
    err = ib.rel_l2(out, ref)
    assert err < tol, (case, precision, err)


def test_msda_module_fused_kernel_equals_unfused_path():
    """The fused kernel (softmax + locations inline) against the plain sampling core fed with torch-computed locations / weights."""
    from iuvl_b200.msda import MSDeformAttn
    g = torch.Generator().manual_seed(21)
    C, M, P, shapes = 128, 4, 4, [(10, 12), (5, 6)]
    L, S, N, Lq = len(shapes), sum(h * w for h, w in shapes), 2, 77
    mod = MSDeformAttn(C, L, M, P).to(DEV)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.3)
        mod.attention_weights.weight.normal_(0, 0.3)
        mod.precision = "fp32"
        q, inp = torch.randn(N, Lq, C, generator=g).to(DEV), torch.randn(N, S, C, generator=g).to(DEV)
        ref = torch.rand(N, Lq, L, 2, generator=g).to(DEV)
        out = mod(q, ref, inp, torch.tensor(shapes), torch.tensor(_starts(shapes)))
        value = (inp @ mod.value_proj.weight.t() + mod.value_proj.bias).view(N, S, M, C // M)
        off = mod.sampling_offsets(q).view(N, Lq, M, L, P, 2)
        aw = torch.softmax(mod.attention_weights(q).view(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device=DEV)
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        core = ms_deform_attn_forward(value.contiguous(), torch.tensor(shapes), torch.tensor(_starts(shapes)), loc, aw)
        ref_out = core @ mod.output_proj.weight.t() + mod.output_proj.bias
    assert ib.rel_l2(out, ref_out) < 1e-4, ib.rel_l2(out, ref_out)


@pytest.mark.parametrize("case", ["small", "heads64"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_deform_encoder_against_reference_goldens(case, precision, tol):
    """The drop-in MSDeformAttnTransformerEncoderOnly (per layer: fused pos-add + cast, MSDeformAttn, residual + LayerNorm, ReLU and
    residual in the FFN GEMMs' epilogues) against outputs of the UNMODIFIED reference classes (transformer_encoder_deform.py:23-161)."""
    from iuvl_b200.msda import MSDeformAttnTransformerEncoderOnly
    z = np.load(os.path.join(GOLDEN, f"deform_encoder_{case}.npz"))
    C, M, NL, F_, P, L = (int(v) for v in z["meta"])
    mod = MSDeformAttnTransformerEncoderOnly(C, M, NL, F_, 0.1, "relu", L, P)
    mod.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}, strict=True)
    mod.to(DEV).eval()
    mod.precision = precision
    with torch.no_grad():
        memory, shapes, starts = mod([torch.from_numpy(z[f"src{i}"]).to(DEV) for i in range(L)],
                                     [torch.from_numpy(z[f"pos{i}"]).to(DEV) for i in range(L)])
    assert shapes.cpu().tolist() == z["shapes"].tolist() and starts.cpu().tolist() == z["starts"].tolist()
    ref = torch.from_numpy(z["memory"])
    assert tuple(memory.shape) == tuple(ref.shape)
    err = ib.rel_l2(memory, ref)
    assert err < tol, (case, precision, err)


def test_deform_encoder_layer_matches_composition_of_ops():
    """One encoder layer in fp32 against the same arithmetic spelled out with torch ops around the drop-in MSDeformAttn."""
    from iuvl_b200.msda import MSDeformAttnTransformerEncoderLayer
    g = torch.Generator().manual_seed(5)
    C, M, P, shapes = 128, 4, 4, [(9, 7), (4, 4)]
    L, S, N = len(shapes), sum(h * w for h, w in shapes), 2
    layer = MSDeformAttnTransformerEncoderLayer(C, 256, 0.1, "relu", L, M, P).to(DEV).eval()
    layer.precision = "fp32"
    with torch.no_grad():
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.3)
        layer.self_attn.attention_weights.weight.normal_(0, 0.3)
        src, pos = torch.randn(N, S, C, generator=g).to(DEV), torch.randn(N, S, C, generator=g).to(DEV)
        ref = torch.rand(N, S, L, 2, generator=g).to(DEV)
        ss, st = torch.tensor(shapes), torch.tensor(_starts(shapes))
        for p_ in (pos, None):
            out = layer(src, p_, ref, ss, st)
            a = layer.self_attn(src if p_ is None else src + p_, ref, src, ss, st)
            y = layer.norm1(src + a)
            want = layer.norm2(y + layer.linear2(torch.relu(layer.linear1(y))))
            assert ib.rel_l2(out, want) < 1e-4, ib.rel_l2(out, want)


def test_msda_one_lane_kernel_stays_correct():
    """The fused kernel's predecessor (every lane of a unit computes the softmax / locations / bilinear weights for itself: the path of
    head widths that are not a power-of-two number of 16-byte groups, and SVB_MSDA_SHARED=0 for A/B runs) passes the same module tests
    as the shared-arithmetic kernel; the variable is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, SVB_MSDA_SHARED="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_msda.py", "-q", "-x", "-k", "(module or against_reference or against_oracle) and not one_lane"],
                       cwd=root, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-500:])


@pytest.mark.parametrize("rows,dim,pos_rows", [(1000, 512, 250), (777, 256, 777), (64, 1280, 0)])
def test_layernorm_post_outputs(rows, dim, pos_rows):
    """svb_layernorm_post: LayerNorm(x [+ add]) once, three outputs (fp32, bf16, bf16 of the sum with a position embedding shared by
    the batch) — the post-norm pairs of the deformable encoder layer (transformer_encoder_deform.py:126-127, 119, 112-114)."""
    from iuvl_b200 import cabi
    g = torch.Generator().manual_seed(rows + dim)
    x = (torch.randn(rows, dim, generator=g) * 1.7 + 0.3).to(DEV)
    add = torch.randn(rows, dim, generator=g).to(DEV)
    w, b = (1 + 0.2 * torch.randn(dim, generator=g)).to(DEV), (0.2 * torch.randn(dim, generator=g)).to(DEV)
    pos = torch.randn(pos_rows, dim, generator=g).to(DEV) if pos_rows else None
    x0, add0 = x.clone(), add.clone()
    for use_add in (False, True):
        out = torch.full((rows, dim), float("nan"), device=DEV)
        ob = torch.full((rows, dim), float("nan"), dtype=torch.bfloat16, device=DEV)
        oq = torch.full((rows, dim), float("nan"), dtype=torch.bfloat16, device=DEV) if pos_rows else None
        cabi.check(cabi.lib().svb_layernorm_post(x.data_ptr(), add.data_ptr() if use_add else None, w.data_ptr(), b.data_ptr(), out.data_ptr(),
                                                 ob.data_ptr(), cabi.ptr(pos), pos_rows, cabi.ptr(oq), rows, dim, 1e-5, cabi.stream_ptr()), "ln_post")
        torch.cuda.synchronize()
        ref = torch.nn.functional.layer_norm((x0 + add0 if use_add else x0).double(), (dim,), w.double(), b.double(), 1e-5)
        assert ib.rel_l2(out, ref) < 2e-6
        assert torch.equal(ob, out.bfloat16())
        if pos_rows:
            idx = torch.arange(rows, device=DEV) % pos_rows
            assert torch.equal(oq, (out + pos[idx]).bfloat16())
        assert torch.equal(x, x0) and torch.equal(add, add0)          # the inputs are only read
    only_b = torch.empty(rows, dim, dtype=torch.bfloat16, device=DEV)           # outputs are optional one by one
    cabi.check(cabi.lib().svb_layernorm_post(x.data_ptr(), add.data_ptr(), w.data_ptr(), b.data_ptr(), None, only_b.data_ptr(), None, 0, None, rows,
                                             dim, 1e-5, cabi.stream_ptr()), "ln_post")
    assert torch.equal(only_b, ob)
    assert cabi.lib().svb_layernorm_post(x.data_ptr(), None, w.data_ptr(), b.data_ptr(), None, None, None, 0, None, rows, dim, 1e-5,
                                         cabi.stream_ptr()) != 0          # no output requested
    assert cabi.lib().svb_layernorm_post(x.data_ptr(), None, w.data_ptr(), b.data_ptr(), out.data_ptr(), None, None, 0, None, rows, 200, 1e-5,
                                         cabi.stream_ptr()) != 0          # a width that is not built
