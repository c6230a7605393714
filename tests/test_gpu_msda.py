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
