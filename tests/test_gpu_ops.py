"""GPU parity of the single operators, called through the C ABI, against fp64 references / the oracle's pieces."""
import math
import os

import pytest
import torch

import iuvl_b200 as ib
from iuvl_b200 import cabi
from tests.util import ref_attention_core

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _linear(mode, A, W, bias=None, gelu=False, resid=None, resid_mod=0, out_dtype=torch.float32, stats=None, rps=0,
            out=None, remap=(0, 0)):
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    rc = cabi.lib().svb_linear(
        mode, A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), M, N, K, cabi.ptr(bias), int(gelu), cabi.ptr(resid),
        resid.stride(0) if resid is not None else 0, resid_mod, out.data_ptr(),
        cabi.DTYPE_BF16 if out.dtype == torch.bfloat16 else cabi.DTYPE_F32, out.stride(0), cabi.ptr(stats), rps,
        remap[0], remap[1], cabi.stream_ptr())
    cabi.check(rc, "svb_linear")
    return out


def _ref_linear(A, W, bias=None, gelu=False, resid=None, resid_mod=0):
    y = A.double() @ W.double().t()
    if bias is not None:
        y = y + bias.double()
    pre = y.clone()
    if gelu:
        y = 0.5 * y * (1 + torch.erf(y / math.sqrt(2)))
    if resid is not None:
        r = resid.double()
        if resid_mod:
            r = r[torch.arange(A.shape[0], device=A.device) % resid_mod]
        y = y + r
    return y, pre


SHAPES = [(128, 256, 64), (256, 512, 128), (4096, 768, 768), (4096, 2304, 768), (2048, 1280, 5120), (4096, 128, 192),
          (1024, 1024, 2048), (200, 136, 72), (4096, 3840, 1280), (384, 96, 320), (128, 2560, 640)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tcgen05_plain(M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    out = _linear(cabi.MODE_BF16, A, W)
    torch.cuda.synchronize()
    ref, _ = _ref_linear(A, W)
    err = ib.rel_l2(out, ref)
    assert err < 2e-5, err          # bf16 operands are exact inputs; only fp32 accumulation order/rounding differs
    assert (out.double() - ref).abs().max() < 1e-4


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_tcgen05_epilogues(out_dtype):
    M, N, K = 8192, 1280, 768
    g = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    tol = 1e-5 if out_dtype == torch.float32 else 4e-3
    # bias + GELU
    out = _linear(cabi.MODE_BF16, A, W, bias=bias, gelu=True, out_dtype=out_dtype)
    ref, _ = _ref_linear(A, W, bias, gelu=True)
    # the bf16 path evaluates erf by Abramowitz-Stegun 7.1.25 (|erf error| <= 2.5e-5): visible only in an fp32 store
    assert ib.rel_l2(out, ref) < (5e-5 if out_dtype == torch.float32 else tol)
    assert (out.double() - ref).abs().max() < (1e-4 if out_dtype == torch.float32 else 0.05)
    # bias + broadcast residual (pos_embed style) and in-place residual, fp32 stream
    if out_dtype == torch.float32:
        pos = torch.randn(4096, N, generator=g).to(DEV)
        out = _linear(cabi.MODE_BF16, A, W, bias=bias, resid=pos, resid_mod=4096)
        ref, _ = _ref_linear(A, W, bias, resid=pos, resid_mod=4096)
        assert ib.rel_l2(out, ref) < tol
        X = torch.randn(M, N, generator=g).to(DEV)
        X0 = X.clone()
        _linear(cabi.MODE_BF16, A, W, bias=bias, resid=X, out=X)
        ref, _ = _ref_linear(A, W, bias, resid=X0)
        assert ib.rel_l2(X, ref) < tol
        # GroupNorm statistics of (acc + bias), 2 samples of 4096 rows
        stats = torch.zeros(2, 2, dtype=torch.float64, device=DEV)
        out = _linear(cabi.MODE_BF16, A, W, bias=bias, stats=stats, rps=4096)
        ref, pre = _ref_linear(A, W, bias)
        s_ref = torch.stack([pre.reshape(2, -1).sum(1), (pre ** 2).reshape(2, -1).sum(1)], 1)
        assert torch.allclose(stats, s_ref, rtol=1e-5, atol=1e-2), (stats, s_ref)


def _linear_fused(A, W, bias, out, gelu=False, resid=None, resid_mod=0, ln=None, out2=None, stat_out=None, remap=(0, 0)):
    M, K = A.shape
    N = W.shape[0]
    ln_stats, ln_c, ln_dim, ln_eps = ln if ln is not None else (None, None, 0, 0.0)
    rc = cabi.lib().svb_linear_fused(
        A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), M, N, K, cabi.ptr(bias), int(gelu), cabi.ptr(resid),
        resid.stride(0) if resid is not None else 0, resid_mod, out.data_ptr(),
        cabi.DTYPE_BF16 if out.dtype == torch.bfloat16 else cabi.DTYPE_F32, out.stride(0), cabi.ptr(ln_stats), cabi.ptr(ln_c), ln_dim,
        ln_eps, cabi.ptr(out2), out2.stride(0) if out2 is not None else 0, cabi.ptr(stat_out), remap[0], remap[1], cabi.stream_ptr())
    cabi.check(rc, "svb_linear_fused")
    return out


@pytest.mark.parametrize("M,D,Kp", [(4096, 1280, 256), (8192, 768, 3072), (1000, 160, 64), (512, 128, 128), (4096, 1024, 1024)])
def test_gemm_layernorm_fold(M, D, Kp):
    """norm -> Linear folded into the GEMMs of the bf16 path: the producer GEMM (residual epilogue) emits bf16(x) and partial
    row sums of x; the consumer GEMM computes act(LayerNorm(x) W^T + b) from the un-normalised bf16 rows.
    Reference: fp64 LayerNorm + Linear of the fp32 x (image_encoder.py:183,195 + common.py:25)."""
    g = torch.Generator(device="cpu").manual_seed(M + D + Kp)
    parts = (D + 127) // 128
    # ---- producer: x = x0 + A Wp^T + bp ----
    A = torch.randn(M, Kp, generator=g).to(DEV).bfloat16()
    Wp = (torch.randn(D, Kp, generator=g) / math.sqrt(Kp)).to(DEV).bfloat16()
    bp = torch.randn(D, generator=g).to(DEV)
    x0 = (torch.randn(M, D, generator=g) * 2 + 0.7).to(DEV)
    x = x0.clone()
    xb = torch.full((M, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    stat = torch.full((M, parts, 2), float("nan"), device=DEV)
    _linear_fused(A, Wp, bp, x, resid=x, out2=xb, stat_out=stat)
    torch.cuda.synchronize()
    ref_x, _ = _ref_linear(A, Wp, bp, resid=x0)
    assert ib.rel_l2(x, ref_x) < 1e-5
    assert torch.equal(xb, x.bfloat16())                      # the bf16 copy is the rounded fp32 stream
    pad = parts * 128 - D
    xs = torch.nn.functional.pad(x.double(), (0, pad)).reshape(M, parts, 128)
    s_ref = torch.stack([xs.sum(-1), (xs ** 2).sum(-1)], -1)
    # (SVB_GEMM2_DBG & 64, an A/B variant, hands the two 128-column slots of a 256-column tile to interleaved chunks: the consumer
    # only ever sums the slots, so there the check is on the sums per tile)
    if int(os.environ.get("SVB_GEMM2_DBG", "0")) & 64 and D % 256 == 0:
        assert torch.allclose(stat.double().reshape(M, parts // 2, 2, 2).sum(2), s_ref.reshape(M, parts // 2, 2, 2).sum(2), rtol=2e-5, atol=1e-3)
    else:
        assert torch.allclose(stat.double(), s_ref, rtol=2e-5, atol=1e-3), (stat[0], s_ref[0])
    # ---- consumer: y = gelu(LayerNorm(x) W^T + b) ----
    N = 3 * D
    W = (torch.randn(N, D, generator=g) / math.sqrt(D)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    gamma = (1 + 0.2 * torch.randn(D, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(D, generator=g)).to(DEV)
    Wg = torch.empty(N, D, dtype=torch.bfloat16, device=DEV)
    csum, bf = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
    cabi.check(cabi.lib().svb_fold_layernorm(W.data_ptr(), b.data_ptr(), gamma.data_ptr(), beta.data_ptr(), Wg.data_ptr(), csum.data_ptr(),
                                             bf.data_ptr(), N, D, cabi.stream_ptr()), "fold")
    assert torch.equal(Wg, (W * gamma).bfloat16())
    assert torch.allclose(csum.double(), Wg.double().sum(1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(bf.double(), b.double() + W.double() @ beta.double(), rtol=1e-5, atol=1e-5)
    ln_ref = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-6)
    for gelu in (False, True):
        y = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
        _linear_fused(xb, Wg, bf, y, gelu=gelu, ln=(stat, csum, D, 1e-6))
        ref = ln_ref @ W.double().t() + b.double()
        if gelu:
            ref = torch.nn.functional.gelu(ref)
        err = ib.rel_l2(y, ref)
        # same error class as the unfused path (bf16 rounding of the normalised rows instead of the raw rows)
        xn = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
        cabi.check(cabi.lib().svb_layernorm(x.data_ptr(), None, gamma.data_ptr(), beta.data_ptr(), xn.data_ptr(), cabi.DTYPE_BF16, M, D,
                                            1e-6, cabi.stream_ptr()), "ln")
        y2 = _linear(cabi.MODE_BF16, xn, W.bfloat16(), bias=b, gelu=gelu, out_dtype=torch.bfloat16)
        err2 = ib.rel_l2(y2, ref)
        assert err < 6e-3 and err < 1.5 * err2 + 1e-3, (err, err2)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (4096, 768, 768), (200, 136, 72), (1024, 128, 320)])
def test_gemm_fp32_simt(M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    out = _linear(cabi.MODE_FP32, A, W, bias=bias, gelu=True, resid=res)
    ref, _ = _ref_linear(A, W, bias, gelu=True, resid=res)
    assert ib.rel_l2(out, ref) < 2e-6
    if M % 128 == 0:
        stats = torch.zeros(M // 128, 2, dtype=torch.float64, device=DEV)
        _linear(cabi.MODE_FP32, A, W, bias=bias, stats=stats, rps=128)
        _, pre = _ref_linear(A, W, bias)
        s_ref = torch.stack([pre.reshape(M // 128, -1).sum(1), (pre ** 2).reshape(M // 128, -1).sum(1)], 1)
        assert torch.allclose(stats, s_ref, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("D,out_dtype", [(768, torch.float32), (1280, torch.bfloat16), (160, torch.float32), (1024, torch.bfloat16)])
def test_layernorm(D, out_dtype):
    rows = 4096 + 3
    x = torch.randn(rows, D, device=DEV) * 3 + 0.5
    w, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    out = torch.empty(rows, D, dtype=out_dtype, device=DEV)
    cabi.check(cabi.lib().svb_layernorm(x.data_ptr(), None, w.data_ptr(), b.data_ptr(), out.data_ptr(),
                                        cabi.DTYPE_BF16 if out_dtype == torch.bfloat16 else cabi.DTYPE_F32, rows, D, 1e-6,
                                        cabi.stream_ptr()))
    ref = torch.nn.functional.layer_norm(x.double(), (D,), w.double(), b.double(), 1e-6)
    assert ib.rel_l2(out, ref) < (2e-6 if out_dtype == torch.float32 else 3e-3)
    # fused shortcut add (image_encoder.py:194): x += add is written back to the fp32 stream, out = LN(x)
    add = (torch.randn(rows, D, device=DEV) * 2).to(out_dtype)
    x0 = x.clone()
    cabi.check(cabi.lib().svb_layernorm(x.data_ptr(), add.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(),
                                        cabi.DTYPE_BF16 if out_dtype == torch.bfloat16 else cabi.DTYPE_F32, rows, D, 1e-6,
                                        cabi.stream_ptr()))
    xs = x0.double() + add.double()
    assert ib.rel_l2(x, xs) < 1e-7
    ref = torch.nn.functional.layer_norm(xs, (D,), w.double(), b.double(), 1e-6)
    assert ib.rel_l2(out, ref) < (2e-6 if out_dtype == torch.float32 else 3e-3)


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_im2col_matches_conv(out_dtype):
    B, C, img, p, D = 2, 3, 1024, 16, 64
    x = torch.randn(B, C, img, img, device=DEV)
    cols = torch.empty(B * 4096, C * p * p, dtype=out_dtype, device=DEV)
    cabi.check(cabi.lib().svb_im2col(x.data_ptr(), cols.data_ptr(), cabi.DTYPE_BF16 if out_dtype == torch.bfloat16 else cabi.DTYPE_F32,
                                     B, C, img, p, cabi.stream_ptr()))
    w = torch.randn(D, C, p, p, device=DEV)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), stride=p).permute(0, 2, 3, 1).reshape(B * 4096, D)
    got = cols.double() @ w.double().reshape(D, -1).t()
    assert ib.rel_l2(got, ref) < (1e-6 if out_dtype == torch.float32 else 5e-3)
    if out_dtype == torch.float32:   # pure data movement: bit exact against unfold
        unf = torch.nn.functional.unfold(x, p, stride=p).transpose(1, 2).reshape(B * 4096, -1)
        assert torch.equal(cols, unf)


def _attention(dtype, qkv, rel_h, rel_w, bias, B, g, ws, heads, hd, impl=0):
    out = torch.empty(B * g * g, heads * hd, dtype=qkv.dtype, device=DEV)
    cabi.check(cabi.lib().svb_attention(impl, dtype, qkv.data_ptr(), out.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                        bias.data_ptr(), B, g, ws, heads, hd, cabi.stream_ptr()), "svb_attention")
    return out


@pytest.mark.parametrize("ws,heads,hd,rel_std", [(14, 2, 64, 0.02), (14, 2, 80, 0.5), (64, 2, 64, 0.5), (64, 1, 80, 0.02)])
def test_attention_simt_fp32(ws, heads, hd, rel_std):
    B, g = 2 if ws == 14 else 1, 64
    D = heads * hd
    gen = torch.Generator(device="cpu").manual_seed(ws + hd)
    qkv = torch.randn(B * g * g, 3 * D, generator=gen)
    L = 2 * ws - 1
    rel_h, rel_w = torch.randn(L, hd, generator=gen) * rel_std, torch.randn(L, hd, generator=gen) * rel_std
    bias = torch.randn(3 * D, generator=gen)
    ref = ref_attention_core(qkv, rel_h, rel_w, bias, B, g, ws, heads)
    out = _attention(cabi.DTYPE_F32, qkv.to(DEV), rel_h.to(DEV), rel_w.to(DEV), bias.to(DEV), B, g, ws, heads, hd)
    assert ib.rel_l2(out, ref) < 5e-6


@pytest.mark.parametrize("ws,heads,hd", [(14, 2, 64), (64, 1, 80)])
def test_attention_simt_bf16(ws, heads, hd):
    B, g = 1, 64
    D = heads * hd
    gen = torch.Generator(device="cpu").manual_seed(11)
    qkv = torch.randn(B * g * g, 3 * D, generator=gen).bfloat16()
    L = 2 * ws - 1
    rel_h, rel_w = torch.randn(L, hd, generator=gen) * 0.1, torch.randn(L, hd, generator=gen) * 0.1
    bias = torch.randn(3 * D, generator=gen)
    ref = ref_attention_core(qkv.float(), rel_h, rel_w, bias.bfloat16().float(), B, g, ws, heads)
    out = _attention(cabi.DTYPE_BF16, qkv.to(DEV), rel_h.to(DEV), rel_w.to(DEV), bias.to(DEV), B, g, ws, heads, hd)
    assert ib.rel_l2(out, ref) < 4e-3


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("levels,g,C", [(0, 64, 512), (1, 64, 256), (2, 64, 128), (0, 32, 1024), (1, 64, 72), (0, 16, 200)])
def test_groupnorm_nchw_unshuffle(levels, g, C, out_dtype):
    """GroupNorm(1,C)+GELU on rows ordered (b,y,x,s1,..) -> NCHW; reference builds the NCHW tensor by explicit
    pixel un-shuffle and calls torch group_norm."""
    B = 2
    rows = B * g * g * 4 ** levels
    x = torch.randn(rows, C, device=DEV) * 2 + 0.3
    gamma, beta = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    xs = x.double().reshape(B, -1)
    stats = torch.stack([xs.sum(1), (xs ** 2).sum(1)], 1).contiguous()
    Wout = g << levels
    out = torch.full((B, C, Wout, Wout), float("nan"), device=DEV, dtype=out_dtype)
    cabi.check(cabi.lib().svb_groupnorm_apply_nchw(x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                                   cabi.DTYPE_BF16 if out_dtype == torch.bfloat16 else cabi.DTYPE_F32, B, g, levels, C,
                                                   1e-5, 1, cabi.stream_ptr()))
    t = x.double()
    if levels == 0:
        nchw = t.reshape(B, g, g, C).permute(0, 3, 1, 2)
    elif levels == 1:
        nchw = t.reshape(B, g, g, 2, 2, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, 2 * g, 2 * g)
    else:
        nchw = t.reshape(B, g, g, 2, 2, 2, 2, C).permute(0, 7, 1, 3, 5, 2, 4, 6).reshape(B, C, 4 * g, 4 * g)
    ref = torch.nn.functional.gelu(torch.nn.functional.group_norm(nchw, 1, gamma.double(), beta.double(), 1e-5))
    assert ib.rel_l2(out, ref) < (1e-5 if out_dtype == torch.float32 else 3e-3)


@pytest.mark.parametrize("gelu", [0, 1])
def test_groupnorm_rows(gelu):
    B, rps, C = 2, 4096 * 4, 384
    x = torch.randn(B * rps, C, device=DEV) + 1
    gamma, beta = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    xs = x.double().reshape(B, -1)
    stats = torch.stack([xs.sum(1), (xs ** 2).sum(1)], 1).contiguous()
    out = torch.empty(B * rps, C, dtype=torch.bfloat16, device=DEV)
    cabi.check(cabi.lib().svb_groupnorm_apply(x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                              cabi.DTYPE_BF16, B * rps, C, rps, 1e-5, gelu, cabi.stream_ptr()))
    mu = xs.mean(1).reshape(B, 1, 1)
    var = xs.var(1, unbiased=False).reshape(B, 1, 1)
    ref = (x.double().reshape(B, rps, C) - mu) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert ib.rel_l2(out, ref.reshape(-1, C)) < 3e-3


# ---------------------------------------------------------------------------------------------------------------
# tcgen05 attention (attention_tc.cu): operands in the encoder's own layouts
def _attention_tc(qkv, rel_h, rel_w, bias, B, g, ws, heads, hd):
    """qkv (B*g*g, 3D) bf16 token order on the device -> (B*g*g, D) bf16 through svb_attention_tc."""
    lib = cabi.lib()
    D = heads * hd
    st = cabi.stream_ptr()
    rows = lib.svb_rel_pack_rows(ws, g)
    pack = torch.zeros(rows, hd, dtype=torch.bfloat16, device=DEV)
    cabi.check(lib.svb_pack_rel_table(rel_h.data_ptr(), pack.data_ptr(), rel_h.shape[0], hd, 0, st), "pack h")
    cabi.check(lib.svb_pack_rel_table(rel_w.data_ptr(), pack.data_ptr(), rel_w.shape[0], hd, 1, st), "pack w")
    if ws != g:
        gp = -(-g // ws) * ws
        src = torch.full((B, gp, gp, 3 * D), float("nan"), dtype=torch.bfloat16, device=DEV)
        src[:, :g, :g] = qkv.reshape(B, g, g, 3 * D)
        cabi.check(lib.svb_fill_pad_rows(src.data_ptr(), bias.data_ptr(), B, g, gp, 3 * D, st), "fill_pad_rows")
    else:
        src = qkv
    out = torch.full((B * g * g, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    cabi.check(lib.svb_attention_tc(src.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, st), "svb_attention_tc")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("ws,heads,hd,rel_std,B", [
    (14, 2, 64, 0.02, 2), (14, 2, 80, 0.5, 2), (14, 3, 80, 0.1, 1),
    (64, 2, 64, 0.5, 1), (64, 1, 80, 0.02, 2), (64, 2, 80, 0.5, 1)])
def test_attention_tcgen05(ws, heads, hd, rel_std, B):
    g = 64
    D = heads * hd
    gen = torch.Generator(device="cpu").manual_seed(ws * 100 + hd + heads)
    qkv = torch.randn(B * g * g, 3 * D, generator=gen).bfloat16()
    L = 2 * ws - 1
    rel_h, rel_w = torch.randn(L, hd, generator=gen) * rel_std, torch.randn(L, hd, generator=gen) * rel_std
    bias = torch.randn(3 * D, generator=gen)
    # the kernel sees bf16 copies of the tables and of the pad rows (= bias)
    ref = ref_attention_core(qkv.float(), rel_h.bfloat16().float(), rel_w.bfloat16().float(), bias.bfloat16().float(), B, g, ws, heads)
    out = _attention_tc(qkv.to(DEV), rel_h.to(DEV), rel_w.to(DEV), bias.to(DEV), B, g, ws, heads, hd)
    assert torch.isfinite(out.float()).all()
    err = ib.rel_l2(out, ref)
    assert err < 6e-3, err          # P and the output are rounded to bf16 (2^-9 relative)
    # per-head error, so that a broken head / tail slice cannot hide in the average
    o3, r3 = out.float().cpu().reshape(-1, heads, hd), ref.float().reshape(-1, heads, hd)
    for hh in range(heads):
        assert ib.rel_l2(o3[:, hh, :64], r3[:, hh, :64]) < 8e-3
        if hd > 64:
            assert ib.rel_l2(o3[:, hh, 64:], r3[:, hh, 64:]) < 8e-3


@pytest.mark.parametrize("hd", [64, 80])
def test_global_softmax_cannot_overflow(hd):
    """Adversarial logits for the single-pass global softmax: in every 64-key tile ONE key beats everything seen before by
    ~100 natural units (the running maximum rises by 100 x 64 tiles), and the rel-pos tables have std 2.0.  torch.softmax
    (image_encoder.py:246-252) has no limit on such inputs; the kernel bounds every exponent BEFORE the first exponential
    (upper bound from the tile's raw maximum), so the default mode must stay finite and within the bf16 bar."""
    g, ws, heads, B = 64, 64, 1, 1
    D = heads * hd
    gen = torch.Generator(device="cpu").manual_seed(hd)
    qkv = torch.randn(B * g * g, 3 * D, generator=gen)
    u = torch.randn(hd, generator=gen)
    u = u / u.norm()
    scale = hd ** -0.5
    q = 20.0 * u[None, :] + 0.3 * torch.randn(g * g, hd, generator=gen)         # every query points along u, |q| ~ 20
    k = 0.5 * torch.randn(g * g, hd, generator=gen)
    for t in range(64):                                                            # key 64 t + 5: logit ~ 100 (t + 1)
        k[64 * t + 5] = u * (100.0 * (t + 1) / (20.0 * scale))
    qkv[:, :hd], qkv[:, D:D + hd] = q, k
    qkv = qkv.bfloat16()
    rel_h, rel_w = torch.randn(127, hd, generator=gen) * 2.0, torch.randn(127, hd, generator=gen) * 2.0
    bias = torch.zeros(3 * D)
    ref = ref_attention_core(qkv.float(), rel_h.bfloat16().float(), rel_w.bfloat16().float(), bias, B, g, ws, heads)
    assert torch.isfinite(ref).all()
    out = _attention_tc(qkv.to(DEV), rel_h.to(DEV), rel_w.to(DEV), bias.to(DEV), B, g, ws, heads, hd)
    assert torch.isfinite(out.float()).all(), "overflow in the single-pass softmax"
    err = ib.rel_l2(out, ref)
    assert err < 1e-2, err


@pytest.mark.parametrize("hd", [64, 80])
def test_windowed_softmax_reference_moves_without_overflow(hd):
    """Adversarial logits for the windowed softmax: in every 32-key chunk of a window after the first, ONE key beats everything before
    it by ~60 natural units (87 in base 2), and the rel-pos tables have std 2.0.  The production kernel takes the exact row maximum
    first (two passes over tensor memory) and must stay finite and within the bf16 bar; the one-pass variant of
    csrc/experiments/attention_win5.cu (reference = maximum of the row's first 32 keys, moved with an in-place rescale of the written
    P columns when a later chunk exceeds it by 2^64) is held to the same test.  torch.softmax (image_encoder.py:246-252) has no limit
    on such inputs."""
    g, ws, heads, B = 64, 14, 1, 1
    D = heads * hd
    gen = torch.Generator(device="cpu").manual_seed(100 + hd)
    qkv = torch.randn(B * g * g, 3 * D, generator=gen)
    u = torch.randn(hd, generator=gen)
    u = u / u.norm()
    scale = hd ** -0.5
    q = 20.0 * u[None, :] + 0.3 * torch.randn(g * g, hd, generator=gen)         # every query points along u, |q| ~ 20
    k = 0.5 * torch.randn(g * g, hd, generator=gen)
    ys, xs = torch.meshgrid(torch.arange(g), torch.arange(g), indexing="ij")
    local = ((ys % ws) * ws + (xs % ws)).reshape(-1)                               # key index inside its window
    for j, kl in enumerate((40, 70, 100, 140, 170, 194)):                          # one key per later chunk: logit ~ 60 (j + 1)
        k[local == kl] = u * (60.0 * (j + 1) / (20.0 * scale))
    qkv[:, :hd], qkv[:, D:D + hd] = q, k
    qkv = qkv.bfloat16()
    rel_h, rel_w = torch.randn(27, hd, generator=gen) * 2.0, torch.randn(27, hd, generator=gen) * 2.0
    bias = torch.zeros(3 * D)
    ref = ref_attention_core(qkv.float(), rel_h.bfloat16().float(), rel_w.bfloat16().float(), bias, B, g, ws, heads)
    assert torch.isfinite(ref).all()
    out = _attention_tc(qkv.to(DEV), rel_h.to(DEV), rel_w.to(DEV), bias.to(DEV), B, g, ws, heads, hd)
    assert torch.isfinite(out.float()).all(), "overflow in the one-pass windowed softmax"
    err = ib.rel_l2(out, ref)
    assert err < 1e-2, err


def _ref_attention_grid(qkv, rel_h, rel_w, qkv_bias, B, gh, gw, ws, heads):
    """fp64 reference of the attention core on a (gh x gw) token grid: ws = 0 global (tables of 2 gh - 1 / 2 gw - 1 rows), ws = 14
    windows of the zero-padded-then-biased grid (image_encoder.py:239-304, 340-376)."""
    from oracle import sam_vit_oracle as orc
    D = qkv.shape[1] // 3
    hd = D // heads
    x = qkv.reshape(B, gh, gw, 3 * D).double()
    if ws:
        Hp, Wp = -(-gh // ws) * ws, -(-gw // ws) * ws
        xp = qkv_bias.double().reshape(1, 1, 1, -1).expand(B, Hp, Wp, 3 * D).clone()
        xp[:, :gh, :gw] = x
        wins, pad_hw = orc.window_partition(xp, ws)
        sh, sw = ws, ws
    else:
        wins, sh, sw = x, gh, gw
    Bp, S = wins.shape[0], sh * sw
    q, k, v = wins.reshape(Bp, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    scores = (q * hd ** -0.5) @ k.transpose(-1, -2)
    Rh, Rw = orc.rel_pos_rows(sh, sh, rel_h.double()), orc.rel_pos_rows(sw, sw, rel_w.double())
    q5 = q.reshape(Bp, heads, sh, sw, hd)
    bh = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)
    bw = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)
    scores = (scores.reshape(Bp, heads, sh, sw, sh, sw) + bh[..., :, None] + bw[..., None, :]).reshape(Bp, heads, S, S)
    out = (torch.softmax(scores, -1) @ v).permute(0, 2, 1, 3).reshape(Bp, sh, sw, D)
    if ws:
        out = orc.window_unpartition(out, ws, pad_hw, (gh, gw))
    return out.reshape(B * gh * gw, D)


@pytest.mark.parametrize("gh,gw,heads,hd,B", [(64, 128, 2, 80, 1), (96, 32, 2, 64, 2), (32, 64, 1, 80, 1), (32, 96, 1, 64, 1), (64, 64, 1, 80, 1)])
def test_attention_tcgen05_other_token_grids(gh, gw, heads, hd, B):
    """Scope row N3 on the tensor cores: windowed and global attention on token grids other than 64 x 64 (svb_attention_window_hw,
    svb_attention_global_hw; e.g. the 64 x 128 grid of the reference's 1024 x 2048 evaluation canvas) against the fp64 reference."""
    lib, st = cabi.lib(), cabi.stream_ptr
    D = heads * hd
    gen = torch.Generator(device="cpu").manual_seed(gh * 7 + gw + hd)
    qkv = torch.randn(B * gh * gw, 3 * D, generator=gen).bfloat16()
    bias = torch.randn(3 * D, generator=gen)
    # ---- windowed (14 x 14) ----
    rel_h, rel_w = torch.randn(27, hd, generator=gen) * 0.3, torch.randn(27, hd, generator=gen) * 0.3
    ref = _ref_attention_grid(qkv.float(), rel_h.bfloat16().float(), rel_w.bfloat16().float(), bias.bfloat16().float(), B, gh, gw, 14, heads)
    pack = torch.zeros(lib.svb_rel_pack_rows(14, 64), hd, dtype=torch.bfloat16, device=DEV)
    rh27, rw27, bias_d = rel_h.to(DEV), rel_w.to(DEV), bias.to(DEV)
    cabi.check(lib.svb_pack_rel_table(rh27.data_ptr(), pack.data_ptr(), 27, hd, 0, st()), "pack h")
    cabi.check(lib.svb_pack_rel_table(rw27.data_ptr(), pack.data_ptr(), 27, hd, 1, st()), "pack w")
    gph, gpw = -(-gh // 14) * 14, -(-gw // 14) * 14
    src = torch.full((B, gph, gpw, 3 * D), float("nan"), dtype=torch.bfloat16, device=DEV)
    src[:, :gh, :gw] = qkv.to(DEV).reshape(B, gh, gw, 3 * D)
    cabi.check(lib.svb_fill_pad_rows_hw(src.data_ptr(), bias_d.data_ptr(), B, gh, gw, 3 * D, st()), "fill_pad_rows_hw")
    out = torch.full((B * gh * gw, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    cabi.check(lib.svb_attention_window_hw(src.data_ptr(), out.data_ptr(), pack.data_ptr(), B, gh, gw, heads, hd, st()), "window_hw")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert ib.rel_l2(out, ref) < 6e-3, ("windowed", ib.rel_l2(out, ref))
    # ---- global: tables of 2 g - 1 rows (what get_rel_pos's linear resize hands to the bias einsums) ----
    rel_h, rel_w = torch.randn(2 * gh - 1, hd, generator=gen) * 0.2, torch.randn(2 * gw - 1, hd, generator=gen) * 0.2
    ref = _ref_attention_grid(qkv.float(), rel_h.bfloat16().float(), rel_w.bfloat16().float(), bias, B, gh, gw, 0, heads)
    nws = int(lib.svb_attention_global_hw_workspace(B, gh, gw, heads, hd))
    ws = torch.empty(nws + 256, dtype=torch.uint8, device=DEV)
    base = (ws.data_ptr() + 255) & ~255
    out = torch.full((B * gh * gw, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    qd, rhd, rwd = qkv.to(DEV), rel_h.to(DEV), rel_w.to(DEV)          # (kept alive: a temporary's block would be reused by the next one)
    cabi.check(lib.svb_attention_global_hw(qd.data_ptr(), out.data_ptr(), rhd.data_ptr(), rwd.data_ptr(), B, gh, gw, heads, hd, base, nws, st()),
               "global_hw")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    err = ib.rel_l2(out, ref)
    assert err < 6e-3, ("global", err)
    o3, r3 = out.float().cpu().reshape(-1, heads, hd), ref.float().reshape(-1, heads, hd)
    for hh in range(heads):
        assert ib.rel_l2(o3[:, hh], r3[:, hh]) < 8e-3


def test_linear_remap_to_padded_grid():
    """qkv GEMM epilogue storing token rows at their position in the window-padded 70x70 grid."""
    g, gp, B, K, N = 64, 70, 2, 64, 256
    gen = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randn(B * g * g, K, generator=gen).to(DEV).bfloat16()
    W = (torch.randn(N, K, generator=gen) / 8).to(DEV).bfloat16()
    bias = torch.randn(N, generator=gen).to(DEV)
    out = torch.zeros(B * gp * gp, N, dtype=torch.bfloat16, device=DEV)
    _linear(cabi.MODE_BF16, A, W, bias=bias, out=out, remap=(g, gp))
    cabi.check(cabi.lib().svb_fill_pad_rows(out.data_ptr(), bias.data_ptr(), B, g, gp, N, cabi.stream_ptr()), "fill")
    torch.cuda.synchronize()
    ref, _ = _ref_linear(A, W, bias=bias)
    full = bias.double().reshape(1, 1, 1, N).expand(B, gp, gp, N).clone()
    full[:, :g, :g] = ref.reshape(B, g, g, N)
    assert ib.rel_l2(out.reshape(B, gp, gp, N), full) < 4e-3
    assert torch.equal(out.reshape(B, gp, gp, N)[:, g:, :, :].float(), bias.bfloat16().float().expand(B, gp - g, gp, N))


@pytest.mark.parametrize("ws,heads,hd,B", [(14, 16, 80, 4), (14, 12, 64, 4), (64, 4, 80, 2)])
def test_attention_tcgen05_is_deterministic(ws, heads, hd, B):
    """Multi-wave grids (CTAs reusing an SM's shared / tensor memory, two windowed CTAs per SM) must give bit-identical
    results run after run: guards the cross-proxy (generic vs TMA) shared-memory reuse rules of attention_tc.cu."""
    g = 64
    D = heads * hd
    gen = torch.Generator(device="cpu").manual_seed(ws + hd + B)
    qkv = torch.randn(B * g * g, 3 * D, generator=gen).bfloat16().to(DEV)
    L = 2 * ws - 1
    rel_h, rel_w = (torch.randn(L, hd, generator=gen) * 0.1).to(DEV), (torch.randn(L, hd, generator=gen) * 0.1).to(DEV)
    bias = torch.randn(3 * D, generator=gen).to(DEV)
    first = _attention_tc(qkv, rel_h, rel_w, bias, B, g, ws, heads, hd)
    assert torch.isfinite(first.float()).all()
    for _ in range(12):
        again = _attention_tc(qkv, rel_h, rel_w, bias, B, g, ws, heads, hd)
        assert torch.equal(again, first)


@pytest.mark.parametrize("ws,B,rel_std,reps", [(14, 12, 0.02, 300), (14, 16, 0.5, 150), (64, 12, 0.5, 60)])
def test_attention_tcgen05_many_launches_are_bit_identical(ws, B, rel_std, reps):
    """ViT-H geometry at the pass sizes of the batch schedule, many launches on the same inputs: a barrier-protocol race shows up as
    a differing output or as a trapped barrier wait (the store warp of the windowed kernel once aliased two phases of one barrier:
    1 launch in ~20 hung at 12 images)."""
    lib = cabi.lib()
    heads, hd, g = 16, 80, 64
    D = heads * hd
    gen = torch.Generator().manual_seed(3)
    pack = (torch.randn(lib.svb_rel_pack_rows(ws, g), hd, generator=gen) * rel_std).bfloat16().to(DEV)
    shape = (B * g * g, 3 * D) if ws == 64 else (B, 70, 70, 3 * D)
    qkv = torch.randn(shape, generator=gen).bfloat16().to(DEV)
    out = torch.empty(B * g * g, D, dtype=torch.bfloat16, device=DEV)
    ref = None
    for r in range(reps):
        out.fill_(0)
        cabi.check(lib.svb_attention_tc(qkv.data_ptr(), out.data_ptr(), pack.data_ptr(), B, g, ws, heads, hd, cabi.stream_ptr()), "attn")
        if r % 10 == 0 or r == reps - 1:
            torch.cuda.synchronize()
            if ref is None:
                ref = out.clone()
                assert torch.isfinite(ref.float()).all()
            else:
                assert torch.equal(ref, out), f"launch {r} differs"


@pytest.mark.parametrize("env", [{"SVB_ATTNW_POLY": "0", "SVB_ATTNG_POLY": "1"}])
def test_attention_ab_variants_stay_correct(env):
    """The measured-and-kept A/B variants of the attention kernels (selected by environment variables that are read once per process)
    pass the same parity tests as the defaults: exponentials all on the MUFU / a quarter on the FMA pipe."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_ops.py", "-q", "-x", "-k", "(test_attention_tcgen05 or test_windowed_softmax) and not variants and not many_launches"],
                       cwd=root, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (env, r.stdout[-1500:], r.stderr[-500:])


@pytest.mark.parametrize("M,N,K,bias", [(4096, 101, 512, False), (65536, 101, 512, False), (1000, 37, 64, True), (520, 128, 192, True),
                                         (8192, 256, 128, True)])
def test_linear_nt_transposed_store(M, N, K, bias):
    """svb_linear_nt: out_t[n, m] = sum_k A[m, k] W[n, k] + bias[n] — the mask-logit einsum (xdecoder.py:459) with the positions along
    the GEMM's M and the fp32 result stored query-major.  Against the fp64 product of the same bf16 operands, and against svb_linear
    with the operands swapped (the path it replaces)."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    b = torch.randn(N, generator=g).to(DEV) if bias else None
    out = torch.full((N, M), float("nan"), device=DEV)
    cabi.check(cabi.lib().svb_linear_nt(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), M, N, K, cabi.ptr(b), out.data_ptr(), M,
                                         cabi.stream_ptr()), "svb_linear_nt")
    torch.cuda.synchronize()
    ref = W.double() @ A.double().t() + (b.double()[:, None] if bias else 0)
    assert torch.isfinite(out).all()
    assert ib.rel_l2(out, ref) < 1e-5
    swapped = _linear(cabi.MODE_BF16, W, A, out_dtype=torch.float32)             # (N, M) = W A^T, no bias
    assert ib.rel_l2(out - (b[:, None] if bias else 0), swapped) < 1e-6
