"""Pins the tcgen05 shared-memory descriptor encodings the attention kernel relies on (single-CTA probe kernel):
K-major / MN-major operands with 128B and 32B swizzle, and the hand-written 128B-swizzled A tile (the P matrix)."""
import itertools
import json
import os

import pytest
import torch

from iuvl_b200 import cabi
from tests.util import ROOT

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run(a, b, K, N, a_sw, b_sw, b_mn, a_manual, al, asb, ak, bl, bsb, bk):
    out = torch.full((128, N), float("nan"), device=DEV)
    cabi.check_probe(cabi.probe_lib().svb_probe_mma(a.data_ptr(), b.data_ptr(), out.data_ptr(), K, N, a_sw, b_sw, b_mn, a_manual,
                                        al, asb, ak, bl, bsb, bk, cabi.stream_ptr()), "probe")
    torch.cuda.synchronize()
    return out


def err(out, ref):
    if not torch.isfinite(out).all():
        return float("inf")
    return float((out.double() - ref).norm() / ref.norm())


def test_descriptor_encodings():
    g = torch.Generator().manual_seed(0)
    report = {}
    # 1. sanity: both K-major SW128, K = 64
    a = torch.randn(128, 64, generator=g).bfloat16().to(DEV)
    b = torch.randn(64, 64, generator=g).bfloat16().to(DEV)
    ref = a.double() @ b.double().t()
    report["kmajor_sw128"] = err(run(a, b, 64, 64, 128, 128, 0, 0, 0, 1024, 32, 0, 1024, 32), ref)
    # 2. A written by hand with the 128B swizzle (P path), K = 128 (two atoms)
    a2 = torch.randn(128, 128, generator=g).bfloat16().to(DEV)
    bmn = torch.randn(128, 64, generator=g).bfloat16().to(DEV)       # [K keys][N = 64]: MN-major B
    ref2 = a2.double() @ bmn.double()
    cands = {}
    for lbo, sbo, kstep in itertools.product((0, 1024, 16384), (1024, 2048), (2048, 1024)):
        cands[f"lbo{lbo}_sbo{sbo}_k{kstep}"] = err(run(a2, bmn, 128, 64, 128, 128, 1, 0, 0, 1024, 32, lbo, sbo, kstep), ref2)
    report["mn_sw128_tma_a"] = cands
    report["mn_sw128_manual_a"] = err(run(a2, bmn, 128, 64, 128, 128, 1, 1, 0, 1024, 32, 0, 1024, 2048), ref2)
    # 2b. A = P in TENSOR MEMORY (tcgen05.st of packed bf16 pairs, A operand taken from TMEM), K = 128, MN-major B
    report["mn_sw128_tmem_a"] = err(run(a2, bmn, 128, 64, 128, 128, 1, 2, 0, 1024, 32, 0, 1024, 2048), ref2)
    # 3. 32B-swizzle tail, K-major both (Q K^T tail: K = 16)
    a3 = torch.randn(128, 16, generator=g).bfloat16().to(DEV)
    b3 = torch.randn(128, 16, generator=g).bfloat16().to(DEV)
    ref3 = a3.double() @ b3.double().t()
    c3 = {}
    for sbo in (256, 128, 512):
        c3[f"sbo{sbo}"] = err(run(a3, b3, 16, 128, 32, 32, 0, 0, 0, sbo, 0, 0, sbo, 0), ref3)
    report["kmajor_sw32"] = c3
    # 4. 32B-swizzle MN-major B tail (V tail: N = 16), A hand-written SW128 with K = 128
    b4 = torch.randn(128, 16, generator=g).bfloat16().to(DEV)
    ref4 = a2.double() @ b4.double()
    c4 = {}
    for lbo, sbo, kstep in itertools.product((0, 256), (256, 128, 512), (512, 256)):
        c4[f"lbo{lbo}_sbo{sbo}_k{kstep}"] = err(run(a2, b4, 128, 16, 128, 32, 1, 1, 0, 1024, 32, lbo, sbo, kstep), ref4)
    report["mn_sw32"] = c4
    report["mn_sw32_tmem_a"] = err(run(a2, b4, 128, 16, 128, 32, 1, 2, 0, 1024, 32, 0, 256, 512), ref4)
    # 5. MN-major B spanning two 64-element atoms along N (V with head_dim 80 as ONE MMA): LBO = byte stride between the atoms
    b5 = torch.randn(128, 80, generator=g).bfloat16().to(DEV)
    ref5 = a2.double() @ b5.double()
    c5 = {}
    for lbo in (16384, 1024, 0):
        c5[f"lbo{lbo}_sbo1024_k2048"] = err(run(a2, b5, 128, 80, 128, 128, 1, 2, 0, 1024, 32, lbo, 1024, 2048), ref5)
    report["mn_sw128_two_atoms_tmem_a"] = c5
    # 6. the same with the 32B swizzle: five 16-element atoms along N, K * 32 bytes apart (no padding columns in shared memory)
    c6 = {}
    for lbo in (4096, 256, 0):
        c6[f"lbo{lbo}_sbo256_k512"] = err(run(a2, b5, 128, 80, 128, 32, 1, 2, 0, 1024, 32, lbo, 256, 512), ref5)
    report["mn_sw32_five_atoms_tmem_a"] = c6
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))
    # the encodings attention_tc.cu uses
    assert report["kmajor_sw128"] < 1e-5
    assert report["mn_sw128_tma_a"]["lbo0_sbo1024_k2048"] < 1e-5
    assert report["mn_sw128_manual_a"] < 1e-5
    assert report["mn_sw128_tmem_a"] < 1e-5
    assert report["mn_sw32_tmem_a"] < 1e-5
    assert report["kmajor_sw32"]["sbo256"] < 1e-5
    assert report["mn_sw32"]["lbo0_sbo256_k512"] < 1e-5
    assert report["mn_sw128_two_atoms_tmem_a"]["lbo16384_sbo1024_k2048"] < 1e-5
    assert report["mn_sw32_five_atoms_tmem_a"]["lbo4096_sbo256_k512"] < 1e-5
