"""CPU tests of the pass schedules (host logic, no device): how a batch is split into passes through the kernels by
svb_encoder_forward (wave quantisation of the block GEMMs) and svb_encoder_forward_host (the same + the exposed upload of the first
pass / download of the last).  The product calls the same code through svb_encoder_pass_schedule; here it runs from the geometry
alone (svb_pass_schedule_model)."""
import ctypes as C

import pytest

from iuvl_b200 import cabi

GEOM = {"vit_b": (768, 3072, 12), "vit_l": (1024, 4096, 24), "vit_h": (1280, 5120, 32)}
IN_BYTES, OUT_BYTES_BF16 = 3 * 1024 * 1024 * 4.0, 15728640 * 2.0          # one 1024^2 fp32 image in, res2..res5 in bf16 out


def schedule(model, batch, chunk, host, sms=148, tokens=4096):
    d, mlp, depth = GEOM[model]
    buf = (C.c_int * 256)()
    rc = cabi.lib().svb_pass_schedule_model(d, mlp, depth, tokens, sms, batch, chunk, int(host), IN_BYTES, OUT_BYTES_BF16, buf, 256)
    assert rc >= 1000, cabi.lib().svb_last_error()
    return [buf[i] for i in range(rc - 1000)]


@pytest.mark.parametrize("model", sorted(GEOM))
@pytest.mark.parametrize("chunk", [1, 2, 8, 16])
def test_every_image_runs_once_and_no_pass_exceeds_the_chunk(model, chunk):
    for batch in (1, 2, 3, 7, 8, 13, 16, 31, 64, 100):
        for host in (False, True):
            s = schedule(model, batch, chunk, host)
            assert sum(s) == batch and all(1 <= c <= chunk for c in s), (model, batch, chunk, host, s)


def test_device_schedule_follows_the_wave_quantisation():
    # ViT-H on 74 CTA pairs: 12 images fill 99.8 % of the rounds' tile slots, 16 images 96 % -> 64 images run as 16 + 4 x 12
    assert schedule("vit_h", 64, 16, False) == [16, 12, 12, 12, 12]
    assert schedule("vit_h", 16, 16, False) == [16] and schedule("vit_b", 16, 16, False) == [16]      # fits the chunk: one pass
    assert schedule("vit_h", 64, 8, False) == [8] * 8
    assert schedule("vit_l", 32, 16, False) == [16, 16]       # (not 15 + 15 + 2: a pass has a fixed cost)


def test_host_schedule_keeps_the_exposed_copies_small():
    s = schedule("vit_h", 64, 16, True)
    assert s == [12, 12, 12, 12, 12, 4]                   # the last pass (whose download nothing overlaps) is the smallest
    assert schedule("vit_b", 16, 16, True) == [4, 8, 4]    # a single pass would expose both copies of all 16 images
    for model in GEOM:
        for batch in (8, 16, 32, 64):
            dev, host = schedule(model, batch, 16, False), schedule(model, batch, 16, True)
            assert host[-1] <= dev[-1] or host[-1] <= 4
            assert host[0] % 4 == 0 and host[-1] % 4 == 0  # end passes in whole groups of 4 images
    assert schedule("vit_h", 3, 2, True) == [2, 1]         # no room for a 4-image end pass: the device schedule


def test_bad_arguments_are_refused():
    buf = (C.c_int * 4)()
    lib = cabi.lib()
    assert 0 < lib.svb_pass_schedule_model(1280, 5120, 32, 4096, 148, 0, 16, 0, IN_BYTES, OUT_BYTES_BF16, buf, 4) < 1000
    assert 0 < lib.svb_pass_schedule_model(1280, 5120, 32, 4096, 148, 64, 1, 0, IN_BYTES, OUT_BYTES_BF16, buf, 4) < 1000   # 64 passes, 4 slots
