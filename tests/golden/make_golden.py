"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Run in the build container only (needs ``/root/reference``; nothing at test/bench time reads it):

    python tests/golden/make_golden.py [case ...]

For every case: build the reference ``ImageEncoderViT`` (``sam/modeling/image_encoder.py:17``) with the
kwargs ``_build_sam`` uses (``sam/build_sam.py:60-73``), load the seeded synthetic ``state_dict``
(``synthetic.make_state_dict``) with ``strict=True``, run it in fp32 / eval / no_grad on the seeded
synthetic images, and store — for each output and for the token stream after the patch embedding and
after every block (forward hooks) — a fixed random sample of values, their flat indices, and the
tensor's L2 norm / mean.  Full tensors are up to 63 MB per image, so samples keep the fixtures small.
"""
from __future__ import annotations

import os
import sys
import time
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import iuvl_b200 as ib  # noqa: E402

# name -> (preset, batch, rel_std, n_samples)
CASES = {
    "tiny64_std": ("tiny64", 2, 0.02, 4096),
    "tiny64_stress": ("tiny64", 1, 0.5, 4096),
    "tiny80_std": ("tiny80", 1, 0.02, 4096),
    "tiny80_stress": ("tiny80", 2, 0.5, 4096),
    "vit_b_std": ("vit_b", 1, 0.02, 4096),
    "vit_b_stress": ("vit_b", 1, 0.5, 4096),
    "vit_l_std": ("vit_l", 1, 0.02, 4096),
    "vit_h_std": ("vit_h", 1, 0.02, 4096),
    "vit_h_stress": ("vit_h", 1, 0.5, 4096),
    # three DISTINCT images per model, 16384 samples per output: the parity cases at the benchmarked batch / pass sizes
    # (ViT-B 16, ViT-L 32, ViT-H 13 and 64 images in passes of 12 / 16) are built from these
    "vit_b_std3": ("vit_b", 3, 0.02, 16384),
    "vit_l_std3": ("vit_l", 3, 0.02, 16384),
    "vit_h_std3": ("vit_h", 3, 0.02, 16384),
    # scope row N3: canvases other than 1024 x 1024 run the reference's bicubic pos_embed / linear rel_pos fallbacks
    # (image_encoder.py:111-114,124-132,319-330); 5-tuples carry the (H, W) of the input
    "tiny64_wide": ("tiny64", 1, 0.1, 4096, (1024, 2048)),
    "tiny80_tall": ("tiny80", 1, 0.1, 4096, (1536, 512)),
    "vit_b_wide": ("vit_b", 1, 0.05, 16384, (1024, 2048)),        # the reference's COCO evaluation canvas (configs/step1.yaml:211-212)
}
WEIGHT_SEED, IMAGE_SEED = 1234, 0


def build_reference(cfg):
    from sam.modeling import ImageEncoderViT  # the unmodified reference class
    enc = ImageEncoderViT(
        depth=cfg.depth, embed_dim=cfg.embed_dim, img_size=cfg.img_size, mlp_ratio=4,
        norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=cfg.num_heads, patch_size=cfg.patch_size,
        qkv_bias=True, use_rel_pos=True, global_attn_indexes=list(cfg.global_attn_indexes), window_size=14,
        out_chans=256)
    return enc.eval()


def sample(t: torch.Tensor, n: int, rs: np.random.RandomState):
    flat = t.detach().reshape(-1).double()
    idx = rs.randint(0, flat.numel(), size=min(n, flat.numel())).astype(np.int64)
    return {"idx": idx, "val": flat[torch.from_numpy(idx)].float().numpy(),
            "norm": np.float64(flat.norm().item()), "mean": np.float64(flat.mean().item()),
            "shape": np.array(t.shape, dtype=np.int64)}


def run_case(name: str):
    preset, batch, rel_std, n = CASES[name][:4]
    hw = CASES[name][4] if len(CASES[name]) > 4 else None
    cfg = ib.PRESETS[preset]
    sd = ib.make_state_dict(cfg, WEIGHT_SEED, rel_std=rel_std)
    x = ib.make_images(batch, cfg, IMAGE_SEED, hw=hw)
    enc = build_reference(cfg)
    missing = enc.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    taps = {}
    hooks = [enc.patch_embed.register_forward_hook(lambda m, i, o: None)]
    for i, blk in enumerate(enc.blocks):
        hooks.append(blk.register_forward_hook(lambda m, inp, out, i=i: taps.__setitem__(f"block{i}", out.detach()[:1].clone())))
    # token stream after patch-embed + pos: the input of block 0
    hooks.append(enc.blocks[0].register_forward_pre_hook(lambda m, inp: taps.__setitem__("embed", inp[0].detach()[:1].clone())))
    t0 = time.time()
    outs = {k: [] for k in ("res2", "res3", "res4", "res5")}
    with torch.no_grad():
        for b in range(batch):   # per image: the reference materialises (h,S,S) scores
            o = enc(x[b:b + 1])
            for k in outs:
                outs[k].append(o[k])
            if b == 0:
                taps0 = dict(taps)
    dt = time.time() - t0
    for h in hooks:
        h.remove()
    rs = np.random.RandomState(12345)
    blob = {"meta_preset": np.array(preset), "meta_batch": np.int64(batch), "meta_rel_std": np.float64(rel_std),
            "meta_weight_seed": np.int64(WEIGHT_SEED), "meta_image_seed": np.int64(IMAGE_SEED),
            "meta_ref_seconds": np.float64(dt), "meta_torch": np.array(torch.__version__),
            "meta_hw": np.array(hw if hw else (cfg.img_size, cfg.img_size), dtype=np.int64)}
    for k, v in outs.items():
        for kk, vv in sample(torch.cat(v, 0), n, rs).items():
            blob[f"out.{k}.{kk}"] = vv
    for k, v in taps0.items():
        for kk, vv in sample(v, 1024, rs).items():
            blob[f"tap.{k}.{kk}"] = vv
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
    print(f"{name}: reference forward {dt:.1f}s, wrote {name}.npz", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for case in (sys.argv[1:] or list(CASES)):
        run_case(case)
