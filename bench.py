#!/usr/bin/env python
"""Headline benchmark: SAM ViT-H 1024x1024 image-encoder forward, images/s (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU (oracle port)

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); images are sharded data-parallel by batch, no
collective on the data path.  A "step" = one encoder forward over one batch of synthetic images per GPU.
Prints ONE JSON line on rank 0 (see the contract in the task statement): `value` = whole-job images/s with inputs
resident in HBM; `e2e` = same metric through the host-buffer C-ABI call (H2D + D2H inside the timed region);
`roofline` = dominant kernel (tcgen05 GEMM) from CUDA events around every launch; `cpu_baseline` = the oracle
timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SAM ViT-H 1024x1024 encoder images/sec"
UNIT = "images/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def traffic_per_launch(model):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, averaged over the GEMM launches of one
    `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_traffic.py from the committed capture); None if absent."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return d.get(model, {}).get("gemm_bytes_per_launch_avg")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # the sampler runs from before the warm-up: keep the samples taken under load (power >= 60 % of the maximum seen)
            busy = [c for c, w in zip(sm, pw) if w >= 0.6 * max(pw)] or sm
            busy.sort()
            out.update(sm_mhz=busy[len(busy) // 2], sm_max_mhz=max(mx), power_w_max=max(pw), reasons=sorted(reasons),
                       samples=len(sm), samples_under_load=len(busy))
        return out


REF_DIR = os.path.join(ROOT, "oracle", "_ref")      # unmodified copy of the reference's sam/ package (git-ignored; made by build())


def reference_encoder(model, torch):
    """The reference's own CPU implementation of the path: `sam.build_sam.sam_model_registry[...]().image_encoder` imported from the
    UNMODIFIED copy of its `sam/` package under oracle/_ref (`__graft_entry__.build()` copies it from /root/reference when that
    exists; it travels to the GPU box like the built .so).  Returns (callable(sd, x) -> outputs, kind)."""
    import iuvl_b200 as ib
    cfg = ib.PRESETS[model]
    if os.path.exists(os.path.join(REF_DIR, "sam", "build_sam.py")):
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from sam.build_sam import sam_model_registry          # the reference's public factory (sam/build_sam.py:108-112)
        enc = sam_model_registry[model](checkpoint=None).image_encoder.eval()
        state = {"loaded": False}

        def run(sd, x):
            if not state["loaded"]:
                enc.load_state_dict(sd, strict=True)
                state["loaded"] = True
            with torch.no_grad():
                return enc(x)
        return run, "reference"
    from oracle import sam_vit_oracle as orc                   # the restatement (kind "port") when the copy is absent
    return (lambda sd, x: orc.encoder_forward_cfg(sd, x, cfg)), "port"


def run_reference(args, rank, world):
    """--impl reference: the reference's own encoder (unmodified sam/ package from oracle/_ref; the oracle port only if that copy
    is absent) on the host CPU with all threads.  Each step = ONE image of the workload (bounded sample)."""
    if rank != 0:
        return
    import torch
    import iuvl_b200 as ib
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = ib.PRESETS[args.model]
    sd = ib.make_state_dict(cfg, 1234)
    x = ib.make_images(1, cfg, 0)
    fwd, kind = reference_encoder(args.model, torch)
    for _ in range(args.warmup):
        fwd(sd, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fwd(sd, x)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = 1.0 / dt
    what = "unmodified reference ImageEncoderViT (oracle/_ref/sam)" if kind == "reference" else "oracle port"
    sample = f"1 image of the {args.model} workload per step, fp32, {what}, torch CPU ops, {torch.get_num_threads()} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_next_rows(torch, cabi, dev, pk):
    """Kernel-alone CUDA-event timings of scope rows N1 / N2 (not part of `value`): the multi-scale deformable attention forward at
    the step1.yaml geometry (4 images, bf16 values; an L2 -> SM gather) and the fused uint8 staging of 16 images (HBM-bound)."""
    import ctypes as C
    from iuvl_b200.msda import ms_deform_attn_forward
    out = {}
    shapes, M, D, P, L, N = [(128, 128), (64, 64), (32, 32)], 8, 64, 4, 3, 4
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, device=dev).bfloat16()
    loc = (torch.rand(N, S, 1, 1, 1, 2, device=dev) + 0.05 * torch.randn(N, S, M, L, P, 2, device=dev)).contiguous()
    aw = torch.softmax(torch.randn(N, S, M, L * P, device=dev), -1).reshape(N, S, M, L, P).contiguous()
    sh, st = torch.tensor(shapes), torch.tensor([0, 128 * 128, 128 * 128 + 64 * 64])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(3):
            ms_deform_attn_forward(value, sh, st, loc, aw)
        e0.record()
        for _ in range(10):
            ms_deform_attn_forward(value, sh, st, loc, aw)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gather = N * S * M * L * P * 4 * D * 2
    out["msda_forward"] = {"workload": "4 x (128^2+64^2+32^2) queries, 8 heads x 64 ch, 3 levels x 4 points, bf16", "us_per_launch": ms * 1e3,
                           "gathered_gbs": gather / ms / 1e6, "bound": "L1/L2->SM gather (value maps are L2-resident)"}
    # the whole MSDeformAttn module forward (4 GEMMs + the fused softmax / location / gather kernel), d_model 512
    from iuvl_b200.msda import MSDeformAttn
    mod = MSDeformAttn(512, 3, 8, 4).to(dev)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.05)
        mod.attention_weights.weight.normal_(0, 0.05)
        xq = torch.randn(N, S, 512, device=dev).bfloat16()
        rp = torch.rand(N, S, 3, 2, device=dev)
        for _ in range(2):
            mod(xq, rp, xq, sh, st)
        e0.record()
        for _ in range(5):
            mod(xq, rp, xq, sh, st)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out["msda_module"] = {"workload": "MSDeformAttn.forward, d_model 512, 4 images x 21504 queries, bf16", "ms_per_layer": ms,
                          "images_per_s_per_layer": N / ms * 1e3}
    del mod, xq
    # row N1, the whole pixel decoder on the four maps the encoder produces (transformer_encoder_deform.py:315-359; step1.yaml geometry)
    from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder
    dec = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=6,
                                   conv_dim=512, mask_dim=512, norm="GN").to(dev).eval()
    with torch.no_grad():
        for layer in dec.transformer.encoder.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
            layer.self_attn.attention_weights.weight.normal_(0, 0.05)
        feats = {f"res{2 + i}": torch.randn(N, c, 256 >> i, 256 >> i, device=dev).bfloat16() for i, c in enumerate((128, 256, 512, 1024))}
        for _ in range(2):
            dec(feats)
        e0.record()
        for _ in range(3):
            dec(feats)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out["pixel_decoder"] = {"workload": "MSDeformAttnPixelDecoder.forward, conv_dim 512, 6 encoder layers, 4 images (res2..res5 of 1024^2 inputs), bf16",
                            "ms_per_forward": ms, "images_per_s": N / ms * 1e3}
    del dec, feats
    # row N4 (first slice): the mask branch of the prediction heads, one call as a decoder layer makes it (xdecoder.py:457-470)
    from iuvl_b200.mask_head import MaskPredictionHead
    head = MaskPredictionHead(512, 512, 101, 8).to(dev).eval()
    with torch.no_grad():
        q = torch.randn(101, N, 512, device=dev)
        mf = torch.randn(N, 512, 256, 256, device=dev).bfloat16()
        for _ in range(2):
            head(q, mf, (64, 64))
        e0.record()
        for _ in range(5):
            head(q, mf, (64, 64))
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out["mask_head"] = {"workload": "mask branch of forward_prediction_heads, 101 queries x 512, mask_features 512 x 256^2, 4 images, 64^2 attention mask, bf16",
                        "ms_per_call": ms}
    # the whole mask path of the predictor (9 layers: masked cross-attention, self-attention, FFN, mask branch; xdecoder.py:191-329)
    from iuvl_b200.mask_head import XDecoderMaskPath
    path = XDecoderMaskPath(512, 512, 101, 8, 2048).to(dev).eval()
    with torch.no_grad():
        xs = [torch.randn(N, 512, sd_, sd_, device=dev) for sd_ in (32, 64, 128)]
        for _ in range(2):
            path(xs, mf)
        e0.record()
        for _ in range(3):
            path(xs, mf)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out["xdecoder_mask_path"] = {"workload": "XDecoder.forward mask path (task seg), 9 layers, 101 queries x 512, 4 images, 256^2 masks, bf16",
                                 "ms_per_forward": ms, "images_per_s": N / ms * 1e3}
    del head, mf, path, xs
    B = 16
    imgs = [torch.randint(0, 256, (3, 1024, 1024), dtype=torch.uint8, device=dev) for _ in range(B)]
    dst = torch.empty(B * 4096, 768, dtype=torch.bfloat16, device=dev)
    ptrs = (C.c_void_p * B)(*[t.data_ptr() for t in imgs])
    hs, wsz = (C.c_int * B)(*([1024] * B)), (C.c_int * B)(*([1024] * B))
    mean, std = (C.c_float * 3)(123.675, 116.28, 103.53), (C.c_float * 3)(58.395, 57.12, 57.375)
    lib = cabi.lib()
    for _ in range(3):
        cabi.check(lib.svb_stage_images_u8(ptrs, hs, wsz, B, 3, 1024, 16, mean, std, dst.data_ptr(), cabi.DTYPE_BF16, cabi.stream_ptr()))
    e0.record()
    for _ in range(10):
        cabi.check(lib.svb_stage_images_u8(ptrs, hs, wsz, B, 3, 1024, 16, mean, std, dst.data_ptr(), cabi.DTYPE_BF16, cabi.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = B * 3 * 1024 * 1024 * (1 + 2)
    out["stage_u8"] = {"workload": "16 uint8 1024^2 images -> normalised bf16 patch rows", "us_per_launch": ms * 1e3, "gbs": byts / ms / 1e6,
                       "frac_of_hbm_peak": byts / ms / 1e6 / pk["hbm_gbs"], "bound": "hbm"}
    return out


def workload_config(args, world):
    return {
        "workload": "SAM %s image-encoder forward, batch %d x 3x1024x1024 per GPU (BASELINE.json configs[3] shape), "
                    "random-init weights" % ({"vit_b": "ViT-B", "vit_l": "ViT-L", "vit_h": "ViT-H"}[args.model], args.batch),
        "encoder": args.model, "batch_per_gpu": args.batch, "global_batch": args.batch * world, "image": 1024,
        "chunk": args.chunk, "out_dtype": args.out_dtype,
        "parallelism": f"dp{world}: images sharded by batch, no collective on the data path",
        "l2": "per-step inputs (batch x 12.6 MB fp32) exceed the 126 MB L2; no explicit flush",
    }


def measure_pipeline(torch, ib, enc, dev, n_img, steps, barrier, max_over_ranks, world):
    """BASELINE config 5 (step1.yaml): uint8 images -> ViT-H encoder -> MSDeformAttn pixel decoder -> X-Decoder mask path
    (XdecoderHead.forward, modeling/body/xdecoder_head.py:55-58) on this rank's shard of `n_img` images; device-resident
    (CUDA events, max over ranks) and end to end (pinned host uint8 in, pinned host masks out)."""
    from iuvl_b200.mask_head import XDecoderMaskPath
    from iuvl_b200.pixel_decoder import MSDeformAttnPixelDecoder
    dec = MSDeformAttnPixelDecoder(transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024, transformer_enc_layers=6,
                                   conv_dim=512, mask_dim=512, norm="GN").to(dev).eval()
    path = XDecoderMaskPath(512, 512, 101, 8, 2048).to(dev).eval()
    mean, std = [123.675, 116.280, 103.530], [58.395, 57.120, 57.375]                 # configs/step1.yaml:320-321
    g = torch.Generator().manual_seed(77)
    host = torch.randint(0, 256, (n_img, 3, 1024, 1024), generator=g, dtype=torch.uint8).pin_memory()
    x_dev = host.to(dev)
    old = (enc.out_dtype, enc.max_chunk)
    enc.out_dtype, enc.max_chunk = torch.bfloat16, 16
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.no_grad():
        for layer in dec.transformer.encoder.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
            layer.self_attn.attention_weights.weight.normal_(0, 0.05)

        def run(xd):
            ev[0].record()
            feats = enc.forward_uint8(list(xd), mean, std)
            ev[1].record()
            _, multi, extra = dec(feats, rows_out=True)
            ev[2].record()
            out = path(multi, None, mask_rows=extra["mask_rows"], mask_shape=extra["mask_shape"])
            ev[3].record()
            return out
        for _ in range(2):
            out = run(x_dev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = run(x_dev)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        stage = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        masks_host = torch.empty(out["pred_masks"].shape, dtype=out["pred_masks"].dtype).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            xd = host.to(dev, non_blocking=True)
            out = run(xd)
            masks_host.copy_(out["pred_masks"], non_blocking=True)
            torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / steps)
    enc.out_dtype, enc.max_chunk = old
    rec = {"workload": "BASELINE configs[4] (step1.yaml): uint8 1024^2 images -> ViT-H encoder -> MSDeformAttn pixel decoder (conv_dim 512, 6 layers) "
                       "-> X-Decoder mask path (9 layers, 101 queries), %d images per GPU, bf16 hand-overs" % n_img,
           "images_per_gpu": n_img, "global_batch": n_img * world, "value": world * n_img / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "stage_ms_rank0": {"encoder": stage[0], "pixel_decoder": stage[1], "mask_path": stage[2]},
           "e2e": {"value": world * n_img / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(host.numel()),
                   "d2h_bytes_per_step": int(masks_host.numel() * masks_host.element_size())},
           "pred_masks": list(out["pred_masks"].shape)}
    del dec, path, x_dev, host
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="vit_h", choices=["vit_b", "vit_l", "vit_h"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--chunk", type=int, default=16, help="most images per pass through the kernels (the library splits the batch)")
    ap.add_argument("--out-dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling leg, the ViT-B / ViT-L records, the pipeline and the next-row timings")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3   # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import iuvl_b200 as ib
    from iuvl_b200 import cabi
    from iuvl_b200.encoder import build_encoder

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pk = peaks()
    out_dtype = torch.bfloat16 if args.out_dtype == "bf16" else torch.float32

    def make_encoder(model):
        cfg_ = ib.PRESETS[model]
        enc_ = build_encoder(cfg_)
        enc_.load_state_dict(ib.make_state_dict(cfg_, 1234))
        enc_.to(dev)
        enc_.precision = "bf16"
        enc_.out_dtype = out_dtype
        enc_.max_chunk = args.chunk
        return cfg_, enc_

    cfg, enc = make_encoder(args.model)
    lib = cabi.lib()

    B = args.batch
    # synthetic inputs of the workload's shape; a different seed per rank (each GPU owns its shard of the batch)
    x_host = torch.empty(B, 3, cfg.img_size, cfg.img_size, dtype=torch.float32).pin_memory()
    g = torch.Generator().manual_seed(1000 + rank)
    for b in range(B):
        x_host[b] = torch.randn(3, cfg.img_size, cfg.img_size, generator=g)
    x_dev = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(enc_, x, steps, warmup):
        """ms per step of `steps` forwards (device-resident input), barrier + synchronize on both sides, max over ranks"""
        with torch.no_grad():
            for _ in range(warmup):
                enc_(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(steps):
                enc_(x)
            e1.record()
            barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    def timed_e2e(enc_, cfg_, xh, steps):
        """the same metric through the host-buffer C-ABI call: H2D + forward + D2H per step, host clock, max over ranks"""
        nb = xh.shape[0]
        out_host = {f"res{k + 2}": torch.empty(nb, cfg_.fpn_dims[k], cfg_.img_size // s_, cfg_.img_size // s_, dtype=enc_.out_dtype).pin_memory()
                    for k, s_ in enumerate((4, 8, 16, 32))}
        with torch.no_grad():
            enc_.forward_host(xh, out_host)          # warm-up (allocates the staging buffers)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                enc_.forward_host(xh, out_host)       # synchronous: returns after the D2H copies completed
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / steps
        dt = max_over_ranks(dt)
        return {"value": world * nb / dt, "unit": UNIT, "h2d_bytes_per_step": int(xh.numel() * 4),
                "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in out_host.values())),
                "ms_per_step": dt * 1e3, "steps": steps, "timer": "host perf_counter around the synchronous C-ABI call, max over ranks"}

    # ---------------- device-resident throughput (`value`): clean timed region, no per-launch events ----------------
    sampler = ClockSampler(local_rank) if rank == 0 else None     # started before the warm-up: nvidia-smi needs time to spin up
    with torch.no_grad():
        for _ in range(args.warmup):
            out = enc(x_dev)
        barrier()
        launches0 = lib.svb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = enc(x_dev)
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        launches = lib.svb_launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        # ---------------- live roofline: the same step again with CUDA events around every launch ----------------
        prof_steps = max(1, min(args.steps, 2))
        lib.svb_profile_start()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(prof_steps):
            out = enc(x_dev)
        p1.record()
        torch.cuda.synchronize()
        ms_prof_total = p0.elapsed_time(p1)
        ms5, fl5, by5, ln5 = (C.c_double * 5)(), (C.c_double * 5)(), (C.c_double * 5)(), (C.c_int64 * 5)()
        lib.svb_profile_stop(ms5, fl5, by5, ln5)
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---------------- end to end through the host-buffer C-ABI call ----------------
    e2e = None if args.no_e2e else timed_e2e(enc, cfg, x_host, max(10, min(args.steps, 20)))

    # ---------------- strong scaling (BASELINE configs[3]: ONE global batch of `batch` images split over the GPUs) ----------------
    strong = None
    if not args.no_extras:
        per = max(1, B // world)
        if world == 1:
            strong = {"global_batch": B, "per_gpu_batch": B, "value": value, "ms_per_step": ms_step,
                      "e2e": e2e["value"] if e2e else None, "note": "identical to the headline at one GPU"}
        else:
            ms_s = timed(enc, x_dev[:per], args.steps, 3)
            e2e_s = None if args.no_e2e else timed_e2e(enc, cfg, x_host[:per], max(10, min(args.steps, 20)))
            strong = {"global_batch": per * world, "per_gpu_batch": per, "value": world * per / (ms_s * 1e-3), "ms_per_step": ms_s,
                      "steps": args.steps, "e2e": e2e_s["value"] if e2e_s else None,
                      "whole_encoder_frac_of_sustained": per / (ms_s * 1e-3) * cfg.flops_per_image() / 1e12 / pk["bf16_sustained"]}

    # ---------------- BASELINE configs[4] as a pipeline record: encoder -> pixel decoder -> mask path, 16 images per GPU ----------------
    pipeline = None
    if not args.no_extras and args.model == "vit_h":
        try:
            pipeline = measure_pipeline(torch, ib, enc, dev, 16, max(3, min(args.steps, 5)), barrier, max_over_ranks, world)
        except Exception as e:  # noqa: BLE001  (never lose the headline line over an extra)
            pipeline = {"error": str(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_img = cfg.flops_per_image()
    gemm_tflops = (fl5[0] / (ms5[0] * 1e-3) / 1e12) if ms5[0] > 0 else None
    n_gemm = ln5[0]
    whole_tf = value / world * flops_img / 1e12
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc2s_kernel (CTA-pair tcgen05 GEMM: all linears + patch-embed + neck convs)",
        "achieved": gemm_tflops, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
        "frac": (gemm_tflops / pk["bf16_sustained"]) if gemm_tflops else None,
        "frac_of_burst": (gemm_tflops / pk["bf16_burst"]) if gemm_tflops else None,
        # the number the metric asks for: the WHOLE encoder step (all kernels) against the dense bf16 roofline
        "whole_encoder_tflops": whole_tf, "whole_encoder_frac": whole_tf / pk["bf16_sustained"],
        "whole_encoder_frac_of_burst": whole_tf / pk["bf16_burst"], "flops_per_image": flops_img,
        "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step); burst = %.1f" % pk["bf16_burst"],
        "launches": int(n_gemm), "avg_launch_ms": (ms5[0] / n_gemm) if n_gemm else None,
        "flops_per_launch_avg": (fl5[0] / n_gemm) if n_gemm else None,
        "traffic": traffic_per_launch(args.model),
        "profiled_steps": prof_steps, "profiled_ms_per_step": ms_prof_total / prof_steps,
        "step_share": {name: ms5[i] / (ms_prof_total if ms_prof_total else 1) for i, name in
                       enumerate(("gemm", "attn_windowed", "attn_global", "norms", "other"))},
        "categories": {name: {"ms_per_step": ms5[i] / prof_steps, "launches_per_step": ln5[i] / prof_steps,
                              "tflops": (fl5[i] / (ms5[i] * 1e-3) / 1e12) if ms5[i] > 0 and fl5[i] > 0 else None,
                              "gbs": (by5[i] / (ms5[i] * 1e-3) / 1e9) if ms5[i] > 0 and by5[i] > 0 else None}
                       for i, name in enumerate(("gemm", "attn_windowed", "attn_global", "norms", "other"))},
        "whole_step": {"tflops": whole_tf, "frac_of_sustained": whole_tf / pk["bf16_sustained"],
                       "frac_of_burst": whole_tf / pk["bf16_burst"], "flops_per_image": flops_img},
    }

    # ---------------- BASELINE configs[1] / [2] on this GPU: ViT-B batch 16, ViT-L batch 32 (single-GPU records) ----------------
    other_configs = None
    if not args.no_extras and world == 1 and args.model == "vit_h":
        other_configs = {}
        del x_dev
        for model, nb in (("vit_b", 16), ("vit_l", 32)):
            try:
                c2, e2 = make_encoder(model)
                xs = x_host[:nb].to(dev)
                ms2 = timed(e2, xs, max(5, args.steps), 3)
                v2 = nb / (ms2 * 1e-3)
                ee = None if args.no_e2e else timed_e2e(e2, c2, x_host[:nb], 10)
                other_configs[f"{model}_b{nb}"] = {
                    "workload": "SAM %s image-encoder forward, batch %d, bf16 (BASELINE.json configs[%d])" % (model, nb, 1 if model == "vit_b" else 2),
                    "value": v2, "unit": UNIT, "ms_per_step": ms2, "e2e": ee["value"] if ee else None,
                    "whole_encoder_tflops": v2 * c2.flops_per_image() / 1e12,
                    "whole_encoder_frac": v2 * c2.flops_per_image() / 1e12 / pk["bf16_sustained"],
                    "whole_encoder_frac_of_burst": v2 * c2.flops_per_image() / 1e12 / pk["bf16_burst"]}
                del e2, xs
            except Exception as e:  # noqa: BLE001
                other_configs[f"{model}_b{nb}"] = {"error": str(e)[:300]}

    # ---------------- the "next" rows of the scope table: kernel-alone timings, for the record ----------------
    next_rows = None
    if not args.no_extras:
        try:
            next_rows = measure_next_rows(torch, cabi, dev, pk)
        except Exception as e:  # noqa: BLE001  (never lose the headline line over an extra)
            next_rows = {"error": str(e)[:200]}
        if pipeline is not None:
            next_rows = dict(next_rows or {})
            next_rows["pipeline"] = pipeline

    cpu_baseline = None
    if not args.no_cpu_baseline:
        # the reference's own CPU encoder (or the oracle port when oracle/_ref is absent): 1 warm-up pass, median of 3 timed passes
        torch.set_num_threads(os.cpu_count() or 1)
        fwd, kind = reference_encoder(args.model, torch)
        sd = ib.make_state_dict(cfg, 1234)
        x1 = x_host[:1].clone()
        fwd(sd, x1)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            fwd(sd, x1)
            ts.append(time.perf_counter() - t0)
        dt_cpu = sorted(ts)[1]
        cpu_baseline = {"value": 1.0 / dt_cpu, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                        "sample": f"1 image of the same workload ({args.model}, fp32, torch CPU ops), 1 warm-up + median of 3 passes "
                                  f"({', '.join('%.1f' % t for t in ts)} s)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "strong": strong, "configs": other_configs,
        "next_rows": next_rows,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
