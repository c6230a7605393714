"""Drop-in replacement for the reference's ``sam.modeling.ImageEncoderViT``.

Same constructor signature (``sam/modeling/image_encoder.py:18-36``), same ``forward(x) -> {'res2','res3','res4','res5'}``
contract (``:107-120, 461-466``) and the same ``state_dict`` keys / shapes, so
``sam.build_sam._build_sam`` (``sam/build_sam.py:60-73``), ``load_state_dict(strict=False)`` (``:96-99``) and
``BaseModel.from_pretrained``'s name+shape alignment (``utils/model.py:31-57``) work unchanged.  The sub-modules
(``nn.Linear``, ``nn.LayerNorm``, ``nn.GroupNorm``, ``nn.Conv2d`` ...) exist as PARAMETER HOLDERS only — the optimizer
grouping code walks ``named_modules()`` and tests their types (``trainer/xdecoder_trainer.py:61-73``) — their own
``forward`` is never used: ``ImageEncoderViT.forward`` hands raw device pointers to the hand-written sm_100a kernels
through the C ABI in ``include/samvit_b200.h``.

Forward only.  No CPU path, no PyTorch-op fallback: a missing extension or a non-CUDA input raises.
"""
from __future__ import annotations

import ctypes as C
from functools import partial
from typing import Dict, Optional, Tuple, Type

import torch
import torch.nn as nn

from . import cabi
from .config import EncoderConfig, PRESETS

_FWD_ONLY = ("this sub-module only holds parameters for the fused CUDA encoder; "
             "call ImageEncoderViT.forward instead")


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - never used on the hot path
        raise RuntimeError(_FWD_ONLY)


class LayerNorm2d(_Holder):
    """Parameter holder for ``sam/modeling/common.py:31-43`` (only inside the never-executed ``orig_neck``)."""

    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps


class MLPBlock(_Holder):
    """``sam/modeling/common.py:13-26``: lin1 -> GELU -> lin2."""

    def __init__(self, embedding_dim: int, mlp_dim: int, act: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()


class Attention(_Holder):
    """``image_encoder.py:200-237``."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:
            assert input_size is not None, "Input size must be provided if using relative positional encoding."
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))


class Block(_Holder):
    """``image_encoder.py:134-179``."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 norm_layer: Type[nn.Module] = nn.LayerNorm, act_layer: Type[nn.Module] = nn.GELU,
                 use_rel_pos: bool = False, rel_pos_zero_init: bool = True, window_size: int = 0,
                 input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, use_rel_pos=use_rel_pos,
                              rel_pos_zero_init=rel_pos_zero_init,
                              input_size=input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(embedding_dim=dim, mlp_dim=int(dim * mlp_ratio), act=act_layer)
        self.window_size = window_size


class PatchEmbed(_Holder):
    """``image_encoder.py:379-400``."""

    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans: int = 3, embed_dim: int = 768) -> None:
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)


class SimpleFPN(_Holder):
    """``image_encoder.py:413-447``: module indexes inside each ``nn.Sequential`` are part of the state_dict keys."""

    def __init__(self, in_dim=768, out_dims=(128, 256, 512, 1024)):
        super().__init__()
        self.down_4_chan = max(out_dims[0] * 2, in_dim // 2)
        self.down_4 = nn.Sequential(
            nn.ConvTranspose2d(in_dim, self.down_4_chan, 2, stride=2), nn.GroupNorm(1, self.down_4_chan), nn.GELU(),
            nn.ConvTranspose2d(self.down_4_chan, self.down_4_chan // 2, 2, stride=2), nn.GroupNorm(1, self.down_4_chan // 2),
            nn.Conv2d(self.down_4_chan // 2, out_dims[0], 1), nn.GroupNorm(1, out_dims[0]), nn.GELU())
        self.down_8_chan = max(out_dims[1], in_dim // 2)
        self.down_8 = nn.Sequential(
            nn.ConvTranspose2d(in_dim, self.down_8_chan, 2, stride=2), nn.GroupNorm(1, self.down_8_chan),
            nn.Conv2d(self.down_8_chan, out_dims[1], 1), nn.GroupNorm(1, out_dims[1]), nn.GELU())
        self.down_16 = nn.Sequential(nn.Conv2d(in_dim, out_dims[2], 1), nn.GroupNorm(1, out_dims[2]), nn.GELU())
        self.down_32_chan = max(out_dims[3], in_dim * 2)
        self.down_32 = nn.Sequential(
            nn.Conv2d(in_dim, self.down_32_chan, 2, stride=2), nn.GroupNorm(1, self.down_32_chan),
            nn.Conv2d(self.down_32_chan, out_dims[3], 1), nn.GroupNorm(1, out_dims[3]), nn.GELU())


class ImageEncoderViT(nn.Module):
    """B200-native SAM ViT image encoder (forward only).  See module docstring."""

    def __init__(
        self,
        img_size: int = 1024,
        patch_size: int = 16,
        in_chans: int = 3,
        embed_dim: int = 768,
        depth: int = 12,
        num_heads: int = 12,
        mlp_ratio: float = 4.0,
        out_chans: int = 256,
        qkv_bias: bool = True,
        norm_layer: Type[nn.Module] = nn.LayerNorm,
        act_layer: Type[nn.Module] = nn.GELU,
        use_abs_pos: bool = True,
        use_rel_pos: bool = False,
        rel_pos_zero_init: bool = True,
        window_size: int = 0,
        global_attn_indexes: Tuple[int, ...] = (),
    ) -> None:
        super().__init__()
        # The kernels implement the configuration family _build_sam instantiates (build_sam.py:60-73).
        if not (use_abs_pos and use_rel_pos and qkv_bias):
            raise NotImplementedError("the B200 encoder implements use_abs_pos=True, use_rel_pos=True, qkv_bias=True "
                                      "(the only configuration sam/build_sam.py constructs)")
        if act_layer is not nn.GELU:
            raise NotImplementedError("only act_layer=nn.GELU (exact erf form) is implemented")
        if window_size <= 0:
            raise NotImplementedError("window_size must be > 0 (build_sam.py uses 14)")
        self.img_size = img_size            # read by sam/utils/onnx.py:35
        self.patch_embed = PatchEmbed(kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size),
                                      in_chans=in_chans, embed_dim=embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, img_size // patch_size, img_size // patch_size, embed_dim))
        self.blocks = nn.ModuleList()
        g = img_size // patch_size
        for i in range(depth):
            self.blocks.append(Block(
                dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, norm_layer=norm_layer,
                act_layer=act_layer, use_rel_pos=use_rel_pos, rel_pos_zero_init=rel_pos_zero_init,
                window_size=window_size if i not in global_attn_indexes else 0, input_size=(g, g)))
        if not isinstance(self.blocks[0].norm1, nn.LayerNorm):
            raise NotImplementedError("norm_layer must build an nn.LayerNorm")
        self.orig_neck = nn.Sequential(
            nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False), LayerNorm2d(out_chans),
            nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False), LayerNorm2d(out_chans))
        self.neck = SimpleFPN(in_dim=embed_dim, out_dims=[128, 256, 512, 1024])

        self.cfg = EncoderConfig(
            embed_dim=embed_dim, depth=depth, num_heads=num_heads,
            global_attn_indexes=tuple(int(i) for i in global_attn_indexes), img_size=img_size, patch_size=patch_size,
            in_chans=in_chans, mlp_ratio=mlp_ratio, out_chans=out_chans, window_size=window_size,
            ln_eps=float(self.blocks[0].norm1.eps), gn_eps=float(self.neck.down_16[1].eps))
        # runtime knobs (not part of the reference signature)
        self.precision = "bf16"             # "bf16" (tcgen05) or "fp32" (validation mode)
        self.out_dtype = torch.float32      # dtype of the returned embeddings (float32 or bfloat16)
        self.max_chunk = 16                 # most images per pass through the kernels (bounds the workspace, ~0.3 GB per image)
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[torch.device] = None
        self._weights_sig = None
        self._plist = None
        self._ws: Dict[Tuple[int, int], torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                cabi.lib().svb_encoder_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _mode(self) -> int:
        if self.precision == "bf16":
            return cabi.MODE_BF16
        if self.precision == "fp32":
            return cabi.MODE_FP32
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision!r}")

    def _out_code(self) -> int:
        if self.out_dtype == torch.float32:
            return cabi.DTYPE_F32
        if self.out_dtype == torch.bfloat16:
            return cabi.DTYPE_BF16
        raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")

    def _ensure_handle(self, device: torch.device) -> None:
        lib = cabi.lib()
        if self._handle is not None and self._handle_device == device:
            return
        if self._handle is not None:
            lib.svb_encoder_destroy(self._handle)
            self._handle = None
        cfg = self.cfg
        c = cabi.SvbConfig()
        c.img_size, c.patch_size, c.in_chans = cfg.img_size, cfg.patch_size, cfg.in_chans
        c.embed_dim, c.depth, c.num_heads, c.mlp_dim = cfg.embed_dim, cfg.depth, cfg.num_heads, cfg.mlp_dim
        c.window_size = cfg.window_size
        c.num_global = len(cfg.global_attn_indexes)
        if c.num_global > 16:
            raise NotImplementedError("at most 16 global-attention blocks")
        for i, v in enumerate(cfg.global_attn_indexes):
            c.global_idx[i] = v
        for i, v in enumerate(cfg.fpn_dims):
            c.fpn_dims[i] = v
        c.ln_eps, c.gn_eps = cfg.ln_eps, cfg.gn_eps
        h = C.c_void_p()
        with torch.cuda.device(device):
            cabi.check(lib.svb_encoder_create(C.byref(c), C.byref(h)), "svb_encoder_create")
        self._handle, self._handle_device, self._weights_sig = h, device, None
        self._ws.clear()

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() replace the parameter storage without touching the version counters
        self._weights_sig = None
        self._plist = None
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        # also reached when a parent module (Sam, the X-Decoder backbone) loads a checkpoint; covers load_state_dict(assign=True)
        self._weights_sig = None
        self._plist = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _sync_weights(self, device: torch.device) -> None:
        """(Re)pack the parameters into the native encoder whenever they changed.  Per forward this is ONE pass over the cached
        parameter list reading the autograd version counters (in-place updates — load_state_dict's copy_, optimizer steps — bump
        them; storage swaps go through _apply above): ~30 us for ViT-H's 489 entries, no state_dict() walk."""
        plist = getattr(self, "_plist", None)
        if plist is None:
            plist = self._plist = [v for _, v in self.state_dict(keep_vars=True).items()]
            self._weights_sig = None
        sig = [v._version for v in plist]
        if sig == self._weights_sig:
            return
        sd = self.state_dict(keep_vars=True)
        self._plist = [v for _, v in sd.items()]
        lib = cabi.lib()
        stream = cabi.stream_ptr()
        keep = []
        for k, v in sd.items():
            t = v.detach()
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=device, dtype=torch.float32).contiguous()
                keep.append(t)
            cabi.check(lib.svb_encoder_load_param(self._handle, k.encode(), t.data_ptr(), t.numel(), stream),
                       f"load_param({k})")
        torch.cuda.current_stream().synchronize()
        self._weights_sig = [v._version for v in self._plist]

    def _workspace(self, chunk: int, mode: int, device: torch.device, hw: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        hw = hw or (self.img_size, self.img_size)
        key = (chunk, mode, hw)
        ws = self._ws.get(key)
        if ws is None or ws.device != device:
            n = cabi.lib().svb_encoder_workspace_bytes_hw(self._handle, chunk, mode, hw[0], hw[1])
            self._ws.clear()
            ws = torch.empty(n + 1024, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def _check_input(self, x: torch.Tensor, square_only: bool = False) -> None:
        if x.dim() != 4 or x.shape[1] != self.cfg.in_chans:
            raise ValueError(f"expected (B,{self.cfg.in_chans},H,W), got {tuple(x.shape)}")
        H, W = int(x.shape[2]), int(x.shape[3])
        if H == self.img_size and W == self.img_size:
            return
        unit = 32 * self.cfg.patch_size
        if square_only or H % unit or W % unit or H <= 0 or W <= 0:
            raise NotImplementedError(
                f"input {(H, W)}: implemented are {self.img_size}x{self.img_size} and, through forward(), canvases whose sides are "
                f"multiples of {unit} (the reference's bicubic pos_embed / linear rel_pos resize fallback, "
                "image_encoder.py:111-114,319-330)")

    def _prepare(self, device: torch.device) -> None:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("ImageEncoderViT (B200) implements the forward pass only: call it under torch.no_grad() "
                               "(sam/build_sam.py:101-105 marks the encoder parameters requires_grad=True)")
        self._ensure_handle(device)
        self._sync_weights(device)

    # ------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        if not x.is_cuda:
            raise RuntimeError("ImageEncoderViT (B200) has no CPU path: the input must be a CUDA tensor")
        self._check_input(x)
        device = x.device
        with torch.cuda.device(device):
            self._prepare(device)
            xf = x.detach()
            # fp32 / fp16 / bf16 inputs are read as they are (the reference's pipeline feeds fp16 images, pipeline/XDecoderPipeline.py:
            # 93-95): the cast to the GEMM operand type happens inside the patch-embedding loader
            xcode = {torch.float32: cabi.DTYPE_F32, torch.float16: cabi.DTYPE_F16, torch.bfloat16: cabi.DTYPE_BF16}.get(xf.dtype)
            if xcode is None:
                raise TypeError(f"ImageEncoderViT (B200): input dtype {xf.dtype} is not float32 / float16 / bfloat16")
            if not xf.is_contiguous():
                xf = xf.contiguous()
            B, H, W = xf.shape[0], int(xf.shape[2]), int(xf.shape[3])
            od = self.cfg.fpn_dims
            outs = [torch.empty(B, od[k], H // s, W // s, dtype=self.out_dtype, device=device)
                    for k, s in enumerate((4, 8, 16, 32))]
            if B == 0:
                return dict(zip(("res2", "res3", "res4", "res5"), outs))
            chunk = max(1, min(int(self.max_chunk), B))
            mode = self._mode()
            native = (H == self.img_size and W == self.img_size)
            if not native:
                # other token grids: the reference's pos_embed / rel_pos resize fallbacks (scope row N3)
                chunk = max(1, min(chunk, max(1, (4 * self.img_size * self.img_size) // (H * W))))
            ws = self._workspace(chunk, mode, device, None if native else (H, W))
            base = (ws.data_ptr() + 1023) & ~1023
            if native and xcode == cabi.DTYPE_F32:
                cabi.check(cabi.lib().svb_encoder_forward(
                    self._handle, xf.data_ptr(), B, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                    outs[3].data_ptr(), self._out_code(), mode, chunk, base, ws.numel() - (base - ws.data_ptr()),
                    cabi.stream_ptr()), "svb_encoder_forward")
            else:
                cabi.check(cabi.lib().svb_encoder_forward_x(
                    self._handle, xf.data_ptr(), xcode, B, H, W, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                    outs[3].data_ptr(), self._out_code(), mode, chunk, base, ws.numel() - (base - ws.data_ptr()),
                    cabi.stream_ptr()), "svb_encoder_forward_x")
        return {"res2": outs[0], "res3": outs[1], "res4": outs[2], "res5": outs[3]}

    def forward_uint8(self, images, pixel_mean, pixel_std) -> Dict[str, torch.Tensor]:
        """The forward fed with the callers' raw inputs (scope row N2): ``images`` is a list of uint8 CUDA tensors (C,h,w) with
        h,w <= img_size; equivalent to ``self(ImageList.from_tensors([(x - pixel_mean) / pixel_std for x in images], 1024).tensor)``
        (modeling/architectures/xdecoder_model.py:481-484) with the normalisation and the zero padding done inside the
        patch-embedding loader."""
        if len(images) == 0:
            raise ValueError("forward_uint8 needs at least one image")
        device = images[0].device
        keep = []
        for t in images:
            if not t.is_cuda or t.dtype != torch.uint8 or t.dim() != 3 or t.shape[0] != self.cfg.in_chans:
                raise ValueError(f"forward_uint8 takes uint8 CUDA tensors of shape ({self.cfg.in_chans},h,w)")
            if t.shape[1] > self.img_size or t.shape[2] > self.img_size:
                raise NotImplementedError(f"image {tuple(t.shape[1:])} exceeds the {self.img_size}x{self.img_size} canvas "
                                          "(larger token grids are a 'next' row)")
            keep.append(t.contiguous())
        B = len(keep)
        mean = [float(v) for v in torch.as_tensor(pixel_mean).flatten().tolist()]
        std = [float(v) for v in torch.as_tensor(pixel_std).flatten().tolist()]
        if len(mean) != self.cfg.in_chans or len(std) != self.cfg.in_chans:
            raise ValueError("pixel_mean / pixel_std must have one entry per channel")
        with torch.cuda.device(device):
            self._prepare(device)
            S, od = self.img_size, self.cfg.fpn_dims
            outs = [torch.empty(B, od[k], S // s, S // s, dtype=self.out_dtype, device=device) for k, s in enumerate((4, 8, 16, 32))]
            chunk = max(1, min(int(self.max_chunk), B))
            mode = self._mode()
            ws = self._workspace(chunk, mode, device)
            base = (ws.data_ptr() + 1023) & ~1023
            ptrs = (C.c_void_p * B)(*[t.data_ptr() for t in keep])
            hs = (C.c_int * B)(*[int(t.shape[1]) for t in keep])
            wds = (C.c_int * B)(*[int(t.shape[2]) for t in keep])
            cm = (C.c_float * len(mean))(*mean)
            cs = (C.c_float * len(std))(*std)
            cabi.check(cabi.lib().svb_encoder_forward_u8(
                self._handle, ptrs, hs, wds, cm, cs, B, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                outs[3].data_ptr(), self._out_code(), mode, chunk, base, ws.numel() - (base - ws.data_ptr()),
                cabi.stream_ptr()), "svb_encoder_forward_u8")
        return {"res2": outs[0], "res3": outs[1], "res4": outs[2], "res5": outs[3]}

    def pass_schedule(self, batch: int, host_path: bool = False):
        """Images per pass through the kernels for a batch of ``batch`` images (``max_chunk`` as set): what ``forward`` (device tensors) or
        ``forward_host`` (host tensors: the first upload / last download are charged too) will run."""
        import ctypes
        device = next(self.parameters()).device
        with torch.cuda.device(device):
            self._ensure_handle(device)
            buf = (ctypes.c_int * 256)()
            odt = cabi.DTYPE_BF16 if self.out_dtype == torch.bfloat16 else cabi.DTYPE_F32
            rc = cabi.lib().svb_encoder_pass_schedule(self._handle, int(batch), int(self.max_chunk), int(bool(host_path)), odt, buf, 256)
            if rc < 1000:
                cabi.check(rc, "svb_encoder_pass_schedule")
            return [int(buf[i]) for i in range(rc - 1000)]

    def forward_host(self, x: torch.Tensor, out: Optional[Dict[str, torch.Tensor]] = None,
                     device: Optional[torch.device] = None) -> Dict[str, torch.Tensor]:
        """End-to-end call with HOST tensors: ``x`` (B,3,S,S) fp32 on the CPU (pinned for full PCIe speed); returns pinned
        CPU tensors.  Host->device and device->host copies are double-buffered against the compute inside the native
        library (``svb_encoder_forward_host``).  Synchronous."""
        if x.is_cuda:
            raise ValueError("forward_host takes a CPU tensor")
        self._check_input(x, square_only=True)
        device = device or next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("the encoder parameters must live on a CUDA device")
        with torch.cuda.device(device):
            self._prepare(device)
            xf = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
            B = xf.shape[0]
            S, od = self.img_size, self.cfg.fpn_dims
            if out is None:
                out = {f"res{k + 2}": torch.empty(B, od[k], S // s, S // s, dtype=self.out_dtype).pin_memory()
                       for k, s in enumerate((4, 8, 16, 32))}
            chunk = max(1, min(int(self.max_chunk), B))
            cabi.check(cabi.lib().svb_encoder_forward_host(
                self._handle, xf.data_ptr(), B, out["res2"].data_ptr(), out["res3"].data_ptr(), out["res4"].data_ptr(),
                out["res5"].data_ptr(), self._out_code(), self._mode(), chunk), "svb_encoder_forward_host")
        return out

    # ---- test hooks ----
    def enable_taps(self, enable: bool = True) -> None:
        device = next(self.parameters()).device
        with torch.cuda.device(device):
            self._ensure_handle(device)
            cabi.check(cabi.lib().svb_encoder_enable_taps(self._handle, int(enable)), "enable_taps")

    def read_tap(self, block: int) -> torch.Tensor:
        """Token stream (1, g, g, D) of the first image of the last chunk: ``block=-1`` after patch-embed + pos."""
        g, D = self.cfg.grid, self.cfg.embed_dim
        t = torch.empty(1, g, g, D, dtype=torch.float32, device=self._handle_device)
        with torch.cuda.device(self._handle_device):
            cabi.check(cabi.lib().svb_encoder_read_tap(self._handle, block, t.data_ptr(), t.numel(), cabi.stream_ptr()),
                       "read_tap")
        return t


# ----------------------------------------------------------------------------------------------
def build_encoder(cfg: EncoderConfig) -> ImageEncoderViT:
    """Construct the encoder with the keyword arguments ``_build_sam`` uses (``sam/build_sam.py:60-73``)."""
    enc = ImageEncoderViT(
        depth=cfg.depth, embed_dim=cfg.embed_dim, img_size=cfg.img_size, mlp_ratio=cfg.mlp_ratio,
        norm_layer=partial(torch.nn.LayerNorm, eps=cfg.ln_eps), num_heads=cfg.num_heads, patch_size=cfg.patch_size,
        qkv_bias=True, use_rel_pos=True, global_attn_indexes=list(cfg.global_attn_indexes),
        window_size=cfg.window_size, out_chans=cfg.out_chans)
    return enc.eval()


sam_encoder_registry = {name: (lambda n=name: build_encoder(PRESETS[n])) for name in ("vit_h", "vit_l", "vit_b")}


def install_into_reference(sam_package=None) -> None:
    """Make the reference's factories build THIS encoder: rebinds ``ImageEncoderViT`` in ``sam.modeling`` and
    ``sam.build_sam`` (``sam/modeling/__init__.py:8``, ``sam/build_sam.py:11``) so ``sam_model_registry[...]`` /
    ``GeneralizedXdecoder.__init__`` (``xdecoder_model.py:72-87``) pick it up with no source change."""
    import importlib
    sam_modeling = importlib.import_module("sam.modeling")
    sam_build = importlib.import_module("sam.build_sam")
    sam_ie = importlib.import_module("sam.modeling.image_encoder")
    for m in (sam_modeling, sam_build, sam_ie):
        setattr(m, "ImageEncoderViT", ImageEncoderViT)
