"""ctypes binding of libs3od_b200.so and the device-side model object.

`B200DPTSegmentation` plays the role of the reference's `DPTSegmentation` nn.Module behind
`BackgroundRemoval.model` (/root/reference/src/s3od/model.py:89-106): `model(x)` takes a float (B,3,S,S) tensor and
returns `{'pred_masks', 'pred_iou'}`.  PyTorch is used only for device memory and streams; every kernel is in the
CUDA library.  There is no CPU or eager fallback: if the library cannot be loaded, or no sm_100 GPU is present, this
module raises.
"""
import ctypes
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import geometry
from .arch import ARCHS, ArchSpec
from .utils import check_padding, get_pad_info
from .weights import pack_weights

_LIB = None


class S3odImage(ctypes.Structure):
    _fields_ = [("d_src", ctypes.c_void_p), ("h", ctypes.c_int32), ("w", ctypes.c_int32), ("new_h", ctypes.c_int32),
                ("new_w", ctypes.c_int32), ("pad_h", ctypes.c_int32), ("pad_w", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("d_xtab", ctypes.c_void_p), ("d_ytab", ctypes.c_void_p)]


class S3odPost(ctypes.Structure):
    _fields_ = [("d_src", ctypes.c_void_p), ("d_all_masks", ctypes.c_void_p), ("d_rgba", ctypes.c_void_p),
                ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("pad_h", ctypes.c_int32), ("pad_w", ctypes.c_int32),
                ("ky", ctypes.c_int32), ("kx", ctypes.c_int32), ("d_ystart", ctypes.c_void_p), ("d_yw", ctypes.c_void_p),
                ("d_xstart", ctypes.c_void_p), ("d_xw", ctypes.c_void_p)]


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libs3od_b200.so")


def load_library() -> ctypes.CDLL:
    """Load the C-ABI library; raises (no fallback) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m s3od_b200.build` (there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.s3od_last_error.restype = ctypes.c_char_p
    lib.s3od_version.restype = ctypes.c_char_p
    lib.s3od_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci, ci]
    lib.s3od_set_tensor.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_size_t]
    lib.s3od_finalize.argtypes = [vp]
    lib.s3od_preprocess_u8.argtypes = [vp, ctypes.POINTER(S3odImage), ci, vp]
    lib.s3od_pack_input_f32.argtypes = [vp, vp, ci, vp]
    lib.s3od_forward.argtypes = [vp, ci, vp, vp, vp]
    lib.s3od_postprocess.argtypes = [vp, vp, vp, ctypes.POINTER(S3odPost), ci, vp, vp, vp]
    lib.s3od_get_stage.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t)]
    lib.s3od_read_stage.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_size_t, vp]
    lib.s3od_profile_enable.argtypes = [vp, ci]
    lib.s3od_profile_read.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
    lib.s3od_launch_count.argtypes = [vp]
    lib.s3od_launch_count.restype = ctypes.c_longlong
    lib.s3od_preprocess_mode.argtypes = [vp]
    lib.s3od_destroy.argtypes = [vp]
    lib.s3od_destroy.restype = None
    lib.s3od_op_gemm_f32.argtypes = [vp, vp, vp, ci, ci, ci, vp]
    lib.s3od_op_layernorm.argtypes = [vp, vp, vp, vp, ci, ci, cf, vp]
    lib.s3od_op_attention.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp]
    lib.s3od_op_conv3x3.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
    lib.s3od_metrics_stats.argtypes = [vp, vp, ci, ci, vp, vp, ctypes.c_size_t, vp]
    lib.s3od_metrics_region.argtypes = [vp, vp, ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
    lib.s3od_metrics_stats_bytes.restype = ctypes.c_size_t
    lib.s3od_metrics_region_bytes.restype = ctypes.c_size_t
    lib.s3od_threshold_f32.argtypes = [vp, vp, ctypes.c_size_t, cf, vp]
    lib.s3od_vis_composite.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, vp]
    lib.s3od_vis_mask_grid.argtypes = [vp, vp, ci, vp, ci, ci, vp]
    lib.s3od_mask_pair_counts.argtypes = [vp, ci, ci, ci, vp, vp]
    lib.s3od_op_convt_rows.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.s3od_op_conv3x3_rows.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
    _LIB = lib
    return lib


def _check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.s3od_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class B200DPTSegmentation:
    """Device model: packed weights + launch plan inside the CUDA library."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], arch: ArchSpec, image_size: int = 1024, device: str = "cuda:0",
                 max_batch: int = 1, micro_batch: Optional[int] = None, normalisation: str = "s3od"):
        if not torch.cuda.is_available():
            raise RuntimeError("s3od_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.arch = arch
        self.image_size = int(image_size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("s3od_b200 runs on CUDA devices only")
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.dev_index)
        self.max_batch = int(max_batch)
        self.micro_batch = int(micro_batch or min(self.max_batch, 8))
        self.K = arch.num_outputs
        self._ctx = ctypes.c_void_p()
        arch_id = 0 if arch.hidden == 768 else 1
        with torch.cuda.device(self.dev_index):
            _check(self.lib, self.lib.s3od_create(ctypes.byref(self._ctx), self.dev_index, arch_id, self.K, self.image_size,
                                                  self.max_batch, self.micro_batch), "s3od_create")
            for name, t in pack_weights(state_dict, arch, self.image_size, normalisation).items():
                t = t.contiguous()
                _check(self.lib, self.lib.s3od_set_tensor(self._ctx, name.encode(), t.data_ptr(), t.numel() * t.element_size()),
                       f"s3od_set_tensor({name})")
            _check(self.lib, self.lib.s3od_finalize(self._ctx), "s3od_finalize")
        # Host-side caches are BOUNDED (a long-lived predictor fed images of every size - the reference's Gradio demo - must
        # not grow without limit; the reference frees everything after each call):
        #   _tab_cache   resize / antialias coefficient tables per geometry, least-recently-used, at most `max_tables`;
        #   _fixed       per-slot buffers whose shape only depends on (max_batch, K, S);
        #   _arena       ONE flat byte buffer per output slot, sized to the largest request so far (+25 % on growth) and
        #                carved into the per-image (K,H,W) fp32 / (H,W,4) u8 views of each call.
        self.max_tables = 64
        self._tab_cache: "OrderedDict[Tuple, Tuple]" = OrderedDict()
        self._fixed: Dict[Tuple, torch.Tensor] = {}
        self._arena: Dict[int, torch.Tensor] = {}
        self._arena_off: Dict[int, int] = {}
        # streams other than the launch stream that read slot buffers (the predictor's copy-out stream): a retired arena is
        # handed back to the allocator only after their queued work
        self.aux_streams: List[torch.cuda.Stream] = []

    # -- nn.Module-ish surface used by BackgroundRemoval ----------------------------------------------------------
    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("s3od_b200 model cannot be moved off the GPU (no CPU fallback)")
        return self

    def eval(self):
        return self

    def __call__(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        return self.forward(x)

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The reference's inner seam: x float (B,3,S,S) on this device -> logits (model.py:99-106)."""
        S = self.image_size
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != S or x.shape[3] != S:
            raise ValueError(f"expected input of shape (B,3,{S},{S}), got {tuple(x.shape)}")
        x = x.to(self.device, torch.float32).contiguous()
        B = x.shape[0]
        with torch.cuda.device(self.dev_index):
            st = _stream_ptr(self.device)
            _check(self.lib, self.lib.s3od_pack_input_f32(self._ctx, x.data_ptr(), B, st), "s3od_pack_input_f32")
            return self._forward_staged(B)

    def _buffer(self, slot, name: str, shape, dtype) -> torch.Tensor:
        """Reusable device buffer of a batch-independent shape.  slot=None allocates a fresh tensor; an integer slot returns
        the same storage on every call - the steady-state hot path then makes no allocator calls at all."""
        if slot is None:
            return torch.empty(shape, dtype=dtype, device=self.device)
        key = (slot, name, tuple(shape), dtype)
        t = self._fixed.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._fixed[key] = t
        return t

    def _arena_reserve(self, slot: int, nbytes: int) -> None:
        """Start carving slot `slot`'s arena for a call that needs `nbytes` in total; grows (and retires the old buffer) if short."""
        cur = self._arena.get(slot)
        if cur is None or cur.numel() < nbytes:
            new = torch.empty(max(nbytes, (cur.numel() * 5 // 4) if cur is not None else 0), dtype=torch.uint8, device=self.device)
            if cur is not None:
                for st in self.aux_streams:
                    cur.record_stream(st)
            self._arena[slot] = new
        self._arena_off[slot] = 0

    def _arena_take(self, slot: int, shape, dtype) -> torch.Tensor:
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = self._arena_off[slot]
        self._arena_off[slot] = off + ((n + 255) & ~255)
        return self._arena[slot][off:off + n].view(dtype).view(shape)

    def cache_bytes(self) -> int:
        """Device bytes held by the reusable output buffers (bounded: 3 slots x the largest call so far)."""
        return sum(t.numel() * t.element_size() for t in self._fixed.values()) + sum(t.numel() for t in self._arena.values())

    def _forward_staged(self, B: int, slot=None) -> Dict[str, torch.Tensor]:
        S, K = self.image_size, self.K
        masks = self._buffer(slot, "mask_logits", (self.max_batch, K, S, S), torch.float32)[:B]
        ious = self._buffer(slot, "iou_logits", (self.max_batch, K), torch.float32)[:B]
        _check(self.lib, self.lib.s3od_forward(self._ctx, B, masks.data_ptr(), ious.data_ptr(), _stream_ptr(self.device)),
               "s3od_forward")
        return {"pred_masks": masks, "pred_iou": ious}

    # -- fused uint8 path -----------------------------------------------------------------------------------------
    def _tables(self, key, builder):
        t = self._tab_cache.get(key)
        if t is None:
            t = builder()
            self._tab_cache[key] = t
            while len(self._tab_cache) > self.max_tables:       # LRU; evicted tables were allocated on the launch stream, so
                self._tab_cache.popitem(last=False)             # the allocator reuses them in stream order
        else:
            self._tab_cache.move_to_end(key)
        return t

    def geometry(self, h: int, w: int):
        """pad_info + preprocess mode / tables for a source of size (h, w)."""
        S = self.image_size
        pad = get_pad_info(np.empty((h, w, 0), np.uint8), S)
        check_padding(pad, S)
        return pad

    def preprocess(self, d_images: Sequence[torch.Tensor], placements: Optional[Sequence[dict]] = None) -> List[dict]:
        """_preprocess (predictor.py:79-94) for uint8 (H,W,3) tensors already on the device; stages the model input.
        `placements` (same keys as pad_info: resized_size, height_pad = top, width_pad = left) overrides the reference
        letterbox geometry - SODPredictor's albumentations front end rounds and pads differently; pixels that fall outside
        the image_size x image_size canvas are dropped."""
        B = len(d_images)
        descs = (S3odImage * B)()
        pads = []
        for i, img in enumerate(d_images):
            h, w = int(img.shape[0]), int(img.shape[1])
            pad = self.geometry(h, w) if placements is None else placements[i]
            new_h, new_w = pad["resized_size"]
            mode = geometry.resize_mode(h, w, new_h, new_w)
            xt = yt = None
            if mode == 2:
                xt, yt = self._tables(("lin", h, w, new_h, new_w), lambda: (
                    torch.from_numpy(geometry.linear_tables(new_w, w, False)).to(self.device),
                    torch.from_numpy(geometry.linear_tables(new_h, h, True)).to(self.device)))
            descs[i] = S3odImage(img.data_ptr(), h, w, new_h, new_w, pad["height_pad"], pad["width_pad"], mode,
                                 xt.data_ptr() if xt is not None else None, yt.data_ptr() if yt is not None else None)
            pads.append(pad)
        with torch.cuda.device(self.dev_index):
            _check(self.lib, self.lib.s3od_preprocess_u8(self._ctx, descs, B, _stream_ptr(self.device)), "s3od_preprocess_u8")
        return pads

    def postprocess(self, masks: torch.Tensor, iou_logits: torch.Tensor, d_images: Sequence[torch.Tensor], pads: List[dict],
                    slot=None):
        """Tail of remove_background (predictor.py:113-132) on the device.  Returns per-image (all_masks, rgba) device
        tensors plus (B,K) ious and (B,) best indices.  With an integer `slot` the outputs live in reusable buffers that
        stay valid until the next call with the same slot."""
        B = len(d_images)
        S, K = self.image_size, self.K
        descs = (S3odPost * B)()
        outs = []
        if slot is not None:
            self._arena_reserve(slot, sum(((K * p["original_size"][0] * p["original_size"][1] * 4 + 255) & ~255) +
                                          ((p["original_size"][0] * p["original_size"][1] * 4 + 255) & ~255) for p in pads))
        for i, (img, pad) in enumerate(zip(d_images, pads)):
            H, W = pad["original_size"]
            hp, wp = pad["height_pad"], pad["width_pad"]
            ch, cw = S - 2 * hp, S - 2 * wp
            ys, yw, xs, xw = self._tables(("aa", ch, cw, H, W), lambda: tuple(
                torch.from_numpy(a).to(self.device) for a in (geometry.aa_tables(ch, H) + geometry.aa_tables(cw, W))))
            if slot is None:
                all_masks = torch.empty((K, H, W), dtype=torch.float32, device=self.device)
                rgba = torch.empty((H, W, 4), dtype=torch.uint8, device=self.device)
            else:
                all_masks = self._arena_take(slot, (K, H, W), torch.float32)
                rgba = self._arena_take(slot, (H, W, 4), torch.uint8)
            descs[i] = S3odPost(img.data_ptr(), all_masks.data_ptr(), rgba.data_ptr(), H, W, hp, wp, yw.shape[1], xw.shape[1],
                                ys.data_ptr(), yw.data_ptr(), xs.data_ptr(), xw.data_ptr())
            outs.append((all_masks, rgba))
        ious = self._buffer(slot, "ious", (self.max_batch, K), torch.float32)[:B]
        best = self._buffer(slot, "best", (self.max_batch,), torch.int32)[:B]
        with torch.cuda.device(self.dev_index):
            _check(self.lib, self.lib.s3od_postprocess(self._ctx, masks.data_ptr(), iou_logits.data_ptr(), descs, B,
                                                       ious.data_ptr(), best.data_ptr(), _stream_ptr(self.device)),
                   "s3od_postprocess")
        return outs, ious, best

    def run_u8(self, d_images: Sequence[torch.Tensor], slot=None):
        """preprocess -> forward -> postprocess for device-resident uint8 images (the device-timed hot path)."""
        pads = self.preprocess(d_images)
        with torch.cuda.device(self.dev_index):
            out = self._forward_staged(len(d_images), slot)
        outs, ious, best = self.postprocess(out["pred_masks"], out["pred_iou"], d_images, pads, slot)
        return out, outs, ious, best

    def stage(self, name: str, dtype: torch.dtype, shape: Tuple[int, ...]) -> torch.Tensor:
        """Copy of an internal activation of the last micro-batch (stage-wise parity tests)."""
        out = torch.empty(shape, dtype=dtype, device=self.device)
        with torch.cuda.device(self.dev_index):
            _check(self.lib, self.lib.s3od_read_stage(self._ctx, name.encode(), out.data_ptr(), out.numel() * out.element_size(),
                                                      _stream_ptr(self.device)), "s3od_read_stage")
        return out

    def profile_enable(self, on: bool = True):
        _check(self.lib, self.lib.s3od_profile_enable(self._ctx, 1 if on else 0), "s3od_profile_enable")

    def profile_read(self):
        """[(label, launches, images, total_ms)] for every kernel of the launch plan since the last read."""
        buf = ctypes.create_string_buffer(1 << 16)
        _check(self.lib, self.lib.s3od_profile_read(self._ctx, buf, len(buf)), "s3od_profile_read")
        rows = []
        for line in buf.value.decode().splitlines():
            label, n, imgs, ms = line.split("\t")
            rows.append((label, int(n), int(imgs), float(ms)))
        return rows

    def threshold(self, soft: torch.Tensor, threshold: float) -> torch.Tensor:
        """(soft > threshold) as float32 on the device (SODPredictor.predict, synth_sod/.../predictor.py:461-470)."""
        soft = soft.contiguous()
        out = torch.empty_like(soft)
        with torch.cuda.device(self.dev_index):
            _check(self.lib, self.lib.s3od_threshold_f32(soft.data_ptr(), out.data_ptr(), soft.numel(), float(threshold),
                                                         _stream_ptr(self.device)), "s3od_threshold_f32")
        return out

    def preprocess_mode(self) -> int:
        """1 = normalisation as one FMA (verified against the table by the library), 0 = table look-up."""
        return int(self.lib.s3od_preprocess_mode(self._ctx))

    def launch_count(self) -> int:
        return int(self.lib.s3od_launch_count(self._ctx))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.s3od_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001 - interpreter shutdown
            pass


def arch_from_name(name: str) -> ArchSpec:
    if name not in ARCHS:
        raise ValueError(f"unknown encoder {name!r}; known: {sorted(ARCHS)}")
    return ARCHS[name]
