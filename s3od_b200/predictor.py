"""Drop-in for `s3od.BackgroundRemoval` (/root/reference/src/s3od/predictor.py:16-139) on the B200-native path.

Same constructor, class attributes, `from_pretrained`, `remove_background` signature and `RemovalResult` fields as the
reference; errors are the same types under the same conditions (ValueError for an unloadable model id, ValueError for
the odd-padding inputs the reference cannot paste).  Additive surface: `remove_background_batch`, and
`encoder_name` / `num_outputs` / `max_batch` / `micro_batch` keyword arguments.

Everything between the uint8 source image and the result arrays runs in the CUDA library: letterbox resize + normalise,
the DINOv3 ViT + DPT head forward, sigmoid / crop / antialiased resize / argmax / RGBA composite.
"""
import threading
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
from PIL import Image

from .arch import ARCHS
from .engine import B200DPTSegmentation


@dataclass
class RemovalResult:
    predicted_mask: np.ndarray
    all_masks: np.ndarray
    all_ious: np.ndarray
    rgba_image: Image.Image


def chunk_schedule(n: int, step: int) -> List[tuple]:
    """[(start, end)] of the device chunks of an n-image batch: full micro-batches, the last one halved down to 4 images.
    The device-to-host copy of a chunk overlaps the compute of the next one, so only the LAST chunk's copy is exposed;
    ending on small chunks keeps that tail short (67 MB of results per 2048^2 image).  One of the small chunks leads, so
    that the exposed upload in front of the first compute is short too."""
    sizes = [step] * (n // step)
    if n % step:
        sizes.append(n % step)
    if sizes:
        last = sizes.pop()
        while last > 4:
            half = last // 2
            sizes.append(half)
            last -= half
        sizes.append(last)
        # the compute of the first chunk cannot start before its images are on the device: lead with a small chunk
        # (16, 8, 4, 4 -> 8, 16, 4, 4) so that only its short upload is exposed
        if len(sizes) >= 3 and sizes[-3] < sizes[0]:
            sizes.insert(0, sizes.pop(-3))
    bounds, s0 = [], 0
    for sz in sizes:
        bounds.append((s0, s0 + sz))
        s0 += sz
    return bounds


class BackgroundRemoval:
    DEFAULT_MODEL_ID = "okupyn/s3od"
    DEFAULT_CHECKPOINT_NAME = "s3od.pt"

    def __init__(self, model_id: Optional[str] = None, image_size: int = 1024, device: Optional[str] = None,
                 encoder_name: str = "dinov3_base", num_outputs: int = 3, max_batch: int = 1,
                 micro_batch: Optional[int] = None):
        self.image_size = image_size
        self.device = device or "cuda"
        if not str(self.device).startswith("cuda"):
            raise RuntimeError("s3od_b200.BackgroundRemoval runs on CUDA (sm_100a) devices only; there is no CPU fallback")
        self._arch = ARCHS[encoder_name]
        if num_outputs != self._arch.num_outputs:
            from dataclasses import replace
            self._arch = replace(self._arch, num_outputs=num_outputs)
        self._max_batch = max_batch
        self._micro_batch = micro_batch
        model_id = model_id or self.DEFAULT_MODEL_ID
        self.model = self._load_model(model_id)
        self.model.to(self.device)
        self.model.eval()
        self.mean = np.array([0.485, 0.456, 0.406])
        self.std = np.array([0.229, 0.224, 0.225])
        # one shared instance may be called from several threads (the reference's Gradio demo does, demo/app.py:18-25): the
        # context's workspace and output slots serve one call at a time
        self._lock = threading.Lock()

    @classmethod
    def from_pretrained(cls, model_id: str, **kwargs):
        return cls(model_id=model_id, **kwargs)

    def _load_model(self, model_id: str) -> B200DPTSegmentation:
        """predictor.py:49-77: hub download, else a local path, else ValueError; strict state_dict ingest."""
        try:
            from huggingface_hub import hf_hub_download
            checkpoint_path = hf_hub_download(repo_id=model_id, filename=self.DEFAULT_CHECKPOINT_NAME)
        except Exception as e:  # noqa: BLE001 - the reference catches everything here
            if Path(model_id).exists():
                checkpoint_path = model_id
            else:
                raise ValueError(f"Could not load model from {model_id}. "
                                 f"Ensure model exists on HuggingFace or provide valid local path. Error: {e}")
        checkpoint = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        return B200DPTSegmentation(checkpoint["state_dict"], self._arch, self.image_size, self.device,
                                   max_batch=self._max_batch, micro_batch=self._micro_batch)

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _to_uint8(image: Union[np.ndarray, Image.Image]) -> np.ndarray:
        if isinstance(image, Image.Image):
            return np.array(image.convert("RGB"))
        if isinstance(image, np.ndarray) and image.dtype == np.uint8 and image.ndim == 3 and image.shape[2] == 3:
            return image                # the only layout the reference path handles end to end (predictor.py:106,131)
        Image.fromarray(image)          # anything else: let PIL raise what it raises in the reference (predictor.py:106)
        raise ValueError(f"expected an RGB uint8 array of shape (H, W, 3), got {getattr(image, 'shape', None)}")

    @torch.no_grad()
    def remove_background(self, image: Union[np.ndarray, Image.Image], threshold: float = 0.5) -> RemovalResult:
        """predictor.py:96-139.  `threshold` is accepted and unused, as in the reference."""
        return self.remove_background_batch([image])[0]

    @torch.no_grad()
    def remove_background_batch(self, images: Sequence[Union[np.ndarray, Image.Image]]) -> List[RemovalResult]:
        """Batched form of remove_background: host uint8 images in, host results out.

        Images go through the device in chunks of the model's micro-batch; the device-to-host copies of a chunk's
        results run on a side stream into pinned memory while the next chunk computes."""
        arrays = [np.ascontiguousarray(self._to_uint8(im)) for im in images]
        with self._lock:
            return self._run_batch(arrays)

    def _run_batch(self, arrays: List[np.ndarray]) -> List[RemovalResult]:
        model = self.model
        dev = model.device
        for a in arrays:                                        # raise before touching the GPU, like the reference's paste
            model.geometry(a.shape[0], a.shape[1])
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._upload_stream = torch.cuda.Stream(dev)
        side, up = self._copy_stream, self._upload_stream
        pending = []
        step = max(1, min(model.max_batch, model.micro_batch))
        if getattr(self, "_slot_free", None) is None:
            self._slot_free = [None, None, None]                # event: the slot's device buffers have been copied out
        bounds = chunk_schedule(len(arrays), step)
        # host -> device copies of every chunk on their own stream, ahead of the compute that consumes them
        uploads = []
        with torch.cuda.stream(up):
            for s0, s1 in bounds:
                d_imgs = [torch.from_numpy(a).to(dev, non_blocking=True) for a in arrays[s0:s1]]
                uploads.append((d_imgs, up.record_event()))
        for ci, (d_imgs, uploaded) in enumerate(uploads):
            slot = ci % 3                                       # three sets of reusable device output buffers
            if self._slot_free[slot] is not None:
                main.wait_event(self._slot_free[slot])
            main.wait_event(uploaded)
            for t in d_imgs:
                t.record_stream(main)
            _, outs, ious, best = model.run_u8(d_imgs, slot=slot)
            done = torch.cuda.Event()
            done.record(main)
            side.wait_event(done)
            with torch.cuda.stream(side):
                h_ious = torch.empty(ious.shape, dtype=ious.dtype, pin_memory=True).copy_(ious, non_blocking=True)
                h_best = torch.empty(best.shape, dtype=best.dtype, pin_memory=True).copy_(best, non_blocking=True)
                host = []
                for all_masks, rgba in outs:
                    hm = torch.empty(all_masks.shape, dtype=all_masks.dtype, pin_memory=True).copy_(all_masks, non_blocking=True)
                    hr = torch.empty(rgba.shape, dtype=rgba.dtype, pin_memory=True).copy_(rgba, non_blocking=True)
                    host.append((hm, hr))
                for t in d_imgs:
                    t.record_stream(side)
                self._slot_free[slot] = side.record_event()
            pending.append((host, h_ious, h_best))
        side.synchronize()
        results: List[RemovalResult] = []
        for host, h_ious, h_best in pending:
            ious_np, best_np = h_ious.numpy(), h_best.numpy()
            for i, (hm, hr) in enumerate(host):
                am = hm.numpy()
                results.append(RemovalResult(predicted_mask=am[int(best_np[i])], all_masks=am, all_ious=ious_np[i].copy(),
                                             rgba_image=Image.fromarray(hr.numpy(), mode="RGBA")))
        return results
