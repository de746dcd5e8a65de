"""Drop-in for `s3od.BackgroundRemoval` (/root/reference/src/s3od/predictor.py:16-139) on the B200-native path.

Same constructor, class attributes, `from_pretrained`, `remove_background` signature and `RemovalResult` fields as the
reference; errors are the same types under the same conditions (ValueError for an unloadable model id, ValueError for
the odd-padding inputs the reference cannot paste).  Additive surface: `remove_background_batch`, and
`encoder_name` / `num_outputs` / `max_batch` / `micro_batch` / `devices` / `result_memory` keyword arguments.

Multi-GPU (the reference is single-device, predictor.py:35): `devices=[0, 1, ...]` keeps one weight replica,
one host thread and one set of streams per GPU; `remove_background_batch` splits the batch contiguously across them
(`sharder.shard_range`) - images are independent, so there is no collective on this path.

Everything between the uint8 source image and the result arrays runs in the CUDA library: letterbox resize + normalise,
the DINOv3 ViT + DPT head forward, sigmoid / crop / antialiased resize / argmax / RGBA composite.
"""
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
from PIL import Image

from .arch import ARCHS
from .engine import B200DPTSegmentation
from .sharder import shard_range


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


_COPY_POOL = None


def _copy_pool() -> ThreadPoolExecutor:
    """Host threads for the large memcpys either side of the PCIe copies (numpy releases the GIL while it copies)."""
    global _COPY_POOL
    if _COPY_POOL is None:
        _COPY_POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="s3od-copy")
    return _COPY_POOL


class _Replica:
    """One GPU: the device model, its upload / copy-out streams, three rotating output slots and a pinned staging ring.

    Chunk c of a batch: [host] pageable images -> pinned staging ring (pin-on-ingest; images that are already pinned skip it)
    -> [upload stream] H2D -> [launch stream] preprocess + forward + postprocess into output slot c % 3 -> [copy stream] D2H
    into pinned result buffers.  The host stages and uploads chunk c + 1 only after it has launched chunk c, so at most two
    chunks of inputs are resident on the device and the staging of the next chunk runs under the compute of the current one."""

    RING = 3

    def __init__(self, model: B200DPTSegmentation):
        self.model = model
        dev = model.device
        self.copy_stream = torch.cuda.Stream(dev)
        self.upload_stream = torch.cuda.Stream(dev)
        model.aux_streams = [self.copy_stream]
        self.slot_free = [None] * 3                 # event: the slot's device buffers have been copied out
        self.stage = [None] * self.RING             # pinned uint8 staging buffers, grown to the largest chunk so far
        self.stage_done = [None] * self.RING        # event: the H2D copies out of the staging buffer have finished

    def _staging(self, ring: int, nbytes: int) -> torch.Tensor:
        if self.stage_done[ring] is not None:
            self.stage_done[ring].synchronize()     # the previous user of this ring entry (three chunks ago) has been uploaded
        buf = self.stage[ring]
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, (buf.numel() * 5 // 4) if buf is not None else 0), dtype=torch.uint8, pin_memory=True)
            self.stage[ring] = buf
        return buf

    def _upload(self, arrays: List[np.ndarray], ring: int):
        """Stage (if pageable) and copy one chunk to the device on the upload stream; returns (device tensors, event)."""
        dev = self.model.device
        srcs = [torch.from_numpy(a) for a in arrays]
        pageable = [i for i, t in enumerate(srcs) if not t.is_pinned()]
        if pageable:
            offs, total = [], 0
            for i in pageable:
                offs.append(total)
                total += (arrays[i].nbytes + 255) & ~255
            buf = self._staging(ring, total)
            jobs = []
            for i, off in zip(pageable, offs):
                view = buf[off:off + arrays[i].nbytes].view(arrays[i].shape)
                srcs[i] = view
                if arrays[i].nbytes >= (1 << 20):
                    jobs.append(_copy_pool().submit(np.copyto, view.numpy(), arrays[i]))
                else:
                    np.copyto(view.numpy(), arrays[i])
            for j in jobs:
                j.result()
        with torch.cuda.stream(self.upload_stream):
            d_imgs = [t.to(dev, non_blocking=True) for t in srcs]
            ev = self.upload_stream.record_event()
        if pageable:
            self.stage_done[ring] = ev
        return d_imgs, ev

    def run(self, arrays: List[np.ndarray], pinned_results: bool) -> List[RemovalResult]:
        model = self.model
        dev = model.device
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            side = self.copy_stream
            step = max(1, min(model.max_batch, model.micro_batch))
            bounds = chunk_schedule(len(arrays), step)
            pending = []
            nxt = self._upload(arrays[bounds[0][0]:bounds[0][1]], 0) if bounds else None
            for ci, (s0, s1) in enumerate(bounds):
                d_imgs, uploaded = nxt
                slot = ci % 3                                       # three sets of reusable device output buffers
                if self.slot_free[slot] is not None:
                    main.wait_event(self.slot_free[slot])
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
                    self.slot_free[slot] = side.record_event()
                pending.append((host, h_ious, h_best))
                # the next chunk is staged and uploaded while this one computes (at most two chunks of inputs on the device)
                if ci + 1 < len(bounds):
                    n0, n1 = bounds[ci + 1]
                    nxt = self._upload(arrays[n0:n1], (ci + 1) % self.RING)
                del d_imgs
            side.synchronize()
        results: List[RemovalResult] = []
        copies = []
        for host, h_ious, h_best in pending:
            ious_np, best_np = h_ious.numpy(), h_best.numpy()
            for i, (hm, hr) in enumerate(host):
                am, rg = hm.numpy(), hr.numpy()
                if not pinned_results:
                    # ordinary (pageable) arrays like the reference returns: the pinned buffers go straight back to torch's
                    # host allocator instead of staying locked for as long as the caller keeps the result
                    am2, rg2 = np.empty_like(am), np.empty_like(rg)
                    copies.append(_copy_pool().submit(np.copyto, am2, am))
                    copies.append(_copy_pool().submit(np.copyto, rg2, rg))
                    am, rg = am2, rg2
                results.append((am, int(best_np[i]), ious_np[i].copy(), rg))
        for c in copies:
            c.result()
        return [RemovalResult(predicted_mask=am[b], all_masks=am, all_ious=io, rgba_image=Image.fromarray(rg, mode="RGBA"))
                for am, b, io, rg in results]


class BackgroundRemoval:
    DEFAULT_MODEL_ID = "okupyn/s3od"
    DEFAULT_CHECKPOINT_NAME = "s3od.pt"

    def __init__(self, model_id: Optional[str] = None, image_size: int = 1024, device: Optional[str] = None,
                 encoder_name: str = "dinov3_base", num_outputs: int = 3, max_batch: int = 1,
                 micro_batch: Optional[int] = None, devices: Optional[Sequence[Union[int, str]]] = None,
                 result_memory: str = "pinned"):
        self.image_size = image_size
        if devices:
            names = [d if isinstance(d, str) else f"cuda:{int(d)}" for d in devices]
            if device is not None and str(device) != names[0]:
                raise ValueError(f"`device` ({device}) must be the first entry of `devices` ({names[0]}) when both are given")
        else:
            names = [device or "cuda"]
        for n in names:
            if not str(n).startswith("cuda"):
                raise RuntimeError("s3od_b200.BackgroundRemoval runs on CUDA (sm_100a) devices only; there is no CPU fallback")
        self.device = names[0]
        self.devices = names
        if result_memory not in ("pinned", "pageable"):
            raise ValueError("result_memory must be 'pinned' (zero-copy views of page-locked buffers) or 'pageable' (plain arrays)")
        self._pinned_results = result_memory == "pinned"
        self._arch = ARCHS[encoder_name]
        if num_outputs != self._arch.num_outputs:
            from dataclasses import replace
            self._arch = replace(self._arch, num_outputs=num_outputs)
        self._max_batch = max_batch
        self._micro_batch = micro_batch
        model_id = model_id or self.DEFAULT_MODEL_ID
        self.models = self._load_model(model_id)
        self.model = self.models[0]                 # the reference attribute (predictor.py:44): the first device's replica
        for m, n in zip(self.models, names):
            m.to(n)
            m.eval()
        self.mean = np.array([0.485, 0.456, 0.406])
        self.std = np.array([0.229, 0.224, 0.225])
        self._replicas = [_Replica(m) for m in self.models]
        # one host thread per device, alive for the predictor's lifetime
        self._workers = ([ThreadPoolExecutor(max_workers=1, thread_name_prefix=f"s3od-{n}") for n in names]
                         if len(names) > 1 else [])
        # one shared instance may be called from several threads (the reference's Gradio demo does, demo/app.py:18-25): the
        # context's workspace and output slots serve one call at a time
        self._lock = threading.Lock()

    @classmethod
    def from_pretrained(cls, model_id: str, **kwargs):
        return cls(model_id=model_id, **kwargs)

    def _load_model(self, model_id: str) -> List[B200DPTSegmentation]:
        """predictor.py:49-77: hub download, else a local path, else ValueError; strict state_dict ingest.  One replica of
        the packed weights per device."""
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
        return [B200DPTSegmentation(checkpoint["state_dict"], self._arch, self.image_size, n,
                                    max_batch=self._max_batch, micro_batch=self._micro_batch) for n in self.devices]

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
        """Batched form of remove_background: host uint8 images (pageable or pinned) in, host results out.

        Images go through each device in chunks of the model's micro-batch; the device-to-host copies of a chunk's results
        run on a side stream into pinned memory while the next chunk computes.  With several devices the batch is split
        contiguously across them and every device is driven by its own host thread."""
        arrays = [np.ascontiguousarray(self._to_uint8(im)) for im in images]
        for a in arrays:                                        # raise before touching the GPU, like the reference's paste
            self.model.geometry(a.shape[0], a.shape[1])
        if not arrays:
            return []
        with self._lock:
            n_dev = min(len(self._replicas), len(arrays))
            if n_dev <= 1:
                return self._replicas[0].run(arrays, self._pinned_results)
            futures = []
            for r in range(n_dev):
                b, e = shard_range(len(arrays), r, n_dev)
                futures.append(self._workers[r].submit(self._run_on, r, arrays[b:e]))
            out: List[RemovalResult] = []
            for f in futures:
                out.extend(f.result())
            return out

    @torch.no_grad()
    def _run_on(self, r: int, arrays: List[np.ndarray]) -> List[RemovalResult]:
        torch.cuda.set_device(self.models[r].device)
        if not getattr(self, "_bound", None):
            self._bound = {}
        if r not in self._bound:                               # this device's host thread stays on the GPU's socket (best effort)
            from .sharder import bind_to_gpu_numa
            self._bound[r] = bind_to_gpu_numa(self.models[r].dev_index)
        return self._replicas[r].run(arrays, self._pinned_results)

    def close(self):
        for w in self._workers:
            w.shutdown(wait=True)
        for m in self.models:
            m.close()
