"""Training-step components of BASELINE.json configs[3] on the B200-native path (SURVEY 8e "training").

Host-side mirror of the reference's training step pieces that sit either side of the network's backward pass
(/root/reference/synth_sod/src/synth_sod/model_training/):
  * `LossModule`            loss.py:236-275 (components of config/loss/focal_iou.yaml) - forward AND backward in three
                            bandwidth-bound CUDA kernels (`s3od_loss_forward_backward`);
  * `ParameterLayout`       the flat fp32 layout of the 107.8 M parameters that receive a gradient, in reverse-autograd
                            order, split into all-reduce buckets; the 28 grad-less tensors (SURVEY F8 - the reference needs
                            `ddp_find_unused_parameters_true` for them, config/backend/2gpu.yaml:8) are excluded;
  * `GradientAllReduce`     the DDP exchange step (train.py:116-125 -> Lightning -> DistributedDataParallel): one
                            `torch.distributed` all-reduce per bucket, issued as soon as the bucket's gradients are complete so
                            that it overlaps the rest of the backward pass (NCCL over NVLink on the GPUs, gloo in the CPU tests);
  * `FusedAdamW`            lightning_module.py:183-193 (AdamW, encoder lr, head lr x 10, weight_decay 0.05) as one fused
                            kernel per parameter group over the flat buffers (`s3od_adamw_step`).
PyTorch supplies device memory, streams and the process group; every kernel is in the CUDA library (no fallback).

What is NOT here yet: the backward kernels of the network itself (dgrad / wgrad of the GEMMs and convolutions, LayerNorm /
GELU / attention / BatchNorm backward) - see DESIGN.md section 6.
"""
import ctypes
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .arch import ArchSpec
from .weights import expected_shapes


class S3odLossConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("focal_weight", "iou_weight", "mse_weight", "full_mask_lambda", "decay_rate",
                                              "alpha", "gamma", "smooth")]


def _lib():
    from .engine import load_library
    lib = load_library()
    if not getattr(lib, "_train_bound", False):
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        lib.s3od_loss_default_config.argtypes = [ctypes.POINTER(S3odLossConfig)]
        lib.s3od_loss_default_config.restype = None
        lib.s3od_loss_workspace_bytes.argtypes = [ci, ci, ci, ci]
        lib.s3od_loss_workspace_bytes.restype = ctypes.c_size_t
        lib.s3od_loss_out_floats.argtypes = [ci, ci]
        lib.s3od_loss_out_floats.restype = ctypes.c_size_t
        lib.s3od_loss_forward_backward.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ctypes.POINTER(S3odLossConfig), vp, vp, vp, vp,
                                                   ctypes.c_size_t, vp]
        lib.s3od_adamw_step.argtypes = [vp, vp, vp, vp, ctypes.c_size_t, ci, cf, cf, cf, cf, cf, cf, vp, vp]
        lib._train_bound = True
    return lib


def _check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.s3od_last_error().decode("utf-8", "replace")
        raise (ValueError if rc == -1 else RuntimeError)(f"{what}: {msg}")


class LossModule:
    """`LossModule(loss_config, full_mask_lambda, decay_rate)` of loss.py:236-240 with the focal_iou.yaml components;
    `forward(outputs, targets, epoch)` returns `(loss, loss_dict)` like loss.py:242-275 (0-dim device tensors), and
    `forward_backward` additionally returns d loss / d pred_masks and d loss / d pred_iou - what autograd would hand to the
    network's backward pass."""

    def __init__(self, focal_weight: float = 20.0, iou_weight: float = 1.0, mse_weight: float = 0.05, full_mask_lambda: float = 0.1,
                 decay_rate: float = 0.2, alpha: float = 0.25, gamma: float = 2.0, smooth: float = 1e-6):
        if not torch.cuda.is_available():
            raise RuntimeError("s3od_b200 training kernels need a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib()
        self.cfg = S3odLossConfig(focal_weight, iou_weight, mse_weight, full_mask_lambda, decay_rate, alpha, gamma, smooth)
        self._ws: Optional[torch.Tensor] = None

    def forward_backward(self, outputs: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor], epoch: int, need_grad: bool = True):
        z = outputs["pred_masks"].float().contiguous()
        t = targets["masks"].float().contiguous()
        B, K, H, W = z.shape
        if t.shape != (B, H, W):
            raise ValueError(f"targets['masks'] must have shape {(B, H, W)}, got {tuple(t.shape)}")
        q = outputs["pred_iou"].float().contiguous() if K > 1 else None
        dev = z.device
        need = self.lib.s3od_loss_workspace_bytes(B, K, H, W)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = torch.empty(self.lib.s3od_loss_out_floats(B, K), dtype=torch.float32, device=dev)
        dz = torch.empty_like(z) if need_grad else None
        dq = torch.zeros(B, K, dtype=torch.float32, device=dev) if K > 1 else None
        with torch.cuda.device(dev):
            _check(self.lib, self.lib.s3od_loss_forward_backward(
                z.data_ptr(), q.data_ptr() if q is not None else None, t.data_ptr(), B, K, H, W, int(epoch), ctypes.byref(self.cfg),
                dz.data_ptr() if dz is not None else None, dq.data_ptr() if dq is not None else None, out.data_ptr(),
                self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream(dev).cuda_stream), "s3od_loss_forward_backward")
        if K == 1:                                      # compute_single_mask_loss (loss.py:166-188): {name: component loss}
            parts = {"focal_loss": out[3], "iou_loss": out[5]}
        else:                                           # loss.py:211-231, 272-275 (tensor entries reduced to their means)
            parts = {"best_iou": out[1], "gt_ious": out[2], "focal_loss_best": out[3], "focal_loss_full": out[4],
                     "iou_loss_best": out[5], "iou_loss_full": out[6], "mse_ious_loss": out[7]}
        extra = {"gt_ious_per_mask": out[8:8 + B * K].view(B, K), "best_indices": out[8 + B * K:8 + B * K + B].long()}
        return out[0], parts, {"pred_masks": dz, "pred_iou": dq}, extra

    def forward(self, outputs, targets, epoch: int):
        loss, parts, _, _ = self.forward_backward(outputs, targets, epoch, need_grad=False)
        return loss, parts

    __call__ = forward


# ------------------------------------------------------------------------------------------------------ parameter layout
def gradless_keys(arch: ArchSpec, prefix: str = "encoder.model.layer.") -> List[str]:
    """The parameter tensors that never receive a gradient on this path: the last encoder layer and the final norm
    (dead for the taps, F3), the mask token, and refinenet4.resConfUnit1 (refinenet4 is called with one input, model.py:228)."""
    shapes = expected_shapes(arch, prefix)
    last = f"{prefix}{arch.layers - 1}."
    dead = [k for k in shapes if k.startswith(last) and arch.layers > arch.layers_needed]
    dead += ["encoder.norm.weight", "encoder.norm.bias", "encoder.embeddings.mask_token"]
    dead += [k for k in shapes if k.startswith("seg_head.scratch.refinenet4.resConfUnit1.")
             and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    return dead


def trainable_keys(arch: ArchSpec, prefix: str = "encoder.model.layer.") -> List[str]:
    """Parameters (not buffers) in module-definition order = the order `model.parameters()` yields them."""
    return [k for k in expected_shapes(arch, prefix) if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


@dataclass
class Segment:
    name: str
    offset: int          # in elements, 4-element (16-byte) aligned
    numel: int
    group: int           # 0 = encoder (lr), 1 = seg_head (lr x 10)  - lightning_module.py:184-190
    bucket: int


class ParameterLayout:
    """Flat fp32 layout of the parameters that receive a gradient, in REVERSE autograd order (the order their gradients become
    ready during the backward pass: mask head first, patch embedding last), cut into all-reduce buckets of about
    `bucket_bytes`.  Bucket i is the contiguous slice [bucket_offsets[i], bucket_offsets[i + 1]) of the flat buffers, so one
    collective per bucket needs no packing kernel.  Every segment starts on a 16-byte boundary (the fused AdamW kernel uses
    128-bit accesses); the padding elements stay zero."""

    def __init__(self, arch: ArchSpec, bucket_bytes: int = 64 << 20, prefix: str = "encoder.model.layer."):
        shapes = expected_shapes(arch, prefix)
        dead = set(gradless_keys(arch, prefix))
        keys = [k for k in trainable_keys(arch, prefix) if k not in dead]
        self.gradless = [k for k in trainable_keys(arch, prefix) if k in dead]
        self.shapes = {k: shapes[k] for k in keys}
        self.segments: List[Segment] = []
        self.bucket_offsets = [0]
        off, bucket, cap = 0, 0, max(1, bucket_bytes // 4)
        for k in reversed(keys):                         # reverse of forward / definition order
            n = int(math.prod(shapes[k]))
            if off - self.bucket_offsets[-1] > 0 and off + n - self.bucket_offsets[-1] > cap:
                self.bucket_offsets.append(off)
                bucket += 1
            self.segments.append(Segment(k, off, n, 0 if k.startswith("encoder.") else 1, bucket))
            off += (n + 3) & ~3
        self.total = off
        self.bucket_offsets.append(off)
        self.by_name = {s.name: s for s in self.segments}
        self.num_buckets = len(self.bucket_offsets) - 1
        self._bucket_members: List[List[str]] = [[] for _ in range(self.num_buckets)]
        for s in self.segments:
            self._bucket_members[s.bucket].append(s.name)
        # head parameters come first in reverse order, so each optimiser group is ONE contiguous range of the flat buffer
        head = [s for s in self.segments if s.group == 1]
        enc = [s for s in self.segments if s.group == 0]
        self.group_ranges = {1: (head[0].offset, ((head[-1].offset + head[-1].numel + 3) & ~3)) if head else (0, 0),
                             0: (enc[0].offset, ((enc[-1].offset + enc[-1].numel + 3) & ~3)) if enc else (0, 0)}
        if head and enc and head[-1].offset > enc[0].offset:
            raise AssertionError("optimiser groups are not contiguous in reverse-autograd order")

    def bucket_members(self, b: int) -> List[str]:
        return self._bucket_members[b]

    def numel_with_grad(self) -> int:
        return sum(s.numel for s in self.segments)

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        s = self.by_name[name]
        return flat[s.offset:s.offset + s.numel].view(self.shapes[name])

    def flatten(self, tensors: Dict[str, torch.Tensor], device=None, dtype=torch.float32) -> torch.Tensor:
        flat = torch.zeros(self.total, dtype=dtype, device=device if device is not None else next(iter(tensors.values())).device)
        for s in self.segments:
            flat[s.offset:s.offset + s.numel].copy_(tensors[s.name].reshape(-1))
        return flat


class GradientAllReduce:
    """The data-parallel exchange step: sum-reduce the flat gradient buffer bucket by bucket, each bucket as soon as every
    gradient in it has been written (`mark_ready`), asynchronously, so the collective runs under the remaining backward
    kernels.  `finish()` returns when all buckets are reduced (stream-ordered on CUDA); the 1 / world of the mean is folded into
    the optimiser (`FusedAdamW.step(grad_scale=1 / world)`), so the payload is reduced exactly once and never rescaled."""

    def __init__(self, layout: ParameterLayout, flat_grad: torch.Tensor, group=None):
        self.layout = layout
        self.flat = flat_grad
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.reset()

    def reset(self):
        self._pending = [set(m) for m in (self.layout.bucket_members(b) for b in range(self.layout.num_buckets))]
        self._works = []
        self._launched = [False] * self.layout.num_buckets
        self.launch_order: List[int] = []

    def mark_ready(self, name: str):
        """The gradient of parameter `name` has been (enqueued to be) written into its slice of the flat buffer."""
        b = self.layout.by_name[name].bucket
        self._pending[b].discard(name)
        if not self._pending[b] and not self._launched[b]:
            self._launch(b)

    def _launch(self, b: int):
        self._launched[b] = True
        self.launch_order.append(b)
        if not self.enabled:
            return
        lo, hi = self.layout.bucket_offsets[b], self.layout.bucket_offsets[b + 1]
        # NCCL: the collective is enqueued on the process group's own stream after the current stream's work so far
        self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for b in range(self.layout.num_buckets):          # anything not marked (should not happen) still gets reduced
            if not self._launched[b]:
                self._launch(b)
        for w in self._works:
            w.wait()                                      # CUDA: the current stream waits for the collective; gloo: blocks
        self._works = []

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if self.enabled else 1


class FusedAdamW:
    """torch.optim.AdamW with the reference's two parameter groups (lightning_module.py:183-193) over the flat buffers of a
    `ParameterLayout`: one fused kernel launch per group and step.  Grad-less parameters are not in the layout, which is also
    what torch does (parameters whose .grad is None are skipped: no update, no weight decay)."""

    def __init__(self, layout: ParameterLayout, flat_param: torch.Tensor, lr: float = 1e-5, head_lr_scale: float = 10.0,
                 weight_decay: float = 0.05, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, bf16_copy: bool = False):
        if flat_param.device.type != "cuda":
            raise RuntimeError("FusedAdamW runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib()
        self.layout = layout
        self.param = flat_param
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self.param_bf16 = torch.zeros(flat_param.numel(), dtype=torch.bfloat16, device=flat_param.device) if bf16_copy else None
        self.lrs = {0: lr, 1: lr * head_lr_scale}
        self.weight_decay, self.betas, self.eps = weight_decay, betas, eps
        self.steps = 0

    def step(self, flat_grad: torch.Tensor, grad_scale: float = 1.0, lr_factor: float = 1.0):
        self.steps += 1
        dev = self.param.device
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            for group, (lo, hi) in self.layout.group_ranges.items():
                if hi <= lo:
                    continue
                bf = self.param_bf16[lo:hi].data_ptr() if self.param_bf16 is not None else None
                _check(self.lib, self.lib.s3od_adamw_step(
                    self.param[lo:hi].data_ptr(), flat_grad[lo:hi].data_ptr(), self.exp_avg[lo:hi].data_ptr(),
                    self.exp_avg_sq[lo:hi].data_ptr(), hi - lo, self.steps, self.lrs[group] * lr_factor, self.betas[0], self.betas[1],
                    self.eps, self.weight_decay, grad_scale, bf, st), "s3od_adamw_step")


# -------------------------------------------------------------------- fused exchange + optimiser step over peer memory
class _RawCuda:
    """__cuda_array_interface__ holder: lets torch view a device allocation made by the library (no copy, no ownership)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerBuffer:
    """A flat device buffer that other processes of the box can map (CUDA IPC over NVLink): `s3od_peer_alloc` + a torch view."""

    def __init__(self, numel: int, dtype: torch.dtype, device: torch.device):
        self.lib = _lib()
        self.device = torch.device(device)
        self.nbytes = numel * torch.empty((), dtype=dtype).element_size()
        p = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.s3od_peer_alloc(ctypes.byref(p), self.nbytes), "s3od_peer_alloc")
        self.ptr = p.value
        self._holder = _RawCuda(self.ptr, self.nbytes)
        self.tensor = torch.as_tensor(self._holder, device=self.device).view(dtype)

    def handle(self) -> bytes:
        buf = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.s3od_peer_export(ctypes.c_void_p(self.ptr), buf), "s3od_peer_export")
        return bytes(buf)

    def free(self):
        if self.ptr:
            self.tensor = None
            with torch.cuda.device(self.device):
                self.lib.s3od_peer_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


def _bind_peer(lib):
    if not getattr(lib, "_peer_bound", False):
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        lib.s3od_peer_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
        lib.s3od_peer_free.argtypes = [vp]
        lib.s3od_peer_export.argtypes = [vp, ctypes.POINTER(ctypes.c_ubyte)]
        lib.s3od_peer_open.argtypes = [ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(vp)]
        lib.s3od_peer_close.argtypes = [vp]
        lib.s3od_ddp_fused_adamw_step.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, vp, vp, ctypes.c_size_t,
                                                  ctypes.c_size_t, ci, cf, cf, cf, cf, cf, vp]
        lib._peer_bound = True
    return lib


class FusedDataParallelAdamW:
    """DistributedDataParallel's gradient all-reduce + AdamW.step as ONE kernel per rank over peer-mapped buffers
    (`s3od_ddp_fused_adamw_step`): every rank reduces, updates and re-broadcasts its own 1/world slice of the flat buffers.

    `grad` / `param` / `param_bf16` are this rank's flat buffers (torch views of peer-exportable allocations): the backward pass
    writes `grad`, the forward pass reads `param_bf16` (GEMM operands) and `param` (fp32 vectors).  `step()` is bracketed by two
    stream-ordered barriers of the process group (one-element all-reduces: no kernel ever waits on another GPU).

    `emulate_world=N` builds N virtual ranks on ONE device (N sets of buffers, N sequential launches) - the same kernel and
    slice arithmetic, used to check the path where fewer GPUs than ranks are available."""

    def __init__(self, layout: ParameterLayout, device, lr: float = 1e-5, head_lr_scale: float = 10.0, weight_decay: float = 0.05,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, group=None, emulate_world: int = 0,
                 replicate_fp32: bool = True):
        """replicate_fp32=True (default) keeps DistributedDataParallel's invariant: every rank holds the full, bit-identical fp32
        parameters.  False shards the fp32 masters like the moments (each rank's `param()` is current only on its own slice;
        gather the slices for a checkpoint) and pushes just the bf16 working copy: 2 instead of 6 bytes per parameter."""
        self.lib = _bind_peer(_lib())
        self.replicate_fp32 = replicate_fp32
        self.layout = layout
        self.device = torch.device(device)
        self.group = group
        self.lrs = {0: lr, 1: lr * head_lr_scale}
        self.weight_decay, self.betas, self.eps = weight_decay, betas, eps
        self.steps = 0
        self.emulated = emulate_world > 0
        n = layout.total
        if self.emulated:
            self.world, self.rank = emulate_world, 0
            self._bufs = [(PeerBuffer(n, torch.float32, self.device), PeerBuffer(n, torch.float32, self.device),
                           PeerBuffer(n, torch.bfloat16, self.device)) for _ in range(self.world)]
            self.exp_avg = [torch.zeros(n, device=self.device) for _ in range(self.world)]
            self.exp_avg_sq = [torch.zeros(n, device=self.device) for _ in range(self.world)]
            ptrs = [[b[j].ptr for b in self._bufs] for j in range(3)]
            self._opened = []
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
            own = (PeerBuffer(n, torch.float32, self.device), PeerBuffer(n, torch.float32, self.device),
                   PeerBuffer(n, torch.bfloat16, self.device))
            self._bufs = [own]
            self.exp_avg = [torch.zeros(n, device=self.device)]
            self.exp_avg_sq = [torch.zeros(n, device=self.device)]
            ptrs = [[0] * self.world for _ in range(3)]
            self._opened = []
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, tuple(b.handle() for b in own), group=group)
                for w in range(self.world):
                    for j in range(3):
                        if w == self.rank:
                            ptrs[j][w] = own[j].ptr
                        else:
                            p = ctypes.c_void_p()
                            hb = (ctypes.c_ubyte * 64).from_buffer_copy(handles[w][j])
                            with torch.cuda.device(self.device):
                                _check(self.lib, self.lib.s3od_peer_open(hb, ctypes.byref(p)), "s3od_peer_open")
                            ptrs[j][w] = p.value
                            self._opened.append(p.value)
            else:
                ptrs = [[own[j].ptr] for j in range(3)]
        self._ptr_arrays = [(ctypes.c_void_p * self.world)(*col) for col in ptrs]
        # per (virtual) rank view of the fp32 parameter pointers: peers are NULL when the masters are sharded
        self._param_ptrs = {r: (ctypes.c_void_p * self.world)(*[p if (replicate_fp32 or w == r) else None for w, p in enumerate(ptrs[1])])
                            for r in (range(self.world) if self.emulated else [self.rank])}
        self._token = torch.zeros(1, device=self.device)

    # this rank's buffers (virtual rank r in emulation)
    def grad(self, r: int = 0) -> torch.Tensor:
        return self._bufs[r][0].tensor

    def param(self, r: int = 0) -> torch.Tensor:
        return self._bufs[r][1].tensor

    def param_bf16(self, r: int = 0) -> torch.Tensor:
        return self._bufs[r][2].tensor

    def _barrier(self):
        if not self.emulated and self.world > 1:
            dist.all_reduce(self._token, group=self.group)        # stream-ordered: enqueued behind everything on this stream

    def step(self, lr_factor: float = 1.0):
        self.steps += 1
        self._barrier()                                           # every rank's gradients are complete
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            ranks = range(self.world) if self.emulated else [self.rank]
            for r in ranks:
                m, v = (self.exp_avg[r], self.exp_avg_sq[r]) if self.emulated else (self.exp_avg[0], self.exp_avg_sq[0])
                for group, (lo, hi) in self.layout.group_ranges.items():
                    if hi <= lo:
                        continue
                    _check(self.lib, self.lib.s3od_ddp_fused_adamw_step(
                        self._ptr_arrays[0], self._param_ptrs[r], self._ptr_arrays[2], self.world, r, m.data_ptr(), v.data_ptr(), lo, hi,
                        self.steps, self.lrs[group] * lr_factor, self.betas[0], self.betas[1], self.eps, self.weight_decay, st),
                        "s3od_ddp_fused_adamw_step")
        self._barrier()                                           # every rank's parameters are visible everywhere

    def link_bytes_per_step(self) -> int:
        """Bytes ONE direction of this GPU's link carries per step: its pushes (4 + 2 B per parameter of its slice to each peer, or 2
        B with sharded masters) plus the gradient slices its peers read from it (4 B) = (world-1)/world x 10 (or 6) B per parameter."""
        return int(self.layout.total * (self.world - 1) / self.world * (4 + (6 if self.replicate_fp32 else 2)))

    def close(self):
        for p in self._opened:
            self.lib.s3od_peer_close(ctypes.c_void_p(p))
        self._opened = []
        for bufs in self._bufs:
            for b in bufs:
                b.free()
        self._bufs = []


# ------------------------------------------------------------------------------------ encoder block: forward + backward
def _bind_block(lib):
    if not getattr(lib, "_block_bound", False):
        vp, ci, cf, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
        lib.s3od_train_transpose.argtypes = [vp, ci, vp, ci, ci, ci, ci, ll, ci, cf, vp]
        lib.s3od_train_scale_cast.argtypes = [vp, vp, vp, ll, ci, vp]
        lib.s3od_train_residual_scale_add.argtypes = [vp, vp, vp, vp, ll, ci, vp]
        lib.s3od_train_add_bias.argtypes = [vp, vp, ll, ci, vp]
        lib.s3od_train_colsum_workspace_bytes.argtypes = [ci, ci]
        lib.s3od_train_colsum_workspace_bytes.restype = ctypes.c_size_t
        lib.s3od_train_colsum.argtypes = [vp, vp, ci, ci, vp, vp, ci, vp, vp]
        lib.s3od_train_colsum2_workspace_bytes.argtypes = [ci, ci]
        lib.s3od_train_colsum2_workspace_bytes.restype = ctypes.c_size_t
        lib.s3od_train_colsum2.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp]
        lib.s3od_train_ln_backward_workspace_bytes.argtypes = [ci, ci]
        lib.s3od_train_ln_backward_workspace_bytes.restype = ctypes.c_size_t
        lib.s3od_train_ln_backward.argtypes = [vp, vp, vp, vp, vp, ci, ci, cf, vp, vp, vp, vp]
        lib.s3od_train_gelu_forward.argtypes = [vp, vp, ll, vp]
        lib.s3od_train_gelu_backward.argtypes = [vp, vp, vp, vp, ll, vp]
        lib.s3od_train_qkv_split_rope.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, cf, vp]
        lib.s3od_train_qkv_merge_rope_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, cf, cf, vp]
        lib.s3od_train_qkv_merge_rope_backward_rows.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, cf, cf, vp]
        lib.s3od_train_attention_forward.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]
        lib.s3od_train_attention_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]
        lib.s3od_train_split_heads.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp]
        lib.s3od_train_rowdot64.argtypes = [vp, vp, vp, ll, vp]
        lib.s3od_train_softmax2_rows.argtypes = [vp, vp, ci, ci, vp]
        lib.s3od_train_softmax_backward.argtypes = [vp, vp, vp, vp, ci, ci, vp]
        lib.s3od_op_gemm_f32_bias.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp]
        lib.s3od_op_wgrad_gemm_f32.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci, ci, vp, vp]
        lib.s3od_op_gemm_f32_splitk.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, vp]
        lib._block_bound = True
    return lib


def wgrad_plan(n_out: int, n_in: int, rows: int, sms: int = 148):
    """(k-splits, padded contraction length) of a weight-gradient GEMM dW[n_out, n_in] = dY^T X over `rows` tokens.  The output has
    only ceil(n_out / 256) * (n_in / 256 or 128) CTA-pair tiles - 9 for a 768 x 768 projection, 3 for a 64-channel convolution - so
    the contraction is split until the tiles fill the SM pairs once (s3od_op_gemm_f32_splitk), keeping >= 8 k-blocks per split."""
    tiles = ((n_out + 127) // 128 + 1) // 2 * (n_in // (256 if n_in % 256 == 0 else 128))
    splits = max(1, min(32, (sms // 2) // max(tiles, 1), rows // 512))
    kpad = (rows + 64 * splits - 1) // (64 * splits) * (64 * splits)
    return splits, kpad


def wgrad_plan_tn(n_out: int, n_in: int, rows: int, sms: int = 148) -> int:
    """k-splits of the transpose-free weight-gradient GEMM (csrc/gemm_tn.cuh, 128 x 256 output tiles, one CTA each): enough that the
    (split, tile) items fill the SMs once, at least 4 blocks of 64 rows per split."""
    tiles = ((n_out + 127) // 128) * ((n_in + 255) // 256)
    return max(1, min(64, sms // max(tiles, 1), rows // 256))


def wgrad_tn(lib, check, dy: torch.Tensor, x: torch.Tensor, rows: int, n_out: int, n_in: int, sms: int, stream) -> torch.Tensor:
    """dW fp32 [n_out, n_in] = dy[rows, n_out]^T x[rows, n_in] on `s3od_op_wgrad_gemm_f32`: bf16 operands exactly as they lie in memory
    (one row per token / pixel), nothing transposed.  n_out % 64 == 0 and n_in % 64 == 0."""
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and dy.is_contiguous() and x.is_contiguous()
    splits = wgrad_plan_tn(n_out, n_in, rows, sms)
    c = torch.empty(n_out, n_in, dtype=torch.float32, device=dy.device)
    ws = torch.empty(splits * n_out * n_in, dtype=torch.float32, device=dy.device) if splits > 1 else None
    check(lib.s3od_op_wgrad_gemm_f32(dy.data_ptr(), n_out, x.data_ptr(), n_in, c.data_ptr(), n_out, n_in, rows, splits,
                                     ws.data_ptr() if ws is not None else None, stream), "s3od_op_wgrad_gemm_f32")
    return c


class EncoderBlockStep:
    """Forward (with saved activations) and backward of ONE encoder block - DINOv3ViTLayer.forward, HF modeling_dinov3_vit.py:424-450
    with the attention of HF:294-334 - on the CUDA library: the four linears and their dgrad / wgrad are tcgen05 GEMMs
    (`s3od_op_gemm_f32`: fp32 C = bf16 A x bf16 B^T; dgrad reads the transposed weights, wgrad contracts over the tokens with dY^T
    and X^T zero-padded to a multiple of 64 rows), the attention is the fused flash kernel forward (`s3od_train_attention_forward`,
    which also writes the log-sum-exp of every score row) and the fused flash backward (`s3od_train_attention_backward`,
    csrc/attention_bwd.cuh: P and dA recomputed tile by tile, no N x N scratch), everything between them is a kernel of
    csrc/train_block.cuh.

    Precision: bf16 operands / fp32 accumulation in every GEMM (the reference trains with float32_matmul_precision "medium",
    train.py:74), fp32 LayerNorm / GELU / softmax / reductions; the gradients that feed a GEMM are rounded to bf16 once."""

    LOG2E = 1.4426950408889634

    def __init__(self, sd: Dict[str, torch.Tensor], layer_prefix: str, arch: ArchSpec, image_size: int, device="cuda:0"):
        from .weights import rope_tables
        self.lib = _bind_block(_lib())
        self.arch, self.dev = arch, torch.device(device)
        self.sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self.D, self.H, self.I = arch.hidden, arch.heads, arch.mlp
        g = image_size // arch.patch
        self.N = g * g + arch.n_prefix
        self.Npad = (self.N + 383) // 384 * 384          # a whole number of 128-row and of 96-row tiles (fused attention backward)
        f = lambda t: t.detach().to(self.dev, torch.float32).contiguous()           # noqa: E731
        a = layer_prefix + "attention."
        D = self.D
        kb = sd.get(a + "k_proj.bias", torch.zeros(D, device=sd[a + "q_proj.bias"].device))
        self.w = {"qkv.w": f(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0)),
                  "qkv.b": f(torch.cat([sd[a + "q_proj.bias"], kb, sd[a + "v_proj.bias"]], 0)),
                  "o.w": f(sd[a + "o_proj.weight"]), "o.b": f(sd[a + "o_proj.bias"]),
                  "up.w": f(sd[layer_prefix + "mlp.up_proj.weight"]), "up.b": f(sd[layer_prefix + "mlp.up_proj.bias"]),
                  "down.w": f(sd[layer_prefix + "mlp.down_proj.weight"]), "down.b": f(sd[layer_prefix + "mlp.down_proj.bias"]),
                  "ln1.w": f(sd[layer_prefix + "norm1.weight"]), "ln1.b": f(sd[layer_prefix + "norm1.bias"]),
                  "ln2.w": f(sd[layer_prefix + "norm2.weight"]), "ln2.b": f(sd[layer_prefix + "norm2.bias"]),
                  "ls1": f(sd[layer_prefix + "layer_scale1.lambda1"]), "ls2": f(sd[layer_prefix + "layer_scale2.lambda1"])}
        self.wb = {k: self.w[k].to(torch.bfloat16) for k in ("qkv.w", "o.w", "up.w", "down.w")}       # forward B operands [N_out, K_in]
        self.wt = {k: self._transpose(self.wb[k], 1, self.wb[k].shape[0], self.wb[k].shape[1], self.wb[k].shape[0]) .view(self.wb[k].shape[1], self.wb[k].shape[0])
                   for k in self.wb}                                                                  # dgrad B operands [K_in, N_out]
        cos, sin = rope_tables(g, g, arch.head_dim, arch.rope_theta)
        self.cos, self.sin = cos.to(self.dev).contiguous(), sin.to(self.dev).contiguous()
        self.saved = None

    # ---- thin wrappers around the C ABI -------------------------------------------------------------------------
    def _st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _ck(self, rc, what):
        _check(self.lib, rc, what)

    def _gemm(self, a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        c = torch.empty(M, N, dtype=torch.float32, device=self.dev)
        self._ck(self.lib.s3od_op_gemm_f32_bias(a.data_ptr(), b.data_ptr(), bias.data_ptr() if bias is not None else None, c.data_ptr(), M, N, K,
                                                self._st()), "s3od_op_gemm_f32_bias")
        return c

    def _wgrad(self, dy: torch.Tensor, x: torch.Tensor, rows: int, n_out: int, n_in: int) -> torch.Tensor:
        """dW fp32 [n_out, n_in] = dy[rows, n_out]^T x[rows, n_in] with the contraction split across the SMs: on the transpose-free
        kernel (`wgrad_tn`) when both widths are multiples of 64, else through transposed, zero-padded copies (`wgrad_plan`)."""
        if n_out % 64 == 0 and n_in % 64 == 0:
            if dy.dtype != torch.bfloat16:
                dy = self._scale_cast(dy)
            if x.dtype != torch.bfloat16:
                x = self._scale_cast(x)
            return wgrad_tn(self.lib, self._ck, dy.contiguous(), x.contiguous(), rows, n_out, n_in, self.sms, self._st())
        splits, kpad = wgrad_plan(n_out, n_in, rows, self.sms)
        a = self._transpose(dy, 1, rows, n_out, kpad)
        b = self._transpose(x, 1, rows, n_in, kpad)
        c = torch.empty(n_out, n_in, dtype=torch.float32, device=self.dev)
        ws = torch.empty(splits * n_out * n_in, dtype=torch.float32, device=self.dev) if splits > 1 else None
        self._ck(self.lib.s3od_op_gemm_f32_splitk(a.data_ptr(), b.data_ptr(), c.data_ptr(), n_out, n_in, kpad, splits,
                                                  ws.data_ptr() if ws is not None else None, self._st()), "s3od_op_gemm_f32_splitk")
        return c

    def _transpose(self, t: torch.Tensor, batch: int, rows: int, cols: int, rows_padded: int, scale: float = 1.0) -> torch.Tensor:
        """[batch][rows][cols] (dense) -> bf16 [batch][cols][rows_padded], zero padded."""
        out = torch.empty(batch, cols, rows_padded, dtype=torch.bfloat16, device=self.dev)
        self._ck(self.lib.s3od_train_transpose(t.data_ptr(), 1 if t.dtype == torch.float32 else 0, out.data_ptr(), batch, rows, cols, rows_padded,
                                               rows * cols, cols, scale, self._st()), "s3od_train_transpose")
        return out

    def _layernorm(self, x: torch.Tensor, w, b) -> torch.Tensor:
        M, D = x.shape
        y = torch.empty(M, D, dtype=torch.bfloat16, device=self.dev)
        self._ck(self.lib.s3od_op_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, D, self.arch.ln_eps, self._st()), "s3od_op_layernorm")
        return y

    def _add_bias(self, a, bias):
        self._ck(self.lib.s3od_train_add_bias(a.data_ptr(), bias.data_ptr(), a.numel(), a.shape[1], self._st()), "s3od_train_add_bias")
        return a

    def _residual(self, x, y, lam):
        out = torch.empty_like(x)
        self._ck(self.lib.s3od_train_residual_scale_add(x.data_ptr(), y.data_ptr(), lam.data_ptr(), out.data_ptr(), x.numel(), x.shape[1], self._st()),
                 "s3od_train_residual_scale_add")
        return out

    def _scale_cast(self, x, colscale=None):
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=self.dev)
        self._ck(self.lib.s3od_train_scale_cast(x.data_ptr(), colscale.data_ptr() if colscale is not None else None, out.data_ptr(), x.numel(),
                                                x.shape[-1], self._st()), "s3od_train_scale_cast")
        return out

    def _colsum(self, a, b=None, colscale=None):
        M, C = a.shape
        ws = torch.empty(self.lib.s3od_train_colsum_workspace_bytes(M, C), dtype=torch.uint8, device=self.dev)
        out = torch.empty(C, dtype=torch.float32, device=self.dev)
        self._ck(self.lib.s3od_train_colsum(a.data_ptr(), b.data_ptr() if b is not None else None, M, C,
                                            colscale.data_ptr() if colscale is not None else None, out.data_ptr(), 0, ws.data_ptr(), self._st()),
                 "s3od_train_colsum")
        return out

    def _colsum2(self, a, b, colscale_a=None):
        """(sum_r a*b, colscale_a * sum_r a) per column, one pass over a."""
        M, C = a.shape
        ws = torch.empty(self.lib.s3od_train_colsum2_workspace_bytes(M, C), dtype=torch.uint8, device=self.dev)
        out_ab, out_a = torch.empty(C, dtype=torch.float32, device=self.dev), torch.empty(C, dtype=torch.float32, device=self.dev)
        self._ck(self.lib.s3od_train_colsum2(a.data_ptr(), b.data_ptr(), M, C, colscale_a.data_ptr() if colscale_a is not None else None,
                                             out_ab.data_ptr(), out_a.data_ptr(), ws.data_ptr(), self._st()), "s3od_train_colsum2")
        return out_ab, out_a

    def _ln_backward(self, x, gamma, dy, dres):
        M, D = x.shape
        ws = torch.empty(self.lib.s3od_train_ln_backward_workspace_bytes(M, D), dtype=torch.uint8, device=self.dev)
        dx = torch.empty_like(x)
        dg, db = torch.empty(D, device=self.dev), torch.empty(D, device=self.dev)
        self._ck(self.lib.s3od_train_ln_backward(x.data_ptr(), gamma.data_ptr(), dy.data_ptr(), dres.data_ptr() if dres is not None else None,
                                                 dx.data_ptr(), M, D, self.arch.ln_eps, dg.data_ptr(), db.data_ptr(), ws.data_ptr(), self._st()),
                 "s3od_train_ln_backward")
        return dx, dg, db

    def _split_rope(self, qkv, B, npad):
        q, k, v = (torch.empty(B, self.H, npad, 64, dtype=torch.bfloat16, device=self.dev) for _ in range(3))
        self._ck(self.lib.s3od_train_qkv_split_rope(qkv.data_ptr(), self.cos.data_ptr(), self.sin.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                                    B, self.N, npad, self.H, self.arch.n_prefix, self.LOG2E / 8.0, self._st()), "s3od_train_qkv_split_rope")
        return q, k, v

    def _split_heads(self, t, B):
        out = torch.empty(B, self.H, self.Npad, 64, dtype=torch.bfloat16, device=self.dev)
        self._ck(self.lib.s3od_train_split_heads(t.data_ptr(), 1 if t.dtype == torch.float32 else 0, out.data_ptr(), B, self.N, self.Npad, self.H,
                                                 self._st()), "s3od_train_split_heads")
        return out

    # ---- forward ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 (B, N, D) -> block output fp32 (B, N, D); keeps what backward needs."""
        B, N, D = x.shape
        assert N == self.N and D == self.D
        w, wb, I, H = self.w, self.wb, self.I, self.H
        M = B * N
        x0 = x.to(self.dev, torch.float32).contiguous().view(M, D)
        with torch.cuda.device(self.dev):
            xn1 = self._layernorm(x0, w["ln1.w"], w["ln1.b"])
            qkv = self._gemm(xn1, wb["qkv.w"], M, 3 * D, D, w["qkv.b"])
            q, k, v = self._split_rope(qkv, B, self.Npad)                   # [B, H, Npad, 64] bf16, zero rows behind the sequence
            ctx = torch.empty(M, D, dtype=torch.bfloat16, device=self.dev)
            lse = torch.full((B * H, self.Npad), float("inf"), dtype=torch.float32, device=self.dev)
            self._ck(self.lib.s3od_train_attention_forward(q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, H, N, self.Npad,
                                                           self._st()), "s3od_train_attention_forward")
            o = self._gemm(ctx, wb["o.w"], M, D, D, w["o.b"])
            x1 = self._residual(x0, o, w["ls1"])
            xn2 = self._layernorm(x1, w["ln2.w"], w["ln2.b"])
            hpre = self._gemm(xn2, wb["up.w"], M, I, D, w["up.b"])
            hmid = torch.empty(M, I, dtype=torch.bfloat16, device=self.dev)
            self._ck(self.lib.s3od_train_gelu_forward(hpre.data_ptr(), hmid.data_ptr(), hpre.numel(), self._st()), "s3od_train_gelu_forward")
            y = self._gemm(hmid, wb["down.w"], M, D, I, w["down.b"])
            x2 = self._residual(x1, y, w["ls2"])
        self.saved = dict(B=B, x0=x0, xn1=xn1, q=q, k=k, v=v, lse=lse, ctx=ctx, o=o, x1=x1, xn2=xn2, hpre=hpre, hmid=hmid, y=y)
        return x2.view(B, N, D)

    # ---- backward -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def backward(self, dx2: torch.Tensor):
        """dx2 = d loss / d output, fp32 (B, N, D) -> (d loss / d input, {reference parameter name suffix: gradient})."""
        s, w, wt = self.saved, self.w, self.wt
        B, N, Npad, D, I, H = s["B"], self.N, self.Npad, self.D, self.I, self.H
        M = B * N
        grads: Dict[str, torch.Tensor] = {}
        dx2 = dx2.to(self.dev, torch.float32).contiguous().view(M, D)
        with torch.cuda.device(self.dev):
            st = self._st()
            # ---- MLP branch: x2 = x1 + ls2 * (down(gelu(up(LN2(x1)))))
            grads["layer_scale2.lambda1"], grads["mlp.down_proj.bias"] = self._colsum2(dx2, s["y"], w["ls2"])      # sum dx2 y, ls2 sum dx2
            dy = self._scale_cast(dx2, w["ls2"])
            dhmid = self._gemm(dy, wt["down.w"], M, I, D)                                   # dgrad: dY W
            grads["mlp.down_proj.weight"] = self._wgrad(dy, s["hmid"], M, D, I)                        # wgrad: dY^T X
            dhpre = torch.empty(M, I, dtype=torch.bfloat16, device=self.dev)
            dhpre32 = torch.empty(M, I, dtype=torch.float32, device=self.dev)
            self._ck(self.lib.s3od_train_gelu_backward(s["hpre"].data_ptr(), dhmid.data_ptr(), dhpre.data_ptr(), dhpre32.data_ptr(), dhmid.numel(), st),
                     "s3od_train_gelu_backward")
            grads["mlp.up_proj.bias"] = self._colsum(dhpre32)
            dxn2 = self._gemm(dhpre, wt["up.w"], M, D, I)
            grads["mlp.up_proj.weight"] = self._wgrad(dhpre, s["xn2"], M, I, D)
            dx1, grads["norm2.weight"], grads["norm2.bias"] = self._ln_backward(s["x1"], w["ln2.w"], dxn2, dx2)
            # ---- attention branch: x1 = x0 + ls1 * o_proj(attn(LN1(x0)))
            grads["layer_scale1.lambda1"], grads["attention.o_proj.bias"] = self._colsum2(dx1, s["o"], w["ls1"])
            do = self._scale_cast(dx1, w["ls1"])
            dctx = self._gemm(do, wt["o.w"], M, D, D)
            grads["attention.o_proj.weight"] = self._wgrad(do, s["ctx"], M, D, D)
            dO = self._split_heads(dctx, B)                                                 # [B, H, Npad, 64] bf16
            Oh = self._split_heads(s["ctx"], B)
            BH = B * H
            Dvec = torch.empty(BH * Npad, dtype=torch.float32, device=self.dev)
            self._ck(self.lib.s3od_train_rowdot64(dO.data_ptr(), Oh.data_ptr(), Dvec.data_ptr(), BH * Npad, st), "s3od_train_rowdot64")
            # fused flash backward (csrc/attention_bwd.cuh): P and dA are recomputed tile by tile from q, k, v and the forward's
            # log-sum-exp; nothing of size N x N touches HBM
            dq = torch.empty(BH, Npad, 64, dtype=torch.float32, device=self.dev)
            dk, dv = torch.empty_like(dq), torch.empty_like(dq)
            self._ck(self.lib.s3od_train_attention_backward(s["q"].data_ptr(), s["k"].data_ptr(), s["v"].data_ptr(), dO.data_ptr(), s["lse"].data_ptr(),
                                                            Dvec.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, H, Npad, st),
                     "s3od_train_attention_backward")
            dqkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=self.dev)
            dqkv32 = torch.empty(M, 3 * D, dtype=torch.float32, device=self.dev)
            self._ck(self.lib.s3od_train_qkv_merge_rope_backward_rows(dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), self.cos.data_ptr(), self.sin.data_ptr(),
                                                                      dqkv.data_ptr(), dqkv32.data_ptr(), B, N, Npad, H, self.arch.n_prefix, 1.0 / 8.0,
                                                                      1.0 / self.LOG2E, st), "s3od_train_qkv_merge_rope_backward_rows")
            bq = self._colsum(dqkv32)
            grads["attention.q_proj.bias"], grads["attention.v_proj.bias"] = bq[:D].clone(), bq[2 * D:].clone()      # k_proj has no bias
            dxn1 = self._gemm(dqkv, wt["qkv.w"], M, D, 3 * D)
            dW = self._wgrad(dqkv, s["xn1"], M, 3 * D, D)
            grads["attention.q_proj.weight"], grads["attention.k_proj.weight"], grads["attention.v_proj.weight"] = dW[:D], dW[D:2 * D], dW[2 * D:]
            dx0, grads["norm1.weight"], grads["norm1.bias"] = self._ln_backward(s["x0"], w["ln1.w"], dxn1, dx1)
        return dx0.view(B, N, D), grads


class EncoderTrainer:
    """Forward + backward of the WHOLE encoder path the segmentation head consumes: patch embedding + prefix tokens
    (DINOv3ViTEmbeddings.forward, HF:75-92), the `layers_needed` blocks (`EncoderBlockStep`) and the four un-normed taps
    (`extract_intermediate_features`, /root/reference/src/s3od/model.py:62-86).  `backward` takes d loss / d tap_j (what the DPT
    head's backward would hand over - that head backward is not built yet) and returns the gradient of every encoder parameter
    under its reference name; with a `ParameterLayout` + flat buffer it writes them straight into the flat gradient buffer and
    marks them ready for `GradientAllReduce` in backward order, so the exchange of the late layers overlaps the backward of the
    early ones.  The (image -> patch row) im2col of the 16 x 16 / stride-16 convolution is a pure permutation of the input and is
    done with a torch view + copy (data layout only); every contraction and reduction is a kernel of this library."""

    def __init__(self, sd: Dict[str, torch.Tensor], arch: ArchSpec, image_size: int, device="cuda:0"):
        self.arch, self.dev, self.S = arch, torch.device(device), image_size
        self.prefix = "encoder.model.layer." if any(k.startswith("encoder.model.layer.") for k in sd) else "encoder.layer."
        self.blocks = [EncoderBlockStep(sd, f"{self.prefix}{i}.", arch, image_size, device) for i in range(arch.layers_needed)]
        self.b0 = self.blocks[0]
        e = "encoder.embeddings."
        f = lambda t: t.detach().to(self.dev, torch.float32).contiguous()           # noqa: E731
        D = arch.hidden
        self.patch_w = f(sd[e + "patch_embeddings.weight"]).reshape(D, -1)            # [D, 3*16*16], k = c*256 + ky*16 + kx
        self.patch_wb = self.patch_w.to(torch.bfloat16)
        self.patch_b = f(sd[e + "patch_embeddings.bias"])
        self.prefix_tokens = torch.cat([f(sd[e + "cls_token"]).reshape(1, D), f(sd[e + "register_tokens"]).reshape(-1, D)], 0)
        self.saved = None

    def _im2col(self, x: torch.Tensor) -> torch.Tensor:
        B, g, ps = x.shape[0], self.S // self.arch.patch, self.arch.patch
        return x.view(B, 3, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, 3 * ps * ps).to(torch.bfloat16).contiguous()

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        """x fp32 (B, 3, S, S) -> [tap_0 .. tap_3], each fp32 (B, P, D) (hidden_states[taps][:, n_prefix:])."""
        b0, D, npre = self.b0, self.arch.hidden, self.arch.n_prefix
        x = x.to(self.dev, torch.float32).contiguous()
        B = x.shape[0]
        P = (self.S // self.arch.patch) ** 2
        with torch.cuda.device(self.dev):
            cols = self._im2col(x)
            tok = b0._gemm(cols, self.patch_wb, B * P, D, cols.shape[1], self.patch_b)
            h = torch.empty(B, P + npre, D, dtype=torch.float32, device=self.dev)
            h[:, :npre] = self.prefix_tokens
            h[:, npre:] = tok.view(B, P, D)
            taps = []
            for i, blk in enumerate(self.blocks):
                h = blk.forward(h)
                if (i + 1) in self.arch.taps:
                    taps.append(h[:, npre:])
        self.saved = dict(cols=cols, B=B, P=P)
        return taps

    @torch.no_grad()
    def backward(self, dtaps: Sequence[torch.Tensor], layout: Optional[ParameterLayout] = None, flat_grad: Optional[torch.Tensor] = None,
                 reducer: Optional["GradientAllReduce"] = None) -> Dict[str, torch.Tensor]:
        b0, D, npre = self.b0, self.arch.hidden, self.arch.n_prefix
        B, P, cols = self.saved["B"], self.saved["P"], self.saved["cols"]
        N = P + npre
        grads: Dict[str, torch.Tensor] = {}

        def emit(name: str, g: torch.Tensor):
            grads[name] = g
            if layout is not None and name in layout.by_name:
                layout.view(flat_grad, name).copy_(g.reshape(layout.shapes[name]))
                if reducer is not None:
                    reducer.mark_ready(name)
        tap_of_layer = {t: j for j, t in enumerate(self.arch.taps)}
        with torch.cuda.device(self.dev):
            dh = torch.zeros(B, N, D, dtype=torch.float32, device=self.dev)
            for i in reversed(range(len(self.blocks))):
                if (i + 1) in tap_of_layer:
                    dh[:, npre:] += dtaps[tap_of_layer[i + 1]].to(self.dev, torch.float32)      # fan-in of the residual stream
                dh, g = self.blocks[i].backward(dh)
                dh = dh.contiguous()
                # reverse definition order inside the block = the order autograd produces them
                for name in ("layer_scale2.lambda1", "mlp.down_proj.bias", "mlp.down_proj.weight", "mlp.up_proj.bias", "mlp.up_proj.weight",
                             "norm2.bias", "norm2.weight", "layer_scale1.lambda1", "attention.o_proj.bias", "attention.o_proj.weight",
                             "attention.v_proj.bias", "attention.v_proj.weight", "attention.k_proj.weight", "attention.q_proj.bias",
                             "attention.q_proj.weight", "norm1.bias", "norm1.weight"):
                    emit(f"{self.prefix}{i}.{name}", g[name])
            # embeddings: patch rows -> conv weight / bias (wgrad with the patches as K), prefix rows -> cls / register tokens
            dtok = dh[:, npre:].reshape(B * P, D).contiguous()
            e = "encoder.embeddings."
            emit(e + "patch_embeddings.bias", b0._colsum(dtok))
            emit(e + "patch_embeddings.weight", b0._wgrad(dtok, cols, B * P, D, cols.shape[1]).view(D, 3, self.arch.patch, self.arch.patch))
            dpre = b0._colsum(dh[:, :npre].reshape(B, npre * D).contiguous()).view(npre, D)      # summed over the batch
            emit(e + "register_tokens", dpre[1:].reshape(1, npre - 1, D).clone())
            emit(e + "cls_token", dpre[:1].reshape(1, 1, D).clone())
        return grads
