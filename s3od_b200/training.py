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
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, group=None, emulate_world: int = 0):
        self.lib = _bind_peer(_lib())
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
                        self._ptr_arrays[0], self._ptr_arrays[1], self._ptr_arrays[2], self.world, r, m.data_ptr(), v.data_ptr(), lo, hi,
                        self.steps, self.lrs[group] * lr_factor, self.betas[0], self.betas[1], self.eps, self.weight_decay, st),
                        "s3od_ddp_fused_adamw_step")
        self._barrier()                                           # every rank's parameters are visible everywhere

    def link_bytes_per_step(self) -> int:
        """Bytes this GPU reads from + writes to its peers per step: (world-1)/world x (4 B gradient in, 4 + 2 B parameter out)."""
        return int(self.layout.total * (self.world - 1) / self.world * (4 + 6))

    def close(self):
        for p in self._opened:
            self.lib.s3od_peer_close(ctypes.c_void_p(p))
        self._opened = []
        for bufs in self._bufs:
            for b in bufs:
                b.free()
        self._bufs = []
