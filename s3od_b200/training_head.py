"""Training step of the DPT segmentation head on the CUDA library: forward in TRAIN mode + backward.

`HeadTrainer` mirrors `DPTSegmentationHead.forward` (/root/reference/src/s3od/model.py:217-238 with :193-211 process_encoder_features,
:334-345 ResidualConvUnit, :383-405 FeatureFusionBlock, :455-467 MultiMaskHead, :185-191 classifier head; the training copy
synth_sod/model_training/model.py:84-101 is identical) the way the reference trains it: BatchNorm in train mode with un-synced
batch statistics, the up-sampling BEFORE the 1x1 out_conv, nothing folded.  `backward` is what autograd derives from it.

Every convolution / transposed convolution is the tcgen05 GEMM (`s3od_op_gemm_f32`, bf16 operands, fp32 accumulation) over an
explicit im2col matrix; dgrad contracts the output gradient with the transposed weights and folds the columns back (col2im), wgrad
contracts over the pixels (transposed operands, zero padded).  The kernels between the GEMMs are csrc/train_head.cuh.  This is a
correctness-first form: activations fp32 NHWC, im2col matrices materialised, nothing fused.  Pure permutations / paddings of
tensors (NCHW <-> NHWC of inputs and outputs, weight re-layouts, zero padding to the GEMM's 64 / 128 granularity) use torch views
and copies; all arithmetic is in the library.  Not reproduced: the running_mean / running_var updates of the BatchNorm buffers
(they do not enter the step's gradients).
"""
import ctypes
import math
from typing import Callable, Dict, List, Optional, Sequence

import torch

from .arch import ArchSpec
from .training import _bind_block, _check, _lib, wgrad_plan, wgrad_plan_tn, wgrad_tn


def _bind_head(lib):
    if not getattr(lib, "_head_bound", False):
        vp, ci, cf, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
        lib.s3od_train_im2col.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
        lib.s3od_train_col2im.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp]
        lib.s3od_train_convt_fold.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp]
        lib.s3od_train_convt_unfold.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
        lib.s3od_train_copy_cols.argtypes = [vp, vp, ll, ci, ci, vp, vp]
        lib.s3od_train_bn_workspace_bytes.argtypes = [ci, ci]
        lib.s3od_train_bn_workspace_bytes.restype = ctypes.c_size_t
        lib.s3od_train_bn_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, cf, vp, vp]
        lib.s3od_train_bn_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, vp, vp]
        lib.s3od_train_relu.argtypes = [vp, vp, ll, vp]
        lib.s3od_train_relu_backward.argtypes = [vp, vp, vp, ll, vp]
        lib.s3od_train_add.argtypes = [vp, vp, vp, ll, vp]
        lib.s3od_train_upsample2x.argtypes = [vp, vp, ci, ci, ci, ci, vp]
        lib.s3od_train_upsample2x_backward.argtypes = [vp, vp, ci, ci, ci, ci, vp]
        lib.s3od_train_small_linear.argtypes = [vp, vp, vp, vp, ll, ci, ci, ci, ci, vp]
        lib.s3od_train_small_linear_backward.argtypes = [vp, vp, vp, vp, vp, vp, ll, ci, ci, ci, ci, vp]
        lib.s3od_train_cast_pad.argtypes = [vp, vp, ll, ci, ci, vp]
        lib.s3od_train_cast_slice.argtypes = [vp, vp, ll, ci, ci, vp]
        lib.s3od_op_conv3x3_wgrad_f32.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, vp, vp]
        lib.s3od_train_small_linear_backward_ws.argtypes = [vp, vp, vp, vp, vp, vp, ll, ci, ci, ci, ci, vp, vp]
        lib.s3od_train_small_linear_workspace_bytes.argtypes = [ll, ci, ci]
        lib.s3od_train_small_linear_workspace_bytes.restype = ctypes.c_size_t
        lib.s3od_train_cast_bf16_f32.argtypes = [vp, vp, ll, vp]
        lib._head_bound = True
    return lib


def _up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class _Ops:
    """Thin tensor-allocating wrappers around the C ABI (one CUDA device)."""

    def __init__(self, device):
        self.lib = _bind_head(_bind_block(_lib()))
        self.dev = torch.device(device)
        self.sms = torch.cuda.get_device_properties(self.dev).multi_processor_count

    def st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def ck(self, rc, what):
        _check(self.lib, rc, what)

    def f32(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.dev)

    def gemm(self, a, b, M, N, K):
        c = self.f32(M, N)
        self.ck(self.lib.s3od_op_gemm_f32(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, self.st()), "s3od_op_gemm_f32")
        return c

    def wgrad(self, dy, x, rows, n_out, n_in, n_in_padded):
        """dW fp32 [n_out, n_in] = dy[rows, n_out]^T x[rows, n_in] with the contraction over the rows (pixels) split across the SMs: the
        transpose-free kernel (`training.wgrad_tn`) when both widths are multiples of 64, else transposed zero-padded copies."""
        if n_out % 64 == 0 and n_in % 64 == 0:
            dyb = dy if dy.dtype == torch.bfloat16 else self.cast(dy)
            xb = x if x.dtype == torch.bfloat16 else self.cast(x)
            return wgrad_tn(self.lib, self.ck, dyb.contiguous(), xb.contiguous(), rows, n_out, n_in, self.sms, self.st())
        splits, kpad = wgrad_plan(n_out, n_in_padded, rows, self.sms)
        a = self.transpose_into(dy, rows, n_out, n_out, kpad)
        b = self.transpose_into(x, rows, n_in, n_in_padded, kpad)
        c = self.f32(n_out, n_in_padded)
        ws = self.f32(splits * n_out * n_in_padded) if splits > 1 else None
        self.ck(self.lib.s3od_op_gemm_f32_splitk(a.data_ptr(), b.data_ptr(), c.data_ptr(), n_out, n_in_padded, kpad, splits,
                                                 ws.data_ptr() if ws is not None else None, self.st()), "s3od_op_gemm_f32_splitk")
        return self.copy_cols(c, n_out, n_in, n_in_padded) if n_in_padded != n_in else c

    def conv3x3_wgrad(self, dyb, xb, B, H, W, cin, cout):
        """dW fp32 [cout, 9 * cin] (tap-major) of a 3x3 / stride 1 / pad 1 convolution from bf16 NHWC dy and x (`s3od_op_conv3x3_wgrad_f32`)."""
        splits = wgrad_plan_tn(cout, 9 * cin, B * H * W, self.sms)
        dw = self.f32(cout, 9 * cin)
        ws = self.f32(splits * cout * 9 * cin) if splits > 1 else None
        self.ck(self.lib.s3od_op_conv3x3_wgrad_f32(dyb.data_ptr(), xb.data_ptr(), dw.data_ptr(), B, H, W, cin, cout, splits,
                                                   ws.data_ptr() if ws is not None else None, self.st()), "s3od_op_conv3x3_wgrad_f32")
        return dw

    def transpose_into(self, t, rows, cols, out_rows, rows_padded):
        """t [rows][cols] (fp32 or bf16, dense) -> zeroed bf16 [out_rows >= cols][rows_padded] holding t^T in its first `cols` rows."""
        out = torch.empty(out_rows, rows_padded, dtype=torch.bfloat16, device=self.dev)      # the kernel writes (and zero-pads) rows < cols
        if out_rows > cols:
            out[cols:].zero_()
        self.ck(self.lib.s3od_train_transpose(t.data_ptr(), 1 if t.dtype == torch.float32 else 0, out.data_ptr(), 1, rows, cols, rows_padded,
                                              rows * cols, cols, 1.0, self.st()), "s3od_train_transpose")
        return out

    def cast(self, x):
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=self.dev)
        self.ck(self.lib.s3od_train_scale_cast(x.data_ptr(), None, out.data_ptr(), x.numel(), x.shape[-1], self.st()), "s3od_train_scale_cast")
        return out

    def cast_pad(self, x, rows, c, cp):
        """fp32 [rows, c] -> bf16 [rows, cp], zero columns behind c."""
        if cp == c:
            return self.cast(x.reshape(rows, c))
        out = torch.empty(rows, cp, dtype=torch.bfloat16, device=self.dev)
        self.ck(self.lib.s3od_train_cast_pad(x.data_ptr(), out.data_ptr(), rows, c, cp, self.st()), "s3od_train_cast_pad")
        return out

    def cast_slice(self, xb, rows, c, cp):
        """bf16 [rows, cp] -> fp32 [rows, c]."""
        if cp == c:
            return self.to_f32(xb.reshape(rows, c))
        out = self.f32(rows, c)
        self.ck(self.lib.s3od_train_cast_slice(xb.data_ptr(), out.data_ptr(), rows, c, cp, self.st()), "s3od_train_cast_slice")
        return out

    def to_f32(self, xb):
        out = torch.empty(xb.shape, dtype=torch.float32, device=self.dev)
        self.ck(self.lib.s3od_train_cast_bf16_f32(xb.data_ptr(), out.data_ptr(), xb.numel(), self.st()), "s3od_train_cast_bf16_f32")
        return out

    def conv3x3(self, xb, w_mat, bias, B, H, W, cin, cout):
        """The inference path's implicit-GEMM convolution (TMA im2col, no materialised columns): bf16 NHWC in / out."""
        yb = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device=self.dev)
        if cin == 64 and cout == 64:
            self.ck(self.lib.s3od_op_conv3x3_rows(xb.data_ptr(), w_mat.data_ptr(), bias.data_ptr() if bias is not None else None, yb.data_ptr(),
                                                  B, H, W, 0, self.st()), "s3od_op_conv3x3_rows")
        else:
            self.ck(self.lib.s3od_op_conv3x3(xb.data_ptr(), w_mat.data_ptr(), bias.data_ptr() if bias is not None else None, yb.data_ptr(),
                                             B, H, W, cin, cout, 0, self.st()), "s3od_op_conv3x3")
        return yb

    def colsum(self, a, b=None):
        M, C = a.shape
        ws = torch.empty(self.lib.s3od_train_colsum_workspace_bytes(M, C), dtype=torch.uint8, device=self.dev)
        out = self.f32(C)
        self.ck(self.lib.s3od_train_colsum(a.data_ptr(), b.data_ptr() if b is not None else None, M, C, None, out.data_ptr(), 0, ws.data_ptr(), self.st()),
                "s3od_train_colsum")
        return out

    def copy_cols(self, src, rows, cols, pitch, bias=None):
        out = self.f32(rows, cols)
        self.ck(self.lib.s3od_train_copy_cols(src.data_ptr(), out.data_ptr(), rows, cols, pitch, bias.data_ptr() if bias is not None else None, self.st()),
                "s3od_train_copy_cols")
        return out

    def relu(self, x):
        y = torch.empty_like(x)
        self.ck(self.lib.s3od_train_relu(x.data_ptr(), y.data_ptr(), x.numel(), self.st()), "s3od_train_relu")
        return y

    def relu_backward(self, dy, x):
        dx = torch.empty_like(x)
        self.ck(self.lib.s3od_train_relu_backward(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), self.st()), "s3od_train_relu_backward")
        return dx

    def add(self, a, b):
        out = torch.empty_like(a)
        self.ck(self.lib.s3od_train_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), self.st()), "s3od_train_add")
        return out


Emit = Callable[[str, torch.Tensor], None]


class _Conv:
    """nn.Conv2d(cin, cout, k, stride, pad) on fp32 NHWC activations."""

    def __init__(self, ops: _Ops, name: str, w: torch.Tensor, b: Optional[torch.Tensor], stride: int = 1, pad: int = 0):
        self.ops, self.name, self.stride, self.pad = ops, name, stride, pad
        self.cout, self.cin, self.k = w.shape[0], w.shape[1], w.shape[2]
        dev = ops.dev
        self.K = self.k * self.k * self.cin                               # im2col width, tap-major: (ky*k + kx)*cin + ci
        assert self.K % 64 == 0, name
        wm = w.detach().to(dev, torch.float32).permute(0, 2, 3, 1).reshape(self.cout, self.K)
        self.Np = _up(self.cout, 128)                                    # GEMM N granularity
        self.wf = torch.zeros(self.Np, self.K, dtype=torch.bfloat16, device=dev)
        self.wf[:self.cout] = wm.to(torch.bfloat16)                      # forward B operand [cout (padded), K]
        self.Kc = _up(self.cout, 64)                                     # dgrad contraction (over cout) granularity
        self.Kp = _up(self.K, 128)
        self.wt = torch.zeros(self.Kp, self.Kc, dtype=torch.bfloat16, device=dev)
        self.wt[:self.K, :self.cout] = wm.t().to(torch.bfloat16)         # dgrad B operand [K (padded), cout (padded)]
        self.bias = b.detach().to(dev, torch.float32).contiguous() if b is not None else None
        # 3x3 / stride 1 / pad 1 convolutions with cin % 64 == 0 never materialise an im2col matrix: forward = the inference path's
        # implicit-GEMM convolution (TMA im2col), dgrad = the same kernel on dy with the flipped, transposed weights, wgrad = the
        # implicit weight-gradient kernel (s3od_op_conv3x3_wgrad_f32).  The kernels want output channels in multiples of 128, so
        # cout (forward, and the channel count of dy) and cin (as the OUTPUT of the dgrad convolution) are zero-padded to that
        # granularity in the packed weights and in the bf16 staging copies - the merged mask heads (64 -> 96 at 1024 x 1024) used
        # to cost a 4.8 GB column matrix and a 10.7 GB fp32 dgrad product per batch of 4.
        self.s3 = self.k == 3 and stride == 1 and pad == 1 and self.cin % 64 == 0 and self.cout % 4 == 0
        self.rows64 = self.s3 and self.cin == 64 and self.cout == 64          # row-streaming kernel, needs W % 128 == 0 at run time
        if self.s3:
            self.cp, self.dp = _up(self.cout, 128), _up(self.cin, 128)
            wfull = torch.zeros(self.cp, self.dp, 3, 3, dtype=torch.float32, device=dev)
            wfull[:self.cout, :self.cin] = w.detach().to(dev, torch.float32)
            self.wf_p = wfull[:, :self.cin].permute(0, 2, 3, 1).reshape(self.cp, self.K).to(torch.bfloat16).contiguous()      # [cp, 9 cin]
            self.bias_p = None
            if self.bias is not None:
                self.bias_p = torch.zeros(self.cp, dtype=torch.float32, device=dev)
                self.bias_p[:self.cout] = self.bias
            # w_d[ci][(ky', kx') * cp + co] = w[co][ci][2 - ky'][2 - kx']
            self.wd_p = wfull.flip(2, 3).permute(1, 2, 3, 0).reshape(self.dp, 9 * self.cp).to(torch.bfloat16).contiguous()
            if self.rows64:
                self.wd = self.wd_p[:64].reshape(64, 9, self.cp)[:, :, :64].reshape(64, 9 * 64).contiguous()

    def _cols(self, x, B, H, W, P):
        o = self.ops
        cols = torch.empty(P, self.K, dtype=torch.bfloat16, device=o.dev)
        o.ck(o.lib.s3od_train_im2col(x.data_ptr(), cols.data_ptr(), B, H, W, self.cin, self.k, self.stride, self.pad, o.st()), "s3od_train_im2col")
        return cols

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        o = self.ops
        B, H, W, C = x.shape
        assert C == self.cin
        OH, OW = (H + 2 * self.pad - self.k) // self.stride + 1, (W + 2 * self.pad - self.k) // self.stride + 1
        P = B * OH * OW
        x = x.contiguous()
        if self.s3:
            xb = o.cast(x)
            self.ctx = (None, xb, (B, H, W), (OH, OW))                       # bf16 x: all the weight gradient needs
            if self.rows64 and W % 128 == 0:
                return o.to_f32(o.conv3x3(xb, self.wf[:64], self.bias, B, H, W, 64, 64))
            yb = o.conv3x3(xb, self.wf_p, self.bias_p, B, H, W, self.cin, self.cp)
            return o.cast_slice(yb, P, self.cout, self.cp).view(B, OH, OW, self.cout)
        cols = self._cols(x, B, H, W, P)
        out = o.gemm(cols, self.wf, P, self.Np, self.K)
        y = o.copy_cols(out, P, self.cout, self.Np, self.bias)
        self.ctx = (cols, None, (B, H, W), (OH, OW))
        return y.view(B, OH, OW, self.cout)

    def backward(self, dy: torch.Tensor, emit: Emit, need_dx: bool = True) -> Optional[torch.Tensor]:
        o = self.ops
        cols, x_saved, (B, H, W), (OH, OW) = self.ctx
        P = B * OH * OW
        dy = dy.reshape(P, self.cout).contiguous()
        if self.bias is not None:
            emit(self.name + ".bias", o.colsum(dy))
        if self.s3:
            use_rows = self.rows64 and W % 128 == 0
            cp = 64 if use_rows else self.cp
            dyb = o.cast_pad(dy, P, self.cout, cp)                                         # bf16, channels zero-padded to the kernel granularity
            dW = o.conv3x3_wgrad(dyb, x_saved, B, H, W, self.cin, cp)[:self.cout]          # the nine shifted windows of x are read by TMA
            emit(self.name + ".weight", dW.reshape(self.cout, self.k, self.k, self.cin).permute(0, 3, 1, 2).contiguous())
            if not need_dx:
                return None
            if use_rows:
                return o.to_f32(o.conv3x3(dyb.view(B, H, W, 64), self.wd, None, B, H, W, 64, 64))
            dxb = o.conv3x3(dyb.view(B, H, W, self.cp), self.wd_p, None, B, H, W, self.cp, self.dp)
            return o.cast_slice(dxb, P, self.cin, self.dp).view(B, H, W, self.cin)
        if cols is None:
            cols = self._cols(x_saved, B, H, W, P)
        dW = o.wgrad(dy, cols, P, self.cout, self.K, self.Kp)                              # dY^T cols, pixels as the contraction
        emit(self.name + ".weight", dW.view(self.cout, self.k, self.k, self.cin).permute(0, 3, 1, 2).contiguous())
        if not need_dx:
            return None
        dyb = torch.zeros(P, self.Kc, dtype=torch.bfloat16, device=o.dev)
        dyb[:, :self.cout] = o.cast(dy)                                                  # zero padded to the 64-granular contraction
        dcols = o.gemm(dyb, self.wt, P, self.Kp, self.Kc)
        dx = o.f32(B, H, W, self.cin)
        o.ck(o.lib.s3od_train_col2im(dcols.data_ptr(), dx.data_ptr(), B, H, W, self.cin, self.k, self.stride, self.pad, self.Kp, 0, o.st()),
             "s3od_train_col2im")
        return dx


class _ConvT:
    """nn.ConvTranspose2d(cin, cout, k, stride, pad) on fp32 NHWC activations (weight layout (cin, cout, k, k))."""

    def __init__(self, ops: _Ops, name: str, w: torch.Tensor, b: torch.Tensor, stride: int, pad: int = 0):
        self.ops, self.name, self.stride, self.pad = ops, name, stride, pad
        self.cin, self.cout, self.k = w.shape[0], w.shape[1], w.shape[2]
        dev = ops.dev
        self.N = self.k * self.k * self.cout                              # column (ky*k + kx)*cout + co
        assert self.N % 128 == 0 and self.cin % 128 == 0, name
        wm = w.detach().to(dev, torch.float32).permute(2, 3, 1, 0).reshape(self.N, self.cin)
        self.wf = wm.to(torch.bfloat16).contiguous()                      # forward B operand [N, cin]
        self.wt = wm.t().to(torch.bfloat16).contiguous()                  # dgrad B operand [cin, N]
        self.bias = b.detach().to(dev, torch.float32).contiguous()
        # the mask head's ConvTranspose2d(128 -> 64, k4, s2, p1): forward on the inference path's row-streaming kernel (conv_rows.cuh)
        self.rows = self.cin == 128 and self.cout == 64 and self.k == 4 and stride == 2 and pad == 1
        if self.rows:
            from .weights import convt_rows_weights
            self.wr = convt_rows_weights(w.detach().to(dev, torch.float32)).to(torch.bfloat16).contiguous()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        o = self.ops
        B, H, W, C = x.shape
        OH, OW = (H - 1) * self.stride - 2 * self.pad + self.k, (W - 1) * self.stride - 2 * self.pad + self.k
        P = B * H * W
        xb = o.cast(x.reshape(P, C))
        if self.rows and W % 128 == 0:
            yb = torch.empty(B, OH, OW, self.cout, dtype=torch.bfloat16, device=o.dev)
            o.ck(o.lib.s3od_op_convt_rows(xb.data_ptr(), self.wr.data_ptr(), self.bias.data_ptr(), yb.data_ptr(), B, H, W, 0, o.st()), "s3od_op_convt_rows")
            self.ctx = (xb, (B, H, W))
            return o.to_f32(yb)
        cols = o.gemm(xb, self.wf, P, self.N, C)
        y = o.f32(B, OH, OW, self.cout)
        o.ck(o.lib.s3od_train_convt_fold(cols.data_ptr(), self.bias.data_ptr(), y.data_ptr(), B, H, W, self.cout, self.k, self.stride, self.pad, self.N,
                                         o.st()), "s3od_train_convt_fold")
        self.ctx = (xb, (B, H, W))
        return y

    def backward(self, dy: torch.Tensor, emit: Emit) -> torch.Tensor:
        o = self.ops
        xb, (B, H, W) = self.ctx
        P = B * H * W
        dy = dy.contiguous()
        emit(self.name + ".bias", o.colsum(dy.reshape(-1, self.cout)))
        dcols = torch.empty(P, self.N, dtype=torch.bfloat16, device=o.dev)
        o.ck(o.lib.s3od_train_convt_unfold(dy.data_ptr(), dcols.data_ptr(), B, H, W, self.cout, self.k, self.stride, self.pad, o.st()),
             "s3od_train_convt_unfold")
        dx = o.gemm(dcols, self.wt, P, self.cin, self.N)
        dW = o.wgrad(dcols, xb, P, self.N, self.cin, self.cin)
        emit(self.name + ".weight", dW.view(self.k, self.k, self.cout, self.cin).permute(3, 2, 0, 1).contiguous())
        return dx.view(B, H, W, self.cin)


class _BN:
    """nn.BatchNorm2d in train mode (batch statistics over B*H*W, biased variance, eps 1e-5)."""

    def __init__(self, ops: _Ops, name: str, gamma: torch.Tensor, beta: torch.Tensor, eps: float):
        self.ops, self.name, self.eps = ops, name, eps
        self.gamma = gamma.detach().to(ops.dev, torch.float32).contiguous()
        self.beta = beta.detach().to(ops.dev, torch.float32).contiguous()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        o = self.ops
        C = x.shape[-1]
        rows = x.numel() // C
        xhat, y = torch.empty_like(x), torch.empty_like(x)
        mean, rstd = o.f32(C), o.f32(C)
        self.ws = torch.empty(o.lib.s3od_train_bn_workspace_bytes(rows, C), dtype=torch.uint8, device=o.dev)
        o.ck(o.lib.s3od_train_bn_forward(x.data_ptr(), self.gamma.data_ptr(), self.beta.data_ptr(), xhat.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                         rstd.data_ptr(), rows, C, self.eps, self.ws.data_ptr(), o.st()), "s3od_train_bn_forward")
        self.ctx = (xhat, rstd, rows, C)
        return y

    def backward(self, dy: torch.Tensor, emit: Emit) -> torch.Tensor:
        o = self.ops
        xhat, rstd, rows, C = self.ctx
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        dg, db = o.f32(C), o.f32(C)
        o.ck(o.lib.s3od_train_bn_backward(dy.data_ptr(), xhat.data_ptr(), self.gamma.data_ptr(), rstd.data_ptr(), dx.data_ptr(), dg.data_ptr(),
                                          db.data_ptr(), rows, C, self.ws.data_ptr(), o.st()), "s3od_train_bn_backward")
        emit(self.name + ".bias", db)
        emit(self.name + ".weight", dg)
        return dx


class _RCU:
    """ResidualConvUnit (model.py:334-345): x + bn2(conv2(relu(bn1(conv1(relu(x))))))."""

    def __init__(self, ops, sd, p, eps):
        self.ops = ops
        self.conv1 = _Conv(ops, p + "conv1", sd[p + "conv1.weight"], sd[p + "conv1.bias"], 1, 1)
        self.conv2 = _Conv(ops, p + "conv2", sd[p + "conv2.weight"], sd[p + "conv2.bias"], 1, 1)
        self.bn1 = _BN(ops, p + "bn1", sd[p + "bn1.weight"], sd[p + "bn1.bias"], eps)
        self.bn2 = _BN(ops, p + "bn2", sd[p + "bn2.weight"], sd[p + "bn2.bias"], eps)

    def forward(self, x):
        o = self.ops
        self.x = x
        a = self.bn1.forward(self.conv1.forward(o.relu(x)))
        self.a = a
        return o.add(self.bn2.forward(self.conv2.forward(o.relu(a))), x)

    def backward(self, dy, emit):
        o = self.ops
        d = self.conv2.backward(self.bn2.backward(dy, emit), emit)
        d = o.relu_backward(d, self.a)
        d = self.conv1.backward(self.bn1.backward(d, emit), emit)
        return o.add(o.relu_backward(d, self.x), dy)


class _Fusion:
    """FeatureFusionBlock (model.py:383-405): [x0 + rcu1(skip)] -> rcu2 -> bilinear x2 -> 1x1 out_conv."""

    def __init__(self, ops, sd, p, eps, has_skip: bool):
        self.ops = ops
        self.rcu1 = _RCU(ops, sd, p + "resConfUnit1.", eps) if has_skip else None       # refinenet4 never uses its resConfUnit1 (F8)
        self.rcu2 = _RCU(ops, sd, p + "resConfUnit2.", eps)
        self.out_conv = _Conv(ops, p + "out_conv", sd[p + "out_conv.weight"], sd[p + "out_conv.bias"], 1, 0)

    def forward(self, x0, skip=None):
        o = self.ops
        out = x0 if skip is None else o.add(x0, self.rcu1.forward(skip))
        out = self.rcu2.forward(out)
        B, H, W, C = out.shape
        self.shape = (B, H, W, C)
        up = o.f32(B, 2 * H, 2 * W, C)
        o.ck(o.lib.s3od_train_upsample2x(out.data_ptr(), up.data_ptr(), B, H, W, C, o.st()), "s3od_train_upsample2x")
        return self.out_conv.forward(up)

    def backward(self, dy, emit):
        o = self.ops
        B, H, W, C = self.shape
        dup = self.out_conv.backward(dy, emit).contiguous()
        d = o.f32(B, H, W, C)
        o.ck(o.lib.s3od_train_upsample2x_backward(dup.data_ptr(), d.data_ptr(), B, H, W, C, o.st()), "s3od_train_upsample2x_backward")
        d = self.rcu2.backward(d, emit)
        if self.rcu1 is None:
            return d, None
        return d, self.rcu1.backward(d, emit)                           # (d x0, d skip)


class HeadTrainer:
    def __init__(self, sd: Dict[str, torch.Tensor], arch: ArchSpec, image_size: int, device="cuda:0"):
        self.arch, self.S, self.g = arch, image_size, image_size // arch.patch
        assert self.g % 2 == 0, "the fusion blocks' size= targets are exact 2x up-samplings only for an even patch grid"
        o = self.ops = _Ops(device)
        h, eps, K = "seg_head.", arch.bn_eps, arch.num_outputs
        self.K = K
        self.proj = [_Conv(o, h + f"projects.{i}", sd[h + f"projects.{i}.weight"], sd[h + f"projects.{i}.bias"]) for i in range(4)]
        self.rs0 = _ConvT(o, h + "resize_layers.0", sd[h + "resize_layers.0.weight"], sd[h + "resize_layers.0.bias"], 4)
        self.rs1 = _ConvT(o, h + "resize_layers.1", sd[h + "resize_layers.1.weight"], sd[h + "resize_layers.1.bias"], 2)
        self.rs3 = _Conv(o, h + "resize_layers.3", sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], 2, 1)
        s = h + "scratch."
        self.rn = [_Conv(o, s + f"layer{i + 1}_rn", sd[s + f"layer{i + 1}_rn.weight"], None, 1, 1) for i in range(4)]
        self.ref = {k: _Fusion(o, sd, s + f"refinenet{k}.", eps, has_skip=k != 4) for k in (1, 2, 3, 4)}
        c = h + "classifier_head."
        f = lambda t: t.detach().to(o.dev, torch.float32).contiguous()                       # noqa: E731
        self.cls = {"w1": f(sd[c + "2.weight"]), "b1": f(sd[c + "2.bias"]), "w2": f(sd[c + "4.weight"]), "b2": f(sd[c + "4.bias"])}
        m = h + "mask_head."
        self.mh_c1 = _Conv(o, m + "output_conv1", sd[m + "output_conv1.weight"], sd[m + "output_conv1.bias"], 1, 1)
        self.mh_up = _ConvT(o, m + "upsample_2x.0", sd[m + "upsample_2x.0.weight"], sd[m + "upsample_2x.0.bias"], 2, 1)
        self.mh_c2 = _Conv(o, m + "upsample_2x.2", sd[m + "upsample_2x.2.weight"], sd[m + "upsample_2x.2.bias"], 1, 1)
        # the K mask heads share their input: one 64 -> 32K convolution, then K grouped 32 -> 1 dot products
        self.inter = sd[m + "mask_heads.0.0.weight"].shape[0]
        self.mh_heads = _Conv(o, m + "mask_heads.merged", torch.cat([sd[m + f"mask_heads.{k}.0.weight"] for k in range(K)], 0),
                              torch.cat([sd[m + f"mask_heads.{k}.0.bias"] for k in range(K)], 0), 1, 1)
        self.w2 = f(torch.cat([sd[m + f"mask_heads.{k}.2.weight"].reshape(1, -1) for k in range(K)], 0))    # [K, 32]
        self.b2 = f(torch.cat([sd[m + f"mask_heads.{k}.2.bias"].reshape(1) for k in range(K)], 0))
        self.mask_prefix = m

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, taps: Sequence[torch.Tensor]) -> Dict[str, torch.Tensor]:
        """taps: four fp32 (B, P, D) tensors -> {'pred_masks' (B, K, S, S), 'pred_iou' (B, K)} in train mode."""
        o, g = self.ops, self.g
        B = taps[0].shape[0]
        with torch.cuda.device(o.dev):
            feats = []
            for i, t in enumerate(taps):
                x = self.proj[i].forward(t.to(o.dev, torch.float32).contiguous().view(B, g, g, -1))      # token-major IS NHWC (model.py:206)
                if i == 0:
                    x = self.rs0.forward(x)
                elif i == 1:
                    x = self.rs1.forward(x)
                elif i == 3:
                    x = self.rs3.forward(x)
                feats.append(x)
            l = [self.rn[i].forward(feats[i]) for i in range(4)]
            p4 = self.ref[4].forward(l[3])
            p3 = self.ref[3].forward(p4, l[2])
            p2 = self.ref[2].forward(p3, l[1])
            p1 = self.ref[1].forward(p2, l[0])
            # classifier head: global average pool -> Linear -> ReLU -> Linear
            Bp, H1, W1, C1 = p1.shape
            self.npix1 = H1 * W1
            pooled = torch.stack([o.colsum(p1[b].reshape(self.npix1, C1)) for b in range(B)]) / self.npix1
            z1 = o.f32(B, 64)
            o.ck(o.lib.s3od_train_small_linear(pooled.data_ptr(), self.cls["w1"].data_ptr(), self.cls["b1"].data_ptr(), z1.data_ptr(), B, 64, C1, C1, 0, o.st()),
                 "s3od_train_small_linear")
            a1 = o.relu(z1)
            iou = o.f32(B, self.K)
            o.ck(o.lib.s3od_train_small_linear(a1.data_ptr(), self.cls["w2"].data_ptr(), self.cls["b2"].data_ptr(), iou.data_ptr(), B, self.K, 64, 64, 0, o.st()),
                 "s3od_train_small_linear")
            # mask head
            f0 = self.mh_c1.forward(p1)
            u = self.mh_up.forward(f0)
            f1 = o.relu(u)
            c2 = self.mh_c2.forward(f1)
            f2 = o.relu(c2)                                                                           # model.py:458-460 is the identity here (F6)
            hk = self.mh_heads.forward(f2)
            r = o.relu(hk)
            S = self.S
            P = B * S * S
            logits = o.f32(P, self.K)
            o.ck(o.lib.s3od_train_small_linear(r.data_ptr(), self.w2.data_ptr(), self.b2.data_ptr(), logits.data_ptr(), P, self.K, self.inter,
                                               self.K * self.inter, self.inter, o.st()), "s3od_train_small_linear")
        self.saved = dict(B=B, pooled=pooled, z1=z1, a1=a1, p1_shape=(Bp, H1, W1, C1), u=u, c2=c2, hk=hk, r=r)
        return {"pred_masks": logits.view(B, S, S, self.K).permute(0, 3, 1, 2).contiguous(), "pred_iou": iou}

    @torch.no_grad()
    def backward(self, d_masks: torch.Tensor, d_iou: Optional[torch.Tensor], emit: Optional[Emit] = None):
        """d_masks (B, K, S, S), d_iou (B, K) -> ([d tap_0 .. d tap_3] fp32 (B, P, D), {reference parameter name: gradient})."""
        o, s, K, S = self.ops, self.saved, self.K, self.S
        B = s["B"]
        grads: Dict[str, torch.Tensor] = {}

        def em(name, gten):
            grads[name] = gten
            if emit is not None:
                emit(name, gten)
        with torch.cuda.device(o.dev):
            P = B * S * S
            dlog = d_masks.to(o.dev, torch.float32).permute(0, 2, 3, 1).contiguous().view(P, K)
            dr = o.f32(P, K * self.inter)
            dw2, db2 = o.f32(K, self.inter), o.f32(K)
            ws = torch.empty(o.lib.s3od_train_small_linear_workspace_bytes(P, K, self.inter), dtype=torch.uint8, device=o.dev)
            o.ck(o.lib.s3od_train_small_linear_backward_ws(dlog.data_ptr(), s["r"].data_ptr(), self.w2.data_ptr(), dr.data_ptr(), dw2.data_ptr(), db2.data_ptr(),
                                                           P, K, self.inter, K * self.inter, self.inter, ws.data_ptr(), o.st()),
                 "s3od_train_small_linear_backward_ws")
            m = self.mask_prefix
            for k in range(K):
                em(m + f"mask_heads.{k}.2.bias", db2[k:k + 1].clone())
                em(m + f"mask_heads.{k}.2.weight", dw2[k].reshape(1, self.inter, 1, 1).clone())
            merged: Dict[str, torch.Tensor] = {}
            d = self.mh_heads.backward(o.relu_backward(dr.view(s["hk"].shape), s["hk"]), lambda n, t: merged.__setitem__(n, t))
            for k in range(K):                                            # split the merged 64 -> 32K convolution back into the K heads
                em(m + f"mask_heads.{k}.0.bias", merged[m + "mask_heads.merged.bias"][k * self.inter:(k + 1) * self.inter].clone())
                em(m + f"mask_heads.{k}.0.weight", merged[m + "mask_heads.merged.weight"][k * self.inter:(k + 1) * self.inter].clone())
            d = self.mh_c2.backward(o.relu_backward(d, s["c2"]), em)
            d = self.mh_up.backward(o.relu_backward(d, s["u"]), em)
            dp1 = self.mh_c1.backward(d, em)
            # classifier head
            if d_iou is not None:
                Bp, H1, W1, C1 = s["p1_shape"]
                dio = d_iou.to(o.dev, torch.float32).contiguous()
                da1, dwc2, dbc2 = o.f32(B, 64), o.f32(K, 64), o.f32(K)
                o.ck(o.lib.s3od_train_small_linear_backward(dio.data_ptr(), s["a1"].data_ptr(), self.cls["w2"].data_ptr(), da1.data_ptr(), dwc2.data_ptr(),
                                                            dbc2.data_ptr(), B, K, 64, 64, 0, o.st()), "s3od_train_small_linear_backward")
                dz1 = o.relu_backward(da1, s["z1"])
                dpool, dwc1, dbc1 = o.f32(B, C1), o.f32(64, C1), o.f32(64)
                o.ck(o.lib.s3od_train_small_linear_backward(dz1.data_ptr(), s["pooled"].data_ptr(), self.cls["w1"].data_ptr(), dpool.data_ptr(), dwc1.data_ptr(),
                                                            dbc1.data_ptr(), B, 64, C1, C1, 0, o.st()), "s3od_train_small_linear_backward")
                c = "seg_head.classifier_head."
                em(c + "4.bias", dbc2); em(c + "4.weight", dwc2); em(c + "2.bias", dbc1); em(c + "2.weight", dwc1)
                dp1 = dp1.contiguous()
                for b in range(B):                                        # the mean's gradient: every pixel of image b gets dpool[b] / npix
                    row = (dpool[b] / self.npix1).contiguous()
                    o.ck(o.lib.s3od_train_add_bias(dp1[b].data_ptr(), row.data_ptr(), dp1[b].numel(), C1, o.st()), "s3od_train_add_bias")
            # fusion blocks in reverse
            dp2, dl0 = self.ref[1].backward(dp1, em)
            dp3, dl1 = self.ref[2].backward(dp2, em)
            dp4, dl2 = self.ref[3].backward(dp3, em)
            dl3, _ = self.ref[4].backward(dp4, em)
            dl = [dl0, dl1, dl2, dl3]
            dtaps: List[torch.Tensor] = [None] * 4
            for i in reversed(range(4)):
                dfe = self.rn[i].backward(dl[i], em)
                if i == 0:
                    dfe = self.rs0.backward(dfe, em)
                elif i == 1:
                    dfe = self.rs1.backward(dfe, em)
                elif i == 3:
                    dfe = self.rs3.backward(dfe, em)
                dtaps[i] = self.proj[i].backward(dfe, em).reshape(B, self.g * self.g, -1)
        return dtaps, grads


class TrainStep:
    """One optimisation step of BASELINE.json configs[3] - `SegmentationLightningModule.training_step` + `configure_optimizers`
    (/root/reference/synth_sod/src/synth_sod/model_training/lightning_module.py:183-209, 234-285) without Lightning:
        predictions = model(images)            EncoderTrainer + HeadTrainer (train mode)
        loss = LossModule(predictions, batch)  training.LossModule (forward + backward kernels)
        loss.backward()                        HeadTrainer.backward -> EncoderTrainer.backward, gradients written into the flat
                                               reverse-autograd buffer and marked ready bucket by bucket
        DDP gradient all-reduce                training.GradientAllReduce (overlaps the rest of the backward), mean folded below
        AdamW.step()                           training.FusedAdamW (encoder lr, head lr x 10, weight decay 0.05)
    The parameters live in ONE flat fp32 buffer (`ParameterLayout`); the per-layer GEMM operands are re-packed from it at the start
    of every step (correctness-first: no persistent bf16 shadow yet).  Train-mode RoPE rescaling is not applied: the
    encoder runs the deterministic eval-mode tables, like the parity oracle."""

    def __init__(self, sd: Dict[str, torch.Tensor], arch: ArchSpec, image_size: int, device="cuda:0", lr: float = 1e-5, group=None,
                 bucket_bytes: int = 64 << 20):
        from .training import FusedAdamW, GradientAllReduce, LossModule, ParameterLayout
        self.arch, self.S, self.dev = arch, image_size, torch.device(device)
        prefix = "encoder.model.layer." if any(k.startswith("encoder.model.layer.") for k in sd) else "encoder.layer."
        self.layout = ParameterLayout(arch, bucket_bytes, prefix)
        self.static = {k: v.detach().to(self.dev) for k, v in sd.items() if k not in self.layout.by_name}      # buffers + grad-less parameters
        self.param = self.layout.flatten({k: v for k, v in sd.items() if k in self.layout.by_name}, device=self.dev)
        self.grad = torch.zeros_like(self.param)
        self.reducer = GradientAllReduce(self.layout, self.grad, group)
        self.opt = FusedAdamW(self.layout, self.param, lr=lr)
        self.loss_module = LossModule()
        self.epoch = 0

    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = dict(self.static)
        for s in self.layout.segments:
            sd[s.name] = self.layout.view(self.param, s.name)
        return sd

    @torch.no_grad()
    def step(self, images: torch.Tensor, masks: torch.Tensor):
        """images fp32 (B, 3, S, S), masks fp32 (B, S, S) in [0, 1] -> (loss, loss parts); parameters are updated in place."""
        from .training import EncoderTrainer
        sd = self.state_dict()
        enc = EncoderTrainer(sd, self.arch, self.S, self.dev)
        head = HeadTrainer(sd, self.arch, self.S, self.dev)
        self.reducer.reset()
        self.grad.zero_()

        def emit(name, g):
            if name in self.layout.by_name:
                self.layout.view(self.grad, name).copy_(g.reshape(self.layout.shapes[name]))
                self.reducer.mark_ready(name)
        out = head.forward(enc.forward(images))
        loss, parts, lg, _ = self.loss_module.forward_backward(out, {"masks": masks.to(self.dev)}, self.epoch)
        dtaps, _ = head.backward(lg["pred_masks"], lg["pred_iou"], emit)
        enc.backward(dtaps, self.layout, self.grad, self.reducer)
        self.reducer.finish()
        self.opt.step(self.grad, grad_scale=1.0 / self.reducer.world)
        return loss, parts
