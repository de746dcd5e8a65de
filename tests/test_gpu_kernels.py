"""Kernel-level checks on a real B200, through the C ABI, against fp32 torch on the same device."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from s3od_b200.engine import load_library
    assert torch.cuda.is_available()
    return load_library()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("M,N,K", [(1, 128, 64), (128, 128, 64), (300, 256, 768), (4101, 768, 768), (8202, 3072, 768), (513, 768, 3072)])
def test_gemm_tcgen05(lib, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    c = torch.full((M, N), float("nan"), device="cuda")
    assert lib.s3od_op_gemm_f32(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, _st()) == 0
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    # bf16 products are exact in fp32; only the accumulation order differs
    assert float((c.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) * max(1.0, K / 768)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gemm_pair_accumulator_handoff_stress(lib, seed):
    """TMEM write-after-read stress for the CTA-pair GEMM's RELAXED remote "accumulator free" arrival (gemm_tc2.cuh): K = 64 is
    ONE k-block, so every main loop is as short as it can be while the epilogue (fp32 stores of a 256 x 256 tile) is at its
    longest - 10 240 CTA tiles (1280 M tiles x 8 N tiles) through two accumulator stages per CTA.  An MMA that overwrote an
    accumulator before the epilogue warps had read it would show up as a wrong block; the result is compared element by element
    with fp64 (bf16 products are exact in fp32, K = 64 sums differ by accumulation order only)."""
    M, N, K = 128 * 1280, 2048, 64
    g = torch.Generator(device="cuda").manual_seed(100 + seed)
    # every tile gets its own magnitude, so a stale / prematurely overwritten accumulator cannot pass as a neighbour's values
    a = (torch.randn(M, K, device="cuda", generator=g) * (1.0 + (torch.arange(M, device="cuda") // 128 % 7).float()[:, None])).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) * (1.0 + (torch.arange(N, device="cuda") // 256 % 5).float()[:, None])).bfloat16()
    c = torch.full((M, N), float("nan"), device="cuda")
    for _ in range(2):                       # the second pass runs with every accumulator stage already used once
        assert lib.s3od_op_gemm_f32(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, _st()) == 0
    torch.cuda.synchronize()
    worst = 0.0
    for r0 in range(0, M, 16384):            # fp64 reference in row blocks (2048 x 16384 doubles at a time)
        ref = a[r0:r0 + 16384].double() @ b.double().t()
        worst = max(worst, float((c[r0:r0 + 16384].double() - ref).abs().max()) / float(ref.abs().max()))
    assert worst <= 2e-5, worst


def test_gemm_rejects_bad_shape(lib):
    a = torch.zeros(8, 64, device="cuda", dtype=torch.bfloat16)
    c = torch.zeros(8, 100, device="cuda")
    assert lib.s3od_op_gemm_f32(a.data_ptr(), a.data_ptr(), c.data_ptr(), 8, 100, 64, _st()) == -1


@pytest.mark.parametrize("M,D", [(1, 768), (37, 768), (4101, 768), (4101, 1024)])
def test_layernorm(lib, M, D):
    g = torch.Generator(device="cuda").manual_seed(M)
    x = torch.randn(M, D, device="cuda", generator=g) * 3 + 0.5
    w = torch.randn(D, device="cuda", generator=g)
    b = torch.randn(D, device="cuda", generator=g)
    y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    assert lib.s3od_op_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, D, 1e-5, _st()) == 0
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (D,), w, b, 1e-5)
    assert float((y.float() - ref.bfloat16().float()).abs().max()) <= 0.04      # at most one bf16 ulp of |y| <= 8
    assert _rel(y, ref) < 3e-3


@pytest.mark.parametrize("B,H,ntok", [(1, 1, 1), (1, 1, 21), (1, 2, 128), (2, 3, 129), (2, 3, 261), (1, 12, 4101)])
def test_attention(lib, B, H, ntok):
    g = torch.Generator(device="cuda").manual_seed(ntok)
    q = (torch.randn(B * H, ntok, 64, device="cuda", generator=g) * 1.5).bfloat16()
    k = (torch.randn(B * H, ntok, 64, device="cuda", generator=g) * 1.5).bfloat16()
    v = torch.randn(B * H, ntok, 64, device="cuda", generator=g).bfloat16()
    qs = (q.float() * (0.125 * 1.4426950408889634)).bfloat16()
    out = torch.full((B * ntok, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    assert lib.s3od_op_attention(qs.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, ntok, _st()) == 0
    torch.cuda.synchronize()
    s = (qs.float() @ k.float().transpose(1, 2)) * 0.6931471805599453
    ref = (torch.softmax(s, -1) @ v.float()).reshape(B, H, ntok, 64).permute(0, 2, 1, 3).reshape(B * ntok, H * 64)
    assert not torch.isnan(out.float()).any()
    assert _rel(out, ref) < 6e-3                      # P and the output are bf16
    assert float((out.float() - ref).abs().max()) < 3e-2


@pytest.mark.parametrize("B,h,w,cin,cout,relu", [(1, 8, 16, 64, 256, 0), (1, 4, 4, 256, 256, 1), (2, 32, 32, 256, 256, 0),
                                                  (1, 20, 24, 128, 256, 1), (1, 2, 2, 1024, 256, 0), (1, 64, 64, 1024, 256, 0)])
def test_conv3x3_implicit_gemm(lib, B, h, w, cin, cout, relu):
    g = torch.Generator(device="cuda").manual_seed(h * w + cin)
    x = torch.randn(B, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    bias = torch.randn(cout, device="cuda", generator=g)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    y = torch.full((B, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    assert lib.s3od_op_conv3x3(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y.data_ptr(), B, h, w, cin, cout, relu, _st()) == 0
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1)
    if relu:
        ref = F.relu(ref)
    assert not torch.isnan(y.float()).any()
    assert float((y.float() - ref).abs().max()) <= 0.02 * max(1.0, float(ref.abs().max()) / 4)
    assert _rel(y, ref) < 3e-3
