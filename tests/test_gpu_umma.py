"""GPU: the tcgen05/TMA convolution kernels against the CUDA-core implicit-GEMM kernels on
the same bf16 inputs (both accumulate in fp32), through the C-ABI hook b2_conv_test, and the
CUDA-core kernels against torch's fp32 conv on the bf16-rounded inputs (oracle arithmetic)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ddp_tutorial_b200 as b2

pytestmark = pytest.mark.gpu
DEV = 0

# (N, Hi, Wi, Ci, Co, k, stride, pad): every conv shape class of ResNet-18/50 on 32x32 input
SHAPES = [
    (4, 8, 8, 64, 64, 3, 1, 1),
    (2, 16, 16, 64, 64, 3, 1, 1),
    (8, 8, 8, 64, 128, 3, 2, 1),
    (8, 8, 8, 64, 128, 1, 2, 0),
    (16, 4, 4, 128, 128, 3, 1, 1),
    (32, 4, 4, 128, 256, 3, 2, 1),
    (64, 2, 2, 256, 256, 3, 1, 1),
    (128, 2, 2, 256, 512, 3, 2, 1),
    (128, 2, 2, 256, 512, 1, 2, 0),
    (128, 1, 1, 512, 512, 3, 1, 1),
    (4, 8, 8, 64, 256, 1, 1, 0),
    (4, 8, 8, 256, 64, 1, 1, 0),
    (256, 1, 1, 2048, 512, 1, 1, 0),
]


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _run(shape, mode, umma, acc, x, w, dy, out):
    N, Hi, Wi, Ci, Co, k, s, p = shape
    b2._lib.call("b2_conv_test", N, Hi, Wi, Ci, Co, k, s, p, mode, umma, acc, _vp(x), _vp(w), _vp(dy), _vp(out), None)


def _data(shape, seed=0):
    N, Hi, Wi, Ci, Co, k, s, p = shape
    Ho, Wo = (Hi + 2 * p - k) // s + 1, (Wi + 2 * p - k) // s + 1
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Hi, Wi, Ci, generator=g).to(DEV).bfloat16()
    w = (torch.randn(Co, k, k, Ci, generator=g) / (k * k * Ci) ** 0.5).to(DEV)
    dy = torch.randn(N, Ho, Wo, Co, generator=g).to(DEV).bfloat16()
    return x, w, dy, Ho, Wo


@pytest.mark.parametrize("shape", SHAPES)
def test_fprop(shape):
    N, Hi, Wi, Ci, Co, k, s, p = shape
    x, w, dy, Ho, Wo = _data(shape)
    y_u = torch.zeros(N, Ho, Wo, Co, device=DEV, dtype=torch.bfloat16)
    y_s = torch.zeros_like(y_u)
    _run(shape, 0, 1, 0, x, w, None, y_u)
    _run(shape, 0, 0, 0, x, w, None, y_s)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float().permute(0, 3, 1, 2), stride=s, padding=p).permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    assert (y_s.float() - ref).abs().max().item() < 1e-2 * scale          # CUDA-core kernel vs torch fp32
    assert (y_u.float() - ref).abs().max().item() < 1e-2 * scale          # tcgen05 kernel vs torch fp32
    assert (y_u.float() - y_s.float()).abs().max().item() < 8e-3 * scale  # <= 1 bf16 ulp apart


@pytest.mark.parametrize("acc", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_dgrad(shape, acc):
    N, Hi, Wi, Ci, Co, k, s, p = shape
    x, w, dy, Ho, Wo = _data(shape, 1)
    base = (torch.randn(N, Hi, Wi, Ci, device=DEV) if acc else torch.zeros(N, Hi, Wi, Ci, device=DEV)).bfloat16()
    dx_u, dx_s = base.clone(), base.clone()
    # a sparse (1x1 stride-2) dgrad is launched in accumulate mode over a zero-filled buffer
    _run(shape, 1, 1, 1 if (acc or (k == 1 and s == 2)) else 0, None, w, dy, dx_u)
    _run(shape, 1, 0, acc, None, w, dy, dx_s)
    xr = torch.zeros(N, Ci, Hi, Wi, device=DEV, requires_grad=True)
    yr = F.conv2d(xr, w.bfloat16().float().permute(0, 3, 1, 2), stride=s, padding=p)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    ref = xr.grad.permute(0, 2, 3, 1) + base.float()
    scale = ref.abs().max().item()
    assert (dx_s.float() - ref).abs().max().item() < 1.5e-2 * scale
    assert (dx_u.float() - ref).abs().max().item() < 1.5e-2 * scale
    assert (dx_u.float() - dx_s.float()).abs().max().item() < 1e-2 * scale


@pytest.mark.parametrize("shape", SHAPES)
def test_wgrad(shape):
    N, Hi, Wi, Ci, Co, k, s, p = shape
    x, w, dy, Ho, Wo = _data(shape, 2)
    dw_u = torch.full((Co, k, k, Ci), 7.0, device=DEV)     # poison: every element must be written
    dw_s = torch.full((Co, k, k, Ci), 7.0, device=DEV)
    _run(shape, 2, 1, 0, x, w, dy, dw_u)
    _run(shape, 2, 0, 0, x, w, dy, dw_s)
    wr = w.permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wr, stride=s, padding=p).backward(dy.float().permute(0, 3, 1, 2))
    ref = wr.grad.permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    assert (dw_s - ref).abs().max().item() < 1e-3 * scale
    assert (dw_u - ref).abs().max().item() < 1e-3 * scale
