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
    # the geometries of the BENCHMARKED configurations (ResNet-18 B=512, ResNet-50 B=256): multi-tile
    # persistent CTAs (256 / 1024 tiles on 148 SMs), 128-wide channel tiles, stride-2 parity classes at > 148 tiles
    (512, 8, 8, 64, 64, 3, 1, 1),
    (512, 16, 16, 192, 64, 1, 1, 0),      # the stem's im2col GEMM (K padded 147 -> 192)
    (512, 8, 8, 64, 128, 3, 2, 1),
    (512, 8, 8, 64, 128, 1, 2, 0),
    (512, 4, 4, 128, 128, 3, 1, 1),
    (512, 2, 2, 256, 256, 3, 1, 1),
    (512, 1, 1, 512, 512, 3, 1, 1),
    (256, 8, 8, 64, 256, 1, 1, 0),
    (256, 8, 8, 256, 64, 1, 1, 0),
    (256, 4, 4, 512, 128, 1, 1, 0),
    (256, 8, 8, 256, 512, 1, 2, 0),
    (256, 2, 2, 1024, 256, 1, 1, 0),
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


@pytest.mark.parametrize("deterministic", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_wgrad(shape, deterministic):
    """deterministic=0: the split-K CTAs add into dW with red.global.add (what the model runs);
    deterministic=1: partials + fixed-order reduce kernel."""
    N, Hi, Wi, Ci, Co, k, s, p = shape
    x, w, dy, Ho, Wo = _data(shape, 2)
    dw_u = torch.full((Co, k, k, Ci), 7.0, device=DEV)     # poison: every element must be written
    dw_s = torch.full((Co, k, k, Ci), 7.0, device=DEV)
    _run(shape, 2, 1, deterministic, x, w, dy, dw_u)
    _run(shape, 2, 0, 0, x, w, dy, dw_s)
    wr = w.permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wr, stride=s, padding=p).backward(dy.float().permute(0, 3, 1, 2))
    ref = wr.grad.permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    assert (dw_s - ref).abs().max().item() < 1e-3 * scale
    assert (dw_u - ref).abs().max().item() < 1e-3 * scale


BN_SHAPES = [s for s in SHAPES if not (s[5] == 1 and s[6] == 2)] + [(64, 8, 8, 64, 128, 1, 2, 0)]


@pytest.mark.parametrize("relu,with_res", [(1, 0), (1, 1), (0, 0)])
@pytest.mark.parametrize("shape", BN_SHAPES)
def test_fprop_fused_batchnorm(shape, relu, with_res):
    """conv + training-mode BatchNorm (+residual) (+ReLU) in ONE launch (grid barrier between the statistics and
    the normalisation) against F.batch_norm(F.conv2d(.)) in fp32 on the same bf16-rounded operands; the
    BatchNorm reads the bf16-rounded conv output, as the CUDA path does."""
    N, Hi, Wi, Ci, Co, k, s, p = shape
    x, w, dy, Ho, Wo = _data(shape, 3)
    g = torch.Generator().manual_seed(11)
    gamma = (1.0 + 0.2 * torch.randn(Co, generator=g)).to(DEV)
    beta = (0.3 * torch.randn(Co, generator=g)).to(DEV)
    res = torch.randn(N, Ho, Wo, Co, generator=g).to(DEV).bfloat16() if with_res else None
    z = torch.full((N, Ho, Wo, Co), 3.0, device=DEV, dtype=torch.bfloat16)
    y = torch.full((N, Ho, Wo, Co), 3.0, device=DEV, dtype=torch.bfloat16)
    mean, invstd = torch.zeros(Co, device=DEV), torch.zeros(Co, device=DEV)
    info = (C.c_int * 3)()
    b2._lib.call("b2_conv_bn_test", N, Hi, Wi, Ci, Co, k, s, p, relu, _vp(x), _vp(w), _vp(res), _vp(gamma), _vp(beta), _vp(z), _vp(y),
                 _vp(mean), _vp(invstd), info, None, None)
    zr = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float().permute(0, 3, 1, 2), stride=s, padding=p)
    zs = zr.abs().max().item()
    assert (z.float().permute(0, 3, 1, 2) - zr).abs().max().item() < 1e-2 * zs
    zb = z.float().permute(0, 3, 1, 2)                        # statistics over what the kernel stored
    m_ref = zb.mean(dim=(0, 2, 3)); v_ref = zb.var(dim=(0, 2, 3), unbiased=False)
    np.testing.assert_allclose(mean.cpu().numpy(), m_ref.cpu().numpy(), rtol=1e-3, atol=1e-4 * zs)
    np.testing.assert_allclose(invstd.cpu().numpy(), torch.rsqrt(v_ref + 1e-5).cpu().numpy(), rtol=2e-3)
    yr = F.batch_norm(zb, None, None, gamma, beta, True, 0.1, 1e-5)
    if with_res:
        yr = yr + res.float().permute(0, 3, 1, 2)
    if relu:
        yr = F.relu(yr)
    ys = yr.abs().max().item()
    assert (y.float().permute(0, 3, 1, 2) - yr).abs().max().item() < 1e-2 * ys, (info[0], info[1], info[2])
