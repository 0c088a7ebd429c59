"""TEST INFRASTRUCTURE ONLY -- torchvision ResNet forward/backward restated op by op with the
bf16 storage points of the CUDA path emulated, so that the benchmarked (bf16) configuration can be
compared tensor by tensor instead of only through the final loss.

The arithmetic is the reference stack's (torch fp32 ops: torchvision/models/resnet.py:59-105
BasicBlock, :108-163 Bottleneck, :266-282 _forward_impl, substituted at
/root/reference/cifar_example.py:60); what is added is a round-to-bf16 at exactly the tensors the
CUDA path stores in bf16:

  * conv weights (the bf16 parameter mirror), the network input (bf16 im2col rows of the stem)
  * every conv output z and every BatchNorm(+residual)(+ReLU) output y -- BatchNorm statistics are
    taken over the ROUNDED z, as the conv epilogue does
  * in backward, the gradient of every such tensor (dz, dy) is rounded to bf16 as well
  * BatchNorm scale/shift, the classifier head, the loss, all accumulations: fp32

Op order follows csrc/model.cu:build_resnet (the downsample branch sits before the block's last
conv), so entry i of the returned list is ops[i] of the C model.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


class _RoundBoth(torch.autograd.Function):
    """value -> bf16 -> fp32 in forward, gradient -> bf16 -> fp32 in backward."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def rb(x: torch.Tensor) -> torch.Tensor:
    return _RoundBoth.apply(x)


def _bn_train(z, gamma, beta, eps=1e-5):
    """training-mode BatchNorm2d on (already rounded) z with biased batch variance, fp32 math."""
    mean = z.mean(dim=(0, 2, 3), keepdim=True)
    var = z.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    return (z - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def forward_ops(model: torch.nn.Module, x: torch.Tensor, emulate: bool = True):
    """Returns (logits, [(name, tensor)] in C-model op order, starting at op 1 (the stem conv)).
    Tensors are NCHW fp32 and part of the autograd graph (retain_grad is set on each)."""
    r = rb if emulate else (lambda t: t)
    w = (lambda p: p.bfloat16().float()) if emulate else (lambda p: p)      # weight mirror: value rounding only
    ops: List[Tuple[str, torch.Tensor]] = []

    def keep(name, t):
        t.retain_grad()
        ops.append((name, t))
        return t

    def conv(name, m, t):
        return keep(name, r(F.conv2d(t, w(m.weight), None, m.stride, m.padding)))

    def bn(name, m, z, relu, res=None):
        y = _bn_train(z, m.weight, m.bias, m.eps)
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        return keep(name, r(y))

    t = x.bfloat16().float() if emulate else x
    t = conv("conv1", model.conv1, t)
    t = bn("bn1", model.bn1, t, True)
    t = keep("maxpool", F.max_pool2d(t, 3, 2, 1))
    for li in range(1, 5):
        layer = getattr(model, f"layer{li}")
        for bi, blk in enumerate(layer):
            pre = f"layer{li}.{bi}"
            idn = t
            bottleneck = hasattr(blk, "conv3")

            def down():
                nonlocal idn
                if blk.downsample is not None:
                    d = conv(pre + ".downsample.0", blk.downsample[0], t)
                    idn = bn(pre + ".downsample.1", blk.downsample[1], d, False)

            o = conv(pre + ".conv1", blk.conv1, t)
            o = bn(pre + ".bn1", blk.bn1, o, True)
            if not bottleneck:
                down()
                o = conv(pre + ".conv2", blk.conv2, o)
                o = bn(pre + ".bn2", blk.bn2, o, True, idn)
            else:
                o = conv(pre + ".conv2", blk.conv2, o)
                o = bn(pre + ".bn2", blk.bn2, o, True)
                down()
                o = conv(pre + ".conv3", blk.conv3, o)
                o = bn(pre + ".bn3", blk.bn3, o, True, idn)
            t = o
    t = keep("avgpool", F.adaptive_avg_pool2d(t, 1))
    logits = F.linear(torch.flatten(t, 1), model.fc.weight, model.fc.bias)
    keep("fc", logits)
    return logits, ops


def step0(model: torch.nn.Module, x: torch.Tensor, y: torch.Tensor, emulate: bool = True) -> Dict[str, object]:
    """One forward + backward (no update).  Returns loss, per-op activations and activation gradients,
    and parameter gradients (torch names)."""
    for p in model.parameters():
        p.grad = None
    logits, ops = forward_ops(model, x, emulate)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    return {
        "loss": float(loss),
        "ops": [(n, t.detach(), t.grad.detach() if t.grad is not None else None) for n, t in ops],
        "grads": {n: p.grad.detach().clone() for n, p in model.named_parameters()},
    }


def local_backward(model: torch.nn.Module, x: torch.Tensor, ours: List[Tuple[torch.Tensor, torch.Tensor]]):
    """Teacher-forced, op-local check of a backward pass.

    ResNet backward on 32x32 inputs amplifies bf16 rounding noise so strongly (the reference stack's own bf16
    gradients have cosine 0.94 to its fp64 gradients) that comparing two complete backward passes tensor by
    tensor only bounds them to that noise.  This routine instead verifies every op in isolation: ``ours[i]`` =
    (value, gradient) of the output tensor of op i+1 AS STORED BY THE CUDA PATH.  For each op the reference
    arithmetic (torch autograd of the op's forward, evaluated on OUR stored input values) is applied to OUR
    stored upstream gradient; the contributions are accumulated per producer tensor exactly as the network
    topology does.  The same walk yields the op-local FORWARD check: each op's output recomputed from OUR stored
    inputs.  Returns (expected gradient per op output, expected parameter gradients, expected output per op)."""
    w = lambda p: p.bfloat16().float()
    exp = [torch.zeros_like(v) for v, _ in ours]
    fwd: List[torch.Tensor] = []
    pgrad: Dict[str, torch.Tensor] = {}
    names = dict(model.named_parameters())
    idx = [0]

    def leaf(k):
        return ours[k][0].detach().clone().requires_grad_(True)

    def vjp(out, ins, params, out_fwd=None):
        """ins: [(leaf tensor, producer op index)], params: [parameter names]; upstream = our grad of this op."""
        k = idx[0]
        fwd.append((out if out_fwd is None else out_fwd).detach())
        up = ours[k][1].reshape(out.shape)
        ts = [t for t, _ in ins] + [names[n] for n in params]
        gs = torch.autograd.grad(out, ts, grad_outputs=up, allow_unused=True)
        for (t, src), g in zip(ins, gs[:len(ins)]):
            if g is not None:
                exp[src] += g.reshape(exp[src].shape)
        for n, g in zip(params, gs[len(ins):]):
            pgrad[n] = g.detach()
        idx[0] += 1
        return k

    def conv(pname, m, src, inp=None):
        t = leaf(src) if inp is None else inp
        out = F.conv2d(t, w(m.weight), None, m.stride, m.padding)
        return vjp(out, [(t, src)] if inp is None else [], [pname + ".weight"])

    def bn(pname, m, src, relu, res_src=None):
        z = leaf(src)
        ins = [(z, src)]
        y = _bn_train(z, m.weight, m.bias, m.eps)
        if res_src is not None:
            r_ = leaf(res_src)
            ins.append((r_, res_src))
            y = y + r_
        out_fwd = F.relu(y) if relu else y
        if relu:
            # backward through the ReLU with the mask of the STORED output (what the CUDA path tests, y > 0): a
            # pre-activation within one bf16 ulp of zero may round to the other side when recomputed here
            y = y * (ours[idx[0]][0].reshape(y.shape) > 0).to(y.dtype)
        return vjp(y, ins, [pname + ".weight", pname + ".bias"], out_fwd)

    k = conv("conv1", model.conv1, None, x.bfloat16().float())
    k = bn("bn1", model.bn1, k, True)
    t = leaf(k)
    k = vjp(F.max_pool2d(t, 3, 2, 1), [(t, k)], [])
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(model, f"layer{li}")):
            pre = f"layer{li}.{bi}"
            t_in = k
            idn = t_in
            bottleneck = hasattr(blk, "conv3")
            o = conv(pre + ".conv1", blk.conv1, t_in)
            o = bn(pre + ".bn1", blk.bn1, o, True)

            def down():
                nonlocal idn
                if blk.downsample is not None:
                    d = conv(pre + ".downsample.0", blk.downsample[0], t_in)
                    idn = bn(pre + ".downsample.1", blk.downsample[1], d, False)

            if not bottleneck:
                down()
                o = conv(pre + ".conv2", blk.conv2, o)
                o = bn(pre + ".bn2", blk.bn2, o, True, idn)
            else:
                o = conv(pre + ".conv2", blk.conv2, o)
                o = bn(pre + ".bn2", blk.bn2, o, True)
                down()
                o = conv(pre + ".conv3", blk.conv3, o)
                o = bn(pre + ".bn3", blk.bn3, o, True, idn)
            k = o
    t = leaf(k)
    k = vjp(F.adaptive_avg_pool2d(t, 1), [(t, k)], [])
    t = leaf(k)
    vjp(F.linear(torch.flatten(t, 1), model.fc.weight, model.fc.bias), [(t, k)], ["fc.weight", "fc.bias"])
    return exp, pgrad, fwd
