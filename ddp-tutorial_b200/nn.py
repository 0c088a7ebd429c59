"""Model / criterion facades with the call shapes of the reference script.

  Net()                         /root/reference/cifar_example_ddp.py:23-40, :82
  resnet18/50(num_classes=10)   torchvision.models.resnet (substituted at cifar_example.py:60)
  CrossEntropyLoss()            cifar_example_ddp.py:85,105-106,110

The objects only carry host-side state (initial values drawn from torch's CPU RNG in the
same order the reference's constructors draw them, names, shapes); all arithmetic of
forward / loss / backward runs in libb2ddp.so.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, Iterator, List, Optional, Tuple

import torch

from ._lib import B2Error, call
from .engine import Engine, _stream_ptr


# --------------------------------------------------------------------------------------
# initial values (host, one-time): replay of the constructors' RNG consumption
# --------------------------------------------------------------------------------------
def _conv_default(co: int, ci: int, k: int, bias: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """nn.Conv2d.reset_parameters: kaiming_uniform_(a=sqrt(5)) then bias U(+-1/sqrt(fan_in))."""
    w = torch.empty(co, ci, k, k)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    b = None
    if bias:
        bound = 1.0 / math.sqrt(ci * k * k)
        b = torch.empty(co).uniform_(-bound, bound)
    return w, b


def _linear_default(out_f: int, in_f: int) -> Tuple[torch.Tensor, torch.Tensor]:
    w = torch.empty(out_f, in_f)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1.0 / math.sqrt(in_f)
    return w, torch.empty(out_f).uniform_(-bound, bound)


def _net_init(num_classes: int) -> "OrderedDict[str, torch.Tensor]":
    p: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    p["conv1.weight"], p["conv1.bias"] = _conv_default(6, 3, 5, True)        # :26
    p["conv2.weight"], p["conv2.bias"] = _conv_default(16, 6, 5, True)       # :28
    p["fc1.weight"], p["fc1.bias"] = _linear_default(120, 400)               # :29
    p["fc2.weight"], p["fc2.bias"] = _linear_default(84, 120)                # :30
    p["fc3.weight"], p["fc3.bias"] = _linear_default(num_classes, 84)        # :31
    return p


def _resnet_plan(bottleneck: bool, layers: List[int]):
    """(construction-order convs, registration-order convs) as (name, co, ci, k) -- torchvision
    builds a stage's downsample conv *before* its first block but registers it after the
    block's own convs (resnet.py:_make_layer)."""
    construct, register = [("conv1", 64, 3, 7)], [("conv1", 64, 3, 7)]
    inplanes, exp = 64, (4 if bottleneck else 1)
    for L, nblocks in enumerate(layers):
        planes = 64 << L
        for bi in range(nblocks):
            stride = 2 if (bi == 0 and L > 0) else 1
            pre = f"layer{L + 1}.{bi}"
            down = stride != 1 or inplanes != planes * exp
            if bottleneck:
                convs = [(pre + ".conv1", planes, inplanes, 1), (pre + ".conv2", planes, planes, 3),
                         (pre + ".conv3", planes * 4, planes, 1)]
            else:
                convs = [(pre + ".conv1", planes, inplanes, 3), (pre + ".conv2", planes, planes, 3)]
            dconv = [(pre + ".downsample.0", planes * exp, inplanes, 1)] if down else []
            construct += dconv + convs
            register += convs + dconv
            inplanes = planes * exp
    return construct, register, inplanes


def _resnet_init(bottleneck: bool, layers: List[int], num_classes: int):
    construct, register, feat = _resnet_plan(bottleneck, layers)
    for _, co, ci, k in construct:            # first pass: constructor draws that get overwritten
        _conv_default(co, ci, k, False)
    fcw, fcb = _linear_default(num_classes, feat)
    p: Dict[str, torch.Tensor] = {}
    for name, co, ci, k in register:          # resnet.py:208-213, modules() order
        w = torch.empty(co, ci, k, k)
        torch.nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        p[name + ".weight"] = w
    p["fc.weight"], p["fc.bias"] = fcw, fcb
    return p


# --------------------------------------------------------------------------------------
# module facade
# --------------------------------------------------------------------------------------
class Module:
    """What the reference needs from an ``nn.Module``: ``.to(gpu)``, ``__call__``,
    ``parameters()``, ``state_dict()`` (cifar_example_ddp.py:82,104,86,119)."""

    kind = ""

    def __init__(self, num_classes: int = 10, precision: str = "fp32"):
        self.num_classes = num_classes
        self.engine = Engine(self.kind, num_classes, precision)
        self._host_params: Dict[str, torch.Tensor] = {}
        self._host_buffers: Dict[str, torch.Tensor] = {}
        self.training = True

    # -- placement ------------------------------------------------------------------
    def to(self, device):
        self.engine.to(device, self._host_params, self._host_buffers)
        return self

    def cuda(self, device=None):
        return self.to(torch.cuda.current_device() if device is None else device)

    def train(self, mode: bool = True):
        if not mode:
            raise B2Error("eval mode is outside the hot path this build covers (SURVEY.md 8f row 3)")
        return self

    # -- forward ----------------------------------------------------------------------
    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        return self.engine.forward(inputs)

    forward = __call__

    # -- parameters ---------------------------------------------------------------------
    def _require_device(self):
        if self.engine.device is None:
            raise B2Error("call .to(gpu) before using the model (cifar_example_ddp.py:82)")

    def parameters(self) -> Iterator[torch.Tensor]:
        self._require_device()
        return iter(self.engine.param_views)

    def named_parameters(self, prefix: str = ""):
        self._require_device()
        return iter([(prefix + s.name, v) for s, v in zip(self.engine.specs, self.engine.param_views)])

    def named_buffers(self, prefix: str = ""):
        self._require_device()
        e = self.engine
        out = []
        for name, off, n in e.buffer_specs:
            out.append((prefix + name, e.flat_buffers[off:off + n]))
        return iter(out)

    def state_dict(self, prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
        """fp32, torch layouts (conv weights O,I,H,W), torch key order incl. BatchNorm buffers."""
        self._require_device()
        e = self.engine
        sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        bufs = {name: (off, n) for name, off, n in e.buffer_specs}
        tracked = torch.tensor(e.steps_tracked(), dtype=torch.long, device=e.device)
        for s, v in zip(e.specs, e.param_views):
            sd[prefix + s.name] = v.detach().contiguous().clone()
            if s.name.endswith(".bias"):
                stem = s.name[:-len(".bias")]
                if stem + ".running_mean" in bufs:
                    for bn in ("running_mean", "running_var"):
                        off, n = bufs[f"{stem}.{bn}"]
                        sd[f"{prefix}{stem}.{bn}"] = e.flat_buffers[off:off + n].clone()
                    sd[f"{prefix}{stem}.num_batches_tracked"] = tracked.clone()
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        self._require_device()
        e = self.engine
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
        names = {s.name for s in e.specs}
        if strict:
            missing = names - set(sd)
            if missing:
                raise B2Error(f"load_state_dict: missing keys {sorted(missing)[:5]}...")
        e.load_named({k: v for k, v in sd.items() if k in names},
                     {k: v for k, v in sd.items() if k.endswith(("running_mean", "running_var"))})


class Net(Module):
    """The reference's LeNet-style ``Net`` (cifar_example_ddp.py:23-40); fp32."""
    kind = "net"

    def __init__(self, num_classes: int = 10):
        super().__init__(num_classes, "fp32")
        self._host_params = _net_init(num_classes)


class ResNet(Module):
    def __init__(self, kind: str, num_classes: int, precision: str):
        self.kind = kind
        super().__init__(num_classes, precision)
        bottleneck = kind == "resnet50"
        layers = [3, 4, 6, 3] if bottleneck else [2, 2, 2, 2]
        self._host_params = _resnet_init(bottleneck, layers, num_classes)
        # BatchNorm: weight=1, bias=0, running_mean=0, running_var=1 -- filled in on .to()

    def to(self, device):
        super().to(device)
        e = self.engine
        for s, v in zip(e.specs, e.param_views):
            if s.name not in self._host_params:
                v.fill_(1.0 if s.name.endswith(".weight") else 0.0)     # bn gamma / beta
        for name, off, n in e.buffer_specs:
            e.flat_buffers[off:off + n].fill_(1.0 if name.endswith("running_var") else 0.0)
        e.mark_weights_dirty()
        return self


def resnet18(num_classes: int = 1000, precision: str = "bf16") -> ResNet:
    return ResNet("resnet18", num_classes, precision)


def resnet50(num_classes: int = 1000, precision: str = "bf16") -> ResNet:
    return ResNet("resnet50", num_classes, precision)


# --------------------------------------------------------------------------------------
# criterion
# --------------------------------------------------------------------------------------
class Loss:
    """Result of ``criterion(outputs, labels)``: supports ``backward()`` and ``item()``
    (cifar_example_ddp.py:105,106,110)."""

    def __init__(self, engine: Engine, value: torch.Tensor, dlogits: torch.Tensor, keep, seq=None):
        self._engine, self._value, self._dlogits, self._keep = engine, value, dlogits, keep
        self._seq = seq
        self._done = False

    def backward(self):
        if self._done:
            raise B2Error("backward() called twice on the same loss (the graph is static; run forward again)")
        if self._seq is not None and self._seq != self._engine._loss_seq:
            raise B2Error("backward() of a stale loss: criterion() was called again after this loss was created "
                          "(the softmax-gradient buffer is shared per batch size)")
        self._engine.backward(self._dlogits, self._seq)
        self._done = True

    def item(self) -> float:
        return float(self._value.item())      # device->host read + sync, as in the reference

    __float__ = item

    def detach(self) -> torch.Tensor:
        return self._value.reshape(())

    @property
    def device_value(self) -> torch.Tensor:
        return self._value


class RunningLoss:
    """Device-side ``running_loss`` of the reference's print cadence (cifar_example_ddp.py:110-114):
    the criterion kernel adds every step's loss to a device accumulator; the host reads it only when
    it prints (every ``print_every`` steps) instead of synchronising on ``loss.item()`` every step."""

    def __init__(self, device):
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.acc = torch.zeros(2, dtype=torch.float32, device=dev)        # [sum of losses, count]

    def mean_and_reset(self) -> float:
        s, n = self.acc.tolist()                                           # one device->host read + sync
        self.acc.zero_()
        return s / n if n else 0.0


class CrossEntropyLoss:
    """``nn.CrossEntropyLoss()`` with default arguments (mean reduction).  ``running``: optional
    :class:`RunningLoss` that accumulates the loss of every call on the device."""

    def __init__(self, reduction: str = "mean", running: Optional[RunningLoss] = None):
        if reduction != "mean":
            raise B2Error("only the reference's default reduction='mean' is implemented")
        self.running = running

    def __call__(self, outputs: torch.Tensor, labels: torch.Tensor) -> Loss:
        engine = getattr(outputs, "_b2_engine", None)
        if engine is None:
            raise B2Error("criterion expects the logits returned by a b2 model")
        if labels.dtype != torch.int64 or labels.dim() != 1 or labels.shape[0] != outputs.shape[0]:
            raise B2Error(f"labels must be int64 [B]; got {labels.dtype} {tuple(labels.shape)}")
        if labels.device != outputs.device:
            raise B2Error(f"labels on {labels.device}, outputs on {outputs.device}")
        labels = labels.contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=outputs.device)
        dlogits = engine.dlogits_buffer(outputs.shape[0])
        run = self.running
        if run is not None and run.acc.device != outputs.device:
            raise B2Error(f"running-loss accumulator on {run.acc.device}, outputs on {outputs.device}")
        call("b2_ce_loss_grad_running", C.c_void_p(outputs.data_ptr()), C.c_void_p(labels.data_ptr()), outputs.shape[0],
             outputs.shape[1], C.c_void_p(loss.data_ptr()), C.c_void_p(dlogits.data_ptr()),
             C.c_void_p(run.acc.data_ptr()) if run is not None else None, _stream_ptr())
        seq = getattr(outputs, "_b2_seq", None)
        engine._loss_seq = seq
        return Loss(engine, loss, dlogits, (outputs, labels), seq)
