"""Host side of the training engine: flat parameter / gradient / buffer vectors, the
static-graph C models (one per batch size) bound to them, streams and step state.

PyTorch is used here only as the device-memory container type (``torch.Tensor`` views of
flat vectors, ``data_ptr()``) and for the CUDA stream handle; every kernel on the step path
is launched by libb2ddp.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import B2Error, call, lib

_KIND = {"net": _lib.MODEL_NET, "resnet18": _lib.MODEL_RESNET18, "resnet50": _lib.MODEL_RESNET50}
_PREC = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}


@dataclass
class TensorSpec:
    name: str
    offset: int
    shape: Tuple[int, ...]
    stride: Tuple[int, ...]

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _CModel:
    """One libb2ddp model handle (fixed batch size)."""

    def __init__(self, kind: str, batch: int, num_classes: int, precision: str, device: int):
        h = C.c_void_p()
        call("b2_model_create", _KIND[kind], batch, num_classes, _PREC[precision], device, C.byref(h))
        self.h = h
        self.batch = batch

    def specs(self) -> Tuple[List[TensorSpec], int, List[Tuple[str, int, int]], int]:
        l = lib()
        out = []
        for i in range(l.b2_model_num_tensors(self.h)):
            name, off, rank = C.c_char_p(), C.c_int64(), C.c_int()
            shape, stride = (C.c_int64 * 4)(), (C.c_int64 * 4)()
            call("b2_model_tensor_info", self.h, i, C.byref(name), C.byref(off), C.byref(rank), shape, stride)
            r = rank.value
            out.append(TensorSpec(name.value.decode(), off.value, tuple(shape[:r]), tuple(stride[:r])))
        bufs = []
        for i in range(l.b2_model_num_buffers(self.h)):
            name, off, n = C.c_char_p(), C.c_int64(), C.c_int64()
            call("b2_model_buffer_info", self.h, i, C.byref(name), C.byref(off), C.byref(n))
            bufs.append((name.value.decode(), off.value, n.value))
        return out, l.b2_model_param_count(self.h), bufs, l.b2_model_buffer_count(self.h)

    def destroy(self):
        if self.h:
            lib().b2_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Engine:
    """Owns the flat fp32 vectors and dispatches forward/backward/step to the library."""

    def __init__(self, kind: str, num_classes: int = 10, precision: str = "fp32"):
        if kind not in _KIND:
            raise ValueError(f"unknown model kind {kind!r}")
        if precision not in _PREC:
            raise ValueError(f"unknown precision {precision!r}")
        self.kind, self.num_classes = kind, num_classes
        self.precision = "fp32" if kind == "net" else precision
        self.device: Optional[torch.device] = None
        self.models: Dict[int, _CModel] = {}
        self.specs: List[TensorSpec] = []
        self.buffer_specs: List[Tuple[str, int, int]] = []
        self.n_params = 0
        self.n_buffers = 0
        self.flat_params: Optional[torch.Tensor] = None
        self.flat_grads: Optional[torch.Tensor] = None
        self.flat_buffers: Optional[torch.Tensor] = None
        self.flat_params_bf16: Optional[torch.Tensor] = None   # bf16 mirror shared by all batch sizes
        self.param_views: List[torch.Tensor] = []
        self._last_batch: Optional[int] = None
        self._loss_seq = None      # sequence id of the forward the latest criterion() call consumed
        self._fwd_seq = 0          # every forward gets a sequence id; backward() checks the Loss carries the latest
        self.ddp = None            # set by DistributedDataParallel
        self.optimizer = None      # set by SGD when built from our parameters
        self._keep = []            # objects that must outlive the vectors (comm allocations)
        self._logits: Dict[int, torch.Tensor] = {}
        self._dlogits: Dict[int, torch.Tensor] = {}

    # ------------------------------------------------------------------ layout
    def _probe_layout(self, device_index: int):
        m = self._model(1, device_index)
        self.specs, self.n_params, self.buffer_specs, self.n_buffers = m.specs()

    def _model(self, batch: int, device_index: Optional[int] = None) -> _CModel:
        m = self.models.get(batch)
        if m is None:
            if device_index is None:
                device_index = self.device.index
            m = _CModel(self.kind, batch, self.num_classes, self.precision, device_index)
            if not getattr(self, "_graphs", True):
                call("b2_model_set_graphs", m.h, 0)
            if self.flat_params is not None:
                self._bind(m)
            if self.ddp is not None and self.optimizer is not None and self.ddp.overlap:
                self.ddp._configure_model(m, self.optimizer)
            elif self.ddp is None and self.optimizer is not None and self.optimizer.fuse_into_backward:
                self.optimizer._configure_model(m)
            self.models[batch] = m
        return m

    def _bind(self, m: _CModel):
        call("b2_model_bind", m.h, C.c_void_p(self.flat_params.data_ptr()), C.c_void_p(self.flat_grads.data_ptr()),
             C.c_void_p(self.flat_buffers.data_ptr()) if self.n_buffers else None)
        if self.flat_params_bf16 is not None:
            call("b2_model_bind_bf16", m.h, C.c_void_p(self.flat_params_bf16.data_ptr()))

    # ------------------------------------------------------------------ placement
    def to(self, device, host_params: Dict[str, torch.Tensor], host_buffers: Dict[str, torch.Tensor]):
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if dev.type != "cuda":
            raise B2Error("this engine runs on CUDA devices only (no CPU path)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if self.device is not None:
            # torch's .to(same device) is a no-op; re-running placement would reset trained
            # parameters to the host init and (under DDP) detach them from the peer-mapped buffers
            if dev == self.device:
                return self
            raise B2Error(f"model already lives on {self.device}; moving it to {dev} is not supported")
        self.device = dev
        self._probe_layout(dev.index)
        self._allocate(lambda n: torch.zeros(n, dtype=torch.float32, device=dev))
        self.load_named(host_params, host_buffers)
        return self

    def _allocate(self, alloc):
        old = (self.flat_params, self.flat_grads, self.flat_buffers)
        self.flat_params = alloc(max(self.n_params, 4))
        self.flat_grads = alloc(max(self.n_params, 4))
        self.flat_buffers = alloc(max(self.n_buffers, 4))
        if self.precision == "bf16" and self.flat_params_bf16 is None:
            self.flat_params_bf16 = torch.zeros(max(self.n_params, 64), dtype=torch.bfloat16, device=self.flat_params.device)
        if old[0] is not None:
            self.flat_params.copy_(old[0]); self.flat_grads.copy_(old[1]); self.flat_buffers.copy_(old[2])
        self.param_views = []
        for s in self.specs:
            v = torch.as_strided(self.flat_params, s.shape, s.stride, s.offset)
            v.grad = torch.as_strided(self.flat_grads, s.shape, s.stride, s.offset)
            v._b2_engine = self
            v._b2_name = s.name
            self.param_views.append(v)
        for m in self.models.values():
            self._bind(m)

    def load_named(self, params: Dict[str, torch.Tensor], buffers: Optional[Dict[str, torch.Tensor]] = None):
        for s, v in zip(self.specs, self.param_views):
            if s.name in params:
                src = params[s.name]
                if tuple(src.shape) != s.shape:
                    raise B2Error(f"shape mismatch for {s.name}: {tuple(src.shape)} vs {s.shape}")
                v.copy_(src.to(device=self.device, dtype=torch.float32))
        if buffers:
            for name, off, n in self.buffer_specs:
                if name in buffers:
                    self.flat_buffers[off:off + n].copy_(buffers[name].to(device=self.device, dtype=torch.float32).reshape(-1))
        self.mark_weights_dirty()

    def mark_weights_dirty(self):
        for m in self.models.values():
            call("b2_model_mark_weights_dirty", m.h)

    def mark_weights_clean(self):
        for m in self.models.values():
            call("b2_model_mark_weights_clean", m.h)

    # ------------------------------------------------------------------ step pieces
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.device is None:
            raise B2Error("model is not on a CUDA device: call .to(gpu) first (cifar_example_ddp.py:82)")
        if x.device != self.device:
            raise B2Error(f"input is on {x.device}, model on {self.device}")
        if x.dtype != torch.float32 or x.dim() != 4 or tuple(x.shape[1:]) != (3, 32, 32):
            raise B2Error(f"expected float32 [B,3,32,32] input, got {x.dtype} {tuple(x.shape)}")
        if x.shape[0] == 0:
            raise B2Error("empty batch")
        x = x.contiguous()
        m = self._model(x.shape[0])
        # persistent output buffer per batch size: stable addresses let the library replay its
        # captured CUDA graph instead of re-launching ~200 kernels (the returned tensor is
        # overwritten by the next forward of the same batch size)
        logits = self._logits.get(x.shape[0])
        if logits is None:
            logits = torch.empty(x.shape[0], self.num_classes, dtype=torch.float32, device=self.device)
            self._logits[x.shape[0]] = logits
        call("b2_forward", m.h, C.c_void_p(x.data_ptr()), C.c_void_p(logits.data_ptr()), _stream_ptr())
        self._last_batch = x.shape[0]
        self._fwd_seq += 1
        logits._b2_engine = self
        logits._b2_seq = self._fwd_seq
        logits._b2_input = x          # keep the input alive until backward has been enqueued
        return logits

    def backward(self, dlogits: torch.Tensor, seq: Optional[int] = None):
        if self._last_batch is None:
            raise B2Error("backward() before forward()")
        if seq is not None and seq != self._fwd_seq:
            raise B2Error("backward() of a stale loss: another forward ran after the one this loss belongs to "
                          "(the graph is static and keeps only the latest forward's activations)")
        if dlogits.shape[0] != self._last_batch:
            raise B2Error("backward(): dlogits batch does not match the latest forward")
        m = self._model(self._last_batch)
        if self.ddp is not None:
            self.ddp._before_backward(m)
        elif self.optimizer is not None and self.optimizer.fuse_into_backward:
            self.optimizer._before_backward(m)
        call("b2_backward", m.h, C.c_void_p(dlogits.data_ptr()), _stream_ptr())
        if self.ddp is not None:
            self.ddp._after_backward()

    def dlogits_buffer(self, batch: int) -> torch.Tensor:
        d = self._dlogits.get(batch)
        if d is None:
            d = torch.empty(batch, self.num_classes, dtype=torch.float32, device=self.device)
            self._dlogits[batch] = d
        return d

    def set_graphs(self, enable: bool):
        for m in self.models.values():
            call("b2_model_set_graphs", m.h, int(enable))
        self._graphs = enable

    def check(self) -> None:
        """Synchronises; raises if any kernel of this engine's models reported a failure."""
        for m in self.models.values():
            call("b2_model_check", m.h)

    def steps_tracked(self) -> int:
        return sum(lib().b2_model_steps_tracked(m.h) for m in self.models.values())

    def read_activation(self, op: int, grad: bool = False) -> torch.Tensor:
        m = self._model(self._last_batch)
        n, shape = C.c_int64(), (C.c_int * 4)()
        call("b2_model_read_activation", m.h, op, int(grad), None, 0, C.byref(n), shape, _stream_ptr())
        out = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
        call("b2_model_read_activation", m.h, op, int(grad), C.c_void_p(out.data_ptr()), n.value, C.byref(n), shape, _stream_ptr())
        return out

    def num_ops(self) -> int:
        return lib().b2_model_num_ops(self._model(self._last_batch or 1).h)
