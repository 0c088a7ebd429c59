"""``init_distributed(args)`` and ``DistributedDataParallel(net, device_ids=[gpu])``
(/root/reference/cifar_example_ddp.py:42-58, :83) without NCCL on the step path.

Bootstrap (one-time, host side) uses ``torch.distributed`` with the **gloo** backend purely
as a rendezvous to exchange CUDA IPC handles; after that every rank has every peer's
gradient / parameter / flag buffers mapped over NVLink and the per-step exchange is a single
libb2ddp kernel per bucket: all-reduce(sum)/W fused with the SGD-momentum update
(csrc/comm.cu).  No ``torch.distributed`` / NCCL call and no separate optimizer kernel runs
per step.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch

from . import _lib
from ._lib import B2Error, call, lib
from .engine import _stream_ptr

_state = {"rank": 0, "world": 1, "gpu": 0, "distributed": False, "comm": False, "pg": False}
_wrappers: List["DistributedDataParallel"] = []      # live wrappers (their peer mappings are closed in destroy())


def init_distributed(args) -> None:
    """Same contract as the reference's ``init_distributed`` (cifar_example_ddp.py:42-58):
    reads RANK / WORLD_SIZE / LOCAL_RANK, sets ``args.world_size``, ``args.gpu``,
    ``args.distributed``, selects the device, brings up the communicator and barriers."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.world_size = int(os.environ["WORLD_SIZE"])
        args.gpu = int(os.environ["LOCAL_RANK"])
        rank = int(os.environ["RANK"])
    else:
        print("Not using distributed mode")
        args.gpu = 0
        args.distributed = False
        _state.update(rank=0, world=1, gpu=0, distributed=False)
        torch.cuda.set_device(0)
        return
    args.distributed = True
    torch.cuda.set_device(args.gpu)
    # the reference hard-codes 127.0.0.1:29500 (:55-56); keep that as the default but do not
    # clobber what the launcher exported
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group(backend="gloo", world_size=args.world_size, rank=rank)
        _state["pg"] = True
    _state.update(rank=rank, world=args.world_size, gpu=args.gpu, distributed=True)
    call("b2_comm_init", rank, args.world_size, args.gpu)
    _state["comm"] = True
    dist.barrier()


def rank() -> int:
    return _state["rank"]


def world_size() -> int:
    return _state["world"]


class _CommBuffer:
    """A library-owned, IPC-exportable device allocation viewed as a flat fp32 torch tensor."""

    def __init__(self, n_floats: int, device: torch.device):
        p = C.c_void_p()
        call("b2_comm_alloc", n_floats * 4, C.byref(p))
        self.ptr, self.n = p.value, n_floats
        self.tensor = _as_tensor(self.ptr, n_floats, device, self)

    def handle(self) -> bytes:
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        call("b2_comm_export", C.c_void_p(self.ptr), buf)
        return buf.raw


class _Holder:
    def __init__(self, ptr, n, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self._owner = owner


def _as_tensor(ptr: int, n: int, device: torch.device, owner) -> torch.Tensor:
    with torch.cuda.device(device):
        t = torch.as_tensor(_Holder(ptr, n, owner), device=device)
    return t


def _exchange(payload):
    import torch.distributed as dist
    out: List = [None] * _state["world"]
    dist.all_gather_object(out, payload)
    return out


class DistributedDataParallel:
    """Wrapper with the constructor shape of ``torch.nn.parallel.DistributedDataParallel``.

    Construction (one-time): moves the model's flat vectors into peer-mappable memory, maps
    every peer's vectors, replicates rank 0's parameters and buffers on all ranks
    (distributed.py:_sync_module_states semantics: rank 0's random init wins).
    Per step: see ``_optimizer_step``.
    """

    def __init__(self, module, device_ids: Optional[List[int]] = None, algo: int = 0, overlap: bool = True,
                 bucket_cap_mb: float = 32.0):
        if not _state["distributed"]:
            raise B2Error("Default process group has not been initialized: call init_distributed(args) "
                          "under torchrun first (cifar_example_ddp.py:69)")
        self.module = module
        e = module.engine
        if e.device is None:
            raise B2Error("move the model to its GPU before wrapping it (cifar_example_ddp.py:82-83)")
        if device_ids is not None and list(device_ids) != [e.device.index]:
            raise B2Error(f"device_ids={device_ids} does not match the model's device {e.device}")
        self.engine = e
        self.algo = algo
        # overlap=True: the fused all-reduce + SGD of each gradient bucket is launched from inside
        # backward() on a side stream as soon as the bucket is complete (the reference overlaps its
        # NCCL all-reduce with backward the same way, reducer.hpp:524); parameters are therefore
        # already updated when backward() returns and optimizer.step() only does bookkeeping.
        # overlap=False: one fused kernel over the whole vector at optimizer.step().
        self.overlap = overlap
        self.bucket_elems = int(bucket_cap_mb * (1 << 20) / 4)
        self._fused_pending = False
        W, r = _state["world"], _state["rank"]
        dev = e.device
        npad = max((e.n_params + 3) // 4 * 4, 4)
        self._grads = _CommBuffer(npad, dev)
        self._params = _CommBuffer(npad, dev)
        self._bufs = _CommBuffer(max((e.n_buffers + 3) // 4 * 4, 4), dev)
        self._flags = _CommBuffer(lib().b2_comm_flag_bytes() // 4, dev)
        bufs = iter([self._params, self._grads, self._bufs])
        e._allocate(lambda n: next(bufs).tensor[:n])
        e._keep.append(self)
        # ---- exchange IPC handles, map peers
        handles = _exchange((self._grads.handle(), self._params.handle(), self._flags.handle(), self._bufs.handle()))
        self._peer_ptrs = []
        g, p, f, b = [], [], [], []
        for q in range(W):
            if q == r:
                g.append(self._grads.ptr); p.append(self._params.ptr); f.append(self._flags.ptr); b.append(self._bufs.ptr)
                continue
            ptrs = []
            for h in handles[q]:
                out = C.c_void_p()
                call("b2_comm_import", h, C.byref(out))
                ptrs.append(out.value)
            self._peer_ptrs.extend(ptrs)
            g.append(ptrs[0]); p.append(ptrs[1]); f.append(ptrs[2]); b.append(ptrs[3])
        arr = lambda xs: (C.c_void_p * W)(*xs)
        call("b2_comm_register", arr(g), arr(p), arr(f), npad)
        # ---- rank 0's parameters and buffers win (distributed.py:883-890)
        torch.cuda.synchronize(dev)
        _exchange("params-ready")
        if r != 0:
            self._params.tensor.copy_(_as_tensor(p[0], npad, dev, self))
            self._bufs.tensor.copy_(_as_tensor(b[0], self._bufs.n, dev, self))
            torch.cuda.synchronize(dev)
        _exchange("params-copied")
        e.mark_weights_dirty()
        e.ddp = self
        _wrappers.append(self)
        if e.optimizer is not None:
            self._attach_optimizer(e.optimizer)

    def _close_peers(self):
        """Unmap the peers' buffers (cudaIpcCloseMemHandle).  The model must not be used afterwards."""
        for ptr in self._peer_ptrs:
            lib().b2_comm_unimport(C.c_void_p(ptr))
        self._peer_ptrs = []

    def _attach_optimizer(self, opt):
        """Called when an SGD is built over this model's parameters (cifar_example_ddp.py:86)."""
        e = self.engine
        if not self.overlap:
            # an SGD built before the wrapper armed its own fused update inside backward(); under DDP the
            # update belongs to the exchange kernel at step(), so disarm it (otherwise two updates per step)
            for m in e.models.values():
                call("b2_model_set_fused_sgd", m.h, 0, None, 0.0, 0.0, 0)
            return
        for m in e.models.values():
            self._configure_model(m, opt)

    def _configure_model(self, m, opt):
        call("b2_model_set_ddp", m.h, 1, C.c_void_p(opt.momentum_buffer.data_ptr()), opt.lr, opt.momentum, self.algo,
             self.bucket_elems)

    def _before_backward(self, m):
        opt = self.engine.optimizer
        if self.overlap and opt is not None:
            call("b2_model_ddp_first_step", m.h, 1 if opt.state_steps == 0 else 0)
            self._fused_pending = True

    # -- module surface -------------------------------------------------------------------
    def __call__(self, inputs):
        return self.module(inputs)

    forward = __call__

    def parameters(self):
        return self.module.parameters()

    def named_parameters(self):
        return self.module.named_parameters(prefix="module.")

    def state_dict(self):
        return self.module.state_dict(prefix="module.")     # DDP-wrapped key names

    def load_state_dict(self, sd, strict: bool = True):
        return self.module.load_state_dict(sd, strict)

    # -- step hooks ------------------------------------------------------------------------
    def _after_backward(self):
        """Called when backward has been enqueued.  The exchange itself is fused with the
        optimizer update and launched from ``optimizer.step()``."""
        return None

    def _optimizer_step(self, opt, first: int):
        e = self.engine
        mirror = e.flat_params_bf16
        if self.overlap:
            if not self._fused_pending:
                raise B2Error("optimizer.step() without a preceding backward()")
            self._fused_pending = False            # exchange + update already ran inside backward()
        else:
            call("b2_allreduce_sgd", C.c_void_p(opt.momentum_buffer.data_ptr()), e.n_params, opt.lr, opt.momentum,
                 first, self.algo, C.c_void_p(mirror.data_ptr()) if mirror is not None else None, _stream_ptr())
        if mirror is not None:
            e.mark_weights_clean()                 # the same kernel refreshed the local bf16 mirror
        else:
            e.mark_weights_dirty()

    def allreduce_grads(self):
        """Unfused form (gradients averaged in place on every rank) -- for tests and for users
        that inspect ``.grad`` between ``backward()`` and ``step()``."""
        call("b2_allreduce", self.engine.n_params, 1.0 / _state["world"], self.algo, _stream_ptr())


def barrier() -> None:
    """``dist.barrier()`` (cifar_example_ddp.py:58) on the peer-memory flags."""
    if _state["comm"] and _state["world"] > 1:
        call("b2_comm_barrier", _stream_ptr())


def check(*models) -> None:
    """Synchronises and raises ``B2Error`` if a peer flag wait of the communicator or a grid barrier of any
    of ``models`` (modules or DDP wrappers) timed out.  (A timeout also traps the kernel, so the next CUDA
    call fails as well; this is the cheap explicit poll for training loops.)"""
    if _state["comm"]:
        call("b2_comm_check")
    for m in models:
        m.engine.check()


def destroy() -> None:
    if _state["comm"]:
        torch.cuda.synchronize()
        if _state["pg"]:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier()                 # nobody unmaps while a peer may still be inside a comm kernel
        for w in _wrappers:
            w._close_peers()
        _wrappers.clear()
        call("b2_comm_finalize")               # frees this rank's buffers: wrapped models are dead after this
        _state["comm"] = False
    if _state["pg"]:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
        _state["pg"] = False
    _state.update(distributed=False, world=1, rank=0)
