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


class _VmmBuffer:
    """A cuMemCreate allocation (shareable as a POSIX file descriptor, bindable to an NVSwitch multicast object)
    viewed as a flat fp32 torch tensor.  ``n_floats`` is rounded up to the allocation granularity."""

    def __init__(self, n_floats: int, device: torch.device, granularity: int):
        self.bytes = (n_floats * 4 + granularity - 1) // granularity * granularity
        p, h, fd = C.c_void_p(), C.c_int64(), C.c_int()
        call("b2_vmm_alloc", device.index, self.bytes, C.byref(p), C.byref(h), C.byref(fd))
        self.ptr, self.handle, self.fd, self.n = p.value, h.value, fd.value, n_floats
        self.tensor = _as_tensor(self.ptr, n_floats, device, self)


def _exchange_fds(my_fds: List[int], root: Optional[int] = None) -> List[List[int]]:
    """File descriptors cannot travel through gloo: pass them over Unix-domain sockets (SCM_RIGHTS).
    root=None: all-to-all (returns, per rank, that rank's descriptors as valid in THIS process; own entry = my_fds);
    root=r: only rank r sends (returns [fds of r] at index r)."""
    import socket
    import torch.distributed as dist
    W, r = _state["world"], _state["rank"]
    base = f"/tmp/b2ddp.{os.environ.get('MASTER_PORT', '0')}.{os.getppid()}"
    path = f"{base}.{r}.sock"
    if os.path.exists(path):
        os.unlink(path)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(W + 1)
    dist.barrier()                                        # every listener exists
    paths = [None] * W
    dist.all_gather_object(paths, path)
    senders = range(W) if root is None else [root]
    if r in senders:
        for q in range(W):
            if q == r:
                continue
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(paths[q])
            socket.send_fds(c, [r.to_bytes(4, "little")], list(my_fds))
            c.close()
    out: List[List[int]] = [[] for _ in range(W)]
    if r in senders:
        out[r] = list(my_fds)
    expect = len([q for q in senders if q != r])
    for _ in range(expect):
        conn, _ = srv.accept()
        msg, fds, _, _ = socket.recv_fds(conn, 16, 16)
        out[int.from_bytes(msg[:4], "little")] = list(fds)
        conn.close()
    dist.barrier()
    srv.close()
    os.unlink(path)
    return out


def _nvls_available(dev: torch.device, want: Optional[bool]):
    """(every rank can bind multicast memory, allocation granularity)."""
    W = _state["world"]
    mc_ok, gran = C.c_int(0), C.c_int64(0)
    want = want if want is not None else os.environ.get("B2_NVLS", "1") != "0"
    ok = False
    if want and W > 1:
        try:
            call("b2_vmm_query", dev.index, C.byref(mc_ok), C.byref(gran))
            ok = bool(mc_ok.value)
        except B2Error:
            ok = False
    return all(_exchange(ok)), gran.value


def _share_vmm(bufs: List["_VmmBuffer"], dev: torch.device):
    """Map every peer's copy of each buffer in ``bufs`` and bind each buffer to a multicast object.
    Returns (ptrs[k][q] = address of rank q's buffer k in this process, mc[k] = multicast address of buffer k)."""
    W, r = _state["world"], _state["rank"]
    K = len(bufs)
    ptrs = [[None] * W for _ in range(K)]
    fds = _exchange_fds([b.fd for b in bufs])
    for k, buf in enumerate(bufs):
        ptrs[k][r] = buf.ptr
        for q in range(W):
            if q == r:
                continue
            out = C.c_void_p()
            call("b2_vmm_import", dev.index, fds[q][k], buf.bytes, C.byref(out))
            lib().b2_vmm_close_fd(fds[q][k])
            ptrs[k][q] = out.value
    # one multicast object per buffer: rank 0 creates, everybody adds its device, then binds + maps
    mch = [C.c_int64() for _ in range(K)]
    if r == 0:
        mfd = [C.c_int() for _ in range(K)]
        for k, buf in enumerate(bufs):
            call("b2_mc_create", W, buf.bytes, C.byref(mch[k]), C.byref(mfd[k]))
        _exchange_fds([m.value for m in mfd], root=0)
        for m in mfd:
            lib().b2_vmm_close_fd(m.value)
    else:
        got = _exchange_fds([], root=0)
        for k in range(K):
            call("b2_mc_import", got[0][k], C.byref(mch[k]))
            lib().b2_vmm_close_fd(got[0][k])
    for k in range(K):
        call("b2_mc_add_device", mch[k].value, dev.index)
    _exchange("mc-added")                          # binding requires ALL devices to have been added
    mc = []
    for k, buf in enumerate(bufs):
        out = C.c_void_p()
        call("b2_mc_bind_and_map", mch[k].value, buf.handle, buf.bytes, dev.index, C.byref(out))
        mc.append(out.value)
        lib().b2_vmm_close_fd(buf.fd)
    _exchange("mc-bound")
    return ptrs, mc


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
                 bucket_cap_mb: float = 32.0, nvls: Optional[bool] = None):
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
        # NVLS (in-switch reduction) needs cuMemCreate memory bound to a multicast object; otherwise cudaMalloc + IPC
        self.nvls, gran = _nvls_available(dev, nvls)
        if self.nvls:
            self._grads = _VmmBuffer(npad, dev, gran)
            self._params = _VmmBuffer(npad, dev, gran)
        else:
            self._grads = _CommBuffer(npad, dev)
            self._params = _CommBuffer(npad, dev)
        self._bufs = _CommBuffer(max((e.n_buffers + 3) // 4 * 4, 4), dev)
        self._flags = _CommBuffer(lib().b2_comm_flag_bytes() // 4, dev)
        bufs = iter([self._params, self._grads, self._bufs])
        e._allocate(lambda n: next(bufs).tensor[:n])
        e._keep.append(self)
        # ---- exchange IPC handles, map peers
        self._peer_ptrs = []
        g, p, f, b = [None] * W, [None] * W, [None] * W, [None] * W
        g[r], p[r], f[r], b[r] = self._grads.ptr, self._params.ptr, self._flags.ptr, self._bufs.ptr

        def import_ipc(h):
            out = C.c_void_p()
            call("b2_comm_import", h, C.byref(out))
            self._peer_ptrs.append(out.value)
            return out.value

        mc_ptrs = None
        if self.nvls:
            ptrs, mc_ptrs = _share_vmm([self._grads, self._params], dev)
            g, p = ptrs[0], ptrs[1]
            handles = _exchange((self._flags.handle(), self._bufs.handle()))
            for q in range(W):
                if q != r:
                    f[q], b[q] = import_ipc(handles[q][0]), import_ipc(handles[q][1])
        else:
            handles = _exchange((self._grads.handle(), self._params.handle(), self._flags.handle(), self._bufs.handle()))
            for q in range(W):
                if q != r:
                    g[q], p[q], f[q], b[q] = (import_ipc(h) for h in handles[q])
        arr = lambda xs: (C.c_void_p * W)(*xs)
        call("b2_comm_register", arr(g), arr(p), arr(f), npad)
        if mc_ptrs is not None:
            call("b2_comm_register_multicast", C.c_void_p(mc_ptrs[0]), C.c_void_p(mc_ptrs[1]))
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
        lib().b2_vmm_release_all()
        _state["comm"] = False
    if _state["pg"]:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
        _state["pg"] = False
    _state.update(distributed=False, world=1, rank=0)
