"""``optim.SGD(net.parameters(), lr=0.001, momentum=0.9)`` (cifar_example_ddp.py:86,101,107).

Single process: one fused streaming kernel over the flat vectors (20 B/param).  Under
``DistributedDataParallel`` the update is part of the gradient-exchange kernel (see
parallel.py): ``step()`` then only joins the communication stream."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import B2Error, call
from .engine import _stream_ptr


class SGD:
    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0,
                 weight_decay: float = 0.0, nesterov: bool = False, fuse_into_backward: bool = True,
                 bucket_cap_mb: float = 32.0):
        if dampening != 0.0 or weight_decay != 0.0 or nesterov:
            raise B2Error("only the reference's configuration (lr, momentum) is implemented")
        params = list(params)
        if not params:
            raise B2Error("optimizer got an empty parameter list")
        engines = {id(getattr(p, "_b2_engine", None)) for p in params}
        engine = getattr(params[0], "_b2_engine", None)
        if engine is None or len(engines) != 1:
            raise B2Error("SGD expects the parameters() of one b2 model")
        if len(params) != len(engine.param_views):
            raise B2Error("SGD must be given all parameters of the model (flat-vector update)")
        self.engine = engine
        self.lr, self.momentum = float(lr), float(momentum)
        self.state_steps = 0
        self.momentum_buffer = torch.zeros_like(engine.flat_params)
        # fuse_into_backward (default): the update of a gradient bucket is launched from inside
        # backward() on a side stream as soon as the bucket is complete, so most of the optimizer
        # pass runs under the rest of backward and step() only does bookkeeping (parameters are
        # already updated when backward() returns).  False: one pass over the flat vectors at step().
        self.fuse_into_backward = fuse_into_backward
        self.bucket_elems = int(bucket_cap_mb * (1 << 20) / 4)
        self._fused_pending = False
        engine.optimizer = self
        if engine.ddp is not None:
            engine.ddp._attach_optimizer(self)
        elif fuse_into_backward:
            for m in engine.models.values():
                self._configure_model(m)

    def _configure_model(self, m):
        call("b2_model_set_fused_sgd", m.h, 1, C.c_void_p(self.momentum_buffer.data_ptr()), self.lr, self.momentum,
             self.bucket_elems)

    def _before_backward(self, m):
        call("b2_model_ddp_first_step", m.h, 1 if self.state_steps == 0 else 0)
        self._fused_pending = True

    def zero_grad(self, set_to_none: bool = True):
        """Gradients live in the flat bucket and are fully overwritten by every backward
        (cifar_example_ddp.py:101 only has to make sure nothing stale is accumulated)."""
        return None

    def step(self):
        e = self.engine
        first = 1 if self.state_steps == 0 else 0
        if e.ddp is not None:
            e.ddp._optimizer_step(self, first)
        elif self.fuse_into_backward:
            if not self._fused_pending:
                raise B2Error("optimizer.step() without a preceding backward()")
            self._fused_pending = False            # the update already ran inside backward()
            if e.flat_params_bf16 is not None:
                e.mark_weights_clean()
            else:
                e.mark_weights_dirty()
        else:
            mirror = e.flat_params_bf16
            call("b2_sgd", C.c_void_p(e.flat_params.data_ptr()), C.c_void_p(e.flat_grads.data_ptr()),
                 C.c_void_p(self.momentum_buffer.data_ptr()), e.n_params, self.lr, self.momentum, 1.0, first,
                 C.c_void_p(mirror.data_ptr()) if mirror is not None else None, _stream_ptr())
            if mirror is not None:
                e.mark_weights_clean()          # the bf16 mirror was written by the same kernel
            else:
                e.mark_weights_dirty()
        self.state_steps += 1
