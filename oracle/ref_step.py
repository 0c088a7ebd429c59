"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DDP-Tutorial hot path.

This module restates, on CPU with the installed torch wheel (the reference's own
and only arithmetic dependency, unpinned by the reference; de-facto pinned here
to torch 2.11.0 / torchvision 0.26.0), the six statements that make up the
reference's training step and the objects they act on:

  * ``Net``                       /root/reference/cifar_example_ddp.py:23-40
                                  (= /root/reference/cifar_example.py:17-34)
  * ``nn.CrossEntropyLoss()``     cifar_example_ddp.py:85   (cifar_example.py:63)
  * ``optim.SGD(lr=1e-3, m=.9)``  cifar_example_ddp.py:86   (cifar_example.py:64)
  * the step body                 cifar_example_ddp.py:101-110 (cifar_example.py:74-83)
  * DDP gradient averaging        cifar_example_ddp.py:83 -> c10d Reducer
                                  (grads / world, all-reduce(sum)); restated as
                                  "each rank forwards its own shard, gradients
                                  are averaged, every replica applies the same
                                  update" (SURVEY.md section 4 / 8e).

Pinning: the reference ships no golden vectors, so this restatement is pinned
against the reference module itself, imported in the build container by
``tests/golden/make_golden.py`` (which needs /root/reference); the vectors it
wrote are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle.py`` everywhere, including on the GPU box where
/root/reference does not exist.

ResNet-18/50 are not in the reference (BASELINE.json substitutes them at
cifar_example.py:60); they come from torchvision.models.resnet (0.26.0,
resnet.py:166-285), the same package the reference imports at cifar_example.py:2.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

LR = 0.001          # cifar_example_ddp.py:86
MOMENTUM = 0.9      # cifar_example_ddp.py:86


# ---------------------------------------------------------------------------
# Net, restated functionally (cifar_example_ddp.py:23-40)
# ---------------------------------------------------------------------------
NET_PARAM_SHAPES: List[Tuple[str, Tuple[int, ...]]] = [
    ("conv1.weight", (6, 3, 5, 5)), ("conv1.bias", (6,)),
    ("conv2.weight", (16, 6, 5, 5)), ("conv2.bias", (16,)),
    ("fc1.weight", (120, 400)), ("fc1.bias", (120,)),
    ("fc2.weight", (84, 120)), ("fc2.bias", (84,)),
    ("fc3.weight", (10, 84)), ("fc3.bias", (10,)),
]


def _default_affine_init(shape: Sequence[int], fan_in: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch's default Conv2d/Linear init (kaiming_uniform(a=sqrt 5) weight, then
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) bias), drawing from the global CPU RNG in
    the same order as the module constructors the reference calls
    (cifar_example_ddp.py:26-31)."""
    w = torch.empty(*shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1.0 / math.sqrt(fan_in)
    b = torch.empty(shape[0]).uniform_(-bound, bound)
    return w, b


def net_init(seed: int) -> Dict[str, torch.Tensor]:
    """Parameters of ``Net()`` constructed right after ``torch.manual_seed(seed)``."""
    torch.manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    p["conv1.weight"], p["conv1.bias"] = _default_affine_init((6, 3, 5, 5), 75)
    p["conv2.weight"], p["conv2.bias"] = _default_affine_init((16, 6, 5, 5), 150)
    p["fc1.weight"], p["fc1.bias"] = _default_affine_init((120, 400), 400)
    p["fc2.weight"], p["fc2.bias"] = _default_affine_init((84, 120), 120)
    p["fc3.weight"], p["fc3.bias"] = _default_affine_init((10, 84), 84)
    return p


def net_forward(p: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """Net.forward, cifar_example_ddp.py:33-40."""
    h = F.max_pool2d(F.relu(F.conv2d(x, p["conv1.weight"], p["conv1.bias"])), 2, 2)
    h = F.max_pool2d(F.relu(F.conv2d(h, p["conv2.weight"], p["conv2.bias"])), 2, 2)
    h = torch.flatten(h, 1)
    h = F.relu(F.linear(h, p["fc1.weight"], p["fc1.bias"]))
    h = F.relu(F.linear(h, p["fc2.weight"], p["fc2.bias"]))
    return F.linear(h, p["fc3.weight"], p["fc3.bias"])


class _FunctionalNet(torch.nn.Module):
    """nn.Module shell around the functional restatement so that the step code
    below is model-agnostic (parameters(), state_dict())."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        super().__init__()
        self._names = [n for n, _ in NET_PARAM_SHAPES]
        for n in self._names:
            self.register_parameter(n.replace(".", "__"), torch.nn.Parameter(params[n].clone()))

    def named_flat(self):
        return [(n, getattr(self, n.replace(".", "__"))) for n in self._names]

    def forward(self, x):
        return net_forward(dict(self.named_flat()), x)


def make_model(name: str, seed: int, dtype: torch.dtype = torch.float32) -> torch.nn.Module:
    """Model constructed under ``torch.manual_seed(seed)``."""
    if name == "net":
        m: torch.nn.Module = _FunctionalNet(net_init(seed))
    elif name in ("resnet18", "resnet50"):
        import torchvision
        torch.manual_seed(seed)
        m = getattr(torchvision.models, name)(num_classes=10)
    else:
        raise ValueError(name)
    m.train()
    return m.to(dtype)


def named_params(model: torch.nn.Module) -> List[Tuple[str, torch.Tensor]]:
    if isinstance(model, _FunctionalNet):
        return [(n, t.detach()) for n, t in model.named_flat()]
    return [(n, t.detach()) for n, t in model.named_parameters()]


# ---------------------------------------------------------------------------
# Synthetic input: N(0,1) images, uniform labels
# ---------------------------------------------------------------------------
def synthetic_batch(gen: torch.Generator, batch: int) -> Tuple[torch.Tensor, torch.Tensor]:
    x = torch.randn(batch, 3, 32, 32, generator=gen)
    y = torch.randint(0, 10, (batch,), generator=gen)
    return x, y


def data_generator(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed + 1)


# ---------------------------------------------------------------------------
# The step (cifar_example.py:74-83 / cifar_example_ddp.py:101-110)
# ---------------------------------------------------------------------------
def cross_entropy_and_grad(logits: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """mean_b(-log softmax(logits)[label]) and its gradient (softmax - onehot)/B,
    i.e. nn.CrossEntropyLoss() forward+backward (cifar_example_ddp.py:85,105-106)."""
    lg = logits.detach().clone().requires_grad_(True)
    loss = F.cross_entropy(lg, labels)
    loss.backward()
    return loss.detach(), lg.grad


def sgd_momentum_update(p: torch.Tensor, g: torch.Tensor, m: Optional[torch.Tensor],
                        lr: float = LR, momentum: float = MOMENTUM) -> Tuple[torch.Tensor, torch.Tensor]:
    """One element-wise SGD(momentum) update exactly as torch.optim.SGD does it
    (sgd.py:_single_tensor_sgd): first step buf = clone(grad); later
    buf = momentum*buf + grad; p = p - lr*buf.  No dampening/nesterov/weight decay."""
    if m is None:
        m = g.clone()
    else:
        m = m * momentum + g
    return p - lr * m, m


def train(model_name: str, *, seed: int = 0, batch: int = 4, steps: int = 50,
          world: int = 1, fixed_batch: bool = True, dtype: torch.dtype = torch.float32,
          keep_grads_of_step0: bool = True, threads: Optional[int] = None) -> Dict[str, object]:
    """Run the reference step ``steps`` times on synthetic data.

    ``world`` > 1 restates DDP: rank r owns rows [r*batch, (r+1)*batch) of the
    global batch, forwards them through its replica (per-rank BatchNorm
    statistics, no SyncBN), gradients are averaged over ranks (Reducer divides by
    world size then all-reduces), all replicas apply the same SGD update.  The
    returned loss list is rank 0's local loss per step (cifar_example_ddp.py:110
    never reduces it) plus ``loss_mean`` = mean over ranks.
    """
    if threads is not None:
        torch.set_num_threads(threads)
    model = make_model(model_name, seed, dtype)
    params = [p for p in model.parameters()]
    opt = torch.optim.SGD(params, lr=LR, momentum=MOMENTUM)
    gen = data_generator(seed)
    gb = batch * world
    x_fixed, y_fixed = synthetic_batch(gen, gb)
    out: Dict[str, object] = {"loss": [], "loss_mean": []}
    init = [(n, t.clone()) for n, t in named_params(model)]
    for step in range(steps):
        if fixed_batch or step == 0:
            x, y = x_fixed, y_fixed
        else:
            x, y = synthetic_batch(gen, gb)
        x = x.to(dtype)
        opt.zero_grad()
        rank_losses = []
        if world == 1:
            loss = F.cross_entropy(model(x), y)
            loss.backward()
            rank_losses.append(float(loss))
        else:
            acc = [torch.zeros_like(p) for p in params]
            for r in range(world):
                for p in params:
                    p.grad = None
                loss = F.cross_entropy(model(x[r * batch:(r + 1) * batch]), y[r * batch:(r + 1) * batch])
                loss.backward()
                rank_losses.append(float(loss))
                for a, p in zip(acc, params):
                    a += p.grad / world          # Reducer: divide first, then sum
            for a, p in zip(acc, params):
                p.grad = a
        if step == 0 and keep_grads_of_step0:
            out["grads0"] = [(n, p.grad.detach().clone()) for (n, _), p in zip(named_params(model), params)]
        opt.step()
        out["loss"].append(rank_losses[0])
        out["loss_mean"].append(sum(rank_losses) / len(rank_losses))
    out["init"] = init
    out["final"] = [(n, t.clone()) for n, t in named_params(model)]
    out["x"], out["y"] = x_fixed, y_fixed
    out["model"] = model
    return out
