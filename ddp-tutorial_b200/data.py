"""Input-side rows of the reference script, B200-native:

  * ``DistributedSampler``  -- same index stream as ``torch.utils.data.DistributedSampler``
    (cifar_example_ddp.py:70,92; torch/utils/data/distributed.py:107-134): epoch-seeded
    permutation, pad to a multiple of the world size, ``indices[rank::world]``.
  * ``DeviceCifarLoader``   -- keeps the uint8 HWC images resident in HBM (50k x 3 KB = 150 MB),
    gathers a batch by index and applies ``ToTensor`` + ``Normalize((.5,.5,.5),(.5,.5,.5))``
    (cifar_example_ddp.py:61-63) in one kernel; no per-batch H2D copy remains.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterator, Optional, Tuple

import torch

from ._lib import B2Error, call
from .engine import _stream_ptr


class DistributedSampler:
    def __init__(self, dataset_len: int, num_replicas: Optional[int] = None, rank: Optional[int] = None,
                 shuffle: bool = True, seed: int = 0, drop_last: bool = False):
        if num_replicas is None or rank is None:
            from . import parallel
            num_replicas = parallel.world_size() if num_replicas is None else num_replicas
            rank = parallel.rank() if rank is None else rank
        if not (0 <= rank < num_replicas):
            raise ValueError(f"Invalid rank {rank}, rank should be in the interval [0, {num_replicas - 1}]")
        self.n, self.num_replicas, self.rank = dataset_len, num_replicas, rank
        self.shuffle, self.seed, self.drop_last, self.epoch = shuffle, seed, drop_last, 0
        if drop_last and self.n % num_replicas != 0:
            self.num_samples = math.ceil((self.n - num_replicas) / num_replicas)
        else:
            self.num_samples = math.ceil(self.n / num_replicas)
        self.total_size = self.num_samples * num_replicas

    def set_epoch(self, epoch: int) -> None:                      # cifar_example_ddp.py:92
        self.epoch = epoch

    def indices(self) -> torch.Tensor:
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            idx = torch.randperm(self.n, generator=g)
        else:
            idx = torch.arange(self.n)
        if not self.drop_last:
            pad = self.total_size - idx.numel()
            if pad > 0:
                reps = math.ceil(pad / idx.numel())
                idx = torch.cat([idx, idx.repeat(reps)[:pad]])
        else:
            idx = idx[:self.total_size]
        return idx[self.rank:self.total_size:self.num_replicas]

    def __iter__(self) -> Iterator[int]:
        return iter(self.indices().tolist())

    def __len__(self) -> int:
        return self.num_samples


def normalize_u8(images_u8: torch.Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> torch.Tensor:
    """uint8 [N,H,W,3] on the device -> fp32 [N,3,H,W], ToTensor + Normalize in one kernel."""
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3 or not images_u8.is_cuda:
        raise B2Error("normalize_u8 expects a CUDA uint8 tensor [N,H,W,3]")
    x = images_u8.contiguous()
    N, H, W, _ = x.shape
    out = torch.empty(N, 3, H, W, dtype=torch.float32, device=x.device)
    m, s = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    call("b2_preprocess_u8", C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), N, H, W, 3, m, s, _stream_ptr())
    return out


class DeviceCifarLoader:
    """Iterates (inputs fp32 [B,3,32,32], labels int64 [B]) batches of a uint8 image set that lives
    in HBM, in the order a ``DistributedSampler`` dictates (the reference's
    ``DataLoader(trainset, batch_size, sampler=sampler_train)``, cifar_example_ddp.py:71)."""

    def __init__(self, images_u8: torch.Tensor, labels: torch.Tensor, batch_size: int, sampler: DistributedSampler, device):
        self.images = images_u8.to(device)
        self.labels = labels.to(device=device, dtype=torch.int64)
        self.batch_size, self.sampler = batch_size, sampler

    def __len__(self) -> int:
        return math.ceil(len(self.sampler) / self.batch_size)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        idx = self.sampler.indices().to(self.images.device)
        for i in range(0, idx.numel(), self.batch_size):
            sel = idx[i:i + self.batch_size]
            yield normalize_u8(self.images.index_select(0, sel)), self.labels.index_select(0, sel)


class Accuracy:
    """Top-1 accuracy counters on the device (replaces ``torchmetrics.Accuracy(dist_sync_on_step=True)``,
    cifar_example_ddp.py:124-136): ``update(outputs, labels)`` per batch, ``compute()`` sums the two
    counters over ranks (host side, once) and returns correct / total."""

    def __init__(self, device):
        self.counts = torch.zeros(2, dtype=torch.int64, device=device)

    def update(self, outputs: torch.Tensor, labels: torch.Tensor) -> None:
        call("b2_argmax_count", C.c_void_p(outputs.data_ptr()), C.c_void_p(labels.contiguous().data_ptr()), outputs.shape[0],
             outputs.shape[1], C.c_void_p(self.counts.data_ptr()), _stream_ptr())

    def compute(self) -> float:
        c = self.counts.cpu()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(c)
        return float(c[0]) / max(1.0, float(c[1]))


class DevicePrefetcher:
    """Overlaps the host->device copy of batch i+1 with the training step on batch i (what the
    reference gets from ``DataLoader`` worker prefetch + ``.to(gpu)``, cifar_example_ddp.py:94-98).

    Wraps any iterable of ``(inputs, labels)`` host tensors (pinned memory for a truly asynchronous
    copy).  Each batch is copied on a private copy stream into one of ``depth`` persistent device
    slots (stable addresses, so the model's captured CUDA graphs are re-used), and the consumer's
    stream waits on the copy's event before using it."""

    def __init__(self, iterable, device, depth: int = 2):
        self.it, self.device, self.depth = iterable, torch.device("cuda", device) if isinstance(device, int) else torch.device(device), depth
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = []

    def _slot(self, k, x, y):
        if len(self.slots) <= k or self.slots[k][0].shape != x.shape:
            s = (torch.empty(x.shape, dtype=x.dtype, device=self.device), torch.empty(y.shape, dtype=y.dtype, device=self.device),
                 torch.cuda.Event(), torch.cuda.Event())
            if len(self.slots) <= k:
                self.slots.append(s)
            else:
                self.slots[k] = s
        return self.slots[k]

    def __iter__(self):
        it = iter(self.it)
        pending = []
        k = 0

        def enqueue():
            nonlocal k
            try:
                x, y = next(it)
            except StopIteration:
                return False
            dx, dy, copied, consumed = self._slot(k % self.depth, x, y)
            self.stream.wait_event(consumed)               # the step that last used this slot is done with it
            with torch.cuda.stream(self.stream):
                dx.copy_(x, non_blocking=True); dy.copy_(y, non_blocking=True)
                copied.record(self.stream)
            pending.append((dx, dy, copied, consumed))
            k += 1
            return True

        for _ in range(self.depth):
            if not enqueue():
                break
        while pending:
            dx, dy, copied, consumed = pending.pop(0)
            torch.cuda.current_stream(self.device).wait_event(copied)
            yield dx, dy
            consumed.record(torch.cuda.current_stream(self.device))   # work enqueued so far has consumed the slot
            enqueue()
