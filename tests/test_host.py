"""CPU: host-side logic and the C-ABI library surface (no compute without a GPU)."""
import ctypes
import os
import re
import types

import pytest
import torch

import ddp_tutorial_b200 as b2
from ddp_tutorial_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2ddp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    l = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/b2ddp.h but not exported"


def test_binding_covers_header():
    assert set(_lib.EXPORTED_SYMBOLS) == set(_declared_symbols())


def test_version_and_error_string():
    l = b2.lib()
    assert l.b2_version() >= 100
    assert isinstance(l.b2_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_gpu():
    """No CPU fallback: creating a model without a device is an error with a message."""
    h = ctypes.c_void_p()
    rc = b2.lib().b2_model_create(0, 4, 10, 0, 0, ctypes.byref(h))
    assert rc != 0 and len(b2.lib().b2_last_error()) > 0
    net = b2.Net()
    with pytest.raises(b2.B2Error):
        net(torch.zeros(4, 3, 32, 32))
    with pytest.raises(Exception):
        net.to(0)


def test_bad_arguments_are_rejected():
    l = b2.lib()
    assert l.b2_sgd(None, None, None, -1, 0.1, 0.9, 1.0, 0, None, None) != 0
    assert l.b2_sgd(None, None, None, 0, 0.1, 0.9, 1.0, 0, None, None) == 0          # empty is a no-op
    assert l.b2_ce_loss_grad(None, None, 4, 10, None, None, None) != 0
    assert l.b2_model_create(0, 0, 10, 0, 0, ctypes.byref(ctypes.c_void_p())) != 0
    assert l.b2_comm_register(None, None, None, 0) != 0
    assert b"" != l.b2_last_error()


def test_net_init_matches_torch_modules():
    import torch.nn as nn
    torch.manual_seed(7)
    n = b2.Net()
    torch.manual_seed(7)
    c1, c2 = nn.Conv2d(3, 6, 5), nn.Conv2d(6, 16, 5)
    f1, f2, f3 = nn.Linear(400, 120), nn.Linear(120, 84), nn.Linear(84, 10)
    ref = {"conv1": c1, "conv2": c2, "fc1": f1, "fc2": f2, "fc3": f3}
    for k, m in ref.items():
        assert torch.equal(n._host_params[k + ".weight"], m.weight)
        assert torch.equal(n._host_params[k + ".bias"], m.bias)


@pytest.mark.parametrize("name", ["resnet18", "resnet50"])
def test_resnet_init_matches_torchvision(name):
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(3)
    mine = getattr(b2, name)(num_classes=10)
    torch.manual_seed(3)
    ref = getattr(tv.models, name)(num_classes=10)
    sd = ref.state_dict()
    convs = [k for k in sd if k.endswith("weight") and sd[k].dim() == 4] + ["fc.weight", "fc.bias"]
    assert set(convs) == set(mine._host_params)
    for k in convs:
        assert torch.equal(sd[k], mine._host_params[k]), k


def test_init_distributed_without_env(monkeypatch, capsys):
    """cifar_example_ddp.py:46-50: no RANK/WORLD_SIZE => 'Not using distributed mode'."""
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    args = types.SimpleNamespace()
    b2.init_distributed(args)
    assert args.gpu == 0 and args.distributed is False
    assert "Not using distributed mode" in capsys.readouterr().out
    with pytest.raises(b2.B2Error):
        b2.DistributedDataParallel(b2.Net(), device_ids=[0])


def test_optimizer_rejects_foreign_parameters():
    with pytest.raises(b2.B2Error):
        b2.SGD([torch.zeros(3)], lr=0.1)
    with pytest.raises(b2.B2Error):
        b2.SGD([], lr=0.1)


@pytest.mark.parametrize("n,world,shuffle,drop_last", [(50000, 8, True, False), (10000, 8, False, False), (103, 4, True, False),
                                                         (103, 4, True, True), (7, 8, True, False), (64, 1, True, False)])
def test_distributed_sampler_matches_torch(n, world, shuffle, drop_last):
    """Index stream identical to torch.utils.data.DistributedSampler (cifar_example_ddp.py:70,92)."""
    from torch.utils.data.distributed import DistributedSampler as TorchSampler
    ds = list(range(n))
    for rank in (0, world - 1):
        ref = TorchSampler(ds, num_replicas=world, rank=rank, shuffle=shuffle, seed=3, drop_last=drop_last)
        mine = b2.DistributedSampler(n, num_replicas=world, rank=rank, shuffle=shuffle, seed=3, drop_last=drop_last)
        for epoch in (0, 1, 5):
            ref.set_epoch(epoch); mine.set_epoch(epoch)
            assert list(ref) == list(mine)
            assert len(ref) == len(mine)
