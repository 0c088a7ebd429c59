"""CPU: host-side logic and the C-ABI library surface (no compute without a GPU)."""
import ctypes
import numpy as np
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


# ---- metadata-only model instances (device < 0): layout and schedules without a GPU ------------------
def _dry_model(kind, batch, prec):
    h = ctypes.c_void_p()
    _lib.call("b2_model_create", {"net": 0, "resnet18": 1, "resnet50": 2}[kind], batch, 10, {"fp32": 0, "bf16": 1}[prec], -1,
              ctypes.byref(h))
    return h


def _param_table(h):
    l = b2.lib()
    out = []
    for i in range(l.b2_model_num_tensors(h)):
        name, off, rank = ctypes.c_char_p(), ctypes.c_int64(), ctypes.c_int()
        shape, stride = (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)()
        _lib.call("b2_model_tensor_info", h, i, ctypes.byref(name), ctypes.byref(off), ctypes.byref(rank), shape, stride)
        out.append((name.value.decode(), off.value, tuple(shape[:rank.value]), tuple(stride[:rank.value])))
    return out


@pytest.mark.parametrize("kind", ["resnet18", "resnet50"])
def test_param_table_matches_torchvision(kind):
    tv = pytest.importorskip("torchvision")
    ref = getattr(tv.models, kind)(num_classes=10)
    h = _dry_model(kind, 64, "bf16")
    table = _param_table(h)
    assert [(n, tuple(p.shape)) for n, p in ref.named_parameters()] == [(n, s) for n, _, s, _ in table]
    # flat layout: ascending, 64-element aligned, non-overlapping; conv weights are O,H,W,I strided views
    end = 0
    for name, off, shape, stride in table:
        assert off % 64 == 0 and off >= end
        numel = int(np.prod(shape))
        end = off + numel
        if len(shape) == 4:
            O, I, H, W = shape
            assert stride == (H * W * I, 1, W * I, I), name
    assert end <= b2.lib().b2_model_param_count(h)
    # buffers: running_mean / running_var of every BatchNorm, torch names
    names = []
    for i in range(b2.lib().b2_model_num_buffers(h)):
        nm, off, n = ctypes.c_char_p(), ctypes.c_int64(), ctypes.c_int64()
        _lib.call("b2_model_buffer_info", h, i, ctypes.byref(nm), ctypes.byref(off), ctypes.byref(n))
        names.append(nm.value.decode())
    assert names == [n for n, _ in ref.named_buffers() if not n.endswith("num_batches_tracked")]
    # a metadata-only instance refuses to execute
    assert b2.lib().b2_forward(h, ctypes.c_void_p(8), None, None) != 0
    b2.lib().b2_model_destroy(h)


def test_net_param_table_and_flatten_order():
    h = _dry_model("net", 4, "fp32")
    table = _param_table(h)
    assert [n for n, *_ in table] == ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "fc1.weight", "fc1.bias",
                                      "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"]
    assert dict((n, s) for n, _, s, _ in table)["fc1.weight"] == (120, 400)
    b2.lib().b2_model_destroy(h)


@pytest.mark.parametrize("kind,bucket", [("resnet18", 2 << 20), ("resnet18", 8 << 20), ("resnet50", 2 << 20), ("net", 4096)])
def test_bucket_plan_partitions_the_flat_vector(kind, bucket):
    """DDP buckets (reducer.hpp:30-31 analogue): contiguous suffixes of the flat gradient vector in the
    order backward completes them; together they cover [0, n) exactly once."""
    h = _dry_model(kind, 64, "bf16" if kind != "net" else "fp32")
    n = b2.lib().b2_model_param_count(h)
    offs, cnts, after = (ctypes.c_int64 * 64)(), (ctypes.c_int64 * 64)(), (ctypes.c_int * 64)()
    nb = b2.lib().b2_model_bucket_plan(h, bucket, 64, offs, cnts, after)
    assert 1 <= nb <= 64
    ivals = sorted((offs[i], offs[i] + cnts[i]) for i in range(nb))
    assert ivals[0][0] == 0 and ivals[-1][1] == n
    for (a0, a1), (b0, b1) in zip(ivals, ivals[1:]):
        assert a1 == b0
    # produced in descending-offset order, by ops in descending index order (backward order)
    assert [offs[i] for i in range(nb)] == sorted([offs[i] for i in range(nb)], reverse=True)
    assert [after[i] for i in range(nb)] == sorted([after[i] for i in range(nb)], reverse=True)
    if kind == "net":
        # every bucket but the last one is at least the requested size
        assert all(cnts[i] >= bucket for i in range(nb - 1))
    else:
        # a bucket always closes right before the stem's backward tail (model.cu:compute_buckets): the last bucket is
        # exactly the stem conv + its BatchNorm, the one before it is the remainder, all others reach the cap
        table = _param_table(h)
        stem_elems = sum(-(-int(np.prod(shape)) // 64) * 64 for name, _, shape, _ in table if name in ("conv1.weight", "bn1.weight", "bn1.bias"))
        assert offs[nb - 1] == 0 and cnts[nb - 1] == stem_elems, (cnts[nb - 1], stem_elems)
        assert all(cnts[i] >= bucket for i in range(nb - 2))
    b2.lib().b2_model_destroy(h)


def test_op_schedule_and_kernel_selection():
    """Static graph of ResNet-18 at the benchmark shape: 21 conv/linear ops, 20 BatchNorms; in bf16 mode
    every 64-multiple conv runs on the tcgen05 path, the 3-channel stem on the im2col+tcgen05 path, the
    classifier on the fp32 head kernels; in fp32 mode nothing does."""
    for prec, expect_umma in (("bf16", 19), ("fp32", 0)):
        h = _dry_model("resnet18", 512, prec)
        nops = b2.lib().b2_model_num_ops(h)
        types, umma, stem, fused, fc = [], 0, 0, 0, 0
        for i in range(nops):
            t, geom, flags = ctypes.c_int(), (ctypes.c_int * 11)(), (ctypes.c_int * 4)()
            _lib.call("b2_model_op_info", h, i, ctypes.byref(t), geom, flags)
            types.append(t.value)
            umma += flags[0]; stem += flags[1]; fused += flags[2]; fc += flags[3]
        assert types.count(1) == 21 and types.count(2) == 20 and types.count(3) == 1 and types.count(4) == 1
        assert umma == expect_umma and stem == (1 if prec == "bf16" else 0) and fc == 1
        assert fused == (20 if prec == "bf16" else 0)
        assert b2.lib().b2_model_workspace_bytes(h) > 0
        b2.lib().b2_model_destroy(h)


def test_committed_launch_list_reproduces_the_traffic_figure(tmp_path):
    """profiles/r02_traffic.json (what bench.py copies into roofline.traffic) is derived from the committed ncu launch
    list by tools/traffic_summary.py: re-derive it and compare, and check the step it describes is a whole step."""
    import json
    import subprocess
    import sys
    out = tmp_path / "traffic.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "traffic_summary.py"),
                        os.path.join(ROOT, "profiles", "r02_launches_dram.csv"), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = json.load(open(out))
    want = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    assert got["step_launches"] == want["step_launches"] == 93
    assert got["conv_launches_per_step"] == 39                         # 20 fprop (incl. the stem) + 19 dgrad
    assert abs(got["conv_dram_bytes_per_launch"] - want["conv_dram_bytes_per_launch"]) < 1.0
    k = got["per_kernel"]
    assert k["void k_wgrad<0>"]["n"] + k["void k_wgrad<1>"]["n"] == 20 and k["k_sgd"]["n"] == 3
    # activations live in L2: the whole step moves well under 1 GB through DRAM
    assert 0.3e9 < got["step_dram_read_bytes"] + got["step_dram_write_bytes"] < 1.0e9
