"""CPU: the oracle restatement (oracle/ref_step.py) against the vectors that
tests/golden/make_golden.py produced by driving the REAL reference module
(/root/reference/cifar_example_ddp.py Net / CrossEntropyLoss / SGD / DDP-over-gloo)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_step as O


def _fp(t):
    t = t.double()
    return [float(t.sum()), float(t.abs().sum()), float((t * t).sum())]


def test_net_init_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "net_b4_w1.npz"))
    p = O.net_init(0)
    for name, shape in O.NET_PARAM_SHAPES:
        assert tuple(p[name].shape) == shape
        assert np.array_equal(p[name].numpy(), g["init." + name]), name


def test_net_fixed_batch_50_steps_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "net_b4_w1.npz"))
    torch.set_num_threads(8)
    r = O.train("net", seed=0, batch=4, steps=50, fixed_batch=True)
    assert np.array_equal(r["x"].numpy(), g["x"]) and np.array_equal(r["y"].numpy(), g["y"])
    # same wheel, same ops: the restatement is bit-identical to the reference run
    np.testing.assert_allclose(np.array(r["loss"]), g["loss"], rtol=0, atol=1e-6)
    for (n, t) in r["grads0"]:
        np.testing.assert_allclose(t.numpy(), g["grad0." + n], rtol=1e-5, atol=1e-7, err_msg=n)
    for (n, t) in r["final"]:
        np.testing.assert_allclose(t.numpy(), g["final." + n], rtol=1e-5, atol=1e-7, err_msg=n)


def test_net_fresh_batches(golden_dir):
    g = np.load(os.path.join(golden_dir, "net_b16_fresh.npz"))
    r = O.train("net", seed=0, batch=16, steps=50, fixed_batch=False)
    np.testing.assert_allclose(np.array(r["loss"]), g["loss"], rtol=0, atol=1e-6)
    for (n, t) in r["final"]:
        np.testing.assert_allclose(t.numpy(), g["final." + n], rtol=1e-5, atol=1e-7, err_msg=n)


def test_ddp_restatement_matches_real_ddp(golden_dir):
    """world=2 restatement (per-rank forward, averaged grads) vs torch DDP over gloo."""
    g = np.load(os.path.join(golden_dir, "net_ddp_w2.npz"))
    r = O.train("net", seed=0, batch=4, steps=10, world=2, fixed_batch=True)
    np.testing.assert_allclose(np.array(r["loss"]), g["loss_rank0"], rtol=0, atol=2e-6)
    for (n, t) in r["final"]:
        np.testing.assert_allclose(t.numpy(), g["final." + n], rtol=1e-5, atol=2e-7, err_msg=n)


def test_ddp_equals_big_batch_for_net():
    """without BatchNorm, W ranks x B == one process on W*B."""
    a = O.train("net", seed=0, batch=4, steps=5, world=2)
    b = O.train("net", seed=0, batch=8, steps=5, world=1)
    np.testing.assert_allclose(np.array(a["loss_mean"]), np.array(b["loss"]), atol=1e-6)
    for (n, t), (_, u) in zip(a["final"], b["final"]):
        np.testing.assert_allclose(t.numpy(), u.numpy(), atol=1e-7, err_msg=n)


def test_sgd_restatement_against_torch_optimizer():
    torch.manual_seed(0)
    p0, g1, g2 = torch.randn(1000), torch.randn(1000), torch.randn(1000)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([p], lr=O.LR, momentum=O.MOMENTUM)
    q, m = p0.clone(), None
    for g in (g1, g2):
        p.grad = g.clone(); opt.step()
        q, m = O.sgd_momentum_update(q, g, m)
    np.testing.assert_allclose(q.numpy(), p.detach().numpy(), rtol=0, atol=1e-7)


def test_ce_restatement():
    torch.manual_seed(1)
    lg, y = torch.randn(16, 10), torch.randint(0, 10, (16,))
    loss, d = O.cross_entropy_and_grad(lg, y)
    sm = torch.softmax(lg.double(), 1)
    sm[torch.arange(16), y] -= 1
    np.testing.assert_allclose(d.numpy(), (sm / 16).float().numpy(), atol=1e-7)
    ref = -(torch.log_softmax(lg.double(), 1)[torch.arange(16), y]).mean()
    assert abs(float(loss) - float(ref)) < 1e-6


@pytest.mark.parametrize("name,batch,steps", [("resnet18", 32, 20)])
def test_resnet_restatement_vs_golden(golden_dir, name, batch, steps):
    with open(os.path.join(golden_dir, f"{name}_b{batch}.json")) as f:
        g = json.load(f)
    import torchvision
    if g["torch"] != torch.__version__ or g["torchvision"] != torchvision.__version__:
        pytest.skip("golden generated with a different wheel")
    torch.set_num_threads(g["threads"])
    r = O.train(name, seed=0, batch=batch, steps=steps, fixed_batch=True)
    # ResNet trajectories are chaotic: same thread count => close, not bit-equal
    np.testing.assert_allclose(np.array(r["loss"][:3]), np.array(g["loss"][:3]), atol=1e-5)
    np.testing.assert_allclose(np.array(r["loss"]), np.array(g["loss"]), atol=5e-3)
    for n, t in r["init"]:
        np.testing.assert_allclose(_fp(t), g["init"][n], rtol=1e-6, err_msg=n)
    for n, t in r["grads0"]:
        np.testing.assert_allclose(_fp(t)[1:], g["grad0"][n][1:], rtol=1e-3, err_msg=n)


# ---- the plain-C restatement (oracle/step_math.c) against torch, and against the Python oracle ----
def _cmath():
    import ctypes
    from oracle import build_c
    lib = ctypes.CDLL(build_c.build())
    return lib, ctypes


def test_c_sgd_matches_torch_optimizer_bitwise():
    lib, C = _cmath()
    torch.manual_seed(0)
    n = 10007
    p0, g1, g2, g3 = (torch.randn(n) for _ in range(4))
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([p], lr=O.LR, momentum=O.MOMENTUM)
    pc, mc = p0.clone().numpy(), np.zeros(n, dtype=np.float32)
    for i, g in enumerate((g1, g2, g3)):
        p.grad = g.clone(); opt.step()
        gn = g.numpy()
        lib.oracle_sgd_momentum(pc.ctypes.data_as(C.c_void_p), gn.ctypes.data_as(C.c_void_p), mc.ctypes.data_as(C.c_void_p),
                                C.c_int64(n), C.c_float(O.LR), C.c_float(O.MOMENTUM), C.c_int(1 if i == 0 else 0))
    # torch's CPU kernel may or may not fuse the final multiply-add: at most one rounding apart
    np.testing.assert_allclose(pc, p.detach().numpy(), rtol=2.4e-7, atol=1e-9)
    np.testing.assert_array_equal(mc, opt.state[p]["momentum_buffer"].numpy())


def test_c_cross_entropy_matches_torch_fp64():
    lib, C = _cmath()
    torch.manual_seed(2)
    B, K = 37, 10
    lg, y = (torch.randn(B, K) * 4).contiguous(), torch.randint(0, K, (B,))
    loss, d = C.c_double(), np.zeros((B, K), dtype=np.float64)
    lib.oracle_ce_loss_grad(lg.numpy().ctypes.data_as(C.c_void_p), y.numpy().ctypes.data_as(C.c_void_p), B, K, C.byref(loss),
                            d.ctypes.data_as(C.c_void_p))
    l64 = lg.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(l64, y)
    ref.backward()
    assert abs(loss.value - float(ref)) < 1e-12
    np.testing.assert_allclose(d, l64.grad.numpy(), atol=1e-14)
    l32, d32 = O.cross_entropy_and_grad(lg, y)
    assert abs(float(l32) - loss.value) < 1e-6
    np.testing.assert_allclose(d32.numpy(), d, atol=1e-7)


def test_bf16_emulation_restates_torchvision_when_rounding_is_off():
    """oracle/bf16_emul.py with emulate=False must BE torchvision's forward/backward (same logits, same
    parameter gradients); with emulate=True it must stay within bf16 noise of it."""
    import torch.nn.functional as F
    from oracle import bf16_emul as E
    for name in ("resnet18", "resnet50"):
        model = O.make_model(name, 0)
        x, y = O.synthetic_batch(O.data_generator(0), 8)
        for p in model.parameters():
            p.grad = None
        loss = F.cross_entropy(model(x), y)
        loss.backward()
        want = {n: p.grad.clone() for n, p in model.named_parameters()}
        got = E.step0(model, x, y, emulate=False)
        assert abs(got["loss"] - float(loss)) < 1e-4
        for n, g in want.items():
            scale = float(g.abs().max()) + 1e-12
            assert float((got["grads"][n] - g).abs().max()) / scale < 1e-2, n
        emu = E.step0(model, x, y, emulate=True)
        # (ResNet-50 at a batch of 8 is too ill-conditioned for a closeness statement: SURVEY.md 0.5)
        assert abs(emu["loss"] - float(loss)) < (5e-2 if name == "resnet18" else 1.0)
        assert len(emu["ops"]) == len(got["ops"])
