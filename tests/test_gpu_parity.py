"""GPU parity: the CUDA path (through the C-ABI, driven by the reference-shaped facade)
against the CPU oracle on the same seeded inputs and against the committed golden vectors.

Tolerances (floating point path, north_star: loss within 1e-3 over 50 steps):
  Net, fp32            : loss 1e-4 abs over 50 steps, grads/params rtol 1e-3 atol 1e-5
  ResNet-18, fp32 mode : step-0 loss 1e-4, loss curve 1e-3 abs over 20 steps (fixed batch)
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import ddp_tutorial_b200 as b2
from oracle import ref_step as O

pytestmark = pytest.mark.gpu
DEV = 0


def _vp(t):
    return C.c_void_p(t.data_ptr())


def run_b2(model, x, y, steps, fixed=True, gen=None, batch=None):
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(model.parameters(), lr=0.001, momentum=0.9)
    losses, grads0 = [], None
    for s in range(steps):
        if not fixed and s > 0:
            x, y = O.synthetic_batch(gen, batch)
        xd, yd = x.to(DEV), y.to(DEV)
        opt.zero_grad()
        out = model(xd)
        loss = crit(out, yd)
        loss.backward()
        if s == 0:
            grads0 = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters()}
        opt.step()
        losses.append(loss.item())
    getattr(model, "engine", getattr(getattr(model, "module", None), "engine", None)).check()
    return losses, grads0


# ------------------------------------------------------------------ kernels through the C-ABI
def test_ce_loss_grad_matches_oracle():
    torch.manual_seed(0)
    for B, Cc in ((4, 10), (512, 10), (33, 7), (1, 10)):
        lg, y = torch.randn(B, Cc) * 3, torch.randint(0, Cc, (B,))
        loss_ref, d_ref = O.cross_entropy_and_grad(lg, y)
        lgd, yd = lg.to(DEV), y.to(DEV)
        loss, d = torch.empty(1, device=DEV), torch.empty(B, Cc, device=DEV)
        b2._lib.call("b2_ce_loss_grad", _vp(lgd), _vp(yd), B, Cc, _vp(loss), _vp(d), None)
        torch.cuda.synchronize()
        assert abs(float(loss) - float(loss_ref)) < 2e-6 * max(1, abs(float(loss_ref)))
        np.testing.assert_allclose(d.cpu().numpy(), d_ref.numpy(), rtol=1e-5, atol=1e-7)


def test_sgd_matches_oracle_to_one_ulp_on_ragged_sizes():
    """p and m may differ from the CPU run by one rounding (fused vs unfused multiply-add in
    torch's CPU kernels): tolerance 2 ulp relative."""
    torch.manual_seed(0)
    for n in (1, 3, 4, 5, 1023, 62006, 1 << 20):
        p0, g1, g2 = torch.randn(n), torch.randn(n), torch.randn(n)
        q, m = p0.clone(), None
        pd = p0.to(DEV).clone()
        md = torch.zeros(n, device=DEV)
        for i, g in enumerate((g1, g2)):
            q, m = O.sgd_momentum_update(q, g, m)
            gd = g.to(DEV)
            b2._lib.call("b2_sgd", _vp(pd), _vp(gd), _vp(md), n, 0.001, 0.9, 1.0, 1 if i == 0 else 0, None, None)
        torch.cuda.synchronize()
        np.testing.assert_allclose(pd.cpu().numpy(), q.numpy(), rtol=2.4e-7, atol=1e-9)
        np.testing.assert_allclose(md.cpu().numpy(), m.numpy(), rtol=2.4e-7, atol=1e-9)


# ------------------------------------------------------------------ Net (the reference's model)
def test_net_forward_backward_step0(golden_dir):
    g = np.load(os.path.join(golden_dir, "net_b4_w1.npz"))
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    for n, p in net.named_parameters():
        assert np.array_equal(p.detach().cpu().numpy(), g["init." + n]), n
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    ref = O.net_forward(O.net_init(0), x)
    out = net(x.to(DEV))
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-5)
    loss = b2.CrossEntropyLoss()(out, y.to(DEV))
    loss.backward()
    assert abs(loss.item() - g["loss"][0]) < 1e-5
    for n, p in net.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), g["grad0." + n], rtol=1e-3, atol=1e-6, err_msg=n)


def test_net_50_steps_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "net_b4_w1.npz"))
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    losses, _ = run_b2(net, torch.from_numpy(g["x"]), torch.from_numpy(g["y"]), 50)
    assert np.max(np.abs(np.array(losses) - g["loss"])) < 1e-4
    for n, p in net.named_parameters():
        np.testing.assert_allclose(p.detach().cpu().numpy(), g["final." + n], rtol=1e-3, atol=1e-5, err_msg=n)


def test_net_fresh_batches_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "net_b16_fresh.npz"))
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    gen = O.data_generator(0)
    x, y = O.synthetic_batch(gen, 16)
    losses, _ = run_b2(net, x, y, 50, fixed=False, gen=gen, batch=16)
    assert np.max(np.abs(np.array(losses) - g["loss"])) < 1e-4


def test_net_batch128_vs_oracle():
    r = O.train("net", seed=0, batch=128, steps=10)
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    losses, grads0 = run_b2(net, r["x"], r["y"], 10)
    assert np.max(np.abs(np.array(losses) - np.array(r["loss"]))) < 1e-4
    for n, gr in r["grads0"]:
        np.testing.assert_allclose(grads0[n].numpy(), gr.numpy(), rtol=2e-3, atol=1e-6, err_msg=n)


def test_state_dict_round_trip_and_names():
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    sd = net.state_dict()
    assert list(sd) == [n for n, _ in O.NET_PARAM_SHAPES]
    assert tuple(sd["conv1.weight"].shape) == (6, 3, 5, 5) and sd["conv1.weight"].is_contiguous()
    other = b2.Net().to(DEV)
    other.load_state_dict(sd)
    for (n, a), (_, b) in zip(net.named_parameters(), other.named_parameters()):
        assert torch.equal(a, b), n


# ------------------------------------------------------------------ ResNet-18 (BASELINE configs)
def _grad_report(grads0, ref_grads):
    worst = []
    for n, gr in ref_grads:
        a, b = grads0[n].double(), gr.double()
        err = float((a - b).abs().max() / (b.abs().max() + 1e-12))
        worst.append((err, n))
    worst.sort(reverse=True)
    return worst


@pytest.mark.parametrize("batch", [32])
def test_resnet18_fp32_step0_grads(batch):
    """Gradients against the fp64 oracle (the fp32 CPU wheel itself is 1.8e-2 off in layer1.0,
    profiles/r01_parity_diag_resnet18_fp32.txt); loss against the fp32 oracle."""
    r = O.train("resnet18", seed=0, batch=batch, steps=1)
    r64 = O.train("resnet18", seed=0, batch=batch, steps=1, dtype=torch.float64)
    torch.manual_seed(0)
    net = b2.resnet18(num_classes=10, precision="fp32").to(DEV)
    losses, grads0 = run_b2(net, r["x"], r["y"], 1)
    assert abs(losses[0] - r["loss"][0]) < 1e-4, (losses[0], r["loss"][0])
    worst = _grad_report(grads0, r64["grads0"])
    assert worst[0][0] < 1e-4, worst[:8]


def test_resnet18_fp32_loss_curve_vs_golden(golden_dir):
    with open(os.path.join(golden_dir, "resnet18_b32.json")) as f:
        g = json.load(f)
    gen = O.data_generator(0)
    x, y = O.synthetic_batch(gen, 32)
    torch.manual_seed(0)
    net = b2.resnet18(num_classes=10, precision="fp32").to(DEV)
    losses, _ = run_b2(net, x, y, g["steps"])
    err = np.abs(np.array(losses) - np.array(g["loss"]))
    assert err[0] < 1e-4 and err.max() < 1e-3, err


def test_resnet18_state_dict_keys_match_torchvision():
    import torchvision
    torch.manual_seed(0)
    net = b2.resnet18(num_classes=10, precision="fp32").to(DEV)
    ref = torchvision.models.resnet18(num_classes=10)
    sd, rsd = net.state_dict(), ref.state_dict()
    assert list(sd) == list(rsd)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(rsd[k].shape), k


def test_resnet50_fp32_step0():
    """ResNet-50 at 1x1 spatial in layer4 is so ill-conditioned that torch's own fp32 CPU
    gradients are 30% (max-norm) off its fp64 gradients at B=32
    (profiles/r01_parity_diag_resnet50_fp32.txt).  Gate: loss within 2e-4 of the fp32 oracle and
    our gradient no further from the fp64 oracle than the fp32 oracle itself is."""
    r = O.train("resnet50", seed=0, batch=32, steps=1)
    r64 = O.train("resnet50", seed=0, batch=32, steps=1, dtype=torch.float64)
    torch.manual_seed(0)
    net = b2.resnet50(num_classes=10, precision="fp32").to(DEV)
    losses, grads0 = run_b2(net, r["x"], r["y"], 1)
    assert abs(losses[0] - r["loss"][0]) < 2e-4, (losses[0], r["loss"][0])
    truth = torch.cat([g.flatten() for _, g in r64["grads0"]])
    mine = torch.cat([grads0[n].flatten() for n, _ in r64["grads0"]]).double()
    cpu32 = torch.cat([g.flatten() for _, g in r["grads0"]]).double()
    rel = lambda a: float((a - truth).norm() / truth.norm())
    assert rel(mine) < 1.5 * rel(cpu32) + 1e-3, (rel(mine), rel(cpu32))
    # the early layers are well conditioned: per-tensor check there
    worst = [w for w in _grad_report(grads0, r64["grads0"]) if w[1].startswith(("conv1", "bn1", "layer1"))]
    ref = {n: e for e, n in _grad_report({n: g for n, g in r["grads0"]}, r64["grads0"])}
    for e, n in worst:
        assert e < 3 * ref[n] + 1e-4, (n, e, ref[n])


def _autocast_grads(name, x, y):
    import torch.nn.functional as F
    m = O.make_model(name, 0)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        loss = F.cross_entropy(m(x).float(), y)
    loss.backward()
    return float(loss), torch.cat([p.grad.flatten() for p in m.parameters()]).double()


def test_resnet18_bf16_step0_within_bf16_noise():
    """bf16 storage cannot meet an fp32-level gate on this net.  Yardstick:
    torch's own bf16 autocast on the CPU, both measured against the fp64 oracle: our loss error
    and gradient direction must be no worse than the reference stack's own bf16 mode."""
    r64 = O.train("resnet18", seed=0, batch=64, steps=1, dtype=torch.float64)
    x, y = r64["x"].float(), r64["y"]
    torch.manual_seed(0)
    net = b2.resnet18(num_classes=10, precision="bf16").to(DEV)
    losses, grads0 = run_b2(net, x, y, 1)
    ac_loss, ac = _autocast_grads("resnet18", x, y)
    truth = torch.cat([g.flatten() for _, g in r64["grads0"]])
    mine = torch.cat([grads0[n].flatten() for n, _ in r64["grads0"]]).double()
    cos = lambda a, b: float((a @ b) / (a.norm() * b.norm()))
    assert abs(losses[0] - r64["loss"][0]) < max(2 * abs(ac_loss - r64["loss"][0]), 5e-3)
    assert cos(mine, truth) > cos(ac, truth) - 0.02, (cos(mine, truth), cos(ac, truth))
    assert cos(mine, truth) > 0.9


def test_resnet50_bf16_runs_and_tracks_oracle():
    """BASELINE configs[3] model on the tcgen05 path: step-0 loss near the fp32 oracle, loss falls."""
    r = O.train("resnet50", seed=0, batch=64, steps=1)
    torch.manual_seed(0)
    net = b2.resnet50(num_classes=10, precision="bf16").to(DEV)
    losses, _ = run_b2(net, r["x"], r["y"], 8)
    assert abs(losses[0] - r["loss"][0]) < 5e-2, (losses[0], r["loss"][0])
    assert losses[-1] < losses[0]
    assert all(np.isfinite(losses))


def test_graph_replay_equals_eager():
    """CUDA-graph replay (incl. the forked weight-gradient stream and PDL edges) must reproduce the
    eager launches: same loss curve for ResNet-18 fp32 (deterministic kernels) with graphs on / off."""
    gen = O.data_generator(0)
    x, y = O.synthetic_batch(gen, 32)
    curves = []
    for graphs in (True, False):
        torch.manual_seed(0)
        net = b2.resnet18(num_classes=10, precision="fp32").to(DEV)
        net.engine.set_graphs(graphs)
        losses, _ = run_b2(net, x, y, 8)
        curves.append(losses)
    assert curves[0] == curves[1], curves


def test_persistent_logits_and_batch_size_switch():
    """One engine serves several batch sizes (one static graph each) over the same parameters."""
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
    gen = O.data_generator(3)
    for B in (4, 16, 4, 7):
        x, y = O.synthetic_batch(gen, B)
        out = net(x.to(DEV))
        assert tuple(out.shape) == (B, 10)
        loss = crit(out, y.to(DEV)); loss.backward(); opt.step()
        assert np.isfinite(loss.item())


def test_error_paths_on_gpu():
    net = b2.Net().to(DEV)
    with pytest.raises(b2.B2Error):
        net(torch.zeros(4, 3, 32, 32))                       # CPU tensor into a CUDA model
    with pytest.raises(b2.B2Error):
        net(torch.zeros(4, 3, 16, 16, device=DEV))           # wrong spatial size
    with pytest.raises(b2.B2Error):
        net(torch.zeros(0, 3, 32, 32, device=DEV))           # empty batch
    out = net(torch.zeros(4, 3, 32, 32, device=DEV))
    with pytest.raises(b2.B2Error):
        b2.CrossEntropyLoss()(out, torch.zeros(3, dtype=torch.long, device=DEV))    # ragged labels
    loss = b2.CrossEntropyLoss()(out, torch.zeros(4, dtype=torch.long, device=DEV))
    loss.backward()
    with pytest.raises(b2.B2Error):
        loss.backward()                                      # static graph: no double backward


def test_u8_preprocess_matches_torchvision_transform():
    """ToTensor + Normalize((.5,.5,.5),(.5,.5,.5)) (cifar_example_ddp.py:61-63) on the device, bit-level."""
    import torchvision.transforms as T
    from PIL import Image
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (16, 32, 32, 3), generator=g, dtype=torch.uint8)
    tf = T.Compose([T.ToTensor(), T.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    ref = torch.stack([tf(Image.fromarray(img.numpy())) for img in u8])
    out = b2.normalize_u8(u8.to(DEV)).cpu()
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=0, atol=1.2e-7)


def test_device_loader_follows_sampler_and_accuracy_counts():
    g = torch.Generator().manual_seed(1)
    imgs = torch.randint(0, 256, (103, 32, 32, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 10, (103,), generator=g)
    sampler = b2.DistributedSampler(103, num_replicas=4, rank=1, shuffle=True, seed=0)
    sampler.set_epoch(2)
    loader = b2.DeviceCifarLoader(imgs, labels, 8, sampler, DEV)
    seen = torch.cat([y for _, y in loader]).cpu()
    assert torch.equal(seen, labels[sampler.indices()])
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    acc = b2.Accuracy(DEV)
    correct = total = 0
    for x, y in loader:
        out = net(x)
        acc.update(out, y)
        correct += int((out.argmax(1) == y).sum()); total += y.numel()
    assert abs(acc.compute() - correct / total) < 1e-12


def test_checkpoint_round_trip_into_reference_modules(tmp_path):
    """torch.save(net.state_dict()) (cifar_example_ddp.py:118-119) loads strictly into torchvision's
    ResNet-18 and reproduces our logits in fp32; keys/shapes are torch's."""
    import torchvision
    torch.manual_seed(0)
    net = b2.resnet18(num_classes=10, precision="fp32").to(DEV)
    x = torch.randn(8, 3, 32, 32)
    ours = net(x.to(DEV)).cpu()           # running stats updated once, like the reference's train-mode forward
    path = tmp_path / "cifar_net.pth"
    torch.save(net.state_dict(), path)
    ref = torchvision.models.resnet18(num_classes=10)
    ref.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
    # the checkpointed weights reproduce our logits; ref.train() forward uses batch stats like ours
    ref.train()
    with torch.no_grad():
        theirs = ref(x)
    np.testing.assert_allclose(ours.numpy(), theirs.numpy(), rtol=1e-3, atol=1e-4)
    assert int(ref.bn1.num_batches_tracked) == 2          # one forward here + one in the reference check


def test_device_prefetcher_delivers_batches_in_order():
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(4, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 10, (4,), generator=g).pin_memory())
               for _ in range(7)]
    got = []
    for x, y in b2.DevicePrefetcher(batches, DEV, depth=2):
        got.append((x.cpu().clone(), y.cpu().clone()))     # the slot is recycled after the next yield
    assert len(got) == 7
    for (x, y), (rx, ry) in zip(got, batches):
        assert torch.equal(x, rx) and torch.equal(y, ry)


def test_sgd_kernel_bit_identical_to_c_oracle():
    """b2_sgd against oracle/step_math.c (plain C, no FP contraction): bit-for-bit on p and m,
    including the bf16 mirror = round-to-nearest-even of the updated fp32 parameter."""
    import ctypes
    from oracle import build_c
    lib = ctypes.CDLL(build_c.build())
    torch.manual_seed(3)
    n = 100003
    p0, grads = torch.randn(n), [torch.randn(n) for _ in range(4)]
    pc, mc = p0.clone().numpy(), np.zeros(n, dtype=np.float32)
    pd, md = p0.to(DEV).clone(), torch.zeros(n, device=DEV)
    npad = (n + 3) // 4 * 4
    mirror = torch.zeros(npad, dtype=torch.bfloat16, device=DEV)
    for i, g in enumerate(grads):
        gn = g.numpy()
        lib.oracle_sgd_momentum(pc.ctypes.data_as(ctypes.c_void_p), gn.ctypes.data_as(ctypes.c_void_p),
                                mc.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n), ctypes.c_float(0.001),
                                ctypes.c_float(0.9), ctypes.c_int(1 if i == 0 else 0))
        gd = g.to(DEV)
        b2._lib.call("b2_sgd", _vp(pd), _vp(gd), _vp(md), n, 0.001, 0.9, 1.0, 1 if i == 0 else 0, _vp(mirror), None)
    torch.cuda.synchronize()
    assert np.array_equal(pd.cpu().numpy(), pc)
    assert np.array_equal(md.cpu().numpy(), mc)
    assert torch.equal(mirror[:n].cpu(), torch.from_numpy(pc).bfloat16())


def test_ce_kernel_against_c_oracle_fp64():
    import ctypes
    from oracle import build_c
    lib = ctypes.CDLL(build_c.build())
    torch.manual_seed(4)
    B, K = 512, 10
    lg, y = (torch.randn(B, K) * 3).contiguous(), torch.randint(0, K, (B,))
    ref_loss, ref_d = ctypes.c_double(), np.zeros((B, K), dtype=np.float64)
    lib.oracle_ce_loss_grad(lg.numpy().ctypes.data_as(ctypes.c_void_p), y.numpy().ctypes.data_as(ctypes.c_void_p), B, K,
                            ctypes.byref(ref_loss), ref_d.ctypes.data_as(ctypes.c_void_p))
    lgd, yd = lg.to(DEV), y.to(DEV)
    loss, d = torch.empty(1, device=DEV), torch.empty(B, K, device=DEV)
    b2._lib.call("b2_ce_loss_grad", _vp(lgd), _vp(yd), B, K, _vp(loss), _vp(d), None)
    torch.cuda.synchronize()
    assert abs(float(loss) - ref_loss.value) < 2e-6
    np.testing.assert_allclose(d.cpu().numpy(), ref_d, rtol=2e-5, atol=2e-9)


@pytest.mark.gpu
def test_kernel_trace_records_one_step():
    """b2_trace_start / b2_trace_stop (include/b2ddp.h): every kernel of a training step -- also inside the replayed
    CUDA graphs -- leaves a (tag, %globaltimer) record; tags decode to source lines of __global__ functions."""
    import ctypes as C
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
    x = torch.randn(4, 3, 32, 32).to(DEV)
    y = torch.randint(0, 10, (4,)).to(DEV)

    def step():
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(6):                       # past the eager warm-up: forward/backward replay from graphs
        step()
    torch.cuda.synchronize()
    launches0 = b2.lib().b2_launch_count()
    b2._lib.call("b2_trace_start", 1024)
    step()
    torch.cuda.synchronize()
    launches = b2.lib().b2_launch_count() - launches0
    rec = (C.c_ulonglong * 2048)()
    n = C.c_int(0)
    b2._lib.call("b2_trace_stop", rec, 1024, C.byref(n))
    begins = [(rec[2 * i], rec[2 * i + 1]) for i in range(n.value) if (rec[2 * i] & 15) == 0]
    assert launches > 10 and abs(len(begins) - launches) <= 2, (len(begins), launches)     # one "begin" per launched kernel
    assert all(1 <= (tag >> 24) <= 4 and ((tag >> 4) & 0xFFFFF) > 0 for tag, _ in begins)
    ts = sorted(t for _, t in begins)
    assert 0 < ts[-1] - ts[0] < 50_000_000                            # one small step: well under 50 ms of %globaltimer
    # switched off again: a further step records nothing and a second stop is an error
    step(); torch.cuda.synchronize()
    with pytest.raises(b2.B2Error):
        b2._lib.call("b2_trace_stop", rec, 1024, C.byref(n))
