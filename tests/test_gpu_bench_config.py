"""GPU parity ON THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs[1..3]): ResNet-18 at the
BASELINE batch 512 in bf16 (the kernels bench.py times: multi-tile persistent conv + fused BatchNorm,
split-K clusters, tcgen05 layer4) and ResNet-50 at batch 256, through the facade -> C-ABI.

Protocol (SURVEY.md 0.5 -- a literal 1e-3/50-step gate is unattainable in bf16; the reference stack's own
bf16 mode is 8e-3 off its fp64 run, tests/golden/resnet18_b512_envelope.json):
  (i)   per-op activations and activation gradients at step 0 against oracle/bf16_emul.py (the reference
        arithmetic with the bf16 storage points emulated): rms-relative and max-relative error per tensor
  (ii)  step-0 loss, per-tensor parameter gradients
  (iii) teacher-forced single step from oracle parameters after k reference steps
  (iv)  50-step free-running curve inside the reference stack's own spread around its fp64 trajectory
  (v)   fp32 mode at B=512: the literal north-star gate against the committed golden curve
Measured errors are also written to gpurun_out/r02_parity_bench_config.json.
"""
import json
import os

import numpy as np
import pytest
import torch

import ddp_tutorial_b200 as b2
from oracle import bf16_emul as E
from oracle import ref_step as O

pytestmark = pytest.mark.gpu
DEV = 0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_report = {}


def _save_report():
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "r02_parity_bench_config.json"), "w") as f:
        json.dump(_report, f, indent=1)


def _rel(a: torch.Tensor, b: torch.Tensor):
    a, b = a.double().flatten(), b.double().flatten()
    d = a - b
    return float(d.norm() / (b.norm() + 1e-30)), float(d.abs().max() / (b.abs().max() + 1e-30))


def _per_op_compare(name, batch, act_rms, act_max, loc_rms, loc_max):
    """Forward: every stored tensor against the emulated oracle.  Backward: op-local, teacher-forced check
    (oracle/bf16_emul.py:local_backward) -- the reference arithmetic of each op applied to OUR stored inputs and
    OUR stored upstream gradient must reproduce OUR stored gradient of every tensor and parameter.  (Comparing two
    complete bf16 backward passes only bounds them to the net's noise amplification: the full-pass numbers are
    reported next to the reference stack's own bf16-vs-fp64 level, not asserted.)"""
    model = O.make_model(name, 0)
    x, y = O.synthetic_batch(O.data_generator(0), batch)
    ref = E.step0(model, x, y, emulate=True)
    torch.manual_seed(0)
    net = getattr(b2, name)(num_classes=10, precision="bf16").to(DEV)
    crit = b2.CrossEntropyLoss()
    out = net(x.to(DEV))
    loss = crit(out, y.to(DEV))
    loss.backward()                      # no optimizer attached: backward only produces gradients
    torch.cuda.synchronize()
    net.engine.check()
    e = net.engine
    assert e.num_ops() == len(ref["ops"]) + 1, (e.num_ops(), len(ref["ops"]))
    ours = []
    for i, (opname, val, grad) in enumerate(ref["ops"]):
        a = e.read_activation(i + 1).cpu().reshape(val.shape)
        g = e.read_activation(i + 1, grad=True).cpu().reshape(val.shape)
        ours.append((a, g))
    exp, exp_pg, exp_fwd = E.local_backward(model, x, ours)
    rows, worst = [], []
    for i, (opname, val, grad) in enumerate(ref["ops"]):
        a, g = ours[i]
        r_rms, r_max = _rel(a, val)
        l_rms, l_max = _rel(a, exp_fwd[i].reshape(a.shape))
        row = {"op": i + 1, "name": opname, "act_rms": r_rms, "act_max": r_max, "act_local_rms": l_rms, "act_local_max": l_max}
        if r_rms > act_rms or r_max > act_max:
            worst.append((opname, "act", r_rms, r_max))
        if l_rms > 4e-3 or l_max > 1.6e-2:        # each op alone, from OUR inputs: one bf16 rounding (2^-8) of the output
            worst.append((opname, "act(local)", l_rms, l_max))
        if opname != "fc":
            row["grad_local_rms"], row["grad_local_max"] = _rel(g, exp[i])
            row["grad_fullpass_rms_vs_emulated"] = _rel(g, grad)[0]
            if row["grad_local_rms"] > loc_rms or row["grad_local_max"] > loc_max:
                worst.append((opname, "grad(local)", row["grad_local_rms"], row["grad_local_max"]))
        rows.append(row)
    prow = []
    pg = {n: p.grad.detach().cpu() for n, p in net.named_parameters()}
    for n, g in exp_pg.items():
        r_rms, r_max = _rel(pg[n], g)
        prow.append({"param": n, "local_rms": r_rms, "local_max": r_max, "fullpass_rms_vs_emulated": _rel(pg[n], ref["grads"][n])[0]})
        if r_rms > loc_rms or r_max > 2 * loc_max:
            worst.append((n, "param grad(local)", r_rms, r_max))
    _report[f"{name}_b{batch}_per_op"] = {"loss": loss.item(), "loss_emulated_oracle": ref["loss"], "ops": rows, "param_grads": prow}
    _save_report()
    assert abs(loss.item() - ref["loss"]) < (2e-3 if name == "resnet18" else 5e-2), (loss.item(), ref["loss"])
    assert not worst, worst[:12]
    assert len(exp_pg) == len(pg)
    return net, ref


def test_resnet18_bf16_b512_per_op_activations_and_gradients():
    """Every stored tensor of the step bench.py times, forward and backward, against the emulated oracle.
    A dead filter tap, a mis-scaled BatchNorm or a wrong tile of the multi-tile kernels shows up as an
    O(1) relative error in its tensor; bf16 re-rounding noise stays at the per-cent level."""
    _per_op_compare("resnet18", 512, act_rms=3e-2, act_max=6e-2, loc_rms=1e-2, loc_max=4e-2)


def test_resnet50_bf16_b256_per_op_activations_and_gradients():
    """BASELINE configs[3] per-GPU workload (ResNet-50, batch 256, bf16).  The end-to-end comparison with the
    emulated oracle is only a sanity bound here: ResNet-50's 1x1-spatial stage amplifies bf16 re-rounding noise to
    ~20% by layer4 (two runs of the REFERENCE in fp32 with different thread counts already differ by 7e-2 in
    loss within 8 steps, resnet50_b256_envelope.json); the op-local forward/backward checks carry the parity claim."""
    _per_op_compare("resnet50", 256, act_rms=4e-1, act_max=6e-1, loc_rms=1e-2, loc_max=4e-2)


def _load_envelope(name):
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.load(f)


def test_resnet18_bf16_b512_50_step_curve_inside_reference_envelope():
    """Free-running 50 steps, fixed batch.  The reference stack's own bf16 mode (torch.autocast) deviates from
    its fp64 trajectory by env = |bf16_autocast - fp64| (max 8.2e-3); ours must stay within 2x that spread
    at every step (plus the fp32-vs-fp64 floor), and within 2x at step 0."""
    env = _load_envelope("resnet18_b512_envelope.json")
    f64 = np.array(env["fp64"])
    spread = np.abs(np.array(env["bf16_autocast"]) - f64)
    floor = np.abs(np.array(env["fp32_t1"]) - f64).max()
    x, y = O.synthetic_batch(O.data_generator(0), 512)
    torch.manual_seed(0)
    net = b2.resnet18(num_classes=10, precision="bf16").to(DEV)
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
    xd, yd = x.to(DEV), y.to(DEV)
    losses = []
    for _ in range(env["steps"]):
        opt.zero_grad()
        loss = crit(net(xd), yd)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    net.engine.check()
    err = np.abs(np.array(losses) - f64)
    _report["resnet18_b512_bf16_curve"] = {"loss": losses, "err_vs_fp64": err.tolist(), "reference_bf16_spread": spread.tolist(),
                                           "reference_fp32_floor": float(floor)}
    _save_report()
    assert err[0] <= 2 * spread[0] + 1e-4, (err[0], spread[0])
    # running maximum of the reference's own spread: chaotic growth makes the pointwise values jitter
    bound = 2 * np.maximum.accumulate(spread) + floor
    assert np.all(err <= bound), list(zip(err.tolist(), bound.tolist()))
    assert losses[-1] < 0.1


def test_resnet18_bf16_b512_teacher_forced_single_step():
    """Take the reference's fp32 parameters after k of its own steps, load them here, do ONE step (fresh
    optimizer: first-step SGD semantics, buf = grad) and compare loss and parameter update with the
    emulated oracle at the same point: isolates single-step error from trajectory drift."""
    import torch.nn.functional as F
    x, y = O.synthetic_batch(O.data_generator(0), 512)
    rep = {}
    for k in (0, 10):
        r = O.train("resnet18", seed=0, batch=512, steps=k, keep_grads_of_step0=False) if k else None
        model = r["model"] if k else O.make_model("resnet18", 0)
        sd = {n: t.detach().clone() for n, t in model.state_dict().items()}
        ref = E.step0(model, x, y, emulate=True)
        ref64 = E.step0(_as_double(model), x.double(), y, emulate=False)
        torch.manual_seed(0)
        net = b2.resnet18(num_classes=10, precision="bf16").to(DEV)
        net.load_state_dict(sd)
        before = net.engine.flat_params.clone()
        crit = b2.CrossEntropyLoss()
        opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
        loss = crit(net(x.to(DEV)), y.to(DEV))
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        names = [n for n, _ in net.named_parameters()]
        upd = {n: (p.detach() - torch.as_strided(before, p.shape, p.stride(), p.storage_offset())).cpu()
               for n, p in net.named_parameters()}
        mine = torch.cat([upd[n].flatten() for n in names]).double()
        emu = torch.cat([(-O.LR * ref["grads"][n]).flatten() for n in names]).double()
        tru = torch.cat([(-O.LR * ref64["grads"][n]).flatten() for n in names]).double()
        cos = lambda a, b: float((a @ b) / (a.norm() * b.norm()))
        rep[k] = {"loss": loss.item(), "loss_emulated": ref["loss"], "loss_fp64": ref64["loss"],
                  "update_rel_err_vs_emulated": float((mine - emu).norm() / emu.norm()), "cos_vs_emulated": cos(mine, emu),
                  "cos_vs_fp64": cos(mine, tru), "cos_emulated_vs_fp64": cos(emu, tru)}
        _report["resnet18_b512_teacher_forced"] = rep
        _save_report()
        assert abs(loss.item() - ref["loss"]) < 2e-3, rep[k]
        assert abs(loss.item() - ref64["loss"]) < 6e-3, rep[k]
        # two bf16 realisations of the same update only agree to the net's noise amplification (the emulated reference
        # itself is at cos 0.94 / 0.905 to fp64 at k = 0 / 10); the claim that matters is the next line
        assert rep[k]["cos_vs_emulated"] > 0.9, rep[k]
        # no further from the true (fp64) update direction than the emulated reference arithmetic is
        assert rep[k]["cos_vs_fp64"] > rep[k]["cos_emulated_vs_fp64"] - 0.03, rep[k]


def _as_double(model):
    import copy
    m = copy.deepcopy(model).double()
    for p in m.parameters():
        p.grad = None
    return m


def test_resnet18_fp32_b512_50_steps_vs_golden(golden_dir):
    """north_star's literal gate at the BASELINE batch: fp32 mode, one fixed batch, 50 steps, against the
    curve the reference's loop produced (tests/golden/resnet18_b512.json).  The reference's own 8-thread
    and 1-thread fp32 runs differ from its fp64 run by 0.98e-3 / 1.36e-3 over these 50 steps, so the bound
    on |ours - golden| is 1e-3 for the first 10 steps and 2.5e-3 overall; against fp64 ours must be no worse
    than 1.5x the reference's own fp32 spread."""
    with open(os.path.join(golden_dir, "resnet18_b512.json")) as f:
        g = json.load(f)
    env = _load_envelope("resnet18_b512_envelope.json")
    x, y = O.synthetic_batch(O.data_generator(0), 512)
    torch.manual_seed(0)
    net = b2.resnet18(num_classes=10, precision="fp32").to(DEV)
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
    xd, yd = x.to(DEV), y.to(DEV)
    losses = []
    for _ in range(g["steps"]):
        opt.zero_grad()
        loss = crit(net(xd), yd)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    net.engine.check()
    err = np.abs(np.array(losses) - np.array(g["loss"]))
    f64 = np.array(env["fp64"])
    err64 = np.abs(np.array(losses) - f64)
    ref_spread = max(np.abs(np.array(env["fp32_t8"]) - f64).max(), np.abs(np.array(env["fp32_t1"]) - f64).max())
    _report["resnet18_b512_fp32_curve"] = {"loss": losses, "err_vs_golden_fp32": err.tolist(), "err_vs_fp64": err64.tolist(),
                                           "reference_fp32_spread_vs_fp64": float(ref_spread)}
    _save_report()
    assert err[0] < 1e-4 and err[:10].max() < 1e-3, err[:10]
    assert err.max() < 2.5e-3, err
    assert err64.max() < 1.5 * ref_spread, (err64.max(), ref_spread)


def test_resnet50_fp32_b32_5_steps_vs_golden(golden_dir):
    """tests/golden/resnet50_b32.json (5 steps of the reference loop, torch fp32 on the CPU).  Only step 0 of that
    curve is meaningful: torch's fp32 CPU gradients for this net are 30% (max-norm) off its own fp64 gradients
    (profiles/r01_parity_diag_resnet50_fp32.txt) -- its loss even RISES at step 1 (2.739 -> 2.828) -- so the later
    steps are gated against the fp64 run of the same loop (computed here by the oracle), which ours must follow."""
    with open(os.path.join(golden_dir, "resnet50_b32.json")) as f:
        g = json.load(f)
    r64 = O.train("resnet50", seed=0, batch=32, steps=g["steps"], dtype=torch.float64, keep_grads_of_step0=False)
    x, y = O.synthetic_batch(O.data_generator(0), 32)
    torch.manual_seed(0)
    net = b2.resnet50(num_classes=10, precision="fp32").to(DEV)
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
    xd, yd = x.to(DEV), y.to(DEV)
    losses = []
    for _ in range(g["steps"]):
        opt.zero_grad()
        loss = crit(net(xd), yd)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    err = np.abs(np.array(losses) - np.array(g["loss"]))
    err64 = np.abs(np.array(losses) - np.array(r64["loss"]))
    gold64 = np.abs(np.array(g["loss"]) - np.array(r64["loss"]))
    _report["resnet50_b32_fp32_curve"] = {"loss": losses, "err_vs_golden_fp32": err.tolist(), "err_vs_fp64": err64.tolist(),
                                          "golden_fp32_vs_fp64": gold64.tolist()}
    _save_report()
    assert err[0] < 2e-4 and err64[0] < 2e-4, (err, err64)
    # chaotic from step 1 on: ours must be no further from the fp64 trajectory than the reference's own fp32 run is
    assert err64.max() < max(2e-2, 1.5 * gold64.max()), (err64, gold64)


def test_resnet50_bf16_b256_curve_inside_reference_envelope():
    env = _load_envelope("resnet50_b256_envelope.json")
    f64 = np.array(env["fp64"])
    spread = np.maximum(np.abs(np.array(env["bf16_autocast"]) - f64), np.abs(np.array(env["fp32_t1"]) - f64))
    x, y = O.synthetic_batch(O.data_generator(0), 256)
    torch.manual_seed(0)
    net = b2.resnet50(num_classes=10, precision="bf16").to(DEV)
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
    xd, yd = x.to(DEV), y.to(DEV)
    losses = []
    for _ in range(env["steps"]):
        opt.zero_grad()
        loss = crit(net(xd), yd)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    net.engine.check()
    err = np.abs(np.array(losses) - f64)
    _report["resnet50_b256_bf16_curve"] = {"loss": losses, "err_vs_fp64": err.tolist(), "reference_spread": spread.tolist()}
    _save_report()
    assert err[0] <= 3 * spread[0] + 5e-3, (err[0], spread[0])       # forward-only bf16 noise of this net: 1-2e-2 in the loss
    assert np.all(err <= 3 * np.maximum.accumulate(spread) + 2e-2), list(zip(err.tolist(), spread.tolist()))
    assert all(np.isfinite(losses))


@pytest.mark.parametrize("name,batch", [("net", 16), ("resnet18", 32)])
def test_fused_update_inside_backward_equals_separate_step_bitwise(name, batch):
    """SGD(fuse_into_backward=True) (default: the update of a gradient bucket is launched from inside backward())
    must produce bit-identical parameters and momentum to the plain form (one pass at optimizer.step()), fp32."""
    x, y = O.synthetic_batch(O.data_generator(0), batch)
    res = []
    for fused in (True, False):
        torch.manual_seed(0)
        net = (b2.Net() if name == "net" else getattr(b2, name)(num_classes=10, precision="fp32")).to(DEV)
        crit = b2.CrossEntropyLoss()
        opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9, fuse_into_backward=fused, bucket_cap_mb=1.0)
        xd, yd = x.to(DEV), y.to(DEV)
        for _ in range(5):
            opt.zero_grad()
            loss = crit(net(xd), yd)
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        res.append((net.engine.flat_params.clone(), opt.momentum_buffer.clone(), loss.item()))
    assert torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][1], res[1][1])
    assert res[0][2] == res[1][2]


def test_stale_loss_and_repeated_placement_are_rejected():
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    p0 = net.engine.flat_params.clone()
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
    x, y = O.synthetic_batch(O.data_generator(0), 4)
    loss = crit(net(x.to(DEV)), y.to(DEV)); loss.backward(); opt.step()
    trained = net.engine.flat_params.clone()
    assert not torch.equal(p0, trained)
    net.to(DEV); net.cuda()                                 # torch semantics: no-op, parameters survive
    assert torch.equal(net.engine.flat_params, trained)
    with pytest.raises(b2.B2Error):
        net.to("cpu")                                       # there is no CPU path
    l1 = crit(net(x.to(DEV)), y.to(DEV))
    l2 = crit(net(x.to(DEV)), y.to(DEV))                    # a second forward invalidates l1's activations
    with pytest.raises(b2.B2Error):
        l1.backward()
    l2.backward()


def test_running_loss_accumulates_on_device_and_label_range_is_checked():
    torch.manual_seed(0)
    net = b2.Net().to(DEV)
    run = b2.RunningLoss(DEV)
    crit = b2.CrossEntropyLoss(running=run)
    x, y = O.synthetic_batch(O.data_generator(0), 8)
    vals = []
    for _ in range(3):
        vals.append(crit(net(x.to(DEV)), y.to(DEV)).item())
    assert abs(run.mean_and_reset() - float(np.mean(vals))) < 1e-6
    assert run.mean_and_reset() == 0.0


def test_dropin_script_runs_and_prints_the_reference_lines(tmp_path):
    """ddp-tutorial_b200/cifar_example_ddp.py executed as a program (argparse entry, cifar_example_ddp.py:138-145),
    single process: synthetic input, Net, 20 steps; the printed running means equal the oracle's."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "ddp-tutorial_b200", "cifar_example_ddp.py"), "--model", "net", "--batch-size", "4",
                        "--epochs", "1", "--steps", "20", "--print-every", "10", "--save", str(tmp_path / "cifar_net.pth")],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Not using distributed mode" in r.stdout and "Finished Training" in r.stdout
    ref = O.train("net", seed=0, batch=4, steps=20, fixed_batch=False)["loss"]
    lines = [l for l in r.stdout.splitlines() if l.startswith("[1,")]
    assert len(lines) == 2, r.stdout
    for line, want in zip(lines, (np.mean(ref[:10]), np.mean(ref[10:]))):
        got = float(line.split("loss:")[1])
        assert abs(got - want) < 2e-3, (line, want)
    sd = torch.load(tmp_path / "cifar_net.pth", map_location="cpu")
    assert list(sd) == [n for n, _ in O.NET_PARAM_SHAPES]


def _fake_cifar(root, n_train_per_file=24, n_test=40, seed=3):
    """cifar-10-batches-py in the python-pickle layout torchvision.datasets.CIFAR10 reads (cifar_example_ddp.py:67-68)."""
    import pickle
    rng = np.random.default_rng(seed)
    base = os.path.join(root, "cifar-10-batches-py")
    os.makedirs(base)
    out = {}
    for name, n in [(f"data_batch_{i}", n_train_per_file) for i in range(1, 6)] + [("test_batch", n_test)]:
        d = {"data": rng.integers(0, 256, size=(n, 3072), dtype=np.uint8), "labels": [int(v) for v in rng.integers(0, 10, size=n)]}
        with open(os.path.join(base, name), "wb") as f:
            pickle.dump(d, f)
        out[name] = d
    return out


def test_dropin_script_real_data_branch_eval_loop_and_checkpoint(tmp_path):
    """The branch of the drop-in script a user with the dataset takes (cifar_example_ddp.py:67-71 loaders, :90-114
    epoch loop with sampler.set_epoch, :118-119 checkpoint, :121-136 accuracy loop), on a small fake
    cifar-10-batches-py: uint8 images resident in HBM, normalised on the device, DistributedSampler order."""
    import subprocess
    import sys
    files = _fake_cifar(str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "ddp-tutorial_b200", "cifar_example_ddp.py"), "--model", "net", "--batch-size", "4",
                        "--epochs", "2", "--print-every", "10", "--data", str(tmp_path), "--save", str(tmp_path / "cifar_net.pth")],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "Finished Training" in out
    # 120 training images / batch 4 = 30 steps per epoch -> 3 print lines per epoch, 2 epochs
    assert len([l for l in out.splitlines() if l.startswith("[1,")]) == 3 and len([l for l in out.splitlines() if l.startswith("[2,")]) == 3, out
    acc_lines = [l for l in out.splitlines() if l.startswith("Accuracy of the network on the 10000 test images:")]
    assert len(acc_lines) == 1, out
    acc = int(acc_lines[0].split(":")[1].strip().rstrip("%").strip())
    assert 0 <= acc <= 100
    # the saved state_dict loads into the reference's module and reproduces the printed accuracy on the same test file
    sd = torch.load(tmp_path / "cifar_net.pth", map_location="cpu")
    assert list(sd) == [n for n, _ in O.NET_PARAM_SHAPES]
    x = torch.from_numpy(files["test_batch"]["data"].reshape(-1, 3, 32, 32)).float().div(255.0).sub(0.5).div(0.5)   # ToTensor + Normalize(.5,.5)
    y = torch.tensor(files["test_batch"]["labels"])
    with torch.no_grad():
        ref_acc = (O.net_forward({k: v.float() for k, v in sd.items()}, x).argmax(1) == y).float().mean().item()   # the reference's Net.forward
    assert abs(ref_acc * 100 - acc) <= 100.0 / len(y) + 1.0, (ref_acc, acc)        # printed with %d (truncation) + one borderline sample
