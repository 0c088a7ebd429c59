"""Diagnostic (GPU box): per-tensor step-0 gradient error of the CUDA path and of the fp32
CPU oracle, both measured against the fp64 CPU oracle.  Usage:
    python tools/diag_grads.py resnet18 32 fp32 > gpurun_out/diag.txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddp_tutorial_b200 as b2
from oracle import ref_step as O

name, batch, prec = sys.argv[1], int(sys.argv[2]), sys.argv[3]
r64 = O.train(name, seed=0, batch=batch, steps=1, dtype=torch.float64)
r32 = O.train(name, seed=0, batch=batch, steps=1)
# torch's own bf16 (CPU autocast) as a noise yardstick
import torch.nn.functional as F
mac = O.make_model(name, 0)
with torch.autocast("cpu", dtype=torch.bfloat16):
    lac = F.cross_entropy(mac(r32["x"]).float(), r32["y"])
lac.backward()
gac = {n: p.grad.detach().double() for n, p in mac.named_parameters()} if name != "net" else None
torch.manual_seed(0)
net = getattr(b2, name)(num_classes=10, precision=prec).to(0) if name != "net" else b2.Net().to(0)
crit = b2.CrossEntropyLoss()
out = net(r32["x"].to(0))
loss = crit(out, r32["y"].to(0))
loss.backward()
print("loss mine %.7f  fp32 %.7f  fp64 %.9f" % (loss.item(), r32["loss"][0], r64["loss"][0]))
mine = {n: p.grad.detach().cpu().double() for n, p in net.named_parameters()}
print("autocast loss %.7f" % float(lac))
print("%-34s %10s %10s %10s %10s" % ("tensor", "|g|max", "mine/64", "cpu32/64", "autocast/64"))
for (n, g64), (_, g32) in zip(r64["grads0"], r32["grads0"]):
    s = g64.abs().max().item() + 1e-30
    e_m = (mine[n] - g64).abs().max().item() / s
    e_c = (g32.double() - g64).abs().max().item() / s
    flag = " <<<" if e_m > 10 * max(e_c, 1e-6) else ""
    e_a = (gac[n] - g64).abs().max().item() / s if gac else float("nan")
    print("%-34s %10.3e %10.3e %10.3e %10.3e%s" % (n, s, e_m, e_c, e_a, flag))
a = torch.cat([mine[n].flatten() for n, _ in r64["grads0"]]); b = torch.cat([g.flatten() for _, g in r64["grads0"]])
print("cosine(mine, fp64) = %.5f" % float((a @ b) / (a.norm() * b.norm())))
if gac:
    a = torch.cat([gac[n].flatten() for n, _ in r64["grads0"]])
    print("cosine(autocast, fp64) = %.5f" % float((a @ b) / (a.norm() * b.norm())))
