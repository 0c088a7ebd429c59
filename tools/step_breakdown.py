"""Where does the step go?  Times (CUDA events, graph replay) forward only, forward+loss+backward,
and the full step for ResNet-18 B=512 bf16, plus variants with the wgrad stream / PDL toggled."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddp_tutorial_b200 as b2

dev = 0
torch.manual_seed(0)
net = b2.resnet18(num_classes=10, precision="bf16").to(dev)
crit = b2.CrossEntropyLoss()
fuse = os.environ.get("FUSE", "1") == "1"
opt = b2.SGD(net.parameters(), lr=0.001, momentum=0.9, fuse_into_backward=fuse)
x = torch.randn(512, 3, 32, 32, device=dev); y = torch.randint(0, 10, (512,), device=dev)

def timeit(fn, n=200, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_cpu = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_cpu

def full():
    opt.zero_grad(); l = crit(net(x), y); l.backward(); opt.step()
def fwd():
    net(x)
def fwd_loss():
    crit(net(x), y)
for _ in range(5): full()
print("full step      gpu %.3f ms  cpu-enqueue %.3f ms" % timeit(full))
print("forward only   gpu %.3f ms  cpu-enqueue %.3f ms" % timeit(fwd))
print("forward+loss   gpu %.3f ms  cpu-enqueue %.3f ms" % timeit(fwd_loss))
