"""Where does a fused conv+BatchNorm launch spend its time?  CTA 0 of k_conv records clock64() at 13 points
(b2_conv_bn_test timeline13); this prints the deltas in microseconds for the ResNet-18 B=512 layer classes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ddp_tutorial_b200 as b2

DEV = 0
NAMES = ["entry", "setup done", "pdl_wait done", "first stage landed", "last MMA issued (tile 0)", "accumulator ready (tile 0)",
         "phase A done", "phase B done", "sums flushed (tile 0)", "all tiles stored", "grid barrier passed", "pass 2 done", "exit"]
SHAPES = [(512, 8, 8, 64, 64, 3, 1, 1), (512, 4, 4, 128, 128, 3, 1, 1), (512, 2, 2, 256, 256, 3, 1, 1), (512, 1, 1, 512, 512, 3, 1, 1),
          (512, 16, 16, 192, 64, 1, 1, 0)]
vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
for shape in SHAPES:
    N, Hi, Wi, Ci, Co, k, s, p = shape
    Ho, Wo = (Hi + 2 * p - k) // s + 1, (Wi + 2 * p - k) // s + 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Hi, Wi, Ci, generator=g).to(DEV).bfloat16()
    w = (torch.randn(Co, k, k, Ci, generator=g) / (k * k * Ci) ** 0.5).to(DEV)
    gamma, beta = torch.ones(Co, device=DEV), torch.zeros(Co, device=DEV)
    z = torch.empty(N, Ho, Wo, Co, device=DEV, dtype=torch.bfloat16); y = torch.empty_like(z)
    mean, invstd = torch.zeros(Co, device=DEV), torch.zeros(Co, device=DEV)
    info, tl = (C.c_int * 3)(), (C.c_longlong * 13)()
    for rep in range(3):
        b2._lib.call("b2_conv_bn_test", N, Hi, Wi, Ci, Co, k, s, p, 1, vp(x), vp(w), None, vp(gamma), vp(beta), vp(z), vp(y), vp(mean),
                     vp(invstd), info, tl, None)
    t = list(tl)
    print(f"shape {shape}: channel tile {info[0]}, grid {info[1]}, tiles {info[2]}")
    for i in range(1, 13):
        if t[i] and t[i - 1]:
            print(f"   {NAMES[i]:32s} +{(t[i] - t[i - 1]) / 1965.0:7.2f} us   (t = {(t[i] - t[0]) / 1965.0:7.2f} us)")
