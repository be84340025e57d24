"""One igemm layer, a few launches (for ncu captures): python tools/one_layer.py n h w cin cout k [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import ops
n, h, w, cin, cout, k = [int(v) for v in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = torch.device("cuda:0")
x = torch.randn((n, h, w, cin), device=dev).to(torch.bfloat16)
wt = torch.randn((cout, cin, k, k)) * (cin * k * k) ** -0.5
cw = ops.ConvW(wt, torch.zeros(cout), dev)
out = torch.empty((n, h, w, cout), device=dev, dtype=torch.bfloat16)
for _ in range(reps):
    ops.igemm(x, cw, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.igemm(x, cw, out=out)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"{n}x{h}x{w} {cin}->{cout} k{k}: {us:.1f} us, {2.0 * n * h * w * cin * cout * k * k / us / 1e6:.1f} TFLOP/s")
