"""Per-launch time of small igemm problems inside a CUDA graph (100 launches per replay)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import ops

dev = torch.device("cuda:0")
cases = [  # (label, n,h,w,cin,cout,k)
    ("M128 K64 N64", 1, 1, 128, 64, 64, 1), ("M128 K512 N256", 1, 1, 128, 512, 256, 1),
    ("M2304 K128 N128", 1, 48, 48, 128, 128, 1), ("M2304 K512 N2048", 1, 48, 48, 512, 2048, 1),
    ("M2304 K1024 N512", 1, 48, 48, 1024, 512, 1), ("M9216 K320 N960", 1, 96, 96, 320, 960, 1),
    ("M9216 K320 N320 3x3", 1, 96, 96, 320, 320, 3), ("M2304 K640 N640 3x3", 1, 48, 48, 640, 640, 3),
    ("M576 K1280 N1280 3x3", 1, 24, 24, 1280, 1280, 3), ("M144 K2560 N1280 3x3", 1, 12, 12, 2560, 1280, 3),
    ("M144 K128 N768", 1, 12, 12, 128, 768, 1),
]
res = {}
for label, n, h, w, cin, cout, k in cases:
    x = torch.randn((n, h, w, cin), device=dev).to(torch.bfloat16)
    wt = torch.randn((cout, cin, k, k)) * (cin * k * k) ** -0.5
    cw = ops.ConvW(wt, torch.zeros(cout), dev)
    out = torch.empty((n, h, w, cout), device=dev, dtype=torch.bfloat16)
    ops.igemm(x, cw, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(100):
            ops.igemm(x, cw, out=out)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 500 * 1e3
    fl = 2.0 * n * h * w * cin * cout * k * k
    res[label] = {"us": round(us, 2), "TF/s": round(fl / us / 1e6, 1)}
    print(label, res[label], flush=True)
