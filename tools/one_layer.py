"""One igemm layer, a few launches (timing / ncu captures):
    python tools/one_layer.py n h w cin cout k [reps] [res|geglu|stats ...]
  res: bf16 residual of the output's shape; geglu: GEGLU pair epilogue (cout = 2 x outputs); stats: fused GroupNorm sums"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import ops
n, h, w, cin, cout, k = [int(v) for v in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 and sys.argv[7].isdigit() else 10
flags = set(a for a in sys.argv[7:] if not a.isdigit())
dev = torch.device("cuda:0")
x = torch.randn((n, h, w, cin), device=dev).to(torch.bfloat16)
wt = torch.randn((cout, cin, k, k)) * (cin * k * k) ** -0.5
epi = ops.EPI_GEGLU if "geglu" in flags else ops.EPI_PLAIN
cw = ops.ConvW(wt, torch.zeros(cout), dev, epi=epi)
ncols = cout // 2 if epi != ops.EPI_PLAIN else cout
out = torch.empty((n, h, w, ncols), device=dev, dtype=torch.bfloat16)
res = torch.randn((n, h, w, ncols), device=dev).to(torch.bfloat16) if "res" in flags else None
kw = dict(out=out, res=res, stats=True if "stats" in flags else False)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.gn_arena_reset(dev)
    ops.igemm(x, cw, **kw)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    big.zero_()                                  # flush L2
    ops.gn_arena_reset(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.igemm(x, cw, **kw)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
us = sorted(ts)[len(ts) // 2]
print(f"{n}x{h}x{w} {cin}->{cout} k{k} {' '.join(sorted(flags))}: {us:.1f} us (median of {reps}, L2 flushed), "
      f"{2.0 * n * h * w * cin * cout * k * k / us / 1e6:.1f} TFLOP/s")
