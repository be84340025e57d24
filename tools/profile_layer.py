"""One igemm launch between cudaProfilerStart/Stop (ncu --profile-from-start off): same arguments as tools/one_layer.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import ops
n, h, w, cin, cout, k = [int(v) for v in sys.argv[1:7]]
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
for _ in range(2):
    ops.gn_arena_reset(dev)
    ops.igemm(x, cw, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.igemm(x, cw, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
