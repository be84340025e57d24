"""Where does an igemm CTA spend its time?  Per-role clock counters (onedc_igemm_set_debug) for one layer shape.
    python tools/igemm_roles.py n h w cin cout k [geglu] [res] [stats]"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import lib, ops

n, h, w, cin, cout, k = [int(v) for v in sys.argv[1:7]]
dev = torch.device("cuda:0")
x = torch.randn((n, h, w, cin), device=dev).to(torch.bfloat16)
wt = torch.randn((cout, cin, k, k)) * (cin * k * k) ** -0.5
flags = set(sys.argv[7:])
epi = ops.EPI_GEGLU if "geglu" in flags else ops.EPI_PLAIN
cw = ops.ConvW(wt, torch.zeros(cout), dev, epi=epi)
ncols = cout // 2 if epi != ops.EPI_PLAIN else cout
out = torch.empty((n, h, w, ncols), device=dev, dtype=torch.bfloat16)
res = torch.randn((n, h, w, ncols), device=dev).to(torch.bfloat16) if "res" in flags else None
_igemm = ops.igemm
ops_igemm = lambda x, cw, out: _igemm(x, cw, out=out, res=res, stats=True if "stats" in flags else False)
for _ in range(3):
    ops.gn_arena_reset(dev)
    ops_igemm(x, cw, out=out)
torch.cuda.synchronize()
dbg = torch.zeros((148, 16), device=dev, dtype=torch.int64)
lib.load().onedc_igemm_set_debug(C.c_void_p(dbg.data_ptr()))
ops.gn_arena_reset(dev)
ops_igemm(x, cw, out=out)
torch.cuda.synchronize()
lib.load().onedc_igemm_set_debug(C.c_void_p(0))
act = dbg[:, 4] > 0
d = dbg[act].double().mean(0).tolist()
print('active CTAs', int(act.sum()))
names = {0: "producer total", 1: "  waits free A slot", 2: "  waits free B slot", 4: "MMA issuer total", 5: "  waits A data",
         6: "  waits B data", 7: "  waits free accumulator", 8: "epilogue total", 9: "  waits accumulator",
         10: "  waits other splits", 11: "  split-K: park accumulators", 12: "  split-K: fence + barrier",
         13: "  split-K: arrive + wait splits", 14: "  split-K: reduce over splits", 15: "  split-K: epilogue + stats"}
print(f"{n}x{h}x{w} {cin}->{cout} k{k} {' '.join(sorted(flags))}  colmode={os.environ.get('ONEDC_COLMODE', '1')}  (mean clocks per CTA)")
for i, nm in names.items():
    print(f"  {nm:28s} {d[i]:10.0f}")
