"""GroupNorm statistics/apply kernel timing at the pipeline's shapes (warm L2 where the tensor fits)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import ops, lib as L

dev = torch.device("cuda:0")
lib = L.load()
shapes = [(768, 768, 128), (384, 384, 256), (192, 192, 512), (96, 96, 320), (48, 48, 640), (24, 24, 1280), (12, 12, 2560)]
res = {}
for (h, w, c) in shapes:
    x = torch.randn((1, h, w, c), device=dev).to(torch.bfloat16)
    g = ops.GroupNorm(torch.ones(c), torch.zeros(c), 1e-6, device=dev)
    out = torch.empty_like(x)
    acc, cnt = ops._gn_buffers(dev)
    stats = torch.empty((1, 32, 2), device=dev, dtype=torch.float32)
    st = ops._stream()

    def f_stats():
        L.check(lib.onedc_groupnorm_stats(x.data_ptr(), c, c, 0, 0, 0, 0, 1, h * w, 32, 1e-6, acc.data_ptr(), stats.data_ptr(),
                                          cnt.data_ptr(), 0, st))

    def f_apply():
        L.check(lib.onedc_groupnorm_apply(x.data_ptr(), c, c, 0, 0, 0, 0, 1, h * w, 32, stats.data_ptr(), 0, 0, 1e-6, g.gamma.data_ptr(),
                                          g.beta.data_ptr(), 1, out.data_ptr(), c, st))

    for name, fn in (("stats", f_stats), ("apply", f_apply)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[f"{h}x{w}x{c} {name}"] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
print(os.environ.get("ONEDC_GN_BLOCKS", "default"), json.dumps(res))
