"""HBM roofline of the elementwise / entropy-index kernels on >= 256 MB synthetic instances (the in-pipeline
tensors are ~1 MB and L2-resident, so the achieved-bandwidth evidence is taken here).
Prints one JSON line per kernel: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json hbm_gbs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import ops
from onedc_b200.entropy_models import EntropyCoder, GaussianEncoder

dev = torch.device("cuda:0")
peak = 6547.2
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("hbm_gbs", peak)


def timeit(fn, iters=int(os.environ.get("EW_ITERS", "20"))):
    for _ in range(3 if iters > 1 else 1):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def report(name, nbytes, t, note=""):
    gbs = nbytes / t / 1e9
    print(json.dumps({"kernel": name, "algorithmic_MB": nbytes / 1e6, "us": t * 1e6, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak,
                      "note": note}), flush=True)


ge = GaussianEncoder()
ge.update(force=True, entropy_coder=EntropyCoder())
lut = ge.device_tables(dev)[0]
# 64 images of 1024x1024 latents (64x64 at 1/16): 64*64*64*128 scales = 33.5 M symbols/step ... use a big plane instead
n, h, w = 8, 1024, 1024                       # 8.4 M pixels x 128 ch bf16 = 2.1 GB of scales|means buffer halves
buf = (torch.randn((n, h, w, 256), device=dev) * 0.7).to(torch.bfloat16)
idx = torch.empty((n, 32, h, w), device=dev, dtype=torch.int16)
nsym = n * 32 * h * w
t = timeit(lambda: ops.scale_to_index(buf[..., :128], lut, 1, idx))
report("scale_to_index", nsym * 4, t, "2 B bf16 scale in + 2 B int16 index out per symbol")
sc = buf[..., :128].contiguous()                                    # scales as their own tensor (pixel stride 128)
t = timeit(lambda: ops.scale_to_index(sc, lut, 1, idx))
report("scale_to_index(ld=128)", nsym * 4, t, "same, scales tensor not interleaved with the means")
del sc
sym = torch.randint(-5, 6, (n, 32, h, w), device=dev, dtype=torch.int16)
yh = torch.zeros((n, h, w, 256), device=dev, dtype=torch.bfloat16)
t = timeit(lambda: ops.dequant_accum(sym, buf[..., 128:], yh[..., :128], 1))
report("dequant_accum", nsym * 6, t, "2 B symbol + 2 B mean in, 2 B y_hat out per symbol")
flat = buf.view(-1)[: 1 << 28]
out32 = torch.empty(flat.numel(), device=dev, dtype=torch.int32)
t = timeit(lambda: ge.build_indexes(flat))
report("build_indexes(bf16)", flat.numel() * 6, t, "generic API: 2 B in + 4 B int32 out (includes the output allocation)")
del buf, idx, sym, yh, flat, out32
torch.cuda.empty_cache()
x = torch.randn((1, 1536, 1536, 128), device=dev).to(torch.bfloat16)           # 604 MB
g = ops.GroupNorm(torch.ones(128), torch.zeros(128), 1e-6, device=dev)
o = torch.empty_like(x)
t = timeit(lambda: g(x, out=o))
report("groupnorm stats+apply+silu", x.numel() * 6, t, "2 B read (stats) + 2 B read + 2 B write (apply) per element")
dw = ops.DepthwiseW(torch.randn(128, 1, 3, 3) * 0.2, torch.zeros(128), dev)
t = timeit(lambda: ops.dwconv3x3(x, dw))
report("dwconv3x3", x.numel() * 4, t, "2 B in + 2 B out per element")
tok = torch.randn((64, 16384, 320), device=dev).to(torch.bfloat16)             # 671 MB
ln = ops.LayerNorm(torch.ones(320), torch.zeros(320), device=dev)
t = timeit(lambda: ln(tok))
report("layernorm", tok.numel() * 4, t, "2 B in + 2 B out per element (includes the output allocation)")
xs = torch.randn((1, 768, 768, 256), device=dev).to(torch.bfloat16)
t = timeit(lambda: ops.upsample2x(xs))
report("upsample2x", xs.numel() * 2 * 5, t, "2 B in + 4x2 B out per input element")
