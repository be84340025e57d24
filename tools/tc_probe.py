"""First-contact probe for the tcgen05 kernels: a ladder of tiny cases, each reporting its error pattern.
Usage: python tools/tc_probe.py igemm|attention    (run under `timeout`: a protocol bug shows up as a hang)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from onedc_b200 import ops

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")


def mk(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


def report(name, out, ref):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs()
    tol = 2e-2 * ref.abs().max().item() + 1e-3
    bad = err > tol
    print(f"[{name}] max_err {err.max().item():.4g} ref_max {ref.abs().max().item():.4g} bad {int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any() and out.dim() >= 2:
        o2, b2 = out.reshape(-1, out.shape[-1]), bad.reshape(-1, out.shape[-1])
        rows = b2.any(1).nonzero().flatten()[:12].tolist()
        cols = b2.any(0).nonzero().flatten()[:24].tolist()
        print(f"    bad rows (first) {rows}  bad cols (first) {cols}")
        r = rows[0]
        print(f"    row {r}: out {o2[r, :8].tolist()} ref {ref.reshape(-1, out.shape[-1])[r, :8].tolist()}")
    return not bool(bad.any())


def igemm_ladder():
    ok = True
    for (m, k, n) in ((128, 64, 64), (128, 64, 16), (128, 128, 64), (128, 512, 256), (256, 64, 64), (1000, 320, 320), (128, 8, 32)):
        x = mk((1, 1, m, k), 1)
        w = mk((n, k), 2, k ** -0.5).float().cpu()
        cw = ops.ConvW(w, None, dev)
        out = ops.igemm(x, cw, impl=0)
        torch.cuda.synchronize()
        ref = x.float() @ cw.w[0, :, :k].float().t()
        ok &= report(f"gemm M{m} K{k} N{n}", out, ref)
    for (h, w_, c, co, s) in ((8, 16, 64, 64, 1), (16, 16, 64, 64, 1), (12, 12, 128, 128, 1), (16, 16, 128, 64, 2)):
        x = mk((1, h, w_, c), 3)
        w = mk((co, c, 3, 3), 4, (9 * c) ** -0.5).float().cpu()
        cw = ops.ConvW(w, None, dev)
        out = ops.igemm(x, cw, stride=s, impl=0)
        torch.cuda.synchronize()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), cw.w.float().reshape(3, 3, co, c).permute(2, 3, 0, 1).contiguous(),
                       stride=s, padding=1).permute(0, 2, 3, 1)
        ok &= report(f"conv3x3 {h}x{w_} C{c}->{co} s{s}", out, ref)
    print("IGEMM_PROBE", "PASS" if ok else "FAIL", flush=True)


def attention_ladder():
    ok = True
    for (b, heads, d, sq, skv) in ((1, 1, 64, 128, 64), (1, 1, 64, 128, 128), (1, 1, 64, 128, 512), (1, 2, 40, 128, 64),
                                   (1, 8, 40, 256, 144), (1, 8, 80, 256, 256), (1, 8, 160, 144, 144)):
        c = heads * d
        q, k, v = mk((b, sq, c), 1), mk((b, skv, c), 2), mk((b, skv, c), 3)
        out = torch.zeros((b, sq, c), device=dev, dtype=torch.bfloat16)
        ops.attention(q, k, v, out, heads, d, impl=0)
        torch.cuda.synchronize()
        qh, kh, vh = (t.float().view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
        ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, sq, c)
        ok &= report(f"attn b{b} h{heads} d{d} sq{sq} skv{skv}", out, ref)
    print("ATTN_PROBE", "PASS" if ok else "FAIL", flush=True)


if __name__ == "__main__":
    {"igemm": igemm_ladder, "attention": attention_ladder}[sys.argv[1]]()
