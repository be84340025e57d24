"""Where do the softmax warps and the MMA warp of the attention kernel spend their clocks?
    python tools/attn_roles.py sq skv d [key_block]      (timed build of the kernel: onedc_attention_set_debug)"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import lib, ops
sq, skv, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
bkv = int(sys.argv[4]) if len(sys.argv) > 4 else 0
L = lib.load()
L.onedc_attention_set_plan(bkv, 0)
dev = torch.device("cuda:0")
heads = 8
c = heads * d
q = torch.randn((1, sq, c), device=dev).to(torch.bfloat16)
kv = torch.randn((1, skv, 2 * c), device=dev).to(torch.bfloat16)
o = torch.zeros((1, sq, c), device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(q, kv[:, :, :c], kv[:, :, c:], o, heads, d)
torch.cuda.synchronize()
nct = (sq + 127) // 128 * heads
dbg = torch.zeros((nct, 5, 8), device=dev, dtype=torch.int64)
L.onedc_attention_set_debug(C.c_void_p(dbg.data_ptr()))
ops.attention(q, kv[:, :, :c], kv[:, :, c:], o, heads, d)
torch.cuda.synchronize()
L.onedc_attention_set_debug(C.c_void_p(0))
dk16 = (d + 15) // 16 * 16
eff = bkv or (32 if 2 * 32 + dk16 + 16 <= 128 else 64)
nb = (skv + eff - 1) // eff
m = dbg.double().mean(0) / nb
names = {2: "mask / max / rescale", 0: "wait S(j+1) + issue its load", 4: "wait P buffer free", 3: "exp + pack + tcgen05.st",
         5: "wait::st + wait::ld", 6: "fence + arrive"}
print(f"sq={sq} skv={skv} d={d} key_block={eff}: {nct} CTAs, {nb} blocks; mean clocks per block, softmax warps 2..5 (timed build)")
for i, nm in names.items():
    print(f"  {nm:30s} " + " ".join(f"{m[w, i].item():8.0f}" for w in range(4)))
print(f"  {'total':30s} " + " ".join(f"{m[w, :7].sum().item():8.0f}" for w in range(4)))
print("MMA warp: " + ", ".join(f"{n} {m[4, i].item():.0f}" for i, n in enumerate(["wait K / V", "wait P event", "issue P V", "issue Q K^T"])))
