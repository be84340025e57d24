"""Time of the flash-attention kernel for the UNet shapes: python tools/attn_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import ops, lib
L = lib.load()
dev = torch.device("cuda:0")
for S, d, skv in ((9216, 40, 9216), (2304, 80, 2304), (576, 160, 576), (9216, 40, 144), (65536, 40, 65536)):
    heads = 8
    c = heads * d
    q = torch.randn((1, S, c), device=dev).to(torch.bfloat16)
    kv = torch.randn((1, skv, 2 * c), device=dev).to(torch.bfloat16)
    o = torch.empty((1, S, c), device=dev, dtype=torch.bfloat16)
    run = lambda: ops.attention(q, kv[:, :, :c], kv[:, :, c:], o, heads, d)
    for bkv in (0,):
        L.onedc_attention_set_plan(bkv, 0)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 50
        print(f"Sq={S} Skv={skv} d={d} key_block={bkv or 'default'}: {us:.1f} us, {4.0 * heads * S * skv * d / us / 1e6:.0f} TFLOP/s")
    L.onedc_attention_set_plan(0, 0)
