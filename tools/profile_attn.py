import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import ops
dev = torch.device("cuda:0")
b, heads, d, s = 1, 8, 40, 9216
c = heads * d
qkv = (torch.randn((b, s, 3 * c), device=dev) * 1.0).to(torch.bfloat16)
out = torch.empty((b, s, c), device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], out, heads, d)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.attention(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], out, heads, d)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
