"""Per-launch table of the implicit-GEMM kernel over one device-resident decode step: shape, time (CUDA events around
each launch, launches queued behind a spin kernel so they run back to back), achieved TFLOP/s.
    python tools/layer_table.py [--size 768] [--batch 1] > gpurun_out/layer_table.txt"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onedc_b200 import bitstream, ops, weights as W          # noqa: E402
from onedc_b200.model import SD15_1step_codec_stage1         # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=768)
ap.add_argument("--batch", type=int, default=1)
args = ap.parse_args()
dev = torch.device("cuda:0")
sds = (W.random_state_dict(W.unet_spec(), 0), W.random_state_dict(W.codec_spec(), 0), W.random_state_dict(W.vae_spec(), 0))
model = SD15_1step_codec_stage1(state_dicts=sds, device=dev)
model.codec_model.update(force=True)
H = Wd = args.size
B = args.batch
streams = [model.codec_model.compress_synthetic(H, Wd, seed=1234 + i)[0] for i in range(B)]
hdr = [bitstream.decode_i(s) for s in streams]
trace = []
z_idx = model.codec_model.parse_z([d["bit_stream_z"] for d in hdr], H, Wd)
model.codec_model._decompress_batch([d["bit_stream_y"] for d in hdr], [d["bit_stream_z"] for d in hdr], H, Wd, trace)
syms = [t["sym"].view(B, 32, H // 16, Wd // 16).to(dev) for t in trace]
for _ in range(2):
    model.decode_resident(z_idx, syms)
torch.cuda.synchronize()
acc = collections.OrderedDict()
REP = 3
for rep in range(REP):
    ops.PROFILE, ops.PROFILE_INFO = [], []
    torch.cuda._sleep(int(0.15 * 1.9e9))
    model.decode_resident(z_idx, syms)
    torch.cuda.synchronize()
    prof = [(a.elapsed_time(b), f) for n, a, b, f in ops.PROFILE if n == "igemm"]
    info = ops.PROFILE_INFO
    ops.PROFILE, ops.PROFILE_INFO = None, None
    assert len(prof) == len(info)
    for i, ((ms, fl), d) in enumerate(zip(prof, info)):
        e = acc.setdefault(i, dict(d, ms=[], flops=fl))
        e["ms"].append(ms)
rows = []
for i, e in acc.items():
    ms = sorted(e["ms"])[len(e["ms"]) // 2]
    rows.append((i, e, ms))
tot = sum(r[2] for r in rows)
print(f"# {len(rows)} igemm launches, {tot:.3f} ms (event sum), {sum(r[1]['flops'] for r in rows) / 1e9:.1f} GFLOP")
print("# idx  n   h    w   cin  cout taps s epi st res | us      GFLOP   TFLOP/s")
for i, e, ms in rows:
    print(f"{i:4d} {e['n']:2d} {e['h']:4d} {e['w']:4d} {e['cin']:5d} {e['cout']:5d} {e['taps']:2d} {e['stride']} {e['epi']} {e['store']} "
          f"{int(e['res'])} | {ms * 1e3:7.1f} {e['flops'] / 1e9:8.2f} {e['flops'] / ms / 1e9:8.1f}")
# grouped by shape
grp = collections.OrderedDict()
for i, e, ms in rows:
    k = (e['n'], e['h'], e['w'], e['cin'], e['cout'], e['taps'], e['stride'], e['epi'], e['store'])
    g = grp.setdefault(k, [0, 0.0, 0.0])
    g[0] += 1
    g[1] += ms
    g[2] += e['flops']
print("\n# grouped by shape, sorted by total time")
print("# count  n    h    w   cin  cout taps s epi st | total_us  share  TFLOP/s")
for k, g in sorted(grp.items(), key=lambda kv: -kv[1][1]):
    print(f"{g[0]:4d}  {k[0]:2d} {k[1]:4d} {k[2]:4d} {k[3]:5d} {k[4]:5d} {k[5]:2d} {k[6]} {k[7]} {k[8]} | {g[1] * 1e3:8.1f} {g[1] / tot * 100:5.1f}% {g[2] / g[1] / 1e9:8.1f}")
