"""In-pipeline (warm L2, back-to-back) kernel times of one graph-replayed decode step via torch.profiler (CUPTI)."""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from onedc_b200 import weights as W
from onedc_b200.model import SD15_1step_codec_stage1

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=768)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--out", default="gpurun_out/kineto_step.json")
a = ap.parse_args()
dev = torch.device("cuda:0")
sds = (W.random_state_dict(W.unet_spec(), 0), W.random_state_dict(W.codec_spec(), 0), W.random_state_dict(W.vae_spec(), 0))
model = SD15_1step_codec_stage1(state_dicts=sds, device=dev)
model.codec_model.update(force=True)
H = Wd = a.size
B = a.batch
gd = model.graphed(B, H, Wd)
z_idx = torch.randint(0, 16384, (B, H // 64, Wd // 64), dtype=torch.int32, device=dev)
syms = [torch.randint(-3, 4, (B, 32, H // 16, Wd // 16), dtype=torch.int16, device=dev) for _ in range(4)]
gd.set_resident_inputs(z_idx, syms)
for _ in range(3):
    gd.run_resident()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        gd.run_resident()
    torch.cuda.synchronize()
tot = collections.defaultdict(float)
cnt = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.split("(")[0]
        tot[name] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        cnt[name] += 1
rows = sorted(tot.items(), key=lambda kv: -kv[1])
total = sum(tot.values())
print(f"total kernel time per step {total / 3 / 1e3:.3f} ms")
for k, v in rows[:16]:
    print(f"{k[:60]:60s} n/step={cnt[k] // 3:4d}  {v / 3 / 1e3:8.3f} ms  {100 * v / total:5.1f}%")
trace = a.out.replace(".json", "_trace.json")
prof.export_chrome_trace(trace)
ev = [e for e in json.load(open(trace))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
per = len(ev) // 3
rows2 = [{"i": i, "name": e["name"].split("(")[0][:48], "us": e["dur"], "grid": e["args"].get("grid"), "block": e["args"].get("block")}
         for i, e in enumerate(ev[2 * per:])]
json.dump(rows2, open(a.out.replace(".json", "_launches.json"), "w"))
os.remove(trace)
json.dump({k: {"ms_per_step": v / 3 / 1e3, "launches_per_step": cnt[k] // 3} for k, v in rows}, open(a.out, "w"), indent=1)
