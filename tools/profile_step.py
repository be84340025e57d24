"""One device-resident 768x768 decode step inside a cudaProfiler range (for `ncu --profile-from-start off`)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import bitstream, weights as W
from onedc_b200.model import SD15_1step_codec_stage1

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=768)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
sds = (W.random_state_dict(W.unet_spec(), 0), W.random_state_dict(W.codec_spec(), 0), W.random_state_dict(W.vae_spec(), 0))
model = SD15_1step_codec_stage1(state_dicts=sds, device=dev)
model.codec_model.update(force=True)
H = Wd = a.size
B = a.batch
z_idx = torch.randint(0, 16384, (B, H // 64, Wd // 64), dtype=torch.int32, device=dev)
syms = [torch.randint(-3, 4, (B, 32, H // 16, Wd // 16), dtype=torch.int16, device=dev) for _ in range(4)]
model.decode_resident(z_idx, syms)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    model.decode_resident(z_idx, syms)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", a.steps, "step(s)")
