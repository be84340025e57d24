"""One small eager decode for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_decode.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onedc_b200 import weights as W
from onedc_b200.model import SD15_1step_codec_stage1

dev = torch.device("cuda:0")
sds = (W.random_state_dict(W.unet_spec(), 0), W.random_state_dict(W.codec_spec(), 0), W.random_state_dict(W.vae_spec(), 0))
model = SD15_1step_codec_stage1(state_dicts=sds, device=dev)
model.codec_model.update(force=True)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 192
stream, _ = model.codec_model.compress_synthetic(size, size, seed=7)
img = model.decode(stream=stream, stages={})          # eager launches, every kernel family of the path
torch.cuda.synchronize()
print("decoded", tuple(img.shape), bool(torch.isfinite(img).all()))
