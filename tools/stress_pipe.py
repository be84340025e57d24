"""Pipelined decode stress (several graphs in flight on separate streams): python tools/stress_pipe.py [reps] [depth] [H W]
Run it under `timeout`: a deadlock between concurrently running kernels shows up as a hang."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from onedc_b200 import lib
from onedc_b200.model import SD15_1step_codec_stage1
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (512, 768)
dev = torch.device("cuda:0")
L = lib.load()
model = SD15_1step_codec_stage1(state_dicts=bench._state_dicts(), device=dev)
model.codec_model.update(force=True)
streams = [model.codec_model.compress_synthetic(H, W, seed=100 + i)[0] for i in range(16)]
print("streams ready", flush=True)
try:
    for r in range(reps):
        t0 = time.time()
        imgs = model.decode_many(streams, depth=depth)
        torch.cuda.synchronize()
        print(f"rep {r}: {len(imgs)} images in {time.time() - t0:.2f} s", flush=True)
    print("completed")
except Exception as e:
    print("failed:", str(e).splitlines()[0])
