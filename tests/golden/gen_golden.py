"""Generates the golden fixtures in tests/golden/ from the UNMODIFIED reference (imported from /root/reference via
oracle/ref_import.py).  Run in the build container only:   python tests/golden/gen_golden.py

  cdf_table.npz        GaussianEncoder.update() tables of the reference (256x103 int32 CDFs, lengths, offsets)
  bf16_index_lut.npy   reference GaussianEncoder.build_indexes on all 65536 bf16 bit patterns (NaN -> 0)
  codec_keys.txt       decode-side state-dict keys + shapes of the reference IntraNoAR
  codec_128x128.npz    seed-0 weights (onedc_b200.weights.random_state_dict), synthetic 128x128 stream:
                       stream bytes, z indices, per-step indices/symbols, reference IntraNoAR.decode outputs
  rans_escape.npz      reference RansEncoder bytes for symbols with escapes (+-300) and the decoded symbols
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from onedc_b200 import weights as W                      # noqa: E402
from oracle.decode import CodecOracle                    # noqa: E402
from oracle.ref_import import build_reference_codec     # noqa: E402

torch.manual_seed(0)
ref = build_reference_codec()
q, l, o = ref.gaussian_encoder.get_cdf_info()
np.savez_compressed(os.path.join(HERE, "cdf_table.npz"), cdf=q, length=l, offset=o)

bits = torch.arange(65536, dtype=torch.int32)
vals = (bits << 16).view(torch.float32)
lut = ref.gaussian_encoder.build_indexes(torch.nan_to_num(vals, nan=0.0)).to(torch.uint8).numpy()
np.save(os.path.join(HERE, "bf16_index_lut.npy"), lut)

sd_ref = ref.state_dict()
with open(os.path.join(HERE, "codec_keys.txt"), "w") as f:
    for k, v in sd_ref.items():
        if not (k.startswith("enc.") or k.startswith("hyper_enc.")):
            f.write(f"{k} {'x'.join(str(s) for s in v.shape)}\n")

sd = W.random_state_dict(W.codec_spec(), 0)
missing, unexpected = ref.load_state_dict(sd, strict=False)
assert not unexpected and all(k.startswith(("enc.", "hyper_enc.")) for k in missing)
orc = CodecOracle(sd)
trace = []
stream, z_idx, y_hat = orc.make_stream(128, 128, seed=1234, trace=trace)
x_hat, y_sem, hw, phw, pad = ref.decode(stream=stream)          # the reference decodes the stream itself
np.savez_compressed(os.path.join(HERE, "codec_128x128.npz"),
                    stream=np.frombuffer(stream, dtype=np.uint8), z_idx=z_idx.numpy(),
                    idx=np.stack([t["idx"].reshape(-1).numpy().astype(np.int16) for t in trace]),
                    sym=np.stack([t["sym"].reshape(-1) for t in trace]),
                    x_hat=x_hat.numpy().astype(np.float16), y_sem=y_sem.numpy().astype(np.float32),
                    y_hat=y_hat.numpy().astype(np.float32))

g = np.random.default_rng(0)
n = 20000
idx = g.integers(0, 256, n).astype(np.int16)
scale = np.exp(np.linspace(np.log(0.11), np.log(64), 256))[idx]
sym = np.rint(g.standard_normal(n) * scale).astype(np.int16)
sym[::997] = 300
sym[5::991] = -300
ref.entropy_coder.reset()
ref.entropy_coder.encoder.encode_with_indexes(sym, idx, 0)
ref.entropy_coder.flush()
data = ref.entropy_coder.get_encoded_stream()
ref.entropy_coder.set_stream(data)
dec = ref.entropy_coder.decoder.decode_stream(idx, 0)
assert np.array_equal(dec, sym)
np.savez_compressed(os.path.join(HERE, "rans_escape.npz"), idx=idx, sym=sym, stream=np.frombuffer(data, dtype=np.uint8))
print("golden fixtures written:", sorted(os.listdir(HERE)))
