"""Host rANS decode speed on a synthetic y stream (73 728 symbols per step as in a 768x768 image)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from onedc_b200.entropy_models import EntropyCoder, GaussianEncoder, StreamDecoder

ec = EntropyCoder(); ge = GaussianEncoder(); ge.update(force=True, entropy_coder=ec)
rng = np.random.default_rng(0)
n = 73728
for lo, hi, name in ((0, 95, "idx 0..94 (random-init prior)"), (0, 256, "idx 0..255"), (0, 30, "idx 0..29 (low entropy)")):
    idx = rng.integers(lo, hi, n).astype(np.int16)
    sc = np.exp(np.linspace(np.log(0.11), np.log(64), 256))[idx]
    sym = np.clip(np.rint(rng.standard_normal(n) * sc), -3000, 3000).astype(np.int16)
    ec.reset(); ec.encode_with_indexes_np(sym, idx, 0); ec.flush()
    data = ec.get_encoded_stream()
    out = np.empty_like(sym)
    best = 1e9
    for rep in range(30):
        sd = StreamDecoder(ec, data)
        t0 = time.perf_counter()
        sd.decode_into(idx.ctypes.data, n, out.ctypes.data)
        best = min(best, time.perf_counter() - t0)
    assert np.array_equal(out, sym)
    print(f"{name}: {len(data)} B ({8 * len(data) / n:.2f} bit/sym), decode {best * 1e6:.0f} us = {best / n * 1e9:.2f} ns/symbol")
